// nlookup sum-check prover: fused MLE sweeps over HBM with a device-side Fiat-Shamir sponge.
//
// Reference behaviour reproduced bit-for-bit:
//   gen_eq_table              /root/reference/src/backend/r1cs_helper.rs:508-544
//   linear_mle_product        /root/reference/src/backend/r1cs_helper.rs:441-506
//   prover_mle_partial_eval   /root/reference/src/backend/r1cs_helper.rs:551-634
//   wit_nlookup_gadget        /root/reference/src/backend/r1cs.rs:2260-2392
//
// Schedule (B200-first; none of this shape exists in the reference, only its results do):
//   * The dense part of the eq table is a tensor product  EQ[i] = A[i >> h] * B[i & (2^h-1)]
//     (h = 10).  It is NEVER materialised: every pass streams only the T table and takes
//     row-wise inner products with the L2-resident 32 KiB B table; the tiny A table is folded
//     by the transcript kernel each round.  Per round:
//         U0[row] = sum_lo T[row,lo] B[lo],  U1[row] = sum_lo T[row + half,lo] B[lo]
//         const = sum A0 U0,  g(1) = sum A1 U1,  xsq = sum (A1-A0)(U1-U0),  x = g(1)-const-xsq
//   * The m "lookup" points  rs[k] * [i == q_k]  of the eq table are carried as a sparse list
//     and added analytically to each round's coefficients (m gathers per round).
//   * Pass i+1 folds T with r_i and accumulates round i+1 in the same sweep (reads 2 elements,
//     writes 1 per output), so HBM traffic is ~4N elements instead of the reference-shaped 10N.
//   * When the live table reaches 2^h entries, one CTA finishes all remaining rounds out of
//     shared memory.
//   * Fiat-Shamir runs on device (warp-cooperative Poseidon), so a whole nlookup is a chain of
//     stream-ordered launches with no host round trip.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <memory>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "p2p.cuh"

namespace reef {

static constexpr int H_BITS = 10;
static constexpr int CHUNK = 1 << H_BITS;  // b-values per CTA in a sweep
static constexpr int MAX_ELL = 48;

// Device-resident state of one nlookup session.
struct NlState {
  u32 sponge[5][12];     // SAFE sponge state between kernels: 5 x 10 lazy 29-bit limbs, plain residues (poseidon_lp.cuh)
  Fq r_mont;             // challenge of the last finished round (Montgomery form)
  Fq rm_mont;            // rs[m] = claim_r^(m+1)
  Fq lq_mont[MAX_ELL];   // last_q[j]  (bit j of the table index), Montgomery form
  Fq out_claim_r;        // canonical outputs
  Fq out_rounds[MAX_ELL][4];
  Fq out_last_claim;
  Fq out_next_v;
  Fq a_scalar;           // A table when it has a single entry (ell <= h)
};

// ---------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ Fq fq_from_u32(uint32_t x) {
  Fq r = fe_zero<FqCfg>();
  r.v[0] = x;
  return r;
}

template <bool U32IN>
__device__ __forceinline__ Fq load_t(const void* t, uint64_t idx) {
  if constexpr (U32IN) return fq_from_u32(((const uint32_t*)t)[idx]);
  else return ld256((const Fq*)t + idx);
}

// x0 + r * (x1 - x0), x canonical, r Montgomery -> canonical
__device__ __forceinline__ Fq fold_one(const Fq& x0, const Fq& x1, const Fq& r_mont) {
  return fe_add<FqCfg>(x0, mont_mul<FqCfg>(r_mont, fe_sub<FqCfg>(x1, x0)));
}

// Block-wide sum (mod p) of NV field elements per thread; result valid on thread 0.
template <int NV, int NTHREADS>
__device__ __forceinline__ void block_sum(Fq* vals, Fq* smem /* NV * NTHREADS/32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) vals[k] = warp_sum_fe<FqCfg>(vals[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) smem[warp * NV + k] = vals[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll 1
    for (int w = 1; w < NTHREADS / 32; w++) {
#pragma unroll
      for (int k = 0; k < NV; k++) vals[k] = fe_add<FqCfg>(vals[k], smem[w * NV + k]);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// transcript plumbing shared by the single-CTA kernels (all of them run LP_PERM_THREADS threads):
// the sponge state lives in shared memory (LpPermShared::S) while a kernel runs and in
// NlState::sponge between kernels.
// ---------------------------------------------------------------------------------------
static constexpr int TR_THREADS = LP_PERM_THREADS;

// Launch as a programmatic dependent of the stream predecessor (see pdl_wait / pdl_trigger in common.cuh): the launch
// latency and the kernel's constant-only prologue overlap the tail of the predecessor.  REEF_PDL=0 turns it off.
static bool pdl_enabled() {
  static const bool on = !(getenv("REEF_PDL") && atoi(getenv("REEF_PDL")) == 0);
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

__device__ __forceinline__ void tr_load_state(LpPermShared* sh, const NlState* st) {
  if (threadIdx.x < 60) sh->S[threadIdx.x / 12][threadIdx.x % 12] = st->sponge[threadIdx.x / 12][threadIdx.x % 12];
}
__device__ __forceinline__ void tr_store_state(const LpPermShared* sh, NlState* st) {
  if (threadIdx.x < 60) st->sponge[threadIdx.x / 12][threadIdx.x % 12] = sh->S[threadIdx.x / 12][threadIdx.x % 12];
}

// One sum-check round of the transcript: absorb [const, x, xsq] at rate positions 0..2, permute,
// squeeze r (r1cs_helper.rs:479-488).  On entry THREAD 0 holds con, g1 = g(1), xsq (canonical);
// every thread gets r in Montgomery form; thread 0 records the round in st->out_rounds[ri].
struct TrScratch {
  Fq e[3];
  Fq r_mont;
};
__device__ __forceinline__ Fq tr_round(LpPermShared* sh, TrScratch* ts, const PoseidonLpTables* T, u32& seq, NlState* st,
                                       uint32_t ri, const Fq& con, const Fq& g1, const Fq& xsq) {
  if (threadIdx.x == 0) {
    ts->e[0] = con;
    ts->e[1] = fe_sub<FqCfg>(fe_sub<FqCfg>(g1, con), xsq);
    ts->e[2] = xsq;
  }
  __syncthreads();
  if (threadIdx.x < 27) sh->S[1 + threadIdx.x / 9][threadIdx.x % 9] += lp_limb_of(ts->e[threadIdx.x / 9].v, threadIdx.x % 9);
  __syncthreads();
  poseidon_permute_lp(sh, T, ++seq);
  if (threadIdx.x == 0) {
    const Fq r = lp_squeeze(sh, 1);
    const Fq rm = to_mont<FqCfg>(r);
    ts->r_mont = rm;
    st->r_mont = rm;
    st->out_rounds[ri][0] = r;
    st->out_rounds[ri][1] = ts->e[2];
    st->out_rounds[ri][2] = ts->e[1];
    st->out_rounds[ri][3] = ts->e[0];
  }
  __syncthreads();
  return ts->r_mont;
}

// ---------------------------------------------------------------------------------------
// k_nl_begin: first absorb + claim_r, powers of claim_r, sparse list, selector table
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TR_THREADS) k_nl_begin(NlState* st, const Fq* __restrict__ query, uint32_t n_query, Fq tag_canon,
                                                         const Fq* __restrict__ prev_q, uint32_t ell, const uint64_t* __restrict__ q,
                                                         uint32_t m, uint64_t* __restrict__ sp_pos, Fq* __restrict__ sp_w,
                                                         const PoseidonLpTables* __restrict__ T, uint32_t rank, uint32_t world) {
  __shared__ LpPermShared sh;
  lp_perm_init(&sh, T);
  pdl_wait();      // predecessor complete and visible; only then let the successor become resident (at most two
  pdl_trigger();   // consecutive kernels of the chain are ever co-resident)
  u32 seq = 0;
  if (threadIdx.x < 9) sh.S[0][threadIdx.x] = lp_limb_of(tag_canon.v, threadIdx.x);
  // last_q[j] = prev_running_q[ell-1-j]   (r1cs.rs:2318-2319 passes the reversed vector); off the transcript's path
  for (uint32_t j = threadIdx.x; j < ell; j += TR_THREADS) st->lq_mont[j] = to_mont<FqCfg>(ld256(prev_q + (ell - 1 - j)));
  for (uint32_t e0 = 0; e0 < n_query; e0 += 4) {
    const uint32_t cnt = n_query - e0 < 4 ? n_query - e0 : 4;
    __syncthreads();
    if (e0) poseidon_permute_lp(&sh, T, ++seq);      // absorb position wrapped: permute before the next four
    if (threadIdx.x < 9 * cnt) {
      const Fq x = ld256(query + e0 + threadIdx.x / 9);
      sh.S[1 + threadIdx.x / 9][threadIdx.x % 9] += lp_limb_of(x.v, threadIdx.x % 9);
    }
  }
  __syncthreads();
  poseidon_permute_lp(&sh, T, ++seq);                // squeeze(1): always permutes after an absorb
  tr_store_state(&sh, st);
  if (threadIdx.x == 0) {
    const Fq claim_c = lp_squeeze(&sh, 1);
    st->out_claim_r = claim_c;
    const Fq claim = to_mont<FqCfg>(claim_c);
    Fq pw = claim;                // rs[0] = claim_r
    for (uint32_t k = 0; k < m; k++) {
      // sharded: rank g owns the indices with (q mod world) == g, at local position q / world
      sp_w[k] = (q[k] % world == rank) ? pw : fe_zero<FqCfg>();
      sp_pos[k] = q[k] / world;
      pw = mont_mul<FqCfg>(pw, claim);
    }
    st->rm_mont = pw;             // rs[m]
  }
}

// ---------------------------------------------------------------------------------------
// k_eq_tables: A[hi] = rs[m] * prod_{j>=h} sel(bit_{j-h}(hi), lq[j]),  B[lo] = prod_{j<h} sel(bit_j(lo), lq[j])
// (both in Montgomery form).  When ell <= h, A is the single entry rs[m] and B spans all bits.
// ---------------------------------------------------------------------------------------
// Sharded use: local index bit t pairs with lq[bit_off + t] and A carries the rank factor
// c_g = prod_{t < bit_off} sel(bit_t(rank), lq[t]).
__global__ void k_eq_tables(const NlState* __restrict__ st, uint32_t ell, uint32_t hb, Fq* __restrict__ A,
                            uint64_t a_len, Fq* __restrict__ B, uint64_t b_len, uint32_t bit_off, uint32_t rank) {
  pdl_wait();      // predecessor complete and visible; only then let the successor become resident (at most two
  pdl_trigger();   // consecutive kernels of the chain are ever co-resident)
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a_len + b_len) return;
  const Fq one = fe_one<FqCfg>();
  const Fq* lqs = st->lq_mont + bit_off;
  if (i < a_len) {
    Fq acc = st->rm_mont;
    for (uint32_t t = 0; t < bit_off; t++) {
      Fq lq = st->lq_mont[t];
      acc = mont_mul<FqCfg>(acc, ((rank >> t) & 1) ? lq : fe_sub<FqCfg>(one, lq));
    }
    for (uint32_t j = hb; j < ell; j++) {
      Fq lq = lqs[j];
      Fq f = ((i >> (j - hb)) & 1) ? lq : fe_sub<FqCfg>(one, lq);
      acc = mont_mul<FqCfg>(acc, f);
    }
    A[i] = acc;
  } else {
    uint64_t lo = i - a_len;
    Fq acc = one;
    for (uint32_t j = 0; j < hb; j++) {
      Fq lq = lqs[j];
      Fq f = ((lo >> j) & 1) ? lq : fe_sub<FqCfg>(one, lq);
      acc = mont_mul<FqCfg>(acc, f);
    }
    B[lo] = acc;
  }
}

// ---------------------------------------------------------------------------------------
// k_sweep: (optional fold with r) + row-wise inner products with B.
//   L_in  length of Tin; L = FOLD ? L_in/2 : L_in is the accumulation length, half = L/2 >= 2^h.
//   One WARP per task = 32*ppl consecutive index pairs b (never straddles a 2^h row, ppl <= 32):
//   lane <-> b (coalesced 1 KiB requests), each lane keeps lazily-accumulated 17-limb sums
//   U0 += T'[b] B[lo], U1 += T'[b+half] B[lo] over its ppl pairs.  The warp total is taken
//   with REDUX on 16-bit half-limbs (2 instructions per limb instead of a 5-level shuffle tree
//   of 256-bit modular additions), ONE lane per task reduces mod p and applies the A factors,
//   and the CTA emits one (const, g(1), xsq) triple:  partials[blk][3] (canonical).
// ---------------------------------------------------------------------------------------
static constexpr int SWEEP_WARPS = 4;

// canonical a_mont * s / R for a single-limb s: 8 products + one Montgomery reduction
__device__ __forceinline__ Fq mont_mul_u32(const Fq& a_mont, uint32_t s) {
  u32 T[16], od[9];
#pragma unroll
  for (int i = 0; i < 16; i++) T[i] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) od[i] = 0;
  mad_row4(T, a_mont.v[0], a_mont.v[2], a_mont.v[4], a_mont.v[6], s);
  mad_row4(od, a_mont.v[1], a_mont.v[3], a_mont.v[5], a_mont.v[7], s);
  u32 sh[9];
  sh[0] = 0;
#pragma unroll
  for (int i = 1; i < 9; i++) sh[i] = od[i - 1];
  acc_add<9>(T, sh);
  Fq r;
  mont_reduce<FqCfg>(r.v, T);
  return r;
}

// x0 + r (x1 - x0) for document codes (single-limb operands): canonical
__device__ __forceinline__ Fq fold_one_u32(uint32_t x0, uint32_t x1, const Fq& r_mont) {
  const bool neg = x1 < x0;
  Fq p = mont_mul_u32(r_mont, neg ? x0 - x1 : x1 - x0);
  Fq a = fq_from_u32(x0);
  return neg ? fe_sub<FqCfg>(a, p) : fe_add<FqCfg>(a, p);
}

// w (10 limbs) += t * b for a single-limb t; capacity 2^32 products
struct Wide10 {
  u32 v[10];
};
__device__ __forceinline__ void wide10_mac_small(Wide10& w, uint32_t t, const u32* b) {
  u32 pr[9], od[9];
#pragma unroll
  for (int i = 0; i < 9; i++) { pr[i] = 0; od[i] = 0; }
  mad_row4(pr, b[0], b[2], b[4], b[6], t);
  mad_row4(od, b[1], b[3], b[5], b[7], t);
  u32 sh[9];
  sh[0] = 0;
#pragma unroll
  for (int i = 1; i < 9; i++) sh[i] = od[i - 1];
  acc_add<9>(pr, sh);
  w.v[9] += acc_add<9>(w.v, pr);
}

// warp total of an NL-limb lazy accumulator, written by lane 0 as per-limb 64-bit column sums
template <int NL>
__device__ __forceinline__ void warp_limb_sums(const u32* v, unsigned long long* out /* shared, NL entries */) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    const u32 lo = __reduce_add_sync(0xffffffffu, v[i] & 0xffffu);
    const u32 hi = __reduce_add_sync(0xffffffffu, v[i] >> 16);
    if ((threadIdx.x & 31) == 0) out[i] = (unsigned long long)lo + ((unsigned long long)hi << 16);
  }
}

template <bool U32IN, bool FOLD>
__global__ void __launch_bounds__(SWEEP_WARPS * 32)
k_sweep(const void* Tin, uint64_t L_in, Fq* Tout, const NlState* __restrict__ st,
        const Fq* __restrict__ A, const Fq* __restrict__ B, Fq* __restrict__ partials, uint32_t ppl) {
  constexpr bool SMALL = U32IN && !FOLD;       // single-limb table entries: 10-limb accumulators
  constexpr int NL = SMALL ? 10 : 17;
  __shared__ unsigned long long cols[SWEEP_WARPS][2][17];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t L = FOLD ? (L_in >> 1) : L_in;
  const uint64_t half = L >> 1;
  const uint64_t task = (uint64_t)blockIdx.x * SWEEP_WARPS + warp;
  const uint64_t base = task * 32u * ppl;
  pdl_wait();      // predecessor complete and visible; only then let the successor become resident (at most two
  pdl_trigger();   // consecutive kernels of the chain are ever co-resident)
  Fq r;
  if constexpr (FOLD) r = st->r_mont;
  typename std::conditional<SMALL, Wide10, Wide17>::type acc0, acc1;
#pragma unroll
  for (int i = 0; i < NL; i++) { acc0.v[i] = 0; acc1.v[i] = 0; }
#pragma unroll 2
  for (uint32_t k = 0; k < ppl; k++) {
    const uint64_t b = base + (uint64_t)k * 32 + lane;
    const uint32_t lo = (uint32_t)(b & (CHUNK - 1));
    const Fq bv = ld256(B + lo);
    if constexpr (SMALL) {
      const uint32_t t0 = ((const uint32_t*)Tin)[b], t1 = ((const uint32_t*)Tin)[b + half];
      wide10_mac_small(acc0, t0, bv.v);
      wide10_mac_small(acc1, t1, bv.v);
    } else {
      Fq t0, t1;
      if constexpr (U32IN) {
        const uint32_t* T = (const uint32_t*)Tin;
        t0 = fold_one_u32(T[b], T[b + L], r);
        t1 = fold_one_u32(T[b + half], T[b + half + L], r);
      } else if constexpr (FOLD) {
        const Fq* T = (const Fq*)Tin;
        Fq x00 = ld256(T + b), x01 = ld256(T + b + L);
        Fq x10 = ld256(T + b + half), x11 = ld256(T + b + half + L);
        t0 = fold_one(x00, x01, r);
        t1 = fold_one(x10, x11, r);
      } else {
        const Fq* T = (const Fq*)Tin;
        t0 = ld256(T + b);
        t1 = ld256(T + b + half);
      }
      if constexpr (FOLD) {
        st256(Tout + b, t0);
        st256(Tout + b + half, t1);
      }
      wide_mac(acc0, t0.v, bv.v);
      wide_mac(acc1, t1.v, bv.v);
    }
  }
  warp_limb_sums<NL>(acc0.v, cols[warp][0]);
  warp_limb_sums<NL>(acc1.v, cols[warp][1]);
  __syncthreads();
  if (warp == 0) {
    const int t = lane < SWEEP_WARPS ? lane : 0;
    Wide17 W0, W1;
    unsigned long long c0 = 0, c1 = 0;
#pragma unroll
    for (int i = 0; i < 17; i++) {
      if (i < NL) { c0 += cols[t][0][i]; c1 += cols[t][1][i]; }
      W0.v[i] = (u32)c0; c0 >>= 32;
      W1.v[i] = (u32)c1; c1 >>= 32;
    }
    const Fq u0 = wide_reduce_div_R<FqCfg>(W0);   // sum T*B  (B carries the Montgomery factor)
    const Fq u1 = wide_reduce_div_R<FqCfg>(W1);
    const uint64_t row = (((uint64_t)blockIdx.x * SWEEP_WARPS + t) * 32u * ppl) >> H_BITS;
    const Fq a0 = ld256(A + row), a1 = ld256(A + row + (half >> H_BITS));
    Fq p[3];
    p[0] = mont_mul<FqCfg>(a0, u0);
    p[1] = mont_mul<FqCfg>(a1, u1);
    p[2] = mont_mul<FqCfg>(fe_sub<FqCfg>(a1, a0), fe_sub<FqCfg>(u1, u0));
#pragma unroll
    for (int k = 0; k < 3; k++) {
      Fq v = lane < SWEEP_WARPS ? p[k] : fe_zero<FqCfg>();
#pragma unroll
      for (int msk = 1; msk < SWEEP_WARPS; msk <<= 1) v = fe_add<FqCfg>(v, shfl_xor_fe(v, msk));
      if (lane == 0) st256(partials + (uint64_t)blockIdx.x * 3 + k, v);
    }
  }
}

// largest number of CTA partial triples any sweep over a table of n entries can emit
static uint64_t sweep_max_blocks(uint64_t n) {
  uint64_t b = (n / 2) / (32 * SWEEP_WARPS);
  return b ? b : 1;
}

template <bool U32IN, bool FOLD>
static int launch_sweep(reef_ctx* c, const void* Tin, uint64_t L_in, Fq* Tout, const NlState* st, const Fq* A,
                        const Fq* B, Fq* partials, uint32_t* nblk_out) {
  const uint64_t L = FOLD ? (L_in >> 1) : L_in;
  const uint64_t half = L >> 1;                  // >= 2^h
  // pairs per lane: as deep as possible (amortises the per-task reduction) while keeping at
  // least ~8 warps per SM in flight (integer-issue bound: 2 warps per scheduler already
  // saturate it, tools/bench_lat.cu; measured best of {2,4,8,14} with tools/sweep_probe.py)
  const uint64_t want_tasks = (uint64_t)c->sm_count * 8;
  uint32_t ppl = 1;
  while (ppl < 32 && half / (32ull * (ppl * 2)) >= want_tasks) ppl *= 2;
  const uint64_t tasks = half / (32ull * ppl);   // >= 32, a multiple of SWEEP_WARPS
  const uint32_t nblk = (uint32_t)(tasks / SWEEP_WARPS);
  REEF_CUDA(launch_dep(k_sweep<U32IN, FOLD>, dim3(nblk), dim3(SWEEP_WARPS * 32), 0, c->stream, Tin, L_in, Tout, st, A, B, partials, ppl));
  *nblk_out = nblk;
  REEF_LAUNCHED();
  return REEF_OK;
}

// ---------------------------------------------------------------------------------------
// k_round: finish round `ri` (0-based): sum CTA partials, add the sparse-point terms, run the
// transcript (absorb [const, x, xsq], squeeze r), fold A, advance the sparse list.
// ---------------------------------------------------------------------------------------
static constexpr int ROUND_THREADS = TR_THREADS;

template <bool U32IN>
__global__ void __launch_bounds__(ROUND_THREADS)
k_round(NlState* st, const Fq* __restrict__ partials, uint32_t nblk, const void* __restrict__ Tcur, uint64_t L,
        Fq* A, uint64_t a_len, uint64_t* sp_pos, Fq* sp_w, uint32_t m, uint32_t ri,
        const PoseidonLpTables* __restrict__ K) {
  __shared__ LpPermShared sh;
  __shared__ TrScratch ts;
  __shared__ Fq red[3 * ROUND_THREADS / 32];
  lp_perm_init(&sh, K);
  pdl_wait();      // predecessor complete and visible; only then let the successor become resident (at most two
  pdl_trigger();   // consecutive kernels of the chain are ever co-resident)
  tr_load_state(&sh, st);
  const uint64_t half = L >> 1;
  Fq acc[3];
  acc[0] = acc[1] = acc[2] = fe_zero<FqCfg>();
  for (uint32_t i = threadIdx.x; i < nblk; i += ROUND_THREADS) {
#pragma unroll
    for (int k = 0; k < 3; k++) acc[k] = fe_add<FqCfg>(acc[k], ld256(partials + (uint64_t)i * 3 + k));
  }
  // sparse points: E = w at index pos.  (t0,t1) = T[b], T[b+half]; e0 = top?0:w, e1 = top?w:0.
  for (uint32_t k = threadIdx.x; k < m; k += ROUND_THREADS) {
    uint64_t pos = sp_pos[k];
    bool top = pos >= half;
    uint64_t b = top ? pos - half : pos;
    Fq w = sp_w[k];
    Fq t0 = load_t<U32IN>(Tcur, b), t1 = load_t<U32IN>(Tcur, b + half);
    Fq wt = mont_mul<FqCfg>(w, top ? t1 : t0);        // canonical
    Fq wd = mont_mul<FqCfg>(w, fe_sub<FqCfg>(t1, t0));
    if (top) {
      acc[1] = fe_add<FqCfg>(acc[1], wt);             // g(1) += t1 * w
      acc[2] = fe_add<FqCfg>(acc[2], wd);             // xsq  += (t1-t0) * w
    } else {
      acc[0] = fe_add<FqCfg>(acc[0], wt);             // const += t0 * w
      acc[2] = fe_sub<FqCfg>(acc[2], wd);             // xsq  += (t1-t0) * (-w)
    }
  }
  block_sum<3, ROUND_THREADS>(acc, red);
  u32 seq = 0;
  const Fq r = tr_round(&sh, &ts, K, seq, st, ri, acc[0], acc[1], acc[2]);
  tr_store_state(&sh, st);
  // fold the A table over its top bit (in place: entry x only depends on x and x + a_len/2)
  if (a_len > 1) {
    const uint64_t ah = a_len >> 1;
    for (uint64_t x = threadIdx.x; x < ah; x += ROUND_THREADS) {
      Fq lo = A[x], hi = A[x + ah];
      A[x] = fe_add<FqCfg>(lo, mont_mul<FqCfg>(r, fe_sub<FqCfg>(hi, lo)));
    }
  }
  for (uint32_t k = threadIdx.x; k < m; k += ROUND_THREADS) {
    uint64_t pos = sp_pos[k];
    bool top = pos >= half;
    Fq f = top ? r : fe_sub<FqCfg>(fe_one<FqCfg>(), r);
    sp_w[k] = mont_mul<FqCfg>(sp_w[k], f);
    sp_pos[k] = top ? pos - half : pos;
  }
}

// ---------------------------------------------------------------------------------------
// k_tail: one CTA finishes the sum-check out of shared memory (live length <= 2^h).
// ---------------------------------------------------------------------------------------
static constexpr int TAIL_THREADS = TR_THREADS;

template <bool U32IN>
__global__ void __launch_bounds__(TAIL_THREADS)
k_tail(NlState* st, const void* __restrict__ Tin, uint64_t L_in, int do_fold, const Fq* __restrict__ A,
       const Fq* __restrict__ B, const uint64_t* __restrict__ sp_pos, const Fq* __restrict__ sp_w, uint32_t m,
       uint32_t ri0, const PoseidonLpTables* __restrict__ K) {
  extern __shared__ __align__(32) unsigned char tail_smem[];
  Fq* Ts = reinterpret_cast<Fq*>(tail_smem);   // canonical
  Fq* Es = Ts + CHUNK;                         // Montgomery form
  Fq* red = Es + CHUNK;                        // 3 * TAIL_THREADS/32
  __shared__ LpPermShared sh;
  __shared__ TrScratch ts;
  lp_perm_init(&sh, K);
  pdl_wait();      // predecessor complete and visible; only then let the successor become resident (at most two
  pdl_trigger();   // consecutive kernels of the chain are ever co-resident)
  tr_load_state(&sh, st);
  uint64_t L = do_fold ? (L_in >> 1) : L_in;   // <= CHUNK
  const Fq a0 = A[0];
  const Fq rf = st->r_mont;
  for (uint64_t b = threadIdx.x; b < L; b += TAIL_THREADS) {
    Fq t = do_fold ? fold_one(load_t<U32IN>(Tin, b), load_t<U32IN>(Tin, b + L), rf) : load_t<U32IN>(Tin, b);
    Ts[b] = t;
    Es[b] = mont_mul<FqCfg>(a0, B[b]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (uint32_t k = 0; k < m; k++) Es[sp_pos[k]] = fe_add<FqCfg>(Es[sp_pos[k]], sp_w[k]);
  }
  __syncthreads();
  uint32_t ri = ri0;
  u32 seq = 0;
  Fq r = fe_zero<FqCfg>();
  while (L > 1) {
    const uint64_t half = L >> 1;
    Fq acc[3];
    acc[0] = acc[1] = acc[2] = fe_zero<FqCfg>();
    for (uint64_t b = threadIdx.x; b < half; b += TAIL_THREADS) {
      Fq t0 = Ts[b], t1 = Ts[b + half], e0 = Es[b], e1 = Es[b + half];
      acc[0] = fe_add<FqCfg>(acc[0], mont_mul<FqCfg>(e0, t0));
      acc[1] = fe_add<FqCfg>(acc[1], mont_mul<FqCfg>(e1, t1));
      acc[2] = fe_add<FqCfg>(acc[2], mont_mul<FqCfg>(fe_sub<FqCfg>(e1, e0), fe_sub<FqCfg>(t1, t0)));
    }
    block_sum<3, TAIL_THREADS>(acc, red);
    r = tr_round(&sh, &ts, K, seq, st, ri, acc[0], acc[1], acc[2]);
    for (uint64_t b = threadIdx.x; b < half; b += TAIL_THREADS) {
      Fq t0 = Ts[b], t1 = Ts[b + half], e0 = Es[b], e1 = Es[b + half];
      Fq tn = fold_one(t0, t1, r);
      Fq en = fe_add<FqCfg>(e0, mont_mul<FqCfg>(r, fe_sub<FqCfg>(e1, e0)));
      Ts[b] = tn;   // each b is owned by exactly one thread; reads of b+half happen before
      Es[b] = en;   // any write to indices >= half (none are written this round)
    }
    __syncthreads();
    L = half;
    ri++;
  }
  tr_store_state(&sh, st);
  if (threadIdx.x == 0) {
    // last_claim = g(r) = xsq r^2 + x r + const      (r1cs.rs:2373-2376)
    const uint32_t last = ri - 1;
    Fq t = fe_add<FqCfg>(mont_mul<FqCfg>(r, st->out_rounds[last][1]), st->out_rounds[last][2]);   // xsq*r + x   (canonical)
    Fq lc = fe_add<FqCfg>(mont_mul<FqCfg>(r, t), st->out_rounds[last][3]);
    st->out_last_claim = lc;
    st->out_next_v = Ts[0];                                       // = T~(sc_rs)  (r1cs.rs:2379-2385)
  }
}

// (f1) outputs of the finished sum-check -> slots of an index-addressed witness buffer (canonical elements):
// rounds occupy ell x 4 consecutive slots in the order (sc_r, xsq, x, const) of reef_nlookup_out.rounds
__global__ void k_wit_scatter(const NlState* __restrict__ st, uint32_t ell, Fq* __restrict__ wit, uint64_t slot_claim_r,
                              uint64_t slot_rounds, uint64_t slot_last_claim, uint64_t slot_next_v) {
  pdl_wait();
  const uint64_t none = ~0ull;
  for (uint32_t i = threadIdx.x; i < 4 * ell; i += blockDim.x)
    if (slot_rounds != none) wit[slot_rounds + i] = st->out_rounds[i >> 2][i & 3];
  if (threadIdx.x == 0) {
    if (slot_claim_r != none) wit[slot_claim_r] = st->out_claim_r;
    if (slot_last_claim != none) wit[slot_last_claim] = st->out_last_claim;
    if (slot_next_v != none) wit[slot_next_v] = st->out_next_v;
  }
}

// ---------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------
static unsigned ceil_div_u(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// Dynamic shared-memory size that makes a single-CTA transcript kernel the ONLY resident CTA of its
// SM: every CTA needs at least the 1 KiB the system reserves per block, so a CTA that holds all the
// opt-in shared memory cannot get a neighbour.  The Fiat-Shamir warps are issue-latency bound; MSM or
// sweep CTAs co-scheduled on the same SM sub-partitions slow the critical path by 5-15 %
// (REEF_TRANSCRIPT_EXCLUSIVE=0 turns the reservation off).
static size_t exclusive_smem(reef_ctx* c, const void* kernel, size_t need) {
  static const bool on = !(getenv("REEF_TRANSCRIPT_EXCLUSIVE") && atoi(getenv("REEF_TRANSCRIPT_EXCLUSIVE")) == 0);
  static std::mutex mu;
  static std::vector<std::pair<const void*, size_t>> cache;
  if (!on) {
    if (need) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
    return need;
  }
  std::lock_guard<std::mutex> lk(mu);
  for (auto& kv : cache)
    if (kv.first == kernel) return kv.second;
  cudaFuncAttributes fa;
  int optin = 0;
  size_t dyn = need;
  if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess &&
      cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device) == cudaSuccess &&
      (size_t)optin > fa.sharedSizeBytes + need)
    dyn = (size_t)optin - fa.sharedSizeBytes;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) {
    cudaGetLastError();
    dyn = need;
    if (need) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
  }
  cache.emplace_back(kernel, dyn);
  return dyn;
}

template <bool U32IN>
static int nlookup_run_t(reef_ctx* c, const NlookupArgs& a) {
  const uint32_t ell = a.ell;
  const uint64_t N = a.n;
  const uint32_t m = a.m;
  const uint32_t hb = ell < (uint32_t)H_BITS ? ell : (uint32_t)H_BITS;
  const uint64_t a_len = (uint64_t)1 << (ell - hb);
  const uint64_t b_len = (uint64_t)1 << hb;
  const uint32_t n_sweeps = ell > (uint32_t)H_BITS ? ell - H_BITS : 0;
  const uint64_t max_blk = n_sweeps ? sweep_max_blocks(N) : 1;

  // scratch layout
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  size_t o_state = take(sizeof(NlState));
  size_t o_query = take((size_t)a.n_query * 32);
  size_t o_prevq = take((size_t)ell * 32);
  size_t o_q = take((size_t)m * 8);
  size_t o_pos = take((size_t)m * 8);
  size_t o_w = take((size_t)m * 32);
  size_t o_A = take(a_len * 32);
  size_t o_B = take(b_len * 32);
  size_t o_part = take(max_blk * 3 * 32);
  size_t o_fold = take(n_sweeps ? (N / 2) * 32 : 32);
  void* base;
  int rc = ctx_scratch(c, off, &base);
  if (rc) return rc;
  char* d = (char*)base;
  NlState* st = (NlState*)(d + o_state);
  Fq* d_query = (Fq*)(d + o_query);
  Fq* d_prevq = (Fq*)(d + o_prevq);
  uint64_t* d_q = (uint64_t*)(d + o_q);
  uint64_t* d_pos = (uint64_t*)(d + o_pos);
  Fq* d_w = (Fq*)(d + o_w);
  Fq* d_A = (Fq*)(d + o_A);
  Fq* d_B = (Fq*)(d + o_B);
  Fq* d_part = (Fq*)(d + o_part);
  Fq* d_fold = (Fq*)(d + o_fold);
  cudaStream_t s = c->stream;

  REEF_CUDA(cudaMemcpyAsync(d_query, a.h_query, (size_t)a.n_query * 32, cudaMemcpyHostToDevice, s));
  REEF_CUDA(cudaMemcpyAsync(d_prevq, a.h_prev_q, (size_t)ell * 32, cudaMemcpyHostToDevice, s));
  if (m) REEF_CUDA(cudaMemcpyAsync(d_q, a.h_q, (size_t)m * 8, cudaMemcpyHostToDevice, s));

  Fq tag = fq_canon_from_le32(a.tag_le);   // canonical: the LP sponge absorbs plain residues
  {
    ProfScope ps(c, PROF_NL_SETUP, N);
    k_nl_begin<<<1, TR_THREADS, exclusive_smem(c, (const void*)k_nl_begin, 0), s>>>(st, d_query, a.n_query, tag, d_prevq, ell, d_q, m, d_pos, d_w, c->d_lp, 0, 1);
    REEF_LAUNCHED();
    REEF_CUDA(launch_dep(k_eq_tables, dim3(ceil_div_u(a_len + b_len, 128)), dim3(128), 0, s, (const NlState*)st, ell, hb, d_A, a_len, d_B, b_len, 0u, 0u));
    REEF_LAUNCHED();
  }
  // the first absorb above (5-7 permutations) never reads the table: an asynchronous upload overlaps it
  if (a.table_ready) REEF_CUDA(cudaStreamWaitEvent(s, a.table_ready, 0));

  // sweep rounds 1 .. ell-h
  const void* t_cur = a.d_table;
  uint64_t L = N;                // accumulation length of the current round
  uint64_t a_cur = a_len;
  for (uint32_t i = 0; i < n_sweeps; i++) {
    uint32_t nblk = 0;
    if (i == 0) {
      {
        ProfScope ps(c, PROF_SWEEP_FIRST, N);
        rc = launch_sweep<U32IN, false>(c, a.d_table, N, nullptr, st, d_A, d_B, d_part, &nblk);
        if (rc) return rc;
      }
      ProfScope ps(c, PROF_ROUND, L);
      REEF_CUDA(launch_dep(k_round<U32IN>, dim3(1), dim3(ROUND_THREADS), exclusive_smem(c, (const void*)k_round<U32IN>, 0), s, st, (const Fq*)d_part, nblk,
                           a.d_table, L, d_A, a_cur, d_pos, d_w, m, i, (const PoseidonLpTables*)c->d_lp));
      REEF_LAUNCHED();
    } else {
      {
        ProfScope ps(c, PROF_SWEEP_FOLD, 2 * L);
        if (i == 1) rc = launch_sweep<U32IN, true>(c, a.d_table, 2 * L, d_fold, st, d_A, d_B, d_part, &nblk);
        else rc = launch_sweep<false, true>(c, d_fold, 2 * L, d_fold, st, d_A, d_B, d_part, &nblk);
        if (rc) return rc;
      }
      ProfScope ps(c, PROF_ROUND, L);
      REEF_CUDA(launch_dep(k_round<false>, dim3(1), dim3(ROUND_THREADS), exclusive_smem(c, (const void*)k_round<false>, 0), s, st, (const Fq*)d_part, nblk,
                           (const void*)d_fold, L, d_A, a_cur, d_pos, d_w, m, i, (const PoseidonLpTables*)c->d_lp));
      REEF_LAUNCHED();
      t_cur = d_fold;
    }
    if (a_cur > 1) a_cur >>= 1;
    L >>= 1;
  }
  // tail: fold with the last sweep challenge (if any) and finish
  const size_t tail_smem = (size_t)(2 * CHUNK + 3 * TAIL_THREADS / 32) * sizeof(Fq);
  std::unique_ptr<ProfScope> tail_scope(new ProfScope(c, PROF_TAIL, L));
  const PoseidonLpTables* lpt = (const PoseidonLpTables*)c->d_lp;
  if (n_sweeps == 0) {
    REEF_CUDA(launch_dep(k_tail<U32IN>, dim3(1), dim3(TAIL_THREADS), exclusive_smem(c, (const void*)k_tail<U32IN>, tail_smem), s, st, a.d_table, N, 0,
                         (const Fq*)d_A, (const Fq*)d_B, (const uint64_t*)d_pos, (const Fq*)d_w, m, 0u, lpt));
  } else if (n_sweeps == 1) {
    // L is now 2^h: the table to fold is still the caller's (length 2L)
    REEF_CUDA(launch_dep(k_tail<U32IN>, dim3(1), dim3(TAIL_THREADS), exclusive_smem(c, (const void*)k_tail<U32IN>, tail_smem), s, st, a.d_table, 2 * L, 1,
                         (const Fq*)d_A, (const Fq*)d_B, (const uint64_t*)d_pos, (const Fq*)d_w, m, n_sweeps, lpt));
  } else {
    REEF_CUDA(launch_dep(k_tail<false>, dim3(1), dim3(TAIL_THREADS), exclusive_smem(c, (const void*)k_tail<false>, tail_smem), s, st, t_cur, 2 * L, 1,
                         (const Fq*)d_A, (const Fq*)d_B, (const uint64_t*)d_pos, (const Fq*)d_w, m, n_sweeps, lpt));
  }
  tail_scope.reset();
  REEF_LAUNCHED();
  if (a.d_wit) {
    k_wit_scatter<<<1, 128, 0, s>>>(st, ell, (Fq*)a.d_wit, a.slot_claim_r, a.slot_rounds, a.slot_last_claim, a.slot_next_v);
    REEF_LAUNCHED();
  }

  // results
  void* hs;
  rc = ctx_stage(c, sizeof(NlState), &hs);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(hs, st, sizeof(NlState), cudaMemcpyDeviceToHost, s));
  REEF_CUDA(cudaStreamSynchronize(s));
  const NlState* h = (const NlState*)hs;
  memcpy(a.out_claim_r, h->out_claim_r.v, 32);
  for (uint32_t i = 0; i < ell; i++)
    for (int k = 0; k < 4; k++) memcpy(a.out_rounds + ((size_t)i * 4 + k) * 32, h->out_rounds[i][k].v, 32);
  memcpy(a.out_last_claim, h->out_last_claim.v, 32);
  memcpy(a.out_next_v, h->out_next_v.v, 32);
  return REEF_OK;
}

int nlookup_run(reef_ctx* c, const NlookupArgs& a) {
  REEF_REQUIRE(a.ell >= 1 && a.ell <= (uint32_t)MAX_ELL, REEF_EINVAL, "nlookup: ell out of range");
  REEF_REQUIRE(a.n == ((uint64_t)1 << a.ell), REEF_EINVAL, "nlookup: table length must be 2^ell");
  for (uint32_t k = 0; k < a.m; k++)
    REEF_REQUIRE(a.h_q[k] < a.n, REEF_EASSERT, "nlookup: lookup index out of range (index out of bounds)");
  return a.table_is_u32 ? nlookup_run_t<true>(c, a) : nlookup_run_t<false>(c, a);
}

// ---------------------------------------------------------------------------------------
// Reference-shaped standalone entry points (materialised tables), used by the parity tests
// that mirror the reference's own unit tests (mle_partial, mle_linear_basic).
// ---------------------------------------------------------------------------------------

// eq[i] = rs[m] * prod_j sel(bit_j(i), last_q[j]); then eq[q_k] += rs[k]
__global__ void k_eq_full(const Fq* __restrict__ rs_mont, const Fq* __restrict__ lq_mont, uint32_t ell, uint32_t m,
                          Fq* __restrict__ out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fq one = fe_one<FqCfg>();
  Fq acc = rs_mont[m];
  for (uint32_t j = 0; j < ell; j++) {
    Fq lq = lq_mont[j];
    Fq f = ((i >> j) & 1) ? lq : fe_sub<FqCfg>(one, lq);
    acc = mont_mul<FqCfg>(acc, f);
  }
  st256(out + i, from_mont<FqCfg>(acc));
}

__global__ void k_eq_scatter(const Fq* __restrict__ rs_mont, const uint64_t* __restrict__ qs, uint32_t m, Fq* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    for (uint32_t k = 0; k < m; k++) out[qs[k]] = fe_add<FqCfg>(out[qs[k]], from_mont<FqCfg>(rs_mont[k]));
}

__global__ void k_to_mont(Fq* x, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = to_mont<FqCfg>(x[i]);
}

int launch_gen_eq_table(reef_ctx* c, const uint8_t* h_rs, const uint64_t* h_qs, uint32_t m, const uint8_t* h_last_q,
                        uint32_t ell, void* d_out) {
  REEF_REQUIRE(ell >= 1 && ell <= (uint32_t)MAX_ELL, REEF_EINVAL, "gen_eq_table: ell out of range");
  const uint64_t n = (uint64_t)1 << ell;
  for (uint32_t k = 0; k < m; k++) REEF_REQUIRE(h_qs[k] < n, REEF_EASSERT, "gen_eq_table: index out of bounds");
  size_t bytes = (size_t)(m + 1) * 32 + (size_t)ell * 32 + (size_t)m * 8 + 512;
  void* base;
  int rc = ctx_scratch(c, bytes, &base);
  if (rc) return rc;
  Fq* d_rs = (Fq*)base;
  Fq* d_lq = d_rs + (m + 1);
  uint64_t* d_qs = (uint64_t*)(d_lq + ell);
  cudaStream_t s = c->stream;
  REEF_CUDA(cudaMemcpyAsync(d_rs, h_rs, (size_t)(m + 1) * 32, cudaMemcpyHostToDevice, s));
  REEF_CUDA(cudaMemcpyAsync(d_lq, h_last_q, (size_t)ell * 32, cudaMemcpyHostToDevice, s));
  if (m) REEF_CUDA(cudaMemcpyAsync(d_qs, h_qs, (size_t)m * 8, cudaMemcpyHostToDevice, s));
  k_to_mont<<<ceil_div_u(m + 1 + ell, 128), 128, 0, s>>>(d_rs, m + 1 + ell);
  REEF_LAUNCHED();
  k_eq_full<<<ceil_div_u(n, 128), 128, 0, s>>>(d_rs, d_lq, ell, m, (Fq*)d_out, n);
  REEF_LAUNCHED();
  k_eq_scatter<<<1, 32, 0, s>>>(d_rs, d_qs, m, (Fq*)d_out);
  REEF_LAUNCHED();
  return REEF_OK;
}

// hybrid table (r1cs.rs:2101-2112): out[i] = pub[i] | fill | doc code (repeated with zero padding to a power of two)
__global__ void k_hybrid_table(const Fq* __restrict__ pub, uint64_t n_pub, Fq fill, uint64_t half_len,
                               const uint32_t* __restrict__ codes, uint64_t n_doc, uint64_t doc_pad, Fq* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * half_len) return;
  Fq v;
  if (i < half_len) {
    v = i < n_pub ? ld256(pub + i) : fill;
  } else {
    const uint64_t j = (i - half_len) % doc_pad;
    v = fq_from_u32(j < n_doc ? codes[j] : 0u);
  }
  st256(out + i, v);
}

int launch_hybrid_table(reef_ctx* c, const void* d_pub, uint64_t n_pub, const uint8_t* fill_le, uint64_t half_len,
                        const uint32_t* d_codes, uint64_t n_doc, void* d_out) {
  uint64_t doc_pad = 1;
  while (doc_pad < n_doc) doc_pad <<= 1;
  const Fq fill = fq_canon_from_le32(fill_le);
  k_hybrid_table<<<ceil_div_u(2 * half_len, 256), 256, 0, c->stream>>>((const Fq*)d_pub, n_pub, fill, half_len, d_codes, n_doc, doc_pad,
                                                                     (Fq*)d_out);
  REEF_LAUNCHED();
  return REEF_OK;
}

// out[b] = in[b] + r (in[b+half] - in[b])
template <bool U32IN>
__global__ void k_fold(const void* __restrict__ in, Fq* __restrict__ out, uint64_t half, Fq r_mont) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= half) return;
  st256(out + b, fold_one(load_t<U32IN>(in, b), load_t<U32IN>(in, b + half), r_mont));
}

// T~(x), x[0] <-> top index bit  (verifier_mle_eval / prover_mle_partial_eval without a hole)
int launch_mle_eval(reef_ctx* c, const void* d_table, int is_u32, uint64_t n, const uint8_t* h_x, uint32_t ell,
                    uint8_t* h_out) {
  REEF_REQUIRE(n == ((uint64_t)1 << ell) && ell >= 1, REEF_EINVAL, "mle_eval: table length must be 2^ell");
  void* base;
  int rc = ctx_scratch2(c, (size_t)(n / 2 + n / 4 + 2) * 32, &base);
  if (rc) return rc;
  Fq* buf0 = (Fq*)base;
  Fq* buf1 = buf0 + n / 2 + 1;
  cudaStream_t s = c->stream;
  const void* cur = d_table;
  uint64_t L = n;
  for (uint32_t i = 0; i < ell; i++) {
    Fq r = fq_mont_from_le32(h_x + (size_t)i * 32);
    uint64_t half = L >> 1;
    Fq* out = (i & 1) ? buf1 : buf0;
    if (i == 0 && is_u32) k_fold<true><<<ceil_div_u(half, 128), 128, 0, s>>>(cur, out, half, r);
    else k_fold<false><<<ceil_div_u(half, 128), 128, 0, s>>>(cur, out, half, r);
    REEF_LAUNCHED();
    cur = out;
    L = half;
  }
  REEF_CUDA(cudaMemcpyAsync(h_out, cur, 32, cudaMemcpyDeviceToHost, s));
  REEF_CUDA(cudaStreamSynchronize(s));
  return REEF_OK;
}

// first loop of linear_mle_product on materialised tables: (xsq, x, const)
__global__ void __launch_bounds__(128) k_round_coeffs(const Fq* __restrict__ T, const Fq* __restrict__ E, uint64_t pw,
                                                      Fq* __restrict__ partials) {
  __shared__ Fq red[3 * 4];
  Fq acc[3];
  acc[0] = acc[1] = acc[2] = fe_zero<FqCfg>();
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < pw; b += (uint64_t)gridDim.x * blockDim.x) {
    Fq t0 = ld256(T + b), t1 = ld256(T + b + pw);
    Fq e0 = to_mont<FqCfg>(ld256(E + b)), e1 = to_mont<FqCfg>(ld256(E + b + pw));
    Fq ts = fe_sub<FqCfg>(t1, t0), es = fe_sub<FqCfg>(e1, e0);
    acc[0] = fe_add<FqCfg>(acc[0], mont_mul<FqCfg>(es, ts));                       // xsq
    acc[1] = fe_add<FqCfg>(acc[1], fe_add<FqCfg>(mont_mul<FqCfg>(es, t0), mont_mul<FqCfg>(e0, ts)));  // x
    acc[2] = fe_add<FqCfg>(acc[2], mont_mul<FqCfg>(e0, t0));                       // const
  }
  block_sum<3, 128>(acc, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < 3; k++) st256(partials + (uint64_t)blockIdx.x * 3 + k, acc[k]);
}

__global__ void __launch_bounds__(128) k_sum_partials(const Fq* __restrict__ partials, uint32_t nblk, Fq* out3) {
  __shared__ Fq red[3 * 4];
  Fq acc[3];
  acc[0] = acc[1] = acc[2] = fe_zero<FqCfg>();
  for (uint32_t i = threadIdx.x; i < nblk; i += 128)
    for (int k = 0; k < 3; k++) acc[k] = fe_add<FqCfg>(acc[k], ld256(partials + (uint64_t)i * 3 + k));
  block_sum<3, 128>(acc, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < 3; k++) out3[k] = acc[k];
}

int launch_mle_round_coeffs(reef_ctx* c, const void* d_t, const void* d_eq, uint32_t ell, uint32_t i, uint8_t* h_out3) {
  REEF_REQUIRE(i >= 1 && i <= ell, REEF_EINVAL, "linear_mle_product: round index out of range");
  const uint64_t pw = (uint64_t)1 << (ell - i);
  unsigned nblk = ceil_div_u(pw, 128);
  if (nblk > 1184) nblk = 1184;
  void* base;
  int rc = ctx_scratch(c, (size_t)(nblk + 1) * 3 * 32, &base);
  if (rc) return rc;
  Fq* part = (Fq*)base;
  cudaStream_t s = c->stream;
  k_round_coeffs<<<nblk, 128, 0, s>>>((const Fq*)d_t, (const Fq*)d_eq, pw, part);
  REEF_LAUNCHED();
  k_sum_partials<<<1, 128, 0, s>>>(part, nblk, part + (size_t)nblk * 3);
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(h_out3, part + (size_t)nblk * 3, 96, cudaMemcpyDeviceToHost, s));
  REEF_CUDA(cudaStreamSynchronize(s));
  return REEF_OK;
}

// second loop of linear_mle_product: fold both tables in place with r
__global__ void k_fold2_inplace(Fq* T, Fq* E, uint64_t pw, Fq r_mont) {
  uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pw) return;
  Fq t = fold_one(ld256(T + b), ld256(T + b + pw), r_mont);
  Fq e = fold_one(ld256(E + b), ld256(E + b + pw), r_mont);
  st256(T + b, t);
  st256(E + b, e);
}

int launch_mle_round_fold(reef_ctx* c, void* d_t, void* d_eq, uint32_t ell, uint32_t i, const uint8_t* h_r) {
  REEF_REQUIRE(i >= 1 && i <= ell, REEF_EINVAL, "linear_mle_product: round index out of range");
  const uint64_t pw = (uint64_t)1 << (ell - i);
  Fq r = fq_mont_from_le32(h_r);
  k_fold2_inplace<<<ceil_div_u(pw, 128), 128, 0, c->stream>>>((Fq*)d_t, (Fq*)d_eq, pw, r);
  REEF_LAUNCHED();
  return REEF_OK;
}


// ---------------------------------------------------------------------------------------
// Hyrax prove_eval's  LZ = L^T * M  (commitment.rs:371-393 -> nova hyrax_pc::prove_eval):
// M is the rows x cols row-major matrix view of the document polynomial, L = eq(q_left).
//   out[j] = sum_i L[i] * M[i][j]
// grid (cols/128, row chunks); coalesced over j; lazy 17-limb accumulation per thread.
// ---------------------------------------------------------------------------------------
static constexpr int LZ_ROWS_PER_CTA = 64;

template <bool U32IN>
__global__ void __launch_bounds__(128) k_lz_partial(const void* __restrict__ M, const Fq* __restrict__ L, uint64_t rows,
                                                    uint64_t cols, Fq* __restrict__ partial) {
  const uint64_t j = (uint64_t)blockIdx.x * 128 + threadIdx.x;
  if (j >= cols) return;
  const uint64_t r0 = (uint64_t)blockIdx.y * LZ_ROWS_PER_CTA;
  const uint64_t r1 = min(rows, r0 + LZ_ROWS_PER_CTA);
  Wide17 acc;
  wide_zero(acc);
#pragma unroll 2
  for (uint64_t i = r0; i < r1; i++) {
    Fq l = ld256(L + i);                               // uniform across the CTA (L1 broadcast)
    if constexpr (U32IN) wide_mac_small(acc, ((const uint32_t*)M)[i * cols + j], l.v);
    else {
      Fq x = ld256((const Fq*)M + i * cols + j);
      wide_mac(acc, x.v, l.v);
    }
  }
  st256(partial + (uint64_t)blockIdx.y * cols + j, wide_reduce_canonical<FqCfg>(acc));
}

__global__ void __launch_bounds__(128) k_lz_sum(const Fq* __restrict__ partial, uint32_t n_chunks, uint64_t cols,
                                                Fq* __restrict__ out) {
  const uint64_t j = (uint64_t)blockIdx.x * 128 + threadIdx.x;
  if (j >= cols) return;
  Fq acc = ld256(partial + j);
  for (uint32_t k = 1; k < n_chunks; k++) acc = fe_add<FqCfg>(acc, ld256(partial + (uint64_t)k * cols + j));
  st256(out + j, acc);
}

int launch_lz(reef_ctx* c, const void* d_matrix, int is_u32, uint64_t rows, uint64_t cols, const uint8_t* h_L,
              uint8_t* h_out) {
  const uint32_t n_chunks = ceil_div_u(rows, LZ_ROWS_PER_CTA);
  void* base;
  int rc = ctx_scratch(c, (size_t)(rows + (uint64_t)n_chunks * cols + cols + 8) * 32, &base);
  if (rc) return rc;
  Fq* d_L = (Fq*)base;
  Fq* d_part = d_L + rows;
  Fq* d_out = d_part + (uint64_t)n_chunks * cols;
  cudaStream_t s = c->stream;
  REEF_CUDA(cudaMemcpyAsync(d_L, h_L, (size_t)rows * 32, cudaMemcpyHostToDevice, s));
  dim3 grid(ceil_div_u(cols, 128), n_chunks);
  if (is_u32) k_lz_partial<true><<<grid, 128, 0, s>>>(d_matrix, d_L, rows, cols, d_part);
  else k_lz_partial<false><<<grid, 128, 0, s>>>(d_matrix, d_L, rows, cols, d_part);
  REEF_LAUNCHED();
  k_lz_sum<<<ceil_div_u(cols, 128), 128, 0, s>>>(d_part, n_chunks, cols, d_out);
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(h_out, d_out, (size_t)cols * 32, cudaMemcpyDeviceToHost, s));
  REEF_CUDA(cudaStreamSynchronize(s));
  return REEF_OK;
}


// ---------------------------------------------------------------------------------------
// Multi-GPU sharded sum-check (protocol stated in oracle/sharded.py).
// Rank g holds T_g[j] = T[j*G + g].  Per round: local (const, g(1), xsq) -> all-gather ->
// identical transcript on every rank.  No call below synchronises the stream; the triples
// travel through caller-provided DEVICE buffers so that NCCL can move them directly.
// ---------------------------------------------------------------------------------------

// sum of this rank's CTA partials + its sparse-point terms -> out3 = (const, g(1), xsq)
template <bool U32IN>
__global__ void __launch_bounds__(ROUND_THREADS)
k_shard_local(const Fq* __restrict__ partials, uint32_t nblk, const void* __restrict__ Tcur, uint64_t L,
              const uint64_t* __restrict__ sp_pos, const Fq* __restrict__ sp_w, uint32_t m, Fq* __restrict__ out3, MbRef mb) {
  __shared__ Fq red[3 * ROUND_THREADS / 32];
  __shared__ Fq post_sh[3];
  const uint64_t half = L >> 1;
  Fq acc[3];
  acc[0] = acc[1] = acc[2] = fe_zero<FqCfg>();
  for (uint32_t i = threadIdx.x; i < nblk; i += ROUND_THREADS) {
#pragma unroll
    for (int k = 0; k < 3; k++) acc[k] = fe_add<FqCfg>(acc[k], ld256(partials + (uint64_t)i * 3 + k));
  }
  for (uint32_t k = threadIdx.x; k < m; k += ROUND_THREADS) {
    uint64_t pos = sp_pos[k];
    bool top = pos >= half;
    uint64_t b = top ? pos - half : pos;
    Fq w = sp_w[k];
    Fq t0 = load_t<U32IN>(Tcur, b), t1 = load_t<U32IN>(Tcur, b + half);
    Fq wt = mont_mul<FqCfg>(w, top ? t1 : t0);
    Fq wd = mont_mul<FqCfg>(w, fe_sub<FqCfg>(t1, t0));
    if (top) {
      acc[1] = fe_add<FqCfg>(acc[1], wt);
      acc[2] = fe_add<FqCfg>(acc[2], wd);
    } else {
      acc[0] = fe_add<FqCfg>(acc[0], wt);
      acc[2] = fe_sub<FqCfg>(acc[2], wd);
    }
  }
  block_sum<3, ROUND_THREADS>(acc, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < 3; k++) {
      if (out3) st256(out3 + k, acc[k]);
      post_sh[k] = acc[k];
    }
  if (mb.peers) {   // fused exchange: this kernel is also the sender of the round's all-gather
    __syncthreads();
    if (threadIdx.x < mb.world) mb_post(mb, threadIdx.x, reinterpret_cast<const uint32_t*>(post_sh), 24);
  }
}

// Thread 0 sums the G gathered triples (rank-major (const, g(1), xsq)); then one transcript round by the
// whole CTA.  `triples` may live in global (gathered) or shared memory.
__device__ __forceinline__ Fq shard_transcript(LpPermShared* sh, TrScratch* ts, u32& seq, NlState* st, const Fq* triples, uint32_t G,
                                               uint32_t ri, const PoseidonLpTables* __restrict__ K) {
  Fq con = fe_zero<FqCfg>(), g1 = con, xsq = con;
  if (threadIdx.x == 0) {
    for (uint32_t g = 0; g < G; g++) {
      con = fe_add<FqCfg>(con, triples[(uint64_t)g * 3 + 0]);
      g1 = fe_add<FqCfg>(g1, triples[(uint64_t)g * 3 + 1]);
      xsq = fe_add<FqCfg>(xsq, triples[(uint64_t)g * 3 + 2]);
    }
  }
  return tr_round(sh, ts, K, seq, st, ri, con, g1, xsq);
}

// sweep regime: transcript with the gathered triples, fold A, advance the sparse list
__global__ void __launch_bounds__(ROUND_THREADS)
k_shard_apply(NlState* st, const Fq* __restrict__ triples, uint32_t G, uint64_t L, Fq* A, uint64_t a_len,
              uint64_t* sp_pos, Fq* sp_w, uint32_t m, uint32_t ri, const PoseidonLpTables* __restrict__ K, MbRef mb) {
  __shared__ LpPermShared sh;
  __shared__ TrScratch ts;
  __shared__ Fq trip_sh[MB_MAX_WORLD * 3];
  lp_perm_init(&sh, K);
  tr_load_state(&sh, st);
  if (mb.peers) {   // fused exchange: this kernel is also the receiver of the round's all-gather
    if (threadIdx.x < mb.world) mb_wait_copy(mb, threadIdx.x, reinterpret_cast<uint32_t*>(trip_sh + 3 * threadIdx.x), 24);
    __syncthreads();
    triples = trip_sh;
  }
  const uint64_t half = L >> 1;
  u32 seq = 0;
  const Fq r = shard_transcript(&sh, &ts, seq, st, triples, G, ri, K);
  tr_store_state(&sh, st);
  if (a_len > 1) {
    const uint64_t ah = a_len >> 1;
    for (uint64_t x = threadIdx.x; x < ah; x += ROUND_THREADS) {
      Fq lo = A[x], hi = A[x + ah];
      A[x] = fe_add<FqCfg>(lo, mont_mul<FqCfg>(r, fe_sub<FqCfg>(hi, lo)));
    }
  }
  for (uint32_t k = threadIdx.x; k < m; k += ROUND_THREADS) {
    uint64_t pos = sp_pos[k];
    bool top = pos >= half;
    Fq f = top ? r : fe_sub<FqCfg>(fe_one<FqCfg>(), r);
    sp_w[k] = mont_mul<FqCfg>(sp_w[k], f);
    sp_pos[k] = top ? pos - half : pos;
  }
}

// enter the small regime: T_s (canonical) and E_s (Montgomery) materialised, length L <= 2^h
template <bool U32IN>
__global__ void __launch_bounds__(TAIL_THREADS)
k_shard_materialize(const NlState* __restrict__ st, const void* __restrict__ Tin, uint64_t L_in, int do_fold,
                    const Fq* __restrict__ A, const Fq* __restrict__ B, const uint64_t* __restrict__ sp_pos,
                    const Fq* __restrict__ sp_w, uint32_t m, Fq* Ts, Fq* Es) {
  const uint64_t L = do_fold ? (L_in >> 1) : L_in;
  const Fq a0 = A[0];
  const Fq rf = st->r_mont;
  for (uint64_t b = threadIdx.x; b < L; b += TAIL_THREADS) {
    Fq t = do_fold ? fold_one(load_t<U32IN>(Tin, b), load_t<U32IN>(Tin, b + L), rf) : load_t<U32IN>(Tin, b);
    Ts[b] = t;
    Es[b] = mont_mul<FqCfg>(a0, B[b]);
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (uint32_t k = 0; k < m; k++) Es[sp_pos[k]] = fe_add<FqCfg>(Es[sp_pos[k]], sp_w[k]);
}

// small regime, local sums of the current round
__global__ void __launch_bounds__(TAIL_THREADS) k_small_local(const Fq* __restrict__ Ts, const Fq* __restrict__ Es, uint64_t L,
                                                              Fq* __restrict__ out3, MbRef mb) {
  __shared__ Fq red[3 * TAIL_THREADS / 32];
  __shared__ Fq post_sh[3];
  const uint64_t half = L >> 1;
  Fq acc[3];
  acc[0] = acc[1] = acc[2] = fe_zero<FqCfg>();
  for (uint64_t b = threadIdx.x; b < half; b += TAIL_THREADS) {
    Fq t0 = Ts[b], t1 = Ts[b + half], e0 = Es[b], e1 = Es[b + half];
    acc[0] = fe_add<FqCfg>(acc[0], mont_mul<FqCfg>(e0, t0));
    acc[1] = fe_add<FqCfg>(acc[1], mont_mul<FqCfg>(e1, t1));
    acc[2] = fe_add<FqCfg>(acc[2], mont_mul<FqCfg>(fe_sub<FqCfg>(e1, e0), fe_sub<FqCfg>(t1, t0)));
  }
  block_sum<3, TAIL_THREADS>(acc, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < 3; k++) {
      if (out3) st256(out3 + k, acc[k]);
      post_sh[k] = acc[k];
    }
  if (mb.peers) {
    __syncthreads();
    if (threadIdx.x < mb.world) mb_post(mb, threadIdx.x, reinterpret_cast<const uint32_t*>(post_sh), 24);
  }
}

// small regime: transcript with the gathered triples, then fold T_s / E_s in place
__global__ void __launch_bounds__(TAIL_THREADS)
k_small_apply(NlState* st, const Fq* __restrict__ triples, uint32_t G, Fq* Ts, Fq* Es, uint64_t L, uint32_t ri,
              const PoseidonLpTables* __restrict__ K, MbRef mb) {
  __shared__ LpPermShared sh;
  __shared__ TrScratch ts;
  __shared__ Fq trip_sh[MB_MAX_WORLD * 3];
  lp_perm_init(&sh, K);
  tr_load_state(&sh, st);
  if (mb.peers) {
    if (threadIdx.x < mb.world) mb_wait_copy(mb, threadIdx.x, reinterpret_cast<uint32_t*>(trip_sh + 3 * threadIdx.x), 24);
    __syncthreads();
    triples = trip_sh;
  }
  const uint64_t half = L >> 1;
  u32 seq = 0;
  const Fq r = shard_transcript(&sh, &ts, seq, st, triples, G, ri, K);
  tr_store_state(&sh, st);
  // half <= 512: every b is owned by one thread; reads of b + half precede no write there
  for (uint64_t b = threadIdx.x; b < half; b += TAIL_THREADS) {
    Fq t0 = Ts[b], t1 = Ts[b + half], e0 = Es[b], e1 = Es[b + half];
    Ts[b] = fold_one(t0, t1, r);
    Es[b] = fe_add<FqCfg>(e0, mont_mul<FqCfg>(r, fe_sub<FqCfg>(e1, e0)));
  }
}

__global__ void k_shard_export(const Fq* __restrict__ Ts, const Fq* __restrict__ Es, Fq* __restrict__ out2, MbRef mb) {
  __shared__ Fq post_sh[2];
  if (threadIdx.x == 0) {
    post_sh[0] = Ts[0];
    post_sh[1] = from_mont<FqCfg>(Es[0]);
    if (out2) {
      st256(out2 + 0, post_sh[0]);
      st256(out2 + 1, post_sh[1]);
    }
  }
  if (mb.peers) {
    __syncthreads();
    if (threadIdx.x < mb.world) mb_post(mb, threadIdx.x, reinterpret_cast<const uint32_t*>(post_sh), 16);
  }
}

// last gamma rounds over the G gathered (T, E) pairs (rank g's pair at index g), identical on
// every rank; then last claim and next running claim.
__global__ void __launch_bounds__(TR_THREADS) k_shard_final(NlState* st, const Fq* __restrict__ pairs, uint32_t G, uint32_t ri0,
                                                            const PoseidonLpTables* __restrict__ K, MbRef mb) {
  // warp 0 does the field work between the transcript rounds; the whole CTA runs the permutations
  __shared__ LpPermShared sh;
  __shared__ TrScratch ts;
  __shared__ Fq Ts[64], Es[64];
  __shared__ Fq pair_sh[MB_MAX_WORLD * 2];
  lp_perm_init(&sh, K);
  tr_load_state(&sh, st);
  const int lane = threadIdx.x & 31;
  const bool A = threadIdx.x < 32;
  if (A) {
    if (mb.peers) {   // fused exchange: receive the G exported (T, E) pairs
      if ((uint32_t)lane < mb.world) mb_wait_copy(mb, lane, reinterpret_cast<uint32_t*>(pair_sh + 2 * lane), 16);
      __syncwarp();
      for (uint32_t g = lane; g < G; g += 32) {
        Ts[g] = pair_sh[g * 2];
        Es[g] = to_mont<FqCfg>(pair_sh[g * 2 + 1]);
      }
    } else {
      for (uint32_t g = lane; g < G; g += 32) {
        Ts[g] = ld256(pairs + (uint64_t)g * 2);
        Es[g] = to_mont<FqCfg>(ld256(pairs + (uint64_t)g * 2 + 1));
      }
    }
    __syncwarp();
  }
  uint32_t ri = ri0;
  u32 seq = 0;
  Fq r = fe_zero<FqCfg>();
  for (uint32_t L = G; L > 1; L >>= 1) {
    const uint32_t half = L >> 1;
    Fq acc[3];
    acc[0] = acc[1] = acc[2] = fe_zero<FqCfg>();
    if (A) {
      for (uint32_t b = lane; b < half; b += 32) {
        Fq t0 = Ts[b], t1 = Ts[b + half], e0 = Es[b], e1 = Es[b + half];
        acc[0] = fe_add<FqCfg>(acc[0], mont_mul<FqCfg>(e0, t0));
        acc[1] = fe_add<FqCfg>(acc[1], mont_mul<FqCfg>(e1, t1));
        acc[2] = fe_add<FqCfg>(acc[2], mont_mul<FqCfg>(fe_sub<FqCfg>(e1, e0), fe_sub<FqCfg>(t1, t0)));
      }
#pragma unroll
      for (int k = 0; k < 3; k++) acc[k] = warp_sum_fe<FqCfg>(acc[k]);
    }
    r = tr_round(&sh, &ts, K, seq, st, ri, acc[0], acc[1], acc[2]);
    if (A) {
      for (uint32_t b = lane; b < half; b += 32) {
        Fq t0 = Ts[b], t1 = Ts[b + half], e0 = Es[b], e1 = Es[b + half];
        Ts[b] = fold_one(t0, t1, r);
        Es[b] = fe_add<FqCfg>(e0, mont_mul<FqCfg>(r, fe_sub<FqCfg>(e1, e0)));
      }
    }
    __syncthreads();
    ri++;
  }
  tr_store_state(&sh, st);
  if (threadIdx.x == 0) {
    const uint32_t last = ri - 1;
    const Fq rr = st->r_mont;
    Fq t = fe_add<FqCfg>(mont_mul<FqCfg>(rr, st->out_rounds[last][1]), st->out_rounds[last][2]);
    st->out_last_claim = fe_add<FqCfg>(mont_mul<FqCfg>(rr, t), st->out_rounds[last][3]);
    st->out_next_v = Ts[0];
  }
}

}  // namespace reef

struct reef_nl_session {
  reef_ctx* ctx;
  void* d_buf;
  size_t buf_bytes;
  reef::MbRef mb;     // mailbox exchange fused into the round kernels (peers == nullptr: caller-side exchange)
  const void* d_table;
  int is_u32;
  uint64_t n_loc;
  uint32_t ell, ell_loc, gamma, rank, world, m;
  reef::NlState* st;
  reef::Fq *d_query, *d_prevq, *d_w, *d_A, *d_B, *d_part, *d_fold, *d_Ts, *d_Es;
  uint64_t *d_q, *d_pos;
  uint32_t round;     // local rounds finished
  uint64_t L;         // local accumulation length of the upcoming round
  uint64_t a_cur;
  int small;          // 1 once T_s / E_s are materialised
  uint32_t nblk;
};

namespace reef {

int nl_shard_begin(reef_ctx* c, const NlookupArgs& a, uint32_t rank, uint32_t world, reef_nl_session** out) {
  static_assert(MB_MAX_WORLD <= 64, "k_shard_final keeps the gathered pairs in 64-entry shared arrays");
  REEF_REQUIRE(world >= 1 && (world & (world - 1)) == 0 && world <= MB_MAX_WORLD && rank < world, REEF_EINVAL,
               "nl_shard_begin: world must be a power of two <= 32");
  uint32_t gamma = 0;
  while ((1u << gamma) < world) gamma++;
  REEF_REQUIRE(a.ell > gamma && a.ell <= (uint32_t)MAX_ELL, REEF_EINVAL, "nl_shard_begin: table too small for this world size");
  const uint32_t ell_loc = a.ell - gamma;
  const uint64_t n_loc = (uint64_t)1 << ell_loc;
  REEF_REQUIRE(a.n == n_loc, REEF_EINVAL, "nl_shard_begin: local table must hold 2^(ell - log2 world) entries");
  const uint32_t hb = ell_loc < (uint32_t)H_BITS ? ell_loc : (uint32_t)H_BITS;
  const uint64_t a_len = (uint64_t)1 << (ell_loc - hb), b_len = (uint64_t)1 << hb;
  const bool sweeps = ell_loc > (uint32_t)H_BITS;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  size_t o_state = take(sizeof(NlState)), o_query = take((size_t)a.n_query * 32), o_prevq = take((size_t)a.ell * 32);
  size_t o_q = take((size_t)a.m * 8 + 8), o_pos = take((size_t)a.m * 8 + 8), o_w = take((size_t)a.m * 32 + 32);
  size_t o_A = take(a_len * 32), o_B = take(b_len * 32);
  size_t o_part = take((sweeps ? sweep_max_blocks(n_loc) : 1) * 3 * 32);
  size_t o_fold = take(sweeps ? (n_loc / 2) * 32 : 32);
  size_t o_Ts = take((size_t)CHUNK * 32), o_Es = take((size_t)CHUNK * 32);
  void* buf = nullptr;
  size_t buf_bytes = off;
  if (c->shard_cache && c->shard_cache_bytes >= off) {
    buf = c->shard_cache;
    buf_bytes = c->shard_cache_bytes;
    c->shard_cache = nullptr;
    c->shard_cache_bytes = 0;
  } else {
    cudaError_t e = cudaMalloc(&buf, off);
    if (e != cudaSuccess) return fail(REEF_ENOMEM, std::string("nl_shard_begin: ") + cudaGetErrorString(e));
  }
  char* d = (char*)buf;
  reef_nl_session* s = new reef_nl_session;
  s->buf_bytes = buf_bytes;
  s->mb = MbRef{nullptr, nullptr, nullptr, 0, 0, 0};
  s->ctx = c;
  s->d_buf = buf;
  s->d_table = a.d_table;
  s->is_u32 = a.table_is_u32;
  s->n_loc = n_loc;
  s->ell = a.ell;
  s->ell_loc = ell_loc;
  s->gamma = gamma;
  s->rank = rank;
  s->world = world;
  s->m = a.m;
  s->st = (NlState*)(d + o_state);
  s->d_query = (Fq*)(d + o_query);
  s->d_prevq = (Fq*)(d + o_prevq);
  s->d_q = (uint64_t*)(d + o_q);
  s->d_pos = (uint64_t*)(d + o_pos);
  s->d_w = (Fq*)(d + o_w);
  s->d_A = (Fq*)(d + o_A);
  s->d_B = (Fq*)(d + o_B);
  s->d_part = (Fq*)(d + o_part);
  s->d_fold = (Fq*)(d + o_fold);
  s->d_Ts = (Fq*)(d + o_Ts);
  s->d_Es = (Fq*)(d + o_Es);
  s->round = 0;
  s->L = n_loc;
  s->a_cur = a_len;
  s->small = 0;
  s->nblk = 0;
  cudaStream_t st = c->stream;
  auto bail = [&](int rc) {
    cudaStreamSynchronize(st);
    cudaFree(buf);
    delete s;
    return rc;
  };
  if (cudaMemcpyAsync(s->d_query, a.h_query, (size_t)a.n_query * 32, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(s->d_prevq, a.h_prev_q, (size_t)a.ell * 32, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      (a.m && cudaMemcpyAsync(s->d_q, a.h_q, (size_t)a.m * 8, cudaMemcpyHostToDevice, st) != cudaSuccess))
    return bail(fail(REEF_ECUDA, "nl_shard_begin: upload failed"));
  Fq tag = fq_canon_from_le32(a.tag_le);   // canonical: the LP sponge absorbs plain residues
  {
    ProfScope ps(c, PROF_NL_SETUP, n_loc);
    k_nl_begin<<<1, TR_THREADS, exclusive_smem(c, (const void*)k_nl_begin, 0), st>>>(s->st, s->d_query, a.n_query, tag, s->d_prevq, a.ell, s->d_q, a.m, s->d_pos, s->d_w, c->d_lp, rank, world);
    g_launches.fetch_add(1);
    k_eq_tables<<<ceil_div_u(a_len + b_len, 128), 128, 0, st>>>(s->st, ell_loc, hb, s->d_A, a_len, s->d_B, b_len, gamma, rank);
    g_launches.fetch_add(1);
  }
  if (cudaGetLastError() != cudaSuccess) return bail(fail(REEF_ECUDA, "nl_shard_begin: launch failed"));
  cudaStreamSynchronize(st);   // the host staging vectors of the caller may go away
  ctx_retain(c);
  *out = s;
  return REEF_OK;
}

template <bool U32IN>
static int nl_shard_round_local_t(reef_nl_session* s, void* d_out3) {
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  const size_t tail_smem = 0;
  (void)tail_smem;
  if (!s->small && (s->L >> 1) >= (uint64_t)CHUNK) {
    uint32_t nblk = 0;
    int rc;
    if (s->round == 0) {
      {
        ProfScope ps(c, PROF_SWEEP_FIRST, s->n_loc);
        rc = launch_sweep<U32IN, false>(c, s->d_table, s->n_loc, nullptr, s->st, s->d_A, s->d_B, s->d_part, &nblk);
        if (rc) return rc;
      }
      ProfScope ps(c, PROF_ROUND, s->L);
      k_shard_local<U32IN><<<1, ROUND_THREADS, 0, st>>>(s->d_part, nblk, s->d_table, s->L, s->d_pos, s->d_w, s->m, (Fq*)d_out3, s->mb);
    } else {
      {
        ProfScope ps(c, PROF_SWEEP_FOLD, 2 * s->L);
        if (s->round == 1) rc = launch_sweep<U32IN, true>(c, s->d_table, 2 * s->L, s->d_fold, s->st, s->d_A, s->d_B, s->d_part, &nblk);
        else rc = launch_sweep<false, true>(c, s->d_fold, 2 * s->L, s->d_fold, s->st, s->d_A, s->d_B, s->d_part, &nblk);
        if (rc) return rc;
      }
      ProfScope ps(c, PROF_ROUND, s->L);
      k_shard_local<false><<<1, ROUND_THREADS, 0, st>>>(s->d_part, nblk, s->d_fold, s->L, s->d_pos, s->d_w, s->m, (Fq*)d_out3, s->mb);
    }
    REEF_LAUNCHED();
    return REEF_OK;
  }
  ProfScope ps_tail(c, PROF_TAIL, s->L);
  if (!s->small) {   // enter the small regime
    if (s->round == 0) k_shard_materialize<U32IN><<<1, TAIL_THREADS, 0, st>>>(s->st, s->d_table, s->n_loc, 0, s->d_A, s->d_B, s->d_pos, s->d_w, s->m, s->d_Ts, s->d_Es);
    else if (s->round == 1) k_shard_materialize<U32IN><<<1, TAIL_THREADS, 0, st>>>(s->st, s->d_table, 2 * s->L, 1, s->d_A, s->d_B, s->d_pos, s->d_w, s->m, s->d_Ts, s->d_Es);
    else k_shard_materialize<false><<<1, TAIL_THREADS, 0, st>>>(s->st, s->d_fold, 2 * s->L, 1, s->d_A, s->d_B, s->d_pos, s->d_w, s->m, s->d_Ts, s->d_Es);
    REEF_LAUNCHED();
    s->small = 1;
  }
  k_small_local<<<1, TAIL_THREADS, 0, st>>>(s->d_Ts, s->d_Es, s->L, (Fq*)d_out3, s->mb);
  REEF_LAUNCHED();
  return REEF_OK;
}

int nl_shard_round_local(reef_nl_session* s, void* d_out3) {
  REEF_REQUIRE(s->round < s->ell_loc, REEF_EASSERT, "nl_shard_round_local: all local rounds are done");
  return s->is_u32 ? nl_shard_round_local_t<true>(s, d_out3) : nl_shard_round_local_t<false>(s, d_out3);
}

int nl_shard_round_finish(reef_nl_session* s, const void* d_triples) {
  REEF_REQUIRE(s->round < s->ell_loc, REEF_EASSERT, "nl_shard_round_finish: all local rounds are done");
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  ProfScope ps(c, s->small ? PROF_TAIL : PROF_ROUND, s->L);
  if (!s->small) {
    k_shard_apply<<<1, ROUND_THREADS, exclusive_smem(c, (const void*)k_shard_apply, 0), st>>>(s->st, (const Fq*)d_triples, s->world, s->L, s->d_A, s->a_cur, s->d_pos, s->d_w, s->m, s->round, c->d_lp, s->mb);
    if (s->a_cur > 1) s->a_cur >>= 1;
  } else {
    k_small_apply<<<1, TAIL_THREADS, exclusive_smem(c, (const void*)k_small_apply, 0), st>>>(s->st, (const Fq*)d_triples, s->world, s->d_Ts, s->d_Es, s->L, s->round, c->d_lp, s->mb);
  }
  REEF_LAUNCHED();
  s->L >>= 1;
  s->round++;
  return REEF_OK;
}

int nl_shard_export(reef_nl_session* s, void* d_out2) {
  REEF_REQUIRE(s->round == s->ell_loc && s->small, REEF_EASSERT, "nl_shard_export: local rounds not finished");
  k_shard_export<<<1, 32, 0, s->ctx->stream>>>(s->d_Ts, s->d_Es, (Fq*)d_out2, s->mb);
  REEF_LAUNCHED();
  return REEF_OK;
}

int nl_shard_finish(reef_nl_session* s, const void* d_pairs, uint8_t* out_claim_r, uint8_t* out_rounds, uint8_t* out_last_claim,
                    uint8_t* out_next_v) {
  REEF_REQUIRE(s->round == s->ell_loc, REEF_EASSERT, "nl_shard_finish: local rounds not finished");
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  {
    ProfScope ps(c, PROF_TAIL, s->world);
    k_shard_final<<<1, TR_THREADS, exclusive_smem(c, (const void*)k_shard_final, 0), st>>>(s->st, (const Fq*)d_pairs, s->world, s->ell_loc, c->d_lp, s->mb);
    REEF_LAUNCHED();
  }
  void* hs;
  int rc = ctx_stage(c, sizeof(NlState), &hs);
  if (rc) return rc;
  REEF_CUDA(cudaMemcpyAsync(hs, s->st, sizeof(NlState), cudaMemcpyDeviceToHost, st));
  REEF_CUDA(cudaStreamSynchronize(st));
  const NlState* h = (const NlState*)hs;
  memcpy(out_claim_r, h->out_claim_r.v, 32);
  for (uint32_t i = 0; i < s->ell; i++)
    for (int k = 0; k < 4; k++) memcpy(out_rounds + ((size_t)i * 4 + k) * 32, h->out_rounds[i][k].v, 32);
  memcpy(out_last_claim, h->out_last_claim.v, 32);
  memcpy(out_next_v, h->out_next_v.v, 32);
  return REEF_OK;
}

// One round with the exchange fused into the kernels: k_shard_local / k_small_local post this rank's
// triple into every peer's mailbox, k_shard_apply / k_small_apply acquire the peers' triples.
int nl_shard_round_p2p(reef_nl_session* s) {
  reef_ctx* c = s->ctx;
  REEF_REQUIRE(c->mb_world == s->world && c->mb_rank == s->rank && s->world <= MB_MAX_WORLD, REEF_EINVAL,
               "nl_shard_round_p2p: the context's mailbox is not connected for this rank / world");
  REEF_REQUIRE(s->round < s->ell_loc, REEF_EASSERT, "nl_shard_round_p2p: all local rounds are done");
  s->mb = MbRef{c->mb_peers_dev, (const unsigned char*)c->mb_mine, c->mb_err_dev, c->mb_world, c->mb_rank, c->mb_seq + 1};
  int rc = nl_shard_round_local(s, nullptr);
  if (!rc) {
    c->mb_seq++;            // the post of this exchange is in flight: the receive must follow
    rc = nl_shard_round_finish(s, nullptr);
  }
  s->mb.peers = nullptr;
  return rc;
}

int nl_shard_finish_p2p(reef_nl_session* s, uint8_t* out_claim_r, uint8_t* out_rounds, uint8_t* out_last_claim, uint8_t* out_next_v) {
  reef_ctx* c = s->ctx;
  REEF_REQUIRE(c->mb_world == s->world && c->mb_rank == s->rank && s->world <= MB_MAX_WORLD, REEF_EINVAL,
               "nl_shard_finish_p2p: the context's mailbox is not connected for this rank / world");
  REEF_REQUIRE(s->round == s->ell_loc && s->small, REEF_EASSERT, "nl_shard_finish_p2p: local rounds not finished");
  s->mb = MbRef{c->mb_peers_dev, (const unsigned char*)c->mb_mine, c->mb_err_dev, c->mb_world, c->mb_rank, c->mb_seq + 1};
  int rc = nl_shard_export(s, nullptr);
  if (!rc) {
    c->mb_seq++;
    rc = nl_shard_finish(s, nullptr, out_claim_r, out_rounds, out_last_claim, out_next_v);
  }
  s->mb.peers = nullptr;
  if (rc) return rc;
  uint32_t e = 0;
  REEF_CUDA(cudaMemcpy(&e, c->mb_err_dev, 4, cudaMemcpyDeviceToHost));   // nl_shard_finish has synchronised the stream
  if (e) {
    cudaMemset(c->mb_err_dev, 0, 4);
    return fail(REEF_ECUDA, "nl_shard_finish_p2p: exchange " + std::to_string(e & 0x7fffffffu) + " failed (a peer never posted, or a peer reported a timeout)");
  }
  return REEF_OK;
}

// Forces the (lazily loaded) kernels of the sharded path into the device before the first P2P
// exchange: loading a kernel synchronises with running work, which must not happen while an
// exchange kernel of this process is waiting for a peer that the same host thread has yet to launch.
int nl_shard_preload() {
  cudaFuncAttributes a;
  const void* fns[] = {(const void*)k_nl_begin, (const void*)k_eq_tables, (const void*)k_shard_local<true>, (const void*)k_shard_local<false>,
                       (const void*)k_shard_apply, (const void*)k_small_local, (const void*)k_small_apply,
                       (const void*)k_shard_materialize<true>, (const void*)k_shard_materialize<false>, (const void*)k_shard_export,
                       (const void*)k_shard_final, (const void*)k_sweep<true, false>, (const void*)k_sweep<true, true>,
                       (const void*)k_sweep<false, false>, (const void*)k_sweep<false, true>};
  for (const void* f : fns) REEF_CUDA(cudaFuncGetAttributes(&a, f));
  return REEF_OK;
}

reef_ctx* nl_shard_ctx(reef_nl_session* s) { return s->ctx; }

void nl_shard_free(reef_nl_session* s) {
  if (!s) return;
  reef_ctx* c = s->ctx;
  cudaStreamSynchronize(c->stream);
  if (c->closed.load()) {
    cudaFree(s->d_buf);
  } else if (!c->shard_cache) {
    c->shard_cache = s->d_buf;
    c->shard_cache_bytes = s->buf_bytes;
  } else if (c->shard_cache_bytes < s->buf_bytes) {
    cudaFree(c->shard_cache);
    c->shard_cache = s->d_buf;
    c->shard_cache_bytes = s->buf_bytes;
  } else {
    cudaFree(s->d_buf);
  }
  delete s;
}

}  // namespace reef
