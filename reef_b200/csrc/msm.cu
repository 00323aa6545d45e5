// Pippenger multi-scalar multiplication on the Pasta curves for sm_100a.
//
// Replaces nova-snark's `vartime_multiscalar_mul` / Pedersen `commit` as reached from
//   RecursiveSNARK::prove_step   /root/reference/src/backend/framework.rs:668-675   (commit(W), commit(T))
//   CompressedSNARK::prove       /root/reference/src/backend/framework.rs:695-698   (IPA rounds)
//   hyrax_gen.commit / prove_eval /root/reference/src/backend/commitment.rs:187, 371-393
// The result of an MSM is a unique group element, so parity is exact against the naive
// double-and-add oracle (oracle/curves.py).
//
// B200-first design (generators are static per PublicParams / per document commitment):
//   * bases are registered once: every window level 2^(c*w) * P_i is precomputed, normalised
//     to affine and kept resident in HBM (n * W * 64 B).  All W windows of a scalar then feed
//     ONE bucket set, so there is no per-window bucket reduction and no Horner chain, and a
//     multi-GPU split by windows needs only a sum of one partial point per GPU.
//   * signed digits (buckets 1 .. 2^(c-1)), counting sort of (bucket, point-ref) pairs with
//     warp-aggregated atomics, then segmented accumulation in fixed-size parts so that a
//     skewed histogram (Reef's all-'a' document: every scalar equal) cannot serialise on one
//     thread; remaining parts are combined by further K-ary passes.
//   * weighted bucket sum  sum_v v * B_v  by bit decomposition: c masked tree reductions
//     (warp-shuffle + shared memory) run as independent CTAs, then 2^t scaling in parallel.
#include <cstring>
#include <memory>
#include <vector>

#include "common.cuh"
#include "ec.cuh"
#include "kernels.h"

namespace reef {

static constexpr uint32_t K_FIRST = 8;      // entries per thread in the first accumulation pass (latency: 8 mixed adds);
                                            // 16 once that still leaves > 2048 threads per SM (halves the partials to combine)
static constexpr uint32_t K_NEXT = 128;     // partial points per WARP in the combine passes (4 per lane + shuffle tree)

// REEF_MSM_TMA=1: first accumulation pass with TMA-staged point tiles (k_accum_first_tma); default: plain gathers
static bool msm_tma_enabled() {
  static const bool on = getenv("REEF_MSM_TMA") && atoi(getenv("REEF_MSM_TMA")) != 0;
  return on;
}
static constexpr size_t MSM_TMA_SMEM = 2 * 4 * 128 * 64;

// "Polite" background contexts (reef_init_prio(.., 0, ..) with REEF_MSM_POLITE=1): the long accumulation grids ask for
// enough dynamic shared memory that only two of their CTAs fit an SM.  Four of them fill the register file (106-136
// registers x 128 threads each), and a sweep CTA of a latency-critical sum-check (17 k registers) then waits for one of
// them to retire; with two there is always room for it.  The commitments of fold i have the whole sum-check of fold
// i+1 to hide behind, so their own slowdown is free; the last fold's commitments use a context that is not polite.
static size_t polite_smem(reef_ctx* c, const void* kernel) {
  if (!c->polite) return 0;
  static std::mutex mu;
  static std::vector<const void*> done;
  const size_t bytes = 100 * 1024;
  std::lock_guard<std::mutex> lk(mu);
  for (const void* k : done)
    if (k == kernel) return bytes;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  done.push_back(kernel);
  return bytes;
}

struct MsmPlan {
  uint32_t c;          // window bits
  uint32_t W;          // windows per scalar
  uint32_t L;          // precomputed levels (W == L * G, last group may be short)
  uint32_t G;          // bucket groups (1 when fully precomputed)
  uint32_t B;          // buckets per group = 2^(c-1)
};

// ---------------------------------------------------------------------------------------
// loads / stores
// ---------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ Affine<C> ld_affine(const Affine<C>* p) {
  Affine<C> r;
  r.x = ld256(&p->x);
  r.y = ld256(&p->y);
  return r;
}
template <class C>
__device__ __forceinline__ void st_affine(Affine<C>* p, const Affine<C>& v) {
  st256(&p->x, v.x);
  st256(&p->y, v.y);
}
template <class C>
__device__ __forceinline__ XYZZ<C> ld_xyzz(const XYZZ<C>* p) {
  XYZZ<C> r;
  r.x = ld256(&p->x);
  r.y = ld256(&p->y);
  r.zz = ld256(&p->zz);
  r.zzz = ld256(&p->zzz);
  return r;
}
template <class C>
__device__ __forceinline__ void st_xyzz(XYZZ<C>* p, const XYZZ<C>& v) {
  st256(&p->x, v.x);
  st256(&p->y, v.y);
  st256(&p->zz, v.zz);
  st256(&p->zzz, v.zzz);
}
template <class C>
__device__ __forceinline__ XYZZ<C> shfl_xor_xyzz(const XYZZ<C>& p, int m) {
  XYZZ<C> r;
  r.x = shfl_xor_fe(p.x, m);
  r.y = shfl_xor_fe(p.y, m);
  r.zz = shfl_xor_fe(p.zz, m);
  r.zzz = shfl_xor_fe(p.zzz, m);
  return r;
}

// ---------------------------------------------------------------------------------------
// registration: level[w][i] = 2^(c*w) * P_i, affine, Montgomery form
// ---------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128) k_precompute(const Affine<C>* __restrict__ in_canon, uint64_t n, uint32_t c,
                                                    uint32_t L, Affine<C>* __restrict__ levels, int* bad) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<C> p = ld_affine(in_canon + i);
  bool inf = affine_is_inf<C>(p);
  p.x = to_mont<C>(p.x);
  p.y = to_mont<C>(p.y);
  if (!inf) {  // y^2 == x^3 + 5
    Fe<C> lhs = mont_sqr<C>(p.y);
    Fe<C> rhs = fe_add<C>(mont_mul<C>(mont_sqr<C>(p.x), p.x), fe_from_u64<C>(5));
    if (!fe_eq<C>(lhs, rhs)) atomicExch(bad, 1);
  }
  st_affine(levels + i, p);
#pragma unroll 1
  for (uint32_t w = 1; w < L; w++) {
    XYZZ<C> a = xyzz_dbl_affine<C>(p);
#pragma unroll 1
    for (uint32_t k = 1; k < c; k++) a = xyzz_dbl<C>(a);
    p = xyzz_to_affine<C>(a);
    st_affine(levels + (uint64_t)w * n + i, p);
  }
}

// ---------------------------------------------------------------------------------------
// digits: entry e = w * n + i  ->  key (bucket or sentinel), val (level point ref | sign)
// ---------------------------------------------------------------------------------------
template <bool U32SCALARS>
__global__ void k_digits(const void* __restrict__ scalars, uint64_t n, uint64_t n_bases, MsmPlan pl, uint32_t w_begin,
                         uint32_t w_end, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[9];
  if constexpr (U32SCALARS) {
    s[0] = ((const uint32_t*)scalars)[i];
#pragma unroll
    for (int k = 1; k < 9; k++) s[k] = 0;
  } else {
    Fe<FqCfg> x = ld256((const Fe<FqCfg>*)scalars + i);
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = x.v[k];
    s[8] = 0;
  }
  const uint32_t c = pl.c, half = 1u << (c - 1), full = 1u << c, sentinel = pl.G * pl.B;
  uint32_t carry = 0;
  for (uint32_t w = 0; w < pl.W; w++) {
    const uint32_t bit = w * c, limb = bit >> 5, sh = bit & 31;
    uint32_t raw = 0;
    if (limb < 8) {
      uint64_t two = (uint64_t)s[limb] | ((uint64_t)s[limb + 1] << 32);
      raw = (uint32_t)(two >> sh) & (full - 1);
    }
    raw += carry;
    uint32_t mag, neg;
    if (raw > half) {
      mag = full - raw;
      neg = 1;
      carry = 1;
    } else {
      mag = raw;
      neg = 0;
      carry = 0;
    }
    if (w >= w_begin && w < w_end) {
      const uint64_t e = (uint64_t)(w - w_begin) * n + i;
      const uint32_t level = w % pl.L, group = w / pl.L;
      keys[e] = mag ? group * pl.B + (mag - 1) : sentinel;
      vals[e] = (uint32_t)((uint64_t)level * n_bases + i) | (neg << 31);
    }
  }
}

// ---------------------------------------------------------------------------------------
// counting sort
// ---------------------------------------------------------------------------------------
// Zero digits carry the sentinel key and are never scattered, so they are not counted either: with witness-like
// scalars (most entries 0 / 1 / 16-bit) nine digits in ten are zero, and counting them meant one same-address
// atomic per warp for almost every warp of the grid.
__global__ void k_hist(const uint32_t* __restrict__ keys, uint64_t n_entries, uint32_t sentinel, uint32_t* __restrict__ counts) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const uint32_t key = keys[e];
  const unsigned act = __activemask();
  const unsigned same = __match_any_sync(act, key);
  if (key != sentinel && (threadIdx.x & 31) == (unsigned)(__ffs(same) - 1)) atomicAdd(counts + key, (uint32_t)__popc(same));
}

// exclusive scan of ceil(cnt[b] / K) (K = 1: plain scan) by one CTA; also the max of cnt.
// Tiles of 4096 counts: every thread takes four consecutive ones (one 16-byte load, coalesced), warp scan by shuffles,
// one barrier pair per tile.  (One thread per contiguous slice of nb / 1024 counts was 168 us at nb = 2^16 -- 64
// uncoalesced dependent loads per thread -- and four of them sat in every row-batched commitment.)
__global__ void __launch_bounds__(1024) k_scan(const uint32_t* __restrict__ cnt, uint32_t nb, uint32_t K,
                                               uint32_t* __restrict__ parts, uint32_t* __restrict__ off,
                                               uint32_t* __restrict__ cursor, uint32_t* __restrict__ total_max) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t wmax[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec = (((uintptr_t)cnt | (uintptr_t)off | (uintptr_t)parts | (uintptr_t)cursor) & 15u) == 0;
  uint32_t carry = 0, mx = 0;
  for (uint32_t base = 0; base < nb; base += 4096) {
    const uint32_t b0 = base + 4u * threadIdx.x;
    const bool whole = vec && b0 + 4 <= nb;
    uint32_t p[4] = {0, 0, 0, 0};
    if (whole) {
      const uint4 v = *reinterpret_cast<const uint4*>(cnt + b0);
      p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (b0 + i < nb) p[i] = cnt[b0 + i];
    }
    if (K != 1) {
#pragma unroll
      for (int i = 0; i < 4; i++) p[i] = (p[i] + K - 1) / K;
    }
    mx = max(max(mx, max(p[0], p[1])), max(p[2], p[3]));
    const uint32_t t = p[0] + p[1] + p[2] + p[3];
    uint32_t inc = t;                                  // inclusive scan inside the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    const uint32_t ws = wsum[lane];                    // every warp scans the 32 warp totals
    uint32_t winc = ws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += v;
    }
    const uint32_t warp_excl = __shfl_sync(0xffffffffu, winc - ws, warp);
    const uint32_t tile_total = __shfl_sync(0xffffffffu, winc, 31);
    uint32_t run = carry + warp_excl + (inc - t);
    const uint32_t o0 = run, o1 = o0 + p[0], o2 = o1 + p[1], o3 = o2 + p[2];
    if (whole) {
      const uint4 o = make_uint4(o0, o1, o2, o3);
      *reinterpret_cast<uint4*>(off + b0) = o;
      if (cursor) *reinterpret_cast<uint4*>(cursor + b0) = o;
      if (parts) *reinterpret_cast<uint4*>(parts + b0) = make_uint4(p[0], p[1], p[2], p[3]);
    } else {
      const uint32_t o[4] = {o0, o1, o2, o3};
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (b0 + i < nb) {
          off[b0 + i] = o[i];
          if (cursor) cursor[b0 + i] = o[i];
          if (parts) parts[b0 + i] = p[i];
        }
    }
    carry += tile_total;
    __syncthreads();                                   // wsum is rewritten by the next tile
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) wmax[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    const uint32_t m = __reduce_max_sync(0xffffffffu, wmax[lane]);
    if (lane == 0) {
      off[nb] = carry;
      total_max[0] = carry;
      total_max[1] = m;
    }
  }
}

__global__ void k_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n_entries,
                          uint32_t sentinel, uint32_t* __restrict__ cursor, uint32_t* __restrict__ sorted) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_entries) return;
  const uint32_t key = keys[e];
  const unsigned act = __activemask();
  const unsigned same = __match_any_sync(act, key);
  if (key == sentinel) return;
  const int lane = threadIdx.x & 31, leader = __ffs(same) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(cursor + key, (uint32_t)__popc(same));
  base = __shfl_sync(same, base, leader);
  sorted[base + __popc(same & ((1u << lane) - 1))] = vals[e];
}

// part p of a segmented list -> (bucket b, first element, count)
__device__ __forceinline__ void locate_part(const uint32_t* __restrict__ part_off, uint32_t nb, uint32_t p,
                                            uint32_t& b) {
  uint32_t lo = 0, hi = nb;  // largest b with part_off[b] <= p
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (part_off[mid] <= p) lo = mid;
    else hi = mid;
  }
  b = lo;
}

// first pass: sorted point refs -> one XYZZ partial per part (mixed additions)
template <class C>
__global__ void __launch_bounds__(128) k_accum_first(const uint32_t* __restrict__ sorted,
                                                     const uint32_t* __restrict__ start, const uint32_t* __restrict__ cnt,
                                                     const uint32_t* __restrict__ part_off, uint32_t nb, uint32_t n_parts,
                                                     uint32_t kfirst, const Affine<C>* __restrict__ levels,
                                                     XYZZ<C>* __restrict__ out) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_parts) return;
  uint32_t b;
  locate_part(part_off, nb, p, b);
  const uint32_t j = p - part_off[b];
  const uint32_t first = start[b] + j * kfirst;
  const uint32_t last = min(start[b] + cnt[b], first + kfirst);
  XYZZ<C> acc = xyzz_inf<C>();
#pragma unroll 1
  for (uint32_t e = first; e < last; e++) {
    const uint32_t v = sorted[e];
    Affine<C> q = ld_affine(levels + (v & 0x7fffffffu));
    xyzz_add_affine<C>(acc, q, (v >> 31) != 0);
  }
  st_xyzz(out + p, acc);
}
// The same pass with the gathered points staged through shared memory by the TMA engine (`cp.async.bulk`, SASS
// UBLKCP): every thread reads the point references of its part, posts one 64-byte bulk copy per point into its own
// slots of a shared tile, all completing on one mbarrier per stage, and adds the points out of shared memory.
// Two stages of TMA_STAGE points per thread are in flight from the start, so the L2 / HBM latency of the gather is
// paid once per stage instead of once per point.  A/B against k_accum_first: profiles/r02_summary.md (REEF_MSM_TMA).
static constexpr uint32_t TMA_STAGE = 4;          // points per thread and stage (2 stages x 128 threads x 4 x 64 B = 64 KiB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <class C>
__global__ void __launch_bounds__(128) k_accum_first_tma(const uint32_t* __restrict__ sorted,
                                                         const uint32_t* __restrict__ start, const uint32_t* __restrict__ cnt,
                                                         const uint32_t* __restrict__ part_off, uint32_t nb, uint32_t n_parts,
                                                         uint32_t kfirst, const Affine<C>* __restrict__ levels,
                                                         XYZZ<C>* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char tma_smem[];
  Affine<C>* tile = reinterpret_cast<Affine<C>*>(tma_smem);                    // [2][TMA_STAGE][128]
  __shared__ __align__(8) uint64_t bar[2];
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 128);
    mbar_init(&bar[1], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t first = 0, last = 0;
  if (p < n_parts) {
    uint32_t b;
    locate_part(part_off, nb, p, b);
    const uint32_t j = p - part_off[b];
    first = start[b] + j * kfirst;
    last = min(start[b] + cnt[b], first + kfirst);
  }
  uint32_t neg[2] = {0, 0};
  // stage s of round t covers entries [first + (2 t + s) * TMA_STAGE, + TMA_STAGE)
  auto post = [&](uint32_t s, uint32_t e0) {
    uint32_t k = 0, bits = 0;
    uint32_t refs[TMA_STAGE];
#pragma unroll
    for (uint32_t i = 0; i < TMA_STAGE; i++) {
      if (e0 + i < last) {
        refs[i] = sorted[e0 + i];
        bits |= (refs[i] >> 31) << i;
        k++;
      }
    }
    mbar_arrive_expect_tx(&bar[s], k * (uint32_t)sizeof(Affine<C>));
#pragma unroll
    for (uint32_t i = 0; i < TMA_STAGE; i++)
      if (i < k) bulk_g2s(tile + ((size_t)s * TMA_STAGE + i) * 128 + threadIdx.x, levels + (refs[i] & 0x7fffffffu), (uint32_t)sizeof(Affine<C>), &bar[s]);
    neg[s] = bits;
  };
  post(0, first);
  post(1, first + TMA_STAGE);
  XYZZ<C> acc = xyzz_inf<C>();
  const uint32_t rounds = (kfirst + 2 * TMA_STAGE - 1) / (2 * TMA_STAGE);      // the same for every thread of the grid
#pragma unroll 1
  for (uint32_t t = 0; t < rounds; t++) {
#pragma unroll 1
    for (uint32_t s = 0; s < 2; s++) {
      const uint32_t e0 = first + (2 * t + s) * TMA_STAGE;
      mbar_wait(&bar[s], t & 1);
#pragma unroll 1
      for (uint32_t i = 0; i < TMA_STAGE; i++) {
        if (e0 + i < last) {
          const Affine<C>* q = tile + ((size_t)s * TMA_STAGE + i) * 128 + threadIdx.x;
          Affine<C> a;
          a.x = q->x;
          a.y = q->y;
          xyzz_add_affine<C>(acc, a, ((neg[s] >> i) & 1) != 0);
        }
      }
      if (t + 1 < rounds) {
        __syncthreads();                                 // every thread is done with stage s of this round
        post(s, e0 + 2 * TMA_STAGE);
      }
    }
  }
  if (p < n_parts) st_xyzz(out + p, acc);
}

// combine passes: LANES lanes per (bucket, chunk of K partials): the lanes stride over the chunk, then a
// log2(LANES)-level shuffle tree.  LANES = 32 (K = 128) when there are few buckets (latency: 4 + 5
// dependent additions), LANES = 4 (K = 64) when there are enough buckets to fill the chip with
// 4-lane groups (throughput: 16 + 2 additions per group instead of 9 warp-wide ones per 128 partials,
// i.e. 2x fewer warp-instructions per partial).
template <class C, int LANES>
__global__ void __launch_bounds__(128) k_accum_next(const XYZZ<C>* __restrict__ in, const uint32_t* __restrict__ in_off,
                                                    const uint32_t* __restrict__ in_cnt,
                                                    const uint32_t* __restrict__ part_off, uint32_t nb, uint32_t n_parts,
                                                    uint32_t K, XYZZ<C>* __restrict__ out) {
  const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t p = gid / LANES;
  const int sub = threadIdx.x & (LANES - 1);
  const bool live = p < n_parts && p < part_off[nb];   // n_parts is a host-side upper bound
  if (LANES == 32 && !live) return;                    // warp-uniform
  XYZZ<C> acc = xyzz_inf<C>();
  uint32_t width = 0;
  if (live) {
    uint32_t b;
    locate_part(part_off, nb, p, b);
    const uint32_t j = p - part_off[b];
    const uint32_t first = in_off[b] + j * K;
    const uint32_t last = min(in_off[b] + in_cnt[b], first + K);
#pragma unroll 1
    for (uint32_t e = first + sub; e < last; e += LANES) xyzz_add<C>(acc, ld_xyzz(in + e));
    width = last - first;
  }
  // groups of one warp may have different widths: every lane runs all levels (infinity adds are cheap)
#pragma unroll 1
  for (int m = 1; m < LANES; m <<= 1) {
    if (LANES == 32 && (uint32_t)m >= width) break;    // warp-uniform for whole-warp groups
    XYZZ<C> o = shfl_xor_xyzz(acc, m);
    xyzz_add<C>(acc, o);
  }
  if (live && sub == 0) st_xyzz(out + p, acc);
}
// buckets[b] = (cnt[b] ? parts[off[b]] : infinity)   (after the last pass every count is <= 1)
template <class C>
__global__ void k_gather_buckets(const XYZZ<C>* __restrict__ in, const uint32_t* __restrict__ off,
                                 const uint32_t* __restrict__ cnt, uint32_t nb, XYZZ<C>* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  XYZZ<C> v = cnt[b] ? ld_xyzz(in + off[b]) : xyzz_inf<C>();
  st_xyzz(buckets + b, v);
}

// ---------------------------------------------------------------------------------------
// weighted bucket sum by bit decomposition:  sum_b (b+1) * bucket[b] = sum_t 2^t * S_t,
//   S_t = sum of buckets whose weight (b+1) has bit t set.
// grid (nblk, c_bits, G): each CTA tree-reduces 256 buckets for one bit of one group.
// ---------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ XYZZ<C> block_sum_xyzz(XYZZ<C> v, XYZZ<C>* sm /* blockDim/32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll 1
  for (int m = 16; m >= 1; m >>= 1) {
    XYZZ<C> o = shfl_xor_xyzz(v, m);
    xyzz_add<C>(v, o);      // both halves compute the same sum (commutative up to representation)
  }
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  XYZZ<C> r = xyzz_inf<C>();
  if (warp == 0) {
    r = lane < nw ? sm[lane] : xyzz_inf<C>();
#pragma unroll 1
    for (int m = 1; m < nw; m <<= 1) {
      XYZZ<C> o = shfl_xor_xyzz(r, m);
      xyzz_add<C>(r, o);
    }
  }
  __syncthreads();
  return r;  // valid on warp 0
}

// Each thread first sums `bpt` buckets of its own (stride 256) before the CTA tree: with many buckets
// the 8-level tree is amortised over bpt times more points (bpt = 1 keeps the latency of small MSMs).
template <class C>
__global__ void __launch_bounds__(256) k_bitsum_partial(const XYZZ<C>* __restrict__ buckets, uint32_t B, uint32_t bpt,
                                                        XYZZ<C>* __restrict__ partial) {
  __shared__ XYZZ<C> sm[8];
  const uint32_t t = blockIdx.y, g = blockIdx.z;
  XYZZ<C> v = xyzz_inf<C>();
#pragma unroll 1
  for (uint32_t k = 0; k < bpt; k++) {
    const uint32_t b = (blockIdx.x * bpt + k) * 256 + threadIdx.x;
    if (b < B && (((b + 1) >> t) & 1)) xyzz_add<C>(v, ld_xyzz(buckets + (uint64_t)g * B + b));
  }
  XYZZ<C> r = block_sum_xyzz<C>(v, sm);
  if (threadIdx.x == 0) st_xyzz(partial + ((uint64_t)g * gridDim.y + t) * gridDim.x + blockIdx.x, r);
}

// Throughput shape of the same step (row-batched commitments: thousands of (row, bit) pairs over a few hundred buckets
// each): ONE WARP per (row, bit).  Every lane first sums its B / 32 buckets of that bit serially, then one 5-level
// shuffle tree -- B / 32 + 5 warp-wide additions per pair instead of the 8 x (256 / 32) of the CTA tree above, whose
// upper levels add mostly copies.  Output layout = k_bitsum_partial with nblk = 1.
template <class C>
__global__ void __launch_bounds__(256) k_bitsum_partial_warp(const XYZZ<C>* __restrict__ buckets, uint32_t B, uint32_t cbits,
                                                             uint32_t n_pairs, XYZZ<C>* __restrict__ partial) {
  const uint32_t pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // = g * cbits + t
  const int lane = threadIdx.x & 31;
  if (pair >= n_pairs) return;                                             // warp-uniform
  const uint32_t g = pair / cbits, t = pair - g * cbits;
  XYZZ<C> v = xyzz_inf<C>();
#pragma unroll 1
  for (uint32_t b = lane; b < B; b += 32)
    if (((b + 1) >> t) & 1) xyzz_add<C>(v, ld_xyzz(buckets + (uint64_t)g * B + b));
#pragma unroll 1
  for (int m = 16; m >= 1; m >>= 1) {
    XYZZ<C> o = shfl_xor_xyzz(v, m);
    xyzz_add<C>(v, o);
  }
  if (lane == 0) st_xyzz(partial + pair, v);
}

// one CTA per group: S_t = sum of nblk partials (one warp per bit), scale by 2^t, sum over t,
// then combine the groups by Horner (group g carries weight 2^(c*L*g)) and normalise.
template <class C>
__global__ void __launch_bounds__(1024) k_bitsum_final(const XYZZ<C>* __restrict__ partial, uint32_t nblk,
                                                       uint32_t cbits, uint32_t G, uint32_t group_shift,
                                                       XYZZ<C>* __restrict__ out_xyzz, Affine<C>* __restrict__ out_affine,
                                                       const XYZZ<C>* __restrict__ extra, uint32_t n_extra) {
  __shared__ XYZZ<C> st[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  XYZZ<C> total = xyzz_inf<C>();
  for (int g = (int)G - 1; g >= 0; g--) {
    // warp t: sum of the nblk partials of bit t, then t doublings
    XYZZ<C> v = xyzz_inf<C>();
    if ((uint32_t)warp < cbits) {
      for (uint32_t k = lane; k < nblk; k += 32) xyzz_add<C>(v, ld_xyzz(partial + ((uint64_t)g * cbits + warp) * nblk + k));
#pragma unroll 1
      for (uint32_t m = 1; m < 32 && m < nblk; m <<= 1) {
        XYZZ<C> o = shfl_xor_xyzz(v, (int)m);
        xyzz_add<C>(v, o);
      }
#pragma unroll 1
      for (int k = 0; k < warp; k++) v = xyzz_dbl<C>(v);
    }
    if (lane == 0) st[warp] = v;
    __syncthreads();
    if (warp == 0) {
      XYZZ<C> r = (uint32_t)lane < cbits ? st[lane] : xyzz_inf<C>();
#pragma unroll 1
      for (uint32_t m = 1; m < 32 && m < cbits; m <<= 1) {
        XYZZ<C> o = shfl_xor_xyzz(r, (int)m);
        xyzz_add<C>(r, o);
      }
      if (g != (int)G - 1) {
#pragma unroll 1
        for (uint32_t k = 0; k < group_shift; k++) total = xyzz_dbl<C>(total);
      }
      xyzz_add<C>(total, r);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    for (uint32_t k = 0; k < n_extra; k++) xyzz_add<C>(total, ld_xyzz(extra + k));
    if (out_xyzz) st_xyzz(out_xyzz, total);
    if (out_affine) {
      Affine<C> a = xyzz_to_affine<C>(total);
      a.x = from_mont<C>(a.x);
      a.y = from_mont<C>(a.y);
      st_affine(out_affine, a);
    }
  }
}

// sum of k XYZZ points given in canonical (non-Montgomery) coordinates -> affine canonical
template <class C>
__global__ void k_combine(const XYZZ<C>* __restrict__ pts_canon, uint32_t k, Affine<C>* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ<C> total = xyzz_inf<C>();
  for (uint32_t i = 0; i < k; i++) {
    XYZZ<C> p = ld_xyzz(pts_canon + i);
    p.x = to_mont<C>(p.x);
    p.y = to_mont<C>(p.y);
    p.zz = to_mont<C>(p.zz);
    p.zzz = to_mont<C>(p.zzz);
    xyzz_add<C>(total, p);
  }
  Affine<C> a = xyzz_to_affine<C>(total);
  a.x = from_mont<C>(a.x);
  a.y = from_mont<C>(a.y);
  st_affine(out, a);
}

template <class C>
__global__ void k_xyzz_from_mont(XYZZ<C>* p) {
  if (threadIdx.x || blockIdx.x) return;
  XYZZ<C> v = ld_xyzz(p);
  v.x = from_mont<C>(v.x);
  v.y = from_mont<C>(v.y);
  v.zz = from_mont<C>(v.zz);
  v.zzz = from_mont<C>(v.zzz);
  st_xyzz(p, v);
}

// ---------------------------------------------------------------------------------------
// Row-batched MSM (Hyrax document commitment, commitment.rs:187): `rows` independent MSMs over
// the SAME generators; bucket key = row * B + bucket, so one sort / one accumulation serves
// all rows.  Optional blinding term blind[r] * bases[blind_idx] per row.
// ---------------------------------------------------------------------------------------
template <bool U32SCALARS>
__global__ void k_digits_rows(const void* __restrict__ scalars, uint64_t rows, uint64_t cols, uint64_t col0,
                              uint64_t n_bases, MsmPlan pl, uint32_t w_used, uint64_t entry_base,
                              uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_terms = rows * cols;
  if (t >= n_terms) return;
  const uint64_t r = t / cols, j = t - r * cols;
  uint32_t s[9];
  if constexpr (U32SCALARS) {
    s[0] = ((const uint32_t*)scalars)[t];
#pragma unroll
    for (int k = 1; k < 9; k++) s[k] = 0;
  } else {
    Fe<FqCfg> x = ld256((const Fe<FqCfg>*)scalars + t);
#pragma unroll
    for (int k = 0; k < 8; k++) s[k] = x.v[k];
    s[8] = 0;
  }
  const uint32_t c = pl.c, half = 1u << (c - 1), full = 1u << c, sentinel = (uint32_t)(rows * pl.B);
  uint32_t carry = 0;
  for (uint32_t w = 0; w < w_used; w++) {
    const uint32_t bit = w * c, limb = bit >> 5, sh = bit & 31;
    uint32_t raw = 0;
    if (limb < 8) {
      uint64_t two = (uint64_t)s[limb] | ((uint64_t)s[limb + 1] << 32);
      raw = (uint32_t)(two >> sh) & (full - 1);
    }
    raw += carry;
    uint32_t mag, neg;
    if (raw > half) {
      mag = full - raw;
      neg = 1;
      carry = 1;
    } else {
      mag = raw;
      neg = 0;
      carry = 0;
    }
    const uint64_t e = entry_base + (uint64_t)w * n_terms + t;
    keys[e] = mag ? (uint32_t)(r * pl.B + (mag - 1)) : sentinel;
    vals[e] = (uint32_t)((uint64_t)w * n_bases + col0 + j) | (neg << 31);
  }
}

// one CTA per row: S_t from the row's bit-partials, 2^t scaling, sum, affine out
template <class C>
__global__ void __launch_bounds__(512) k_rows_final(const XYZZ<C>* __restrict__ partial, uint32_t nblk, uint32_t cbits,
                                                    Affine<C>* __restrict__ out, XYZZ<C>* __restrict__ out_xyzz) {
  __shared__ XYZZ<C> st[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t g = blockIdx.x;
  XYZZ<C> v = xyzz_inf<C>();
  if ((uint32_t)warp < cbits) {
    for (uint32_t k = lane; k < nblk; k += 32) xyzz_add<C>(v, ld_xyzz(partial + ((uint64_t)g * cbits + warp) * nblk + k));
#pragma unroll 1
    for (uint32_t m = 1; m < 32 && m < nblk; m <<= 1) {
      XYZZ<C> o = shfl_xor_xyzz(v, (int)m);
      xyzz_add<C>(v, o);
    }
#pragma unroll 1
    for (int k = 0; k < warp; k++) v = xyzz_dbl<C>(v);
  }
  if (lane == 0) st[warp] = v;
  __syncthreads();
  if (warp == 0) {
    XYZZ<C> r = (uint32_t)lane < cbits ? st[lane] : xyzz_inf<C>();
#pragma unroll 1
    for (uint32_t m = 1; m < 16 && m < cbits; m <<= 1) {
      XYZZ<C> o = shfl_xor_xyzz(r, (int)m);
      xyzz_add<C>(r, o);
    }
    if (lane == 0) {
      if (out_xyzz) {                      // a handful of rows: the inversion happens on the host (see msm_run_t)
        st_xyzz(out_xyzz + g, r);
      } else {
        Affine<C> a = xyzz_to_affine<C>(r);
        a.x = from_mont<C>(a.x);
        a.y = from_mont<C>(a.y);
        st_affine(out + g, a);
      }
    }
  }
}

// The same for nblk = 1 (one partial per bit) on ONE WARP per row: lane t holds S_t, doubles it t times, then a shuffle
// tree.  A CTA of c warps per row used one lane of each warp (1024 rows: 3.5 waves of two CTAs per SM, 239 us); four
// rows per 128-thread CTA are resident all at once.
template <class C>
__global__ void __launch_bounds__(128) k_rows_final_warp(const XYZZ<C>* __restrict__ partial, uint64_t rows, uint32_t cbits,
                                                         XYZZ<C>* __restrict__ out_xyzz) {
  const int lane = threadIdx.x & 31;
  const uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= rows) return;                               // warp-uniform
  XYZZ<C> v = (uint32_t)lane < cbits ? ld_xyzz(partial + g * cbits + lane) : xyzz_inf<C>();
#pragma unroll 1
  for (uint32_t k = 1; k < cbits; k++)
    if ((uint32_t)lane >= k && (uint32_t)lane < cbits) v = xyzz_dbl<C>(v);
#pragma unroll 1
  for (uint32_t m = 1; m < 32 && m < cbits; m <<= 1) {
    XYZZ<C> o = shfl_xor_xyzz(v, (int)m);
    xyzz_add<C>(v, o);
  }
  if (lane == 0) st_xyzz(out_xyzz + g, v);
}

// XYZZ -> affine canonical, one WARP per row with lane 0 at work: the inversion (binary Euclid, ~110 k cycles) is a
// data-dependent branch sequence, 32 of them in one warp serialise (210 us for 1024 rows); alone in its warp each
// runs at its own pace, and it no longer holds a whole CTA of k_rows_final on its SM
template <class C>
__global__ void __launch_bounds__(128) k_rows_affine(const XYZZ<C>* __restrict__ in, uint64_t rows, Affine<C>* __restrict__ out) {
  const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows || (threadIdx.x & 31) != 0) return;
  Affine<C> a = xyzz_to_affine<C>(ld_xyzz(in + r));
  a.x = from_mont<C>(a.x);
  a.y = from_mont<C>(a.y);
  st_affine(out + r, a);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

MsmPlanPublic msm_make_plan(uint64_t n, uint32_t scalar_bits, uint64_t max_level_bytes) {
  MsmPlanPublic p;
  uint32_t lg = 0;
  while (((uint64_t)1 << (lg + 1)) <= n) lg++;
  static const int delta = getenv("REEF_MSM_C_DELTA") ? atoi(getenv("REEF_MSM_C_DELTA")) : 0;   // tuning aid: window bits = log2(n) - 4 + delta
  int c = (int)lg - 4 + delta;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  if ((uint32_t)c > scalar_bits + 1) c = (int)scalar_bits + 1;
  p.c = (uint32_t)c;
  p.W = scalar_bits / p.c + 1;
  p.L = p.W;
  while (p.L > 1 && (uint64_t)p.L * n * 64 > max_level_bytes) p.L = (p.L + 1) / 2;
  p.G = (p.W + p.L - 1) / p.L;
  p.B = 1u << (p.c - 1);
  return p;
}

template <class C>
static int bases_register_t(reef_ctx* c, const uint8_t* h_bases, uint64_t n, const MsmPlanPublic& pl, void** d_levels_out) {
  void* d_levels = nullptr;
  const size_t bytes = (size_t)pl.L * n * sizeof(Affine<C>);
  cudaError_t e = cudaMalloc(&d_levels, bytes);
  if (e != cudaSuccess) return fail(REEF_ENOMEM, std::string("reef_bases_register: ") + cudaGetErrorString(e));
  void* base;
  int rc = ctx_scratch(c, (size_t)n * 64 + 256, &base);
  if (rc) {
    cudaFree(d_levels);
    return rc;
  }
  int* d_bad = (int*)((char*)base + (size_t)n * 64);
  cudaStream_t s = c->stream;
  e = cudaMemcpyAsync(base, h_bases, (size_t)n * 64, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, 4, s);
  if (e == cudaSuccess) {
    k_precompute<C><<<cdiv(n, 128), 128, 0, s>>>((const Affine<C>*)base, n, pl.c, pl.L, (Affine<C>*)d_levels, d_bad);
    e = cudaGetLastError();
  }
  int bad = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    cudaFree(d_levels);
    return fail(REEF_ECUDA, std::string("reef_bases_register: ") + cudaGetErrorString(e));
  }
  if (bad) {
    cudaFree(d_levels);
    return fail(REEF_EINVAL, "reef_bases_register: a base point is not on the curve");
  }
  *d_levels_out = d_levels;
  return REEF_OK;
}

int msm_bases_register(reef_ctx* c, int curve, const uint8_t* h_bases, uint64_t n, const MsmPlanPublic& pl, void** d_levels) {
  return curve == 0 ? bases_register_t<FpCfg>(c, h_bases, n, pl, d_levels) : bases_register_t<FqCfg>(c, h_bases, n, pl, d_levels);
}

// Runs windows [w_begin, w_end) of the MSM.  Result: affine canonical (64 B) to h_out_affine
// and/or XYZZ canonical coordinates (128 B) to h_out_xyzz (for a cross-GPU combine).
template <class C>
static int msm_run_t(reef_ctx* c, const MsmRunArgs& a) {
  const MsmPlanPublic& P = a.plan;
  MsmPlan pl{P.c, P.W, P.L, P.G, P.B};
  const uint64_t n = a.n;
  const uint32_t nw = a.w_end - a.w_begin;
  const uint64_t n_entries = (uint64_t)nw * n;
  REEF_REQUIRE(n_entries < ((uint64_t)1 << 31), REEF_EINVAL, "reef_msm: n * windows exceeds 2^31 entries");
  REEF_REQUIRE((uint64_t)P.L * a.n_bases < ((uint64_t)1 << 31), REEF_EINVAL, "reef_msm: too many precomputed points");
  const uint32_t nb = P.G * P.B;
  cudaStream_t s = c->stream;

  // scratch carve-up
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const uint64_t max_parts1 = n_entries / K_FIRST + nb + 1;
  const uint64_t max_parts2 = max_parts1 / 64 + nb + 1;   // smallest chunk size of a combine pass
  size_t o_keys = take(n_entries * 4), o_vals = take(n_entries * 4), o_sorted = take(n_entries * 4);
  size_t o_cnt = take((size_t)(nb + 2) * 4), o_start = take((size_t)(nb + 2) * 4), o_cursor = take((size_t)(nb + 2) * 4);
  size_t o_pcnt[2] = {take((size_t)(nb + 2) * 4), take((size_t)(nb + 2) * 4)};
  size_t o_poff[2] = {take((size_t)(nb + 2) * 4), take((size_t)(nb + 2) * 4)};
  size_t o_tm = take(64);
  size_t o_parts[2] = {take(max_parts1 * sizeof(XYZZ<C>)), take(max_parts2 * sizeof(XYZZ<C>))};
  size_t o_buckets = take((size_t)nb * sizeof(XYZZ<C>));
  const uint32_t bpt = P.B >= 256u * 32u ? 8u : 1u;
  const uint32_t nblk = cdiv(P.B, 256 * bpt);
  size_t o_bitpart = take((size_t)P.G * P.c * nblk * sizeof(XYZZ<C>));
  size_t o_res = take(sizeof(XYZZ<C>) + sizeof(Affine<C>));
  size_t o_extra = take((size_t)(a.n_extra + 1) * sizeof(XYZZ<C>));
  size_t o_gather = take((size_t)32 * sizeof(XYZZ<C>));   // MB_MAX_WORLD partials: the same scratch size with and without the exchange
  void* base;
  int rc = ctx_scratch(c, off, &base);
  if (rc) return rc;
  char* d = (char*)base;
  uint32_t* keys = (uint32_t*)(d + o_keys);
  uint32_t* vals = (uint32_t*)(d + o_vals);
  uint32_t* sorted = (uint32_t*)(d + o_sorted);
  uint32_t* cnt = (uint32_t*)(d + o_cnt);
  uint32_t* start = (uint32_t*)(d + o_start);
  uint32_t* cursor = (uint32_t*)(d + o_cursor);
  uint32_t* pcnt[2] = {(uint32_t*)(d + o_pcnt[0]), (uint32_t*)(d + o_pcnt[1])};
  uint32_t* poff[2] = {(uint32_t*)(d + o_poff[0]), (uint32_t*)(d + o_poff[1])};
  uint32_t* tm = (uint32_t*)(d + o_tm);
  XYZZ<C>* parts[2] = {(XYZZ<C>*)(d + o_parts[0]), (XYZZ<C>*)(d + o_parts[1])};
  XYZZ<C>* buckets = (XYZZ<C>*)(d + o_buckets);
  XYZZ<C>* bitpart = (XYZZ<C>*)(d + o_bitpart);
  XYZZ<C>* res_xyzz = (XYZZ<C>*)(d + o_res);
  Affine<C>* res_aff = (Affine<C>*)(res_xyzz + 1);
  XYZZ<C>* extra = (XYZZ<C>*)(d + o_extra);

  std::unique_ptr<ProfScope> scope(new ProfScope(c, PROF_MSM_SORT, n_entries));
  REEF_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(nb + 2) * 4, s));
  if (a.scalars_u32) k_digits<true><<<cdiv(n, 256), 256, 0, s>>>(a.d_scalars, n, a.n_bases, pl, a.w_begin, a.w_end, keys, vals);
  else k_digits<false><<<cdiv(n, 256), 256, 0, s>>>(a.d_scalars, n, a.n_bases, pl, a.w_begin, a.w_end, keys, vals);
  REEF_LAUNCHED();
  k_hist<<<cdiv(n_entries, 256), 256, 0, s>>>(keys, n_entries, nb, cnt);
  REEF_LAUNCHED();
  // bucket starts (plain scan), then parts of the first pass
  k_scan<<<1, 1024, 0, s>>>(cnt, nb, 1, nullptr, start, cursor, tm);
  REEF_LAUNCHED();
  k_scatter<<<cdiv(n_entries, 256), 256, 0, s>>>(keys, vals, n_entries, nb, cursor, sorted);
  REEF_LAUNCHED();
  const uint32_t kfirst = n_entries / (2 * K_FIRST) >= (uint64_t)c->sm_count * 2048 ? 2 * K_FIRST : K_FIRST;
  k_scan<<<1, 1024, 0, s>>>(cnt, nb, kfirst, pcnt[0], poff[0], nullptr, tm + 2);
  REEF_LAUNCHED();
  scope.reset();
  uint32_t h_tm[4];
  REEF_CUDA(cudaMemcpyAsync(h_tm, tm, 16, cudaMemcpyDeviceToHost, s));
  REEF_CUDA(cudaStreamSynchronize(s));
  uint32_t n_parts = h_tm[2], max_cnt = h_tm[3];   // parts of pass 1, largest per-bucket part count
  scope.reset(new ProfScope(c, PROF_MSM_ACCUM, n_entries));
  if (n_parts) {
    if (msm_tma_enabled()) {
      static bool attr_set = false;
      if (!attr_set) {
        REEF_CUDA(cudaFuncSetAttribute((const void*)k_accum_first_tma<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MSM_TMA_SMEM));
        attr_set = true;
      }
      k_accum_first_tma<C><<<cdiv(n_parts, 128), 128, MSM_TMA_SMEM, s>>>(sorted, start, cnt, poff[0], nb, n_parts, kfirst,
                                                                        (const Affine<C>*)a.d_levels, parts[0]);
    } else {
      k_accum_first<C><<<cdiv(n_parts, 128), 128, polite_smem(c, (const void*)k_accum_first<C>), s>>>(sorted, start, cnt, poff[0], nb, n_parts, kfirst,
                                                                                                    (const Affine<C>*)a.d_levels, parts[0]);
    }
    REEF_LAUNCHED();
  }
  int cur = 0;
  while (max_cnt > 1) {
    const int nxt = cur ^ 1;
    // enough buckets to fill the chip with 4-lane groups -> throughput shape, else whole-warp groups
    const bool narrow = (uint64_t)nb * 4 >= (uint64_t)c->sm_count * 32 * 2;
    const uint32_t kn = narrow ? 64u : K_NEXT;
    k_scan<<<1, 1024, 0, s>>>(pcnt[cur], nb, kn, pcnt[nxt], poff[nxt], nullptr, tm + 2);
    REEF_LAUNCHED();
    const uint32_t n_next = (n_parts + kn - 1) / kn + nb;   // upper bound; exact count read on device
    if (narrow) k_accum_next<C, 4><<<cdiv((uint64_t)n_next * 4, 128), 128, polite_smem(c, (const void*)k_accum_next<C, 4>), s>>>(parts[cur], poff[cur], pcnt[cur], poff[nxt], nb, n_next, kn, parts[nxt]);
    else k_accum_next<C, 32><<<cdiv((uint64_t)n_next * 32, 128), 128, polite_smem(c, (const void*)k_accum_next<C, 32>), s>>>(parts[cur], poff[cur], pcnt[cur], poff[nxt], nb, n_next, kn, parts[nxt]);
    REEF_LAUNCHED();
    max_cnt = (max_cnt + kn - 1) / kn;
    n_parts = n_next;
    cur = nxt;
  }
  scope.reset();
  scope.reset(new ProfScope(c, PROF_MSM_REDUCE, nb));
  k_gather_buckets<C><<<cdiv(nb, 256), 256, 0, s>>>(parts[cur], poff[cur], pcnt[cur], nb, buckets);
  REEF_LAUNCHED();
  k_bitsum_partial<C><<<dim3(nblk, P.c, P.G), 256, 0, s>>>(buckets, P.B, bpt, bitpart);
  REEF_LAUNCHED();
  if (a.n_extra) REEF_CUDA(cudaMemcpyAsync(extra, a.h_extra_xyzz_mont, (size_t)a.n_extra * sizeof(XYZZ<C>), cudaMemcpyHostToDevice, s));
  // The single result leaves the device as XYZZ and becomes affine on the HOST (one inversion by the shared
  // __host__ __device__ field code): the binary-Euclid inversion is ~110 k cycles of one GPU thread at the very end
  // of a latency chain, a few microseconds on a host core.  (The multi-GPU combine stays on the device.)
  const bool host_affine = a.h_out_affine && !a.p2p_combine;
  const bool want_xyzz = a.h_out_xyzz || a.p2p_combine || host_affine;
  k_bitsum_final<C><<<1, 1024, 0, s>>>(bitpart, nblk, P.c, P.G, P.c * P.L, want_xyzz ? res_xyzz : nullptr, nullptr, extra, a.n_extra);
  REEF_LAUNCHED();
  scope.reset();
  XYZZ<C> h_mont;
  if (host_affine) REEF_CUDA(cudaMemcpyAsync(&h_mont, res_xyzz, sizeof(XYZZ<C>), cudaMemcpyDeviceToHost, s));
  if (a.p2p_combine) {
    // multi-GPU: every rank holds the partial of its own windows; one 128-byte all-gather through the peer
    // mailboxes (a kernel of this library, P2P stores over NVLink) and the combine, all stream-ordered
    XYZZ<C>* gathered = (XYZZ<C>*)(d + o_gather);
    k_xyzz_from_mont<C><<<1, 32, 0, s>>>(res_xyzz);
    REEF_LAUNCHED();
    rc = launch_p2p_allgather(c, res_xyzz, (uint32_t)(sizeof(XYZZ<C>) / 4), gathered);
    if (rc) return rc;
    k_combine<C><<<1, 32, 0, s>>>(gathered, c->mb_world, res_aff);
    REEF_LAUNCHED();
    if (a.h_out_affine) REEF_CUDA(cudaMemcpyAsync(a.h_out_affine, res_aff, sizeof(Affine<C>), cudaMemcpyDeviceToHost, s));
  } else if (a.h_out_xyzz) {
    k_xyzz_from_mont<C><<<1, 32, 0, s>>>(res_xyzz);
    REEF_LAUNCHED();
    REEF_CUDA(cudaMemcpyAsync(a.h_out_xyzz, res_xyzz, sizeof(XYZZ<C>), cudaMemcpyDeviceToHost, s));
  }
  REEF_CUDA(cudaStreamSynchronize(s));
  if (host_affine) {
    Affine<C> aff = xyzz_to_affine<C>(h_mont);
    aff.x = from_mont<C>(aff.x);
    aff.y = from_mont<C>(aff.y);
    memcpy(a.h_out_affine, &aff, sizeof(Affine<C>));
  }
  return REEF_OK;
}

int msm_run(reef_ctx* c, int curve, const MsmRunArgs& a) { return curve == 0 ? msm_run_t<FpCfg>(c, a) : msm_run_t<FqCfg>(c, a); }

template <class C>
static int msm_rows_run_t(reef_ctx* c, const MsmRowsArgs& a) {
  const MsmPlanPublic& P = a.plan;
  MsmPlan pl{P.c, P.W, P.L, 1, P.B};
  REEF_REQUIRE(P.G == 1, REEF_EINVAL, "reef_msm_rows: generators must be fully precomputed");
  const uint64_t n_terms = a.rows * a.cols;
  uint32_t w_used = a.scalar_bits / P.c + 1;
  if (w_used > P.W) w_used = P.W;
  const uint64_t n_entries = n_terms * w_used + (a.d_blinds ? a.rows * P.W : 0);
  REEF_REQUIRE(n_entries < ((uint64_t)1 << 31), REEF_EINVAL, "reef_msm_rows: too many digit entries");
  REEF_REQUIRE(a.rows * P.B < ((uint64_t)1 << 31), REEF_EINVAL, "reef_msm_rows: too many buckets");
  const uint32_t nb = (uint32_t)(a.rows * P.B);
  cudaStream_t s = c->stream;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const uint64_t max_parts1 = n_entries / K_FIRST + nb + 1;
  const uint64_t max_parts2 = max_parts1 / 16 + nb + 1;   // smallest chunk size of a combine pass
  size_t o_keys = take(n_entries * 4), o_vals = take(n_entries * 4), o_sorted = take(n_entries * 4);
  size_t o_cnt = take((size_t)(nb + 2) * 4), o_start = take((size_t)(nb + 2) * 4), o_cursor = take((size_t)(nb + 2) * 4);
  size_t o_pcnt[2] = {take((size_t)(nb + 2) * 4), take((size_t)(nb + 2) * 4)};
  size_t o_poff[2] = {take((size_t)(nb + 2) * 4), take((size_t)(nb + 2) * 4)};
  size_t o_tm = take(64);
  size_t o_parts[2] = {take(max_parts1 * sizeof(XYZZ<C>)), take(max_parts2 * sizeof(XYZZ<C>))};
  size_t o_buckets = take((size_t)nb * sizeof(XYZZ<C>));
  const uint32_t bpt = P.B >= 256u * 32u ? 8u : 1u;
  const uint32_t nblk = cdiv(P.B, 256 * bpt);
  size_t o_bitpart = take((size_t)a.rows * P.c * nblk * sizeof(XYZZ<C>));
  // few rows (the W / T commitments of one fold as two rows): results leave as XYZZ, affine on the host
  const bool host_affine = a.rows <= 16 && !a.d_rows_out;
  size_t o_out = take((size_t)a.rows * (host_affine ? sizeof(XYZZ<C>) : sizeof(Affine<C>)));
  size_t o_rowx = take(host_affine ? 0 : (size_t)a.rows * sizeof(XYZZ<C>));
  void* base;
  int rc = ctx_scratch(c, off, &base);
  if (rc) return rc;
  char* d = (char*)base;
  uint32_t* keys = (uint32_t*)(d + o_keys);
  uint32_t* vals = (uint32_t*)(d + o_vals);
  uint32_t* sorted = (uint32_t*)(d + o_sorted);
  uint32_t* cnt = (uint32_t*)(d + o_cnt);
  uint32_t* start = (uint32_t*)(d + o_start);
  uint32_t* cursor = (uint32_t*)(d + o_cursor);
  uint32_t* pcnt[2] = {(uint32_t*)(d + o_pcnt[0]), (uint32_t*)(d + o_pcnt[1])};
  uint32_t* poff[2] = {(uint32_t*)(d + o_poff[0]), (uint32_t*)(d + o_poff[1])};
  uint32_t* tm = (uint32_t*)(d + o_tm);
  XYZZ<C>* parts[2] = {(XYZZ<C>*)(d + o_parts[0]), (XYZZ<C>*)(d + o_parts[1])};
  XYZZ<C>* buckets = (XYZZ<C>*)(d + o_buckets);
  XYZZ<C>* bitpart = (XYZZ<C>*)(d + o_bitpart);
  Affine<C>* d_out = (Affine<C>*)(d + o_out);

  std::unique_ptr<ProfScope> scope(new ProfScope(c, PROF_MSM_SORT, n_entries));
  REEF_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(nb + 2) * 4, s));
  if (a.scalars_u32) k_digits_rows<true><<<cdiv(n_terms, 256), 256, 0, s>>>(a.d_scalars, a.rows, a.cols, 0, a.n_bases, pl, w_used, 0, keys, vals);
  else k_digits_rows<false><<<cdiv(n_terms, 256), 256, 0, s>>>(a.d_scalars, a.rows, a.cols, 0, a.n_bases, pl, w_used, 0, keys, vals);
  REEF_LAUNCHED();
  if (a.d_blinds) {
    k_digits_rows<false><<<cdiv(a.rows, 256), 256, 0, s>>>(a.d_blinds, a.rows, 1, a.blind_base, a.n_bases, pl, P.W, n_terms * w_used, keys, vals);
    REEF_LAUNCHED();
  }
  k_hist<<<cdiv(n_entries, 256), 256, 0, s>>>(keys, n_entries, nb, cnt);
  REEF_LAUNCHED();
  k_scan<<<1, 1024, 0, s>>>(cnt, nb, 1, nullptr, start, cursor, tm);
  REEF_LAUNCHED();
  k_scatter<<<cdiv(n_entries, 256), 256, 0, s>>>(keys, vals, n_entries, nb, cursor, sorted);
  REEF_LAUNCHED();
  const uint32_t kfirst = n_entries / (2 * K_FIRST) >= (uint64_t)c->sm_count * 2048 ? 2 * K_FIRST : K_FIRST;
  k_scan<<<1, 1024, 0, s>>>(cnt, nb, kfirst, pcnt[0], poff[0], nullptr, tm + 2);
  REEF_LAUNCHED();
  scope.reset();
  uint32_t h_tm[4];
  REEF_CUDA(cudaMemcpyAsync(h_tm, tm, 16, cudaMemcpyDeviceToHost, s));
  REEF_CUDA(cudaStreamSynchronize(s));
  uint32_t n_parts = h_tm[2], max_cnt = h_tm[3];
  scope.reset(new ProfScope(c, PROF_MSM_ACCUM, n_entries));
  if (n_parts) {
    k_accum_first<C><<<cdiv(n_parts, 128), 128, polite_smem(c, (const void*)k_accum_first<C>), s>>>(sorted, start, cnt, poff[0], nb, n_parts, kfirst,
                                                                                                  (const Affine<C>*)a.d_levels, parts[0]);
    REEF_LAUNCHED();
  }
  int cur = 0;
  while (max_cnt > 1) {
    const int nxt = cur ^ 1;
    const bool narrow = (uint64_t)nb * 4 >= (uint64_t)c->sm_count * 32 * 2;
    // many rows: a few heavy buckets per row (the high window of short scalars, the all-equal characters of Reef's
    // `aaaa...b` document) sit between thousands of light ones, so a warp of eight 4-lane groups runs as long as its
    // heaviest group: 16 partials per group (4 + 2 additions deep) instead of 64 (16 + 2) keeps the heavy warps short
    // and spreads a heavy bucket over four times as many groups; the extra pass is one scan + one launch
    const uint32_t kn = narrow ? 16u : K_NEXT;
    k_scan<<<1, 1024, 0, s>>>(pcnt[cur], nb, kn, pcnt[nxt], poff[nxt], nullptr, tm + 2);
    REEF_LAUNCHED();
    const uint32_t n_next = (n_parts + kn - 1) / kn + nb;
    if (narrow) k_accum_next<C, 4><<<cdiv((uint64_t)n_next * 4, 128), 128, polite_smem(c, (const void*)k_accum_next<C, 4>), s>>>(parts[cur], poff[cur], pcnt[cur], poff[nxt], nb, n_next, kn, parts[nxt]);
    else k_accum_next<C, 32><<<cdiv((uint64_t)n_next * 32, 128), 128, polite_smem(c, (const void*)k_accum_next<C, 32>), s>>>(parts[cur], poff[cur], pcnt[cur], poff[nxt], nb, n_next, kn, parts[nxt]);
    REEF_LAUNCHED();
    max_cnt = (max_cnt + kn - 1) / kn;
    n_parts = n_next;
    cur = nxt;
  }
  scope.reset();
  scope.reset(new ProfScope(c, PROF_MSM_REDUCE, nb));
  k_gather_buckets<C><<<cdiv(nb, 256), 256, 0, s>>>(parts[cur], poff[cur], pcnt[cur], nb, buckets);
  REEF_LAUNCHED();
  if (nblk == 1 && P.B <= 2048 && a.rows * P.c >= (uint64_t)c->sm_count * 8) {
    const uint32_t n_pairs = (uint32_t)(a.rows * P.c);                      // many rows: one warp per (row, bit)
    k_bitsum_partial_warp<C><<<cdiv((uint64_t)n_pairs * 32, 256), 256, 0, s>>>(buckets, P.B, P.c, n_pairs, bitpart);
  } else {
    k_bitsum_partial<C><<<dim3(nblk, P.c, (unsigned)a.rows), 256, 0, s>>>(buckets, P.B, bpt, bitpart);
  }
  REEF_LAUNCHED();
  // one warp per bit: c warps are all a row needs (a 512-thread CTA kept half an SM's registers idle)
  const unsigned fin_threads = 32u * (P.c < 16u ? P.c : 16u);
  if (host_affine) {
    k_rows_final<C><<<(unsigned)a.rows, fin_threads, 0, s>>>(bitpart, nblk, P.c, nullptr, (XYZZ<C>*)d_out);
    REEF_LAUNCHED();
  } else {
    XYZZ<C>* row_xyzz = (XYZZ<C>*)(d + o_rowx);
    if (nblk == 1 && P.c <= 32u && a.rows >= (uint64_t)c->sm_count) k_rows_final_warp<C><<<cdiv(a.rows, 4), 128, 0, s>>>(bitpart, a.rows, P.c, row_xyzz);
    else k_rows_final<C><<<(unsigned)a.rows, fin_threads, 0, s>>>(bitpart, nblk, P.c, nullptr, row_xyzz);
    REEF_LAUNCHED();
    k_rows_affine<C><<<cdiv(a.rows, 4), 128, 0, s>>>(row_xyzz, a.rows, d_out);
    REEF_LAUNCHED();
  }
  scope.reset();
  if (host_affine) {
    XYZZ<C> h_rows[16];
    REEF_CUDA(cudaMemcpyAsync(h_rows, d_out, (size_t)a.rows * sizeof(XYZZ<C>), cudaMemcpyDeviceToHost, s));
    REEF_CUDA(cudaStreamSynchronize(s));
    for (uint64_t r = 0; r < a.rows; r++) {
      Affine<C> aff = xyzz_to_affine<C>(h_rows[r]);
      aff.x = from_mont<C>(aff.x);
      aff.y = from_mont<C>(aff.y);
      memcpy(a.h_out + r * sizeof(Affine<C>), &aff, sizeof(Affine<C>));
    }
    return REEF_OK;
  }
  REEF_CUDA(cudaMemcpyAsync(a.h_out, d_out, (size_t)a.rows * sizeof(Affine<C>), cudaMemcpyDeviceToHost, s));
  if (a.d_rows_out) *a.d_rows_out = d_out;
  else REEF_CUDA(cudaStreamSynchronize(s));     // a caller that keeps working on the rows synchronises itself
  return REEF_OK;
}

int msm_rows_run(reef_ctx* c, int curve, const MsmRowsArgs& a) {
  return curve == 0 ? msm_rows_run_t<FpCfg>(c, a) : msm_rows_run_t<FqCfg>(c, a);
}

template <class C>
static int msm_combine_t(reef_ctx* c, const uint8_t* h_pts, uint32_t k, uint8_t* h_out) {
  void* base;
  int rc = ctx_scratch(c, (size_t)k * 128 + 256, &base);
  if (rc) return rc;
  cudaStream_t s = c->stream;
  Affine<C>* d_out = (Affine<C>*)((char*)base + (((size_t)k * 128 + 255) & ~(size_t)255));
  REEF_CUDA(cudaMemcpyAsync(base, h_pts, (size_t)k * 128, cudaMemcpyHostToDevice, s));
  k_combine<C><<<1, 32, 0, s>>>((const XYZZ<C>*)base, k, d_out);
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(h_out, d_out, 64, cudaMemcpyDeviceToHost, s));
  REEF_CUDA(cudaStreamSynchronize(s));
  return REEF_OK;
}

// Forces the lazily loaded MSM kernels into the device (see nl_shard_preload: loading a kernel synchronises
// with running work, which must not happen while an exchange kernel is waiting for a peer).
template <class C>
static int msm_preload_t() {
  cudaFuncAttributes a;
  const void* fns[] = {(const void*)k_accum_first<C>, (const void*)k_accum_next<C, 4>, (const void*)k_accum_next<C, 32>,
                       (const void*)k_gather_buckets<C>, (const void*)k_bitsum_partial<C>, (const void*)k_bitsum_final<C>,
                       (const void*)k_combine<C>, (const void*)k_xyzz_from_mont<C>};
  for (const void* f : fns) REEF_CUDA(cudaFuncGetAttributes(&a, f));
  return REEF_OK;
}
int msm_preload() {
  cudaFuncAttributes a;
  const void* fns[] = {(const void*)k_digits<true>, (const void*)k_digits<false>, (const void*)k_hist, (const void*)k_scan, (const void*)k_scatter};
  for (const void* f : fns) REEF_CUDA(cudaFuncGetAttributes(&a, f));
  int rc = msm_preload_t<FpCfg>();
  return rc ? rc : msm_preload_t<FqCfg>();
}

// Montgomery-form level table (plan L levels) from canonical affine points that are already on the device
// (no curve-membership read-back: the caller's points are products of this library).
int msm_levels_from_dev(reef_ctx* c, int curve, const void* d_canon, uint64_t n, const MsmPlanPublic& pl, void* d_levels, int* d_bad) {
  if (curve == 0) k_precompute<FpCfg><<<cdiv(n, 128), 128, 0, c->stream>>>((const Affine<FpCfg>*)d_canon, n, pl.c, pl.L, (Affine<FpCfg>*)d_levels, d_bad);
  else k_precompute<FqCfg><<<cdiv(n, 128), 128, 0, c->stream>>>((const Affine<FqCfg>*)d_canon, n, pl.c, pl.L, (Affine<FqCfg>*)d_levels, d_bad);
  REEF_LAUNCHED();
  return REEF_OK;
}

int msm_combine(reef_ctx* c, int curve, const uint8_t* h_pts, uint32_t k, uint8_t* h_out) {
  REEF_REQUIRE(k >= 1, REEF_EINVAL, "reef_msm_combine: no partial points");
  size_t need = (size_t)k * 128 + 1024;
  void* base;
  int rc = ctx_scratch(c, need, &base);
  if (rc) return rc;
  return curve == 0 ? msm_combine_t<FpCfg>(c, h_pts, k, h_out) : msm_combine_t<FqCfg>(c, h_pts, k, h_out);
}

}  // namespace reef
