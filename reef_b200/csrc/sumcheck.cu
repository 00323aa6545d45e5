// Spartan-side sum-check sweeps and the sparse R1CS products of row a3.
//
// Replaces, inside nova-snark's CompressedSNARK::prove as reached from
//   /root/reference/src/backend/framework.rs:695-698  (S = spartan::RelaxedR1CSSNARK<G, ipa_pc::EvaluationEngine<G>>, framework.rs:5-8)
//   /root/reference/src/backend/commitment.rs:261-268 (SpartanSNARK::cap_prove)
// the per-round work of `SumcheckProof::prove_quad` (inner sum-check, comb = A*B) and
// `prove_cubic_with_additive_term` (outer sum-check, comb = A*(B*C - D) with A = eq(tau, .),
// B = Az, C = Bz, D = u*Cz + E), plus `bound_poly_var_top`.  nova-snark is NOT under
// /root/reference (git dependency without a pinned revision, Cargo.toml:12): the round polynomial
// convention restated here -- evaluations at 0, 2 (and 3) of the round polynomial, top variable
// bound first, Z[i] <- Z[i] + r (Z[i + n/2] - Z[i]) -- is the published upstream algorithm and is
// PARITY-UNPINNED against the fork Reef builds with (oracle/spartan.py says the same).  The
// transcript stays with the caller: one call per round returns the evaluations, the next call
// takes the challenge.
//
// B200 shape: the tables stay resident; round i+1 binds with r_i and accumulates its evaluations
// in the SAME sweep (each table is read once and written once per round).  Tables that only
// ever appear as the left factor of a product are kept in Montgomery form so that every product
// is one lazily accumulated 256x256 multiply (no reduction per term).
#include <cstring>
#include <vector>

#include "common.cuh"
#include "ec.cuh"
#include "kernels.h"

namespace reef {

static constexpr int SC_WARPS = 4;
static constexpr int SC_MAX_TABLES = 4;

template <class C>
struct ScTables {
  Fe<C>* t[SC_MAX_TABLES];
};

template <class C>
__device__ __forceinline__ Fe<C> sc_fold(const Fe<C>& x0, const Fe<C>& x1, const Fe<C>& r_mont) {
  return fe_add<C>(x0, mont_mul<C>(r_mont, fe_sub<C>(x1, x0)));
}

// warp total of a 17-limb lazy accumulator -> canonical (sum / R) on lane 0
template <class C>
__device__ __forceinline__ Fe<C> sc_warp_total(const Wide17& w) {
  Wide17 tot;
  unsigned long long carry = 0;
#pragma unroll
  for (int i = 0; i < 17; i++) {
    const u32 lo = __reduce_add_sync(0xffffffffu, w.v[i] & 0xffffu);
    const u32 hi = __reduce_add_sync(0xffffffffu, w.v[i] >> 16);
    carry += (unsigned long long)lo + ((unsigned long long)hi << 16);
    tot.v[i] = (u32)carry;
    carry >>= 32;
  }
  return wide_reduce_div_R<C>(tot);
}

// KIND 2: tables A (Montgomery), B (canonical); evaluations of sum A*B at 0 and 2.
// KIND 4: tables A, B (Montgomery), C, D (canonical); evaluations of sum A*(B*C - D) at 0, 2, 3.
// L_in = length of the tables on entry; with FOLD they are first bound to L = L_in/2 with r.
// partials[warp][NE] canonical.
template <class C, int KIND, bool FOLD>
__global__ void __launch_bounds__(SC_WARPS * 32)
k_sc_round(ScTables<C> tabs, uint64_t L_in, Fe<C> r_mont, Fe<C>* __restrict__ partials, uint32_t ppl) {
  constexpr int NE = KIND == 2 ? 2 : 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t L = FOLD ? (L_in >> 1) : L_in;
  const uint64_t half = L >> 1;
  const uint64_t wid = (uint64_t)blockIdx.x * SC_WARPS + warp;
  Wide17 acc[NE];
#pragma unroll
  for (int e = 0; e < NE; e++) wide_zero(acc[e]);
#pragma unroll 1
  for (uint32_t k = 0; k < ppl; k++) {
    const uint64_t i = (wid * ppl + k) * 32 + lane;
    if (i >= half) break;
    Fe<C> lo[KIND], hi[KIND];
#pragma unroll
    for (int t = 0; t < KIND; t++) {
      Fe<C>* T = tabs.t[t];
      if constexpr (FOLD) {
        lo[t] = sc_fold<C>(ld256(T + i), ld256(T + i + L), r_mont);
        hi[t] = sc_fold<C>(ld256(T + i + half), ld256(T + i + half + L), r_mont);
        st256(T + i, lo[t]);
        st256(T + i + half, hi[t]);
      } else {
        lo[t] = ld256(T + i);
        hi[t] = ld256(T + i + half);
      }
    }
    // points 0, 2, 3 of each table's line through (lo, hi): lo, 2hi - lo, 3hi - 2lo
    Fe<C> pt[KIND];
#pragma unroll
    for (int t = 0; t < KIND; t++) pt[t] = lo[t];
#pragma unroll
    for (int e = 0; e < NE; e++) {
      if (e > 0) {
#pragma unroll
        for (int t = 0; t < KIND; t++) {
          const Fe<C> d = fe_sub<C>(hi[t], lo[t]);
          pt[t] = e == 1 ? fe_add<C>(hi[t], d) : fe_add<C>(pt[t], d);
        }
      }
      if constexpr (KIND == 2) {
        wide_mac(acc[e], pt[0].v, pt[1].v);                       // (A R) * B
      } else {
        const Fe<C> bc = mont_mul<C>(pt[1], pt[2]);               // (B R) * C / R = B C
        const Fe<C> w = fe_sub<C>(bc, pt[3]);
        wide_mac(acc[e], pt[0].v, w.v);                           // (A R) * (B C - D)
      }
    }
  }
#pragma unroll
  for (int e = 0; e < NE; e++) {
    const Fe<C> tot = sc_warp_total<C>(acc[e]);
    if (lane == 0) st256(partials + wid * NE + e, tot);
  }
}

template <class C, int NE>
__global__ void __launch_bounds__(128) k_sc_sum(const Fe<C>* __restrict__ partials, uint32_t n_warps, Fe<C>* __restrict__ out) {
  __shared__ Fe<C> red[NE * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Fe<C> acc[NE];
#pragma unroll
  for (int e = 0; e < NE; e++) acc[e] = fe_zero<C>();
  for (uint32_t i = threadIdx.x; i < n_warps; i += 128)
#pragma unroll
    for (int e = 0; e < NE; e++) acc[e] = fe_add<C>(acc[e], ld256(partials + (uint64_t)i * NE + e));
#pragma unroll
  for (int e = 0; e < NE; e++) acc[e] = warp_sum_fe<C>(acc[e]);
  if (lane == 0)
#pragma unroll
    for (int e = 0; e < NE; e++) red[warp * NE + e] = acc[e];
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int e = 0; e < NE; e++) {
      Fe<C> s = red[e];
      for (int w = 1; w < 4; w++) s = fe_add<C>(s, red[w * NE + e]);
      st256(out + e, s);
    }
  }
}

// in-place conversions of a table: canonical -> Montgomery (x R) and back
template <class C>
__global__ void k_sc_to_mont(Fe<C>* t, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) st256(t + i, to_mont<C>(ld256(t + i)));
}

// final bind of length-2 tables -> one value per table, canonical
template <class C>
__global__ void k_sc_final(ScTables<C> tabs, int kind, int n_mont, Fe<C> r_mont, Fe<C>* __restrict__ out) {
  const int t = threadIdx.x;
  if (t >= kind) return;
  Fe<C> v = sc_fold<C>(tabs.t[t][0], tabs.t[t][1], r_mont);
  if (t < n_mont) v = from_mont<C>(v);
  st256(out + t, v);
}

// ---------------------------------------------------------------------------------------
// sparse matrix-vector product over the field (rows of A, B, C times z; framework.rs:668-675
// -> nova-snark's R1CSShape::multiply_vec): CSR, one warp per row, lazily accumulated.
// ---------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128) k_spmv(const uint64_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                                              const Fe<C>* __restrict__ val_mont, const Fe<C>* __restrict__ z,
                                              uint64_t n_rows, Fe<C>* __restrict__ out) {
  const uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;   // warp-uniform
  const uint64_t b = row_ptr[row], e = row_ptr[row + 1];
  Wide17 acc;
  wide_zero(acc);
  for (uint64_t k = b + lane; k < e; k += 32) {
    const Fe<C> v = ld256(val_mont + k);
    const Fe<C> x = ld256(z + col[k]);
    wide_mac(acc, v.v, x.v);
  }
  const Fe<C> tot = sc_warp_total<C>(acc);
  if (lane == 0) st256(out + row, tot);
}

// ---------------------------------------------------------------------------------------
// IPA generator folding (commitment.rs:371-393 -> nova-snark ipa_pc: `ck.fold(&r_inverse, &r)`):
//   out[i] = s_lo * G[i] + s_hi * G[i + n/2],  the same two scalars for every i.
// One thread per output point: joint (Shamir) double-and-add over the 255 scalar bits with the
// three addends G_lo, G_hi, G_lo + G_hi; uniform control flow; Fermat inversion for the final
// affine form (no divergent binary GCD across the warp).
// ---------------------------------------------------------------------------------------
struct Scalar256 {
  uint32_t w[8];
};

template <class C>
__global__ void __launch_bounds__(128) k_ipa_fold(const Affine<C>* __restrict__ in_canon, uint64_t half, Scalar256 s_lo,
                                                  Scalar256 s_hi, Affine<C>* __restrict__ out_canon) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  Affine<C> g0, g1;
  g0.x = to_mont<C>(ld256(&in_canon[i].x));
  g0.y = to_mont<C>(ld256(&in_canon[i].y));
  g1.x = to_mont<C>(ld256(&in_canon[i + half].x));
  g1.y = to_mont<C>(ld256(&in_canon[i + half].y));
  XYZZ<C> both = xyzz_from_affine<C>(g0);
  xyzz_add_affine<C>(both, g1, false);
  XYZZ<C> acc = xyzz_inf<C>();
#pragma unroll 1
  for (int b = 255; b >= 0; b--) {
    acc = xyzz_dbl<C>(acc);
    const uint32_t b0 = (s_lo.w[b >> 5] >> (b & 31)) & 1u, b1 = (s_hi.w[b >> 5] >> (b & 31)) & 1u;
    if (b0 & b1) xyzz_add<C>(acc, both);
    else if (b0) xyzz_add_affine<C>(acc, g0, false);
    else if (b1) xyzz_add_affine<C>(acc, g1, false);
  }
  Affine<C> r;
  if (xyzz_is_inf<C>(acc)) {
    r.x = fe_zero<C>();
    r.y = fe_zero<C>();
  } else {
    r = xyzz_to_affine_with_inv<C>(acc, fe_pow_pm2<C>(acc.zzz));
    r.x = from_mont<C>(r.x);
    r.y = from_mont<C>(r.y);
  }
  st256(&out_canon[i].x, r.x);
  st256(&out_canon[i].y, r.y);
}

// eq(r, x) over x in {0,1}^k, r[0] <-> top index bit (what bound_poly_var_top consumes first); canonical output
template <class C>
__global__ void k_eq_table(const Fe<C>* __restrict__ r_mont, uint32_t k, Fe<C>* __restrict__ out, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fe<C> one = fe_one<C>();
  Fe<C> acc = one;
  for (uint32_t j = 0; j < k; j++) {
    const Fe<C> rj = r_mont[j];
    acc = mont_mul<C>(acc, ((i >> (k - 1 - j)) & 1) ? rj : fe_sub<C>(one, rj));
  }
  st256(out + i, from_mont<C>(acc));
}

// out = a * x + y (canonical in / out)
template <class C>
__global__ void k_axpy(Fe<C> a_mont, const Fe<C>* __restrict__ x, const Fe<C>* __restrict__ y, uint64_t n, Fe<C>* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<C> v = mont_mul<C>(a_mont, ld256(x + i));
  if (y) v = fe_add<C>(v, ld256(y + i));
  st256(out + i, v);
}

// NIFS cross term of two relaxed R1CS instance/witness pairs (nova-snark `R1CSShape::commit_T`, reached from every
// prove_step, framework.rs:668-675):  T = Az1 o Bz2 + Az2 o Bz1 - u1 Cz2 - u2 Cz1   (canonical in / out)
template <class C>
__global__ void k_cross_term(const Fe<C>* __restrict__ abc1, const Fe<C>* __restrict__ abc2, Fe<C> u1_mont, Fe<C> u2_mont, uint64_t n,
                             Fe<C>* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fe<C> az1 = to_mont<C>(ld256(abc1 + i)), az2 = to_mont<C>(ld256(abc2 + i));
  Fe<C> t = fe_add<C>(mont_mul<C>(az1, ld256(abc2 + n + i)), mont_mul<C>(az2, ld256(abc1 + n + i)));
  t = fe_sub<C>(t, mont_mul<C>(u1_mont, ld256(abc2 + 2 * n + i)));
  t = fe_sub<C>(t, mont_mul<C>(u2_mont, ld256(abc1 + 2 * n + i)));
  st256(out + i, t);
}

// ---------------------------------------------------------------------------------------
// Inner-product argument rounds (commitment.rs:371-393 -> nova-snark ipa_pc, [UPSTREAM, unpinned]):
//   c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo>;   a' = a_lo r + a_hi r^-1,  b' = b_lo r^-1 + b_hi r
// ---------------------------------------------------------------------------------------
template <class S>
__global__ void __launch_bounds__(256) k_ipa_dots(const Fe<S>* __restrict__ a, const Fe<S>* __restrict__ b, uint64_t half,
                                                  Fe<S>* __restrict__ out2) {
  __shared__ Fe<S> red[2][8];
  Fe<S> cl = fe_zero<S>(), cr = fe_zero<S>();
  for (uint64_t i = threadIdx.x; i < half; i += 256) {
    const Fe<S> alo = to_mont<S>(ld256(a + i)), ahi = to_mont<S>(ld256(a + half + i));
    cl = fe_add<S>(cl, mont_mul<S>(alo, ld256(b + half + i)));
    cr = fe_add<S>(cr, mont_mul<S>(ahi, ld256(b + i)));
  }
  cl = warp_sum_fe<S>(cl);
  cr = warp_sum_fe<S>(cr);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = cl;
    red[1][threadIdx.x >> 5] = cr;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    Fe<S> t = red[threadIdx.x][0];
    for (int w = 1; w < 8; w++) t = fe_add<S>(t, red[threadIdx.x][w]);
    st256(out2 + threadIdx.x, t);
  }
}

template <class S>
__global__ void k_ipa_fold_scalars(const Fe<S>* __restrict__ a, const Fe<S>* __restrict__ b, uint64_t half, Fe<S> r_mont,
                                   Fe<S> rinv_mont, Fe<S>* __restrict__ a_out, Fe<S>* __restrict__ b_out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  st256(a_out + i, fe_add<S>(mont_mul<S>(r_mont, ld256(a + i)), mont_mul<S>(rinv_mont, ld256(a + half + i))));
  st256(b_out + i, fe_add<S>(mont_mul<S>(rinv_mont, ld256(b + i)), mont_mul<S>(r_mont, ld256(b + half + i))));
}

// ---- IPA over STATIC (registered) generators: the generators are never folded.  After rounds r_0 .. r_(i-1) the
// folded generator G'_j is  sum_{k = j mod m} w[k] G_k  with  w[k] = prod_t (bit_t(k) ? r_t : r_t^-1)  (bit_t = t-th index
// bit from the top), so  L = <a_lo, G'_hi> = sum_k [k mod m >= h] a[(k mod m) - h] w[k] G_k  and likewise R: every round is
// one two-row MSM over the ORIGINAL generators, whose window levels were precomputed once at registration.
template <class S>
__global__ void k_ipa_fill_one(Fe<S>* __restrict__ w, uint64_t n) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) st256(w + k, fe_one<S>());
}
template <class S>
__global__ void k_ipa_weights(Fe<S>* __restrict__ w, uint64_t n, uint32_t shift, Fe<S> r_mont, Fe<S> rinv_mont) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) st256(w + k, mont_mul<S>(ld256(w + k), ((k >> shift) & 1) ? r_mont : rinv_mont));
}
// rows of the round's MSM: out[0][k] = L scalars, out[1][k] = R scalars (canonical), k <= n (entry n = c_L / c_R for gen_c)
template <class S>
__global__ void k_ipa_scalars(const Fe<S>* __restrict__ a, const Fe<S>* __restrict__ w_mont, uint64_t m, uint64_t n,
                              const Fe<S>* __restrict__ c2, Fe<S>* __restrict__ out) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  if (k == n) {
    st256(out + n, ld256(c2));
    st256(out + (n + 1) + n, ld256(c2 + 1));
    return;
  }
  const uint64_t h = m >> 1, j = k & (m - 1);
  const Fe<S> w = ld256(w_mont + k);
  const Fe<S> zero = fe_zero<S>();
  if (j >= h) {
    st256(out + k, mont_mul<S>(w, ld256(a + (j - h))));
    st256(out + (n + 1) + k, zero);
  } else {
    st256(out + k, zero);
    st256(out + (n + 1) + k, mont_mul<S>(w, ld256(a + (j + h))));
  }
}
template <class S>
__global__ void k_ipa_from_mont(const Fe<S>* __restrict__ w_mont, uint64_t n, Fe<S>* __restrict__ out) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) st256(out + k, from_mont<S>(ld256(w_mont + k)));
}

}  // namespace reef

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct reef_ipa {
  reef_ctx* ctx;
  int curve;
  uint64_t n0, n;            // initial / current length
  void* d_buf;
  char *d_a[2], *d_b[2], *d_G[2];   // double-buffered: scalars canonical (32 B), generators affine canonical (64 B)
  int cur;
  char *d_gc, *d_tmpG, *d_tmpS, *d_lv, *d_c;
  int* d_bad;
  // static-generator mode (reef_ipa_begin_bases)
  int static_gens;
  reef::MsmPlanPublic plan;
  char *d_levels_s, *d_w, *d_scal;   // levels[L][n0 + 1] (gen_c appended), weights (Montgomery), 2 x (n0 + 1) scalars
  uint32_t rounds_done, k_bits;
};

struct reef_sumcheck {
  reef_ctx* ctx;
  int field;        // 0 = Fq, 1 = Fp
  int kind;         // 2 or 4
  uint64_t len;     // current table length (before the pending bind)
  uint32_t rounds_done;
  void* d_buf;
  void* tabs[reef::SC_MAX_TABLES];
  void* d_partials;
  void* d_out;
};

namespace reef {

static unsigned sc_cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

template <class C>
static Fe<C> fe_mont_from_le32(const uint8_t* b) {
  Fe<C> x;
  for (int i = 0; i < 8; i++)
    x.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
  return to_mont<C>(x);
}

template <class C>
static int sc_round_t(reef_sumcheck* s, const uint8_t* r_prev, uint8_t* out_evals) {
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  const bool fold = r_prev != nullptr;
  const uint64_t L = fold ? s->len >> 1 : s->len;
  const uint64_t half = L >> 1;
  ScTables<C> tabs;
  for (int t = 0; t < SC_MAX_TABLES; t++) tabs.t[t] = (Fe<C>*)s->tabs[t];
  Fe<C> r = fe_zero<C>();
  if (fold) r = fe_mont_from_le32<C>(r_prev);
  // pairs per lane: keep ~8 warps per SM in flight, at most 16 pairs per lane
  uint32_t ppl = 1;
  while (ppl < 16 && half / (32ull * ppl * 2) >= (uint64_t)c->sm_count * 8) ppl *= 2;
  const uint64_t n_warps_needed = (half + 32ull * ppl - 1) / (32ull * ppl);
  const unsigned nblk = sc_cdiv(n_warps_needed, SC_WARPS);
  const uint32_t n_warps = nblk * SC_WARPS;
  const int ne = s->kind == 2 ? 2 : 3;
  Fe<C>* part = (Fe<C>*)s->d_partials;
  Fe<C>* d_out = (Fe<C>*)s->d_out;
  if (s->kind == 2) {
    if (fold) k_sc_round<C, 2, true><<<nblk, SC_WARPS * 32, 0, st>>>(tabs, s->len, r, part, ppl);
    else k_sc_round<C, 2, false><<<nblk, SC_WARPS * 32, 0, st>>>(tabs, s->len, r, part, ppl);
    REEF_LAUNCHED();
    k_sc_sum<C, 2><<<1, 128, 0, st>>>(part, n_warps, d_out);
  } else {
    if (fold) k_sc_round<C, 4, true><<<nblk, SC_WARPS * 32, 0, st>>>(tabs, s->len, r, part, ppl);
    else k_sc_round<C, 4, false><<<nblk, SC_WARPS * 32, 0, st>>>(tabs, s->len, r, part, ppl);
    REEF_LAUNCHED();
    k_sc_sum<C, 3><<<1, 128, 0, st>>>(part, n_warps, d_out);
  }
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(out_evals, d_out, (size_t)ne * 32, cudaMemcpyDeviceToHost, st));
  REEF_CUDA(cudaStreamSynchronize(st));
  s->len = L;
  s->rounds_done++;
  return REEF_OK;
}

template <class C>
static int sc_final_t(reef_sumcheck* s, const uint8_t* r_last, uint8_t* out_claims) {
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  ScTables<C> tabs;
  for (int t = 0; t < SC_MAX_TABLES; t++) tabs.t[t] = (Fe<C>*)s->tabs[t];
  const int n_mont = s->kind == 2 ? 1 : 2;
  k_sc_final<C><<<1, 32, 0, st>>>(tabs, s->kind, n_mont, fe_mont_from_le32<C>(r_last), (Fe<C>*)s->d_out);
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(out_claims, s->d_out, (size_t)s->kind * 32, cudaMemcpyDeviceToHost, st));
  REEF_CUDA(cudaStreamSynchronize(st));
  s->len = 1;
  return REEF_OK;
}

template <class C>
static int sc_begin_t(reef_sumcheck* s, const uint8_t* const* tables, uint64_t n) {
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  const int n_mont = s->kind == 2 ? 1 : 2;
  for (int t = 0; t < s->kind; t++) {
    REEF_CUDA(cudaMemcpyAsync(s->tabs[t], tables[t], (size_t)n * 32, cudaMemcpyHostToDevice, st));
    if (t < n_mont) {
      k_sc_to_mont<C><<<sc_cdiv(n, 256), 256, 0, st>>>((Fe<C>*)s->tabs[t], n);
      REEF_LAUNCHED();
    }
  }
  REEF_CUDA(cudaStreamSynchronize(st));
  return REEF_OK;
}

template <class C>
static int spmv_t(reef_ctx* c, const uint64_t* row_ptr, const uint32_t* col, const uint8_t* vals, uint64_t n_rows,
                  uint64_t nnz, const uint8_t* z, uint64_t n_cols, uint8_t* out) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  const size_t o_rp = take((n_rows + 1) * 8), o_col = take(nnz * 4 + 4), o_val = take(nnz * 32 + 32), o_z = take(n_cols * 32),
               o_out = take(n_rows * 32);
  void* base;
  int rc = ctx_scratch(c, off, &base);
  if (rc) return rc;
  char* d = (char*)base;
  cudaStream_t st = c->stream;
  REEF_CUDA(cudaMemcpyAsync(d + o_rp, row_ptr, (n_rows + 1) * 8, cudaMemcpyHostToDevice, st));
  if (nnz) {
    REEF_CUDA(cudaMemcpyAsync(d + o_col, col, nnz * 4, cudaMemcpyHostToDevice, st));
    REEF_CUDA(cudaMemcpyAsync(d + o_val, vals, nnz * 32, cudaMemcpyHostToDevice, st));
    k_sc_to_mont<C><<<sc_cdiv(nnz, 256), 256, 0, st>>>((Fe<C>*)(d + o_val), nnz);
    REEF_LAUNCHED();
  }
  REEF_CUDA(cudaMemcpyAsync(d + o_z, z, n_cols * 32, cudaMemcpyHostToDevice, st));
  k_spmv<C><<<sc_cdiv(n_rows * 32, 128), 128, 0, st>>>((const uint64_t*)(d + o_rp), (const uint32_t*)(d + o_col),
                                                       (const Fe<C>*)(d + o_val), (const Fe<C>*)(d + o_z), n_rows,
                                                       (Fe<C>*)(d + o_out));
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(out, d + o_out, n_rows * 32, cudaMemcpyDeviceToHost, st));
  REEF_CUDA(cudaStreamSynchronize(st));
  return REEF_OK;
}

template <class C>
static int ipa_fold_t(reef_ctx* c, const uint8_t* bases, uint64_t n, const uint8_t* s_lo, const uint8_t* s_hi, uint8_t* out) {
  const uint64_t half = n / 2;
  void* base;
  int rc = ctx_scratch(c, (size_t)n * 64 + (size_t)half * 64 + 512, &base);
  if (rc) return rc;
  Affine<C>* d_in = (Affine<C>*)base;
  Affine<C>* d_out = d_in + n;
  Scalar256 a, b;
  for (int i = 0; i < 8; i++) {
    a.w[i] = (uint32_t)s_lo[4 * i] | ((uint32_t)s_lo[4 * i + 1] << 8) | ((uint32_t)s_lo[4 * i + 2] << 16) | ((uint32_t)s_lo[4 * i + 3] << 24);
    b.w[i] = (uint32_t)s_hi[4 * i] | ((uint32_t)s_hi[4 * i + 1] << 8) | ((uint32_t)s_hi[4 * i + 2] << 16) | ((uint32_t)s_hi[4 * i + 3] << 24);
  }
  cudaStream_t st = c->stream;
  REEF_CUDA(cudaMemcpyAsync(d_in, bases, (size_t)n * 64, cudaMemcpyHostToDevice, st));
  k_ipa_fold<C><<<sc_cdiv(half, 128), 128, 0, st>>>(d_in, half, a, b, d_out);
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(out, d_out, (size_t)half * 64, cudaMemcpyDeviceToHost, st));
  REEF_CUDA(cudaStreamSynchronize(st));
  return REEF_OK;
}

// canonical check against the modulus of the chosen field
static bool le32_lt_modulus(const uint8_t* x, int field) {
  uint32_t p[8];
  for (int i = 0; i < 8; i++) p[i] = field == 0 ? modulus_limb<FqCfg>(i) : modulus_limb<FpCfg>(i);
  for (int i = 7; i >= 0; i--) {
    const uint32_t w = (uint32_t)x[4 * i] | ((uint32_t)x[4 * i + 1] << 8) | ((uint32_t)x[4 * i + 2] << 16) | ((uint32_t)x[4 * i + 3] << 24);
    if (w < p[i]) return true;
    if (w > p[i]) return false;
  }
  return false;
}

static int check_canon_field(const uint8_t* x, uint64_t n, int field, const char* what) {
  for (uint64_t i = 0; i < n; i++)
    if (!le32_lt_modulus(x + i * 32, field)) return fail(REEF_EINVAL, std::string(what) + ": element is not a canonical field element");
  return REEF_OK;
}

}  // namespace reef

using namespace reef;

extern "C" {

int reef_sumcheck_begin(reef_ctx* c, int field, int kind, const uint8_t* const* tables, uint64_t n, reef_sumcheck** out) {
  REEF_REQUIRE(c && tables && out, REEF_EINVAL, "reef_sumcheck_begin: NULL argument");
  REEF_REQUIRE(field == 0 || field == 1, REEF_EINVAL, "reef_sumcheck_begin: field must be 0 (Fq) or 1 (Fp)");
  REEF_REQUIRE(kind == 2 || kind == 4, REEF_EINVAL, "reef_sumcheck_begin: kind must be 2 (quadratic) or 4 (cubic with additive term)");
  REEF_REQUIRE(n >= 2 && (n & (n - 1)) == 0, REEF_EASSERT, "reef_sumcheck_begin: table length must be a power of two >= 2");
  for (int t = 0; t < kind; t++) {
    REEF_REQUIRE(tables[t] != nullptr, REEF_EINVAL, "reef_sumcheck_begin: NULL table");
    int rc = check_canon_field(tables[t], n, field, "reef_sumcheck_begin");
    if (rc) return rc;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  reef_sumcheck* s = new reef_sumcheck;
  memset(s, 0, sizeof(*s));
  s->ctx = c;
  s->field = field;
  s->kind = kind;
  s->len = n;
  const size_t tab_bytes = ((size_t)n * 32 + 255) & ~(size_t)255;
  const size_t part_bytes = ((size_t)(n / 64 + SC_WARPS * 4) * 3 * 32 + 255) & ~(size_t)255;
  cudaError_t e = cudaMalloc(&s->d_buf, tab_bytes * kind + part_bytes + 256);
  if (e != cudaSuccess) {
    delete s;
    return fail(REEF_ENOMEM, std::string("reef_sumcheck_begin: ") + cudaGetErrorString(e));
  }
  for (int t = 0; t < kind; t++) s->tabs[t] = (char*)s->d_buf + tab_bytes * t;
  s->d_partials = (char*)s->d_buf + tab_bytes * kind;
  s->d_out = (char*)s->d_partials + part_bytes;
  int rc = field == 0 ? sc_begin_t<FqCfg>(s, tables, n) : sc_begin_t<FpCfg>(s, tables, n);
  if (rc) {
    cudaFree(s->d_buf);
    delete s;
    return rc;
  }
  ctx_retain(c);
  *out = s;
  return REEF_OK;
}

int reef_sumcheck_round(reef_sumcheck* s, const uint8_t* r_prev, uint8_t* out_evals) {
  REEF_REQUIRE(s && out_evals, REEF_EINVAL, "reef_sumcheck_round: NULL argument");
  REEF_REQUIRE((s->rounds_done == 0) == (r_prev == nullptr), REEF_EASSERT,
               "reef_sumcheck_round: the first round takes no challenge, every later round takes the previous one");
  REEF_REQUIRE((r_prev ? s->len >> 1 : s->len) >= 2, REEF_EASSERT, "reef_sumcheck_round: no variable left to sum over");
  if (r_prev) {
    int rc = check_canon_field(r_prev, 1, s->field, "reef_sumcheck_round: challenge");
    if (rc) return rc;
  }
  reef_ctx* c = s->ctx;
  REEF_CTX_LIVE(c, "reef_sumcheck_round");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return s->field == 0 ? sc_round_t<FqCfg>(s, r_prev, out_evals) : sc_round_t<FpCfg>(s, r_prev, out_evals);
}

int reef_sumcheck_final(reef_sumcheck* s, const uint8_t* r_last, uint8_t* out_claims) {
  REEF_REQUIRE(s && r_last && out_claims, REEF_EINVAL, "reef_sumcheck_final: NULL argument");
  REEF_REQUIRE(s->len == 2 && s->rounds_done > 0, REEF_EASSERT, "reef_sumcheck_final: rounds are not finished");
  int rc = check_canon_field(r_last, 1, s->field, "reef_sumcheck_final: challenge");
  if (rc) return rc;
  reef_ctx* c = s->ctx;
  REEF_CTX_LIVE(c, "reef_sumcheck_final");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return s->field == 0 ? sc_final_t<FqCfg>(s, r_last, out_claims) : sc_final_t<FpCfg>(s, r_last, out_claims);
}

void reef_sumcheck_free(reef_sumcheck* s) {
  if (!s) return;
  reef_ctx* c = s->ctx;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(s->d_buf);
  }
  delete s;
  ctx_release(c);
}

int reef_r1cs_spmv(reef_ctx* c, int field, const uint64_t* row_ptr, const uint32_t* col_idx, const uint8_t* vals,
                   uint64_t n_rows, uint64_t n_cols, const uint8_t* z, uint8_t* out) {
  REEF_REQUIRE(c && row_ptr && z && out, REEF_EINVAL, "reef_r1cs_spmv: NULL argument");
  REEF_REQUIRE(field == 0 || field == 1, REEF_EINVAL, "reef_r1cs_spmv: field must be 0 (Fq) or 1 (Fp)");
  REEF_REQUIRE(n_rows >= 1 && n_cols >= 1, REEF_EINVAL, "reef_r1cs_spmv: empty matrix");
  REEF_REQUIRE(row_ptr[0] == 0, REEF_EINVAL, "reef_r1cs_spmv: row_ptr[0] must be 0");
  const uint64_t nnz = row_ptr[n_rows];
  for (uint64_t r = 0; r < n_rows; r++) REEF_REQUIRE(row_ptr[r] <= row_ptr[r + 1], REEF_EINVAL, "reef_r1cs_spmv: row_ptr must be non-decreasing");
  REEF_REQUIRE(nnz == 0 || (col_idx && vals), REEF_EINVAL, "reef_r1cs_spmv: NULL entries");
  for (uint64_t k = 0; k < nnz; k++) REEF_REQUIRE(col_idx[k] < n_cols, REEF_EASSERT, "reef_r1cs_spmv: column index out of bounds");
  int rc = check_canon_field(z, n_cols, field, "reef_r1cs_spmv: z");
  if (rc) return rc;
  if (nnz) {
    rc = check_canon_field(vals, nnz, field, "reef_r1cs_spmv: values");
    if (rc) return rc;
  }
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return field == 0 ? spmv_t<FqCfg>(c, row_ptr, col_idx, vals, n_rows, nnz, z, n_cols, out)
                    : spmv_t<FpCfg>(c, row_ptr, col_idx, vals, n_rows, nnz, z, n_cols, out);
}

int reef_ipa_fold_bases(reef_ctx* c, int curve, const uint8_t* bases, uint64_t n, const uint8_t s_lo[32],
                        const uint8_t s_hi[32], uint8_t* out) {
  REEF_REQUIRE(c && bases && s_lo && s_hi && out, REEF_EINVAL, "reef_ipa_fold_bases: NULL argument");
  REEF_REQUIRE(curve == 0 || curve == 1, REEF_EINVAL, "reef_ipa_fold_bases: curve must be 0 (Pallas) or 1 (Vesta)");
  REEF_REQUIRE(n >= 2 && (n & 1) == 0, REEF_EASSERT, "reef_ipa_fold_bases: the generator vector must have even length");
  // coordinates live in the curve's base field: Pallas -> Fp (field 1), Vesta -> Fq (field 0)
  int rc = check_canon_field(bases, n * 2, curve == 0 ? 1 : 0, "reef_ipa_fold_bases: point coordinate");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return curve == 0 ? ipa_fold_t<FpCfg>(c, bases, n, s_lo, s_hi, out) : ipa_fold_t<FqCfg>(c, bases, n, s_lo, s_hi, out);
}

}  // extern "C"

template <class C>
static int eq_table_t(reef_ctx* c, const uint8_t* r, uint32_t k, uint8_t* out) {
  const uint64_t n = (uint64_t)1 << k;
  void* base;
  int rc = ctx_scratch(c, (size_t)n * 32 + (size_t)(k + 1) * 32 + 256, &base);
  if (rc) return rc;
  Fe<C>* d_out = (Fe<C>*)base;
  Fe<C>* d_r = d_out + n;
  std::vector<Fe<C>> rm(k ? k : 1);
  for (uint32_t j = 0; j < k; j++) rm[j] = fe_mont_from_le32<C>(r + (size_t)j * 32);
  cudaStream_t st = c->stream;
  if (k) REEF_CUDA(cudaMemcpyAsync(d_r, rm.data(), (size_t)k * 32, cudaMemcpyHostToDevice, st));
  k_eq_table<C><<<sc_cdiv(n, 128), 128, 0, st>>>(d_r, k, d_out, n);
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  REEF_CUDA(cudaStreamSynchronize(st));
  return REEF_OK;
}

template <class C>
static int axpy_t(reef_ctx* c, const uint8_t* a, const uint8_t* x, const uint8_t* y, uint64_t n, uint8_t* out) {
  void* base;
  int rc = ctx_scratch(c, (size_t)n * 96 + 256, &base);
  if (rc) return rc;
  Fe<C>* d_x = (Fe<C>*)base;
  Fe<C>* d_y = d_x + n;
  Fe<C>* d_o = d_y + n;
  cudaStream_t st = c->stream;
  REEF_CUDA(cudaMemcpyAsync(d_x, x, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  if (y) REEF_CUDA(cudaMemcpyAsync(d_y, y, (size_t)n * 32, cudaMemcpyHostToDevice, st));
  k_axpy<C><<<sc_cdiv(n, 128), 128, 0, st>>>(fe_mont_from_le32<C>(a), d_x, y ? d_y : nullptr, n, d_o);
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(out, d_o, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  REEF_CUDA(cudaStreamSynchronize(st));
  return REEF_OK;
}

template <class C>
static int cross_term_t(reef_ctx* c, const uint8_t* abc1, const uint8_t* abc2, const uint8_t* u1, const uint8_t* u2, uint64_t n, uint8_t* out) {
  void* base;
  int rc = ctx_scratch(c, (size_t)n * 32 * 7 + 256, &base);
  if (rc) return rc;
  Fe<C>* d_1 = (Fe<C>*)base;
  Fe<C>* d_2 = d_1 + 3 * n;
  Fe<C>* d_o = d_2 + 3 * n;
  cudaStream_t st = c->stream;
  REEF_CUDA(cudaMemcpyAsync(d_1, abc1, (size_t)n * 96, cudaMemcpyHostToDevice, st));
  REEF_CUDA(cudaMemcpyAsync(d_2, abc2, (size_t)n * 96, cudaMemcpyHostToDevice, st));
  k_cross_term<C><<<sc_cdiv(n, 128), 128, 0, st>>>(d_1, d_2, fe_mont_from_le32<C>(u1), fe_mont_from_le32<C>(u2), n, d_o);
  REEF_LAUNCHED();
  REEF_CUDA(cudaMemcpyAsync(out, d_o, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
  REEF_CUDA(cudaStreamSynchronize(st));
  return REEF_OK;
}

extern "C" {

int reef_nova_cross_term(reef_ctx* c, int field, const uint8_t* abc1, const uint8_t* abc2, const uint8_t u1[32], const uint8_t u2[32], uint64_t n,
                         uint8_t* out) {
  REEF_REQUIRE(c && abc1 && abc2 && u1 && u2 && out && n >= 1, REEF_EINVAL, "reef_nova_cross_term: NULL / empty argument");
  REEF_REQUIRE(field == 0 || field == 1, REEF_EINVAL, "reef_nova_cross_term: field must be 0 (Fq) or 1 (Fp)");
  int rc = check_canon_field(abc1, 3 * n, field, "reef_nova_cross_term: Az1|Bz1|Cz1");
  if (!rc) rc = check_canon_field(abc2, 3 * n, field, "reef_nova_cross_term: Az2|Bz2|Cz2");
  if (!rc) rc = check_canon_field(u1, 1, field, "reef_nova_cross_term: u1");
  if (!rc) rc = check_canon_field(u2, 1, field, "reef_nova_cross_term: u2");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return field == 0 ? cross_term_t<FqCfg>(c, abc1, abc2, u1, u2, n, out) : cross_term_t<FpCfg>(c, abc1, abc2, u1, u2, n, out);
}

int reef_eq_table(reef_ctx* c, int field, const uint8_t* r, uint32_t k, uint8_t* out) {
  REEF_REQUIRE(c && (r || k == 0) && out, REEF_EINVAL, "reef_eq_table: NULL argument");
  REEF_REQUIRE(field == 0 || field == 1, REEF_EINVAL, "reef_eq_table: field must be 0 (Fq) or 1 (Fp)");
  REEF_REQUIRE(k <= 30, REEF_EINVAL, "reef_eq_table: k out of range for a host output buffer");
  int rc = check_canon_field(r, k, field, "reef_eq_table: r");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return field == 0 ? eq_table_t<FqCfg>(c, r, k, out) : eq_table_t<FpCfg>(c, r, k, out);
}

int reef_vec_axpy(reef_ctx* c, int field, const uint8_t a[32], const uint8_t* x, const uint8_t* y, uint64_t n, uint8_t* out) {
  REEF_REQUIRE(c && a && x && out && n >= 1, REEF_EINVAL, "reef_vec_axpy: NULL / empty argument");
  REEF_REQUIRE(field == 0 || field == 1, REEF_EINVAL, "reef_vec_axpy: field must be 0 (Fq) or 1 (Fp)");
  int rc = check_canon_field(a, 1, field, "reef_vec_axpy: a");
  if (!rc) rc = check_canon_field(x, n, field, "reef_vec_axpy: x");
  if (!rc && y) rc = check_canon_field(y, n, field, "reef_vec_axpy: y");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return field == 0 ? axpy_t<FqCfg>(c, a, x, y, n, out) : axpy_t<FpCfg>(c, a, x, y, n, out);
}

// ---- IPA session (see include/reef_b200.h)
int reef_ipa_begin(reef_ctx* c, int curve, const uint8_t* gens, const uint8_t gen_c[64], const uint8_t* a, const uint8_t* b, uint64_t n,
                   reef_ipa** out) {
  REEF_REQUIRE(c && gens && gen_c && a && b && out, REEF_EINVAL, "reef_ipa_begin: NULL argument");
  REEF_REQUIRE(curve == 0 || curve == 1, REEF_EINVAL, "reef_ipa_begin: curve must be 0 (Pallas) or 1 (Vesta)");
  REEF_REQUIRE(n >= 1 && (n & (n - 1)) == 0, REEF_EASSERT, "reef_ipa_begin: the vector length must be a power of two");
  const int sfield = curve == 0 ? 0 : 1, cfield = curve == 0 ? 1 : 0;   // Pallas: scalars Fq, coordinates Fp
  int rc = check_canon_field(a, n, sfield, "reef_ipa_begin: a");
  if (!rc) rc = check_canon_field(b, n, sfield, "reef_ipa_begin: b");
  if (!rc) rc = check_canon_field(gens, 2 * n, cfield, "reef_ipa_begin: generator coordinate");
  if (!rc) rc = check_canon_field(gen_c, 2, cfield, "reef_ipa_begin: generator coordinate");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  reef_ipa* s = new reef_ipa;
  memset(s, 0, sizeof(*s));
  s->ctx = c;
  s->curve = curve;
  s->n0 = s->n = n;
  const size_t half = n / 2 + 1;
  const size_t bytes = 2 * (n * 32) * 2 + 2 * (n * 64) + 64 + half * 64 + half * 32 + half * 64 + 64 + 256;
  cudaError_t e = cudaMalloc(&s->d_buf, bytes);
  if (e != cudaSuccess) {
    delete s;
    return fail(REEF_ENOMEM, std::string("reef_ipa_begin: ") + cudaGetErrorString(e));
  }
  char* p = (char*)s->d_buf;
  for (int k = 0; k < 2; k++) { s->d_a[k] = p; p += n * 32; }
  for (int k = 0; k < 2; k++) { s->d_b[k] = p; p += n * 32; }
  for (int k = 0; k < 2; k++) { s->d_G[k] = p; p += n * 64; }
  s->d_gc = p; p += 64;
  s->d_tmpG = p; p += half * 64;
  s->d_tmpS = p; p += half * 32;
  s->d_lv = p; p += half * 64;
  s->d_c = p; p += 64;
  s->d_bad = (int*)p;
  cudaStream_t st = c->stream;
  e = cudaMemcpyAsync(s->d_a[0], a, n * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_b[0], b, n * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_G[0], gens, n * 64, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_gc, gen_c, 64, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_bad, 0, 4, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    cudaFree(s->d_buf);
    delete s;
    return fail(REEF_ECUDA, std::string("reef_ipa_begin: ") + cudaGetErrorString(e));
  }
  ctx_retain(c);
  *out = s;
  return REEF_OK;
}

}  // extern "C"

// static-generator session: levels of the registered generators + the levels of gen_c as one more base
template <class SC, class CC>
static int ipa_begin_bases_t(reef_ctx* c, const reef_bases* gb, const uint8_t* gen_c, const uint8_t* a, const uint8_t* b, uint64_t n,
                             reef_ipa** out) {
  reef_ipa* s = new reef_ipa;
  memset(s, 0, sizeof(*s));
  s->ctx = c;
  s->curve = gb->curve;
  s->n0 = s->n = n;
  s->static_gens = 1;
  s->plan = gb->plan;
  s->k_bits = 0;
  while (((uint64_t)1 << s->k_bits) < n) s->k_bits++;
  const uint32_t L = gb->plan.L;
  const size_t lev_bytes = (size_t)L * (n + 1) * 64;
  const size_t bytes = 4 * n * 32 + lev_bytes + n * 32 + 2 * (n + 1) * 32 + 64 + (size_t)L * 64 + 64 + 256;
  cudaError_t e = cudaMalloc(&s->d_buf, bytes);
  if (e != cudaSuccess) {
    delete s;
    return fail(REEF_ENOMEM, std::string("reef_ipa_begin_bases: ") + cudaGetErrorString(e));
  }
  char* p = (char*)s->d_buf;
  s->d_levels_s = p; p += lev_bytes;
  for (int k = 0; k < 2; k++) { s->d_a[k] = p; p += n * 32; }
  for (int k = 0; k < 2; k++) { s->d_b[k] = p; p += n * 32; }
  s->d_w = p; p += n * 32;
  s->d_scal = p; p += 2 * (n + 1) * 32;
  s->d_gc = p; p += 64;
  s->d_lv = p; p += (size_t)L * 64;      // levels of gen_c alone
  s->d_c = p; p += 64;
  s->d_bad = (int*)p;
  cudaStream_t st = c->stream;
  e = cudaMemcpyAsync(s->d_a[0], a, n * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_b[0], b, n * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->d_gc, gen_c, 64, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_bad, 0, 4, st);
  int rc = REEF_OK;
  if (e == cudaSuccess) rc = msm_levels_from_dev(c, s->curve, s->d_gc, 1, gb->plan, s->d_lv, s->d_bad);
  // level l of the session = [the first n registered points of level l | level l of gen_c]
  if (e == cudaSuccess && !rc)
    e = cudaMemcpy2DAsync(s->d_levels_s, (n + 1) * 64, gb->d_levels, (size_t)gb->n * 64, n * 64, L, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && !rc)
    e = cudaMemcpy2DAsync(s->d_levels_s + n * 64, (n + 1) * 64, s->d_lv, 64, 64, L, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && !rc) {
    k_ipa_fill_one<SC><<<sc_cdiv(n, 128), 128, 0, st>>>((Fe<SC>*)s->d_w, n);
    e = cudaGetLastError();
  }
  int bad = 0;
  if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(&bad, s->d_bad, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess || rc || bad) {
    cudaFree(s->d_buf);
    delete s;
    if (rc) return rc;
    if (bad) return fail(REEF_EINVAL, "reef_ipa_begin_bases: gen_c is not on the curve");
    return fail(REEF_ECUDA, std::string("reef_ipa_begin_bases: ") + cudaGetErrorString(e));
  }
  ctx_retain(c);
  *out = s;
  return REEF_OK;
}

extern "C" int reef_ipa_begin_bases(reef_ctx* c, const reef_bases* gens, const uint8_t gen_c[64], const uint8_t* a, const uint8_t* b, uint64_t n,
                                    reef_ipa** out) {
  REEF_REQUIRE(c && gens && gen_c && a && b && out, REEF_EINVAL, "reef_ipa_begin_bases: NULL argument");
  REEF_REQUIRE(gens->ctx == c, REEF_EINVAL, "reef_ipa_begin_bases: generators belong to another context");
  REEF_REQUIRE(n >= 1 && (n & (n - 1)) == 0, REEF_EASSERT, "reef_ipa_begin_bases: the vector length must be a power of two");
  REEF_REQUIRE(n <= gens->n, REEF_EASSERT, "reef_ipa_begin_bases: not enough generators");
  REEF_REQUIRE(gens->scalar_bits == 255 && gens->plan.G == 1, REEF_EINVAL,
               "reef_ipa_begin_bases: generators must be registered for 255-bit scalars with all window levels precomputed");
  const int sfield = gens->curve == 0 ? 0 : 1, cfield = gens->curve == 0 ? 1 : 0;
  int rc = check_canon_field(a, n, sfield, "reef_ipa_begin_bases: a");
  if (!rc) rc = check_canon_field(b, n, sfield, "reef_ipa_begin_bases: b");
  if (!rc) rc = check_canon_field(gen_c, 2, cfield, "reef_ipa_begin_bases: generator coordinate");
  if (rc) return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CTX_LIVE(c, "reef_ipa_begin_bases");
  REEF_CUDA(cudaSetDevice(c->device));
  return gens->curve == 0 ? ipa_begin_bases_t<FqCfg, FpCfg>(c, gens, gen_c, a, b, n, out) : ipa_begin_bases_t<FpCfg, FqCfg>(c, gens, gen_c, a, b, n, out);
}

// one round over the static generators: ONE two-row MSM (L, R) over n0 + 1 precomputed bases
template <class SC>
static int ipa_round_static_t(reef_ipa* s, uint8_t* out_L, uint8_t* out_R) {
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  const uint64_t m = s->n, h = m / 2, n = s->n0;
  const char *a = s->d_a[s->cur], *b = s->d_b[s->cur];
  k_ipa_dots<SC><<<1, 256, 0, st>>>((const Fe<SC>*)a, (const Fe<SC>*)b, h, (Fe<SC>*)s->d_c);
  REEF_LAUNCHED();
  k_ipa_scalars<SC><<<sc_cdiv(n + 1, 128), 128, 0, st>>>((const Fe<SC>*)a, (const Fe<SC>*)s->d_w, m, n, (const Fe<SC>*)s->d_c, (Fe<SC>*)s->d_scal);
  REEF_LAUNCHED();
  MsmRowsArgs ar;
  ar.plan = s->plan;
  ar.d_levels = s->d_levels_s;
  ar.n_bases = n + 1;
  ar.d_scalars = s->d_scal;
  ar.scalars_u32 = 0;
  ar.scalar_bits = 255;
  ar.rows = 2;
  ar.cols = n + 1;
  ar.d_blinds = nullptr;
  ar.blind_base = n + 1;
  uint8_t lr[128];
  ar.h_out = lr;
  int rc = msm_rows_run(c, s->curve, ar);
  if (rc) return rc;
  memcpy(out_L, lr, 64);
  memcpy(out_R, lr + 64, 64);
  return REEF_OK;
}

template <class SC>
static int ipa_round_t(reef_ipa* s, uint8_t* out_L, uint8_t* out_R) {
  if (s->static_gens) return ipa_round_static_t<SC>(s, out_L, out_R);
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  const uint64_t h = s->n / 2;
  const char *a = s->d_a[s->cur], *b = s->d_b[s->cur], *G = s->d_G[s->cur];
  k_ipa_dots<SC><<<1, 256, 0, st>>>((const Fe<SC>*)a, (const Fe<SC>*)b, h, (Fe<SC>*)s->d_c);
  REEF_LAUNCHED();
  const MsmPlanPublic pl = msm_make_plan(h + 1, 255, 1);     // 1 byte of level budget: no precomputed levels (L = 1)
  for (int side = 0; side < 2; side++) {
    // L = <a_lo, G_hi> + c_L * gen_c ;  R = <a_hi, G_lo> + c_R * gen_c
    const char* g_src = side == 0 ? G + h * 64 : G;
    const char* a_src = side == 0 ? a : a + h * 32;
    REEF_CUDA(cudaMemcpyAsync(s->d_tmpG, g_src, h * 64, cudaMemcpyDeviceToDevice, st));
    REEF_CUDA(cudaMemcpyAsync(s->d_tmpG + h * 64, s->d_gc, 64, cudaMemcpyDeviceToDevice, st));
    REEF_CUDA(cudaMemcpyAsync(s->d_tmpS, a_src, h * 32, cudaMemcpyDeviceToDevice, st));
    REEF_CUDA(cudaMemcpyAsync(s->d_tmpS + h * 32, s->d_c + side * 32, 32, cudaMemcpyDeviceToDevice, st));
    int rc = msm_levels_from_dev(c, s->curve, s->d_tmpG, h + 1, pl, s->d_lv, s->d_bad);
    if (rc) return rc;
    MsmRunArgs ar;
    ar.plan = pl;
    ar.d_levels = s->d_lv;
    ar.n_bases = h + 1;
    ar.d_scalars = s->d_tmpS;
    ar.scalars_u32 = 0;
    ar.n = h + 1;
    ar.w_begin = 0;
    ar.w_end = pl.W;
    ar.h_out_affine = side == 0 ? out_L : out_R;
    ar.h_out_xyzz = nullptr;
    ar.h_extra_xyzz_mont = nullptr;
    ar.n_extra = 0;
    rc = msm_run(c, s->curve, ar);
    if (rc) return rc;
  }
  return REEF_OK;
}

template <class SC, class CC>
static int ipa_fold_session_t(reef_ipa* s, const uint8_t* r, const uint8_t* r_inv);

extern "C" {

int reef_ipa_round(reef_ipa* s, uint8_t out_L[64], uint8_t out_R[64]) {
  REEF_REQUIRE(s && out_L && out_R, REEF_EINVAL, "reef_ipa_round: NULL argument");
  REEF_REQUIRE(s->n >= 2, REEF_EASSERT, "reef_ipa_round: the vectors are already folded to length 1");
  reef_ctx* c = s->ctx;
  REEF_CTX_LIVE(c, "reef_ipa_round");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return s->curve == 0 ? ipa_round_t<FqCfg>(s, out_L, out_R) : ipa_round_t<FpCfg>(s, out_L, out_R);
}

}  // extern "C"

template <class SC, class CC>
static int ipa_fold_session_t(reef_ipa* s, const uint8_t* r, const uint8_t* r_inv) {
  reef_ctx* c = s->ctx;
  cudaStream_t st = c->stream;
  const uint64_t h = s->n / 2;
  const int nx = s->cur ^ 1;
  const Fe<SC> rm = fe_mont_from_le32<SC>(r), rim = fe_mont_from_le32<SC>(r_inv);
  k_ipa_fold_scalars<SC><<<sc_cdiv(h, 128), 128, 0, st>>>((const Fe<SC>*)s->d_a[s->cur], (const Fe<SC>*)s->d_b[s->cur], h, rm, rim,
                                                          (Fe<SC>*)s->d_a[nx], (Fe<SC>*)s->d_b[nx]);
  REEF_LAUNCHED();
  if (s->static_gens) {                          // the generators stay; their fold lives in the weights
    k_ipa_weights<SC><<<sc_cdiv(s->n0, 128), 128, 0, st>>>((Fe<SC>*)s->d_w, s->n0, s->k_bits - 1 - s->rounds_done, rm, rim);
    REEF_LAUNCHED();
    s->rounds_done++;
    s->cur = nx;
    s->n = h;
    return REEF_OK;
  }
  Scalar256 lo, hi;                              // ck' = ck_lo * r^-1 + ck_hi * r
  for (int i = 0; i < 8; i++) {
    lo.w[i] = (uint32_t)r_inv[4 * i] | ((uint32_t)r_inv[4 * i + 1] << 8) | ((uint32_t)r_inv[4 * i + 2] << 16) | ((uint32_t)r_inv[4 * i + 3] << 24);
    hi.w[i] = (uint32_t)r[4 * i] | ((uint32_t)r[4 * i + 1] << 8) | ((uint32_t)r[4 * i + 2] << 16) | ((uint32_t)r[4 * i + 3] << 24);
  }
  k_ipa_fold<CC><<<sc_cdiv(h, 128), 128, 0, st>>>((const Affine<CC>*)s->d_G[s->cur], h, lo, hi, (Affine<CC>*)s->d_G[nx]);
  REEF_LAUNCHED();
  s->cur = nx;
  s->n = h;
  return REEF_OK;
}

extern "C" {

int reef_ipa_fold(reef_ipa* s, const uint8_t r[32], const uint8_t r_inv[32]) {
  REEF_REQUIRE(s && r && r_inv, REEF_EINVAL, "reef_ipa_fold: NULL argument");
  REEF_REQUIRE(s->n >= 2, REEF_EASSERT, "reef_ipa_fold: the vectors are already folded to length 1");
  const int sfield = s->curve == 0 ? 0 : 1;
  int rc = check_canon_field(r, 1, sfield, "reef_ipa_fold: r");
  if (!rc) rc = check_canon_field(r_inv, 1, sfield, "reef_ipa_fold: r_inv");
  if (rc) return rc;
  reef_ctx* c = s->ctx;
  REEF_CTX_LIVE(c, "reef_ipa_fold");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  return s->curve == 0 ? ipa_fold_session_t<FqCfg, FpCfg>(s, r, r_inv) : ipa_fold_session_t<FpCfg, FqCfg>(s, r, r_inv);
}

int reef_ipa_finish(reef_ipa* s, uint8_t out_a[32], uint8_t out_b[32], uint8_t out_g[64]) {
  REEF_REQUIRE(s && out_a && out_b && out_g, REEF_EINVAL, "reef_ipa_finish: NULL argument");
  REEF_REQUIRE(s->n == 1, REEF_EASSERT, "reef_ipa_finish: rounds are not finished");
  reef_ctx* c = s->ctx;
  REEF_CTX_LIVE(c, "reef_ipa_finish");
  std::lock_guard<std::mutex> lk(c->mu);
  REEF_CUDA(cudaSetDevice(c->device));
  REEF_CUDA(cudaMemcpyAsync(out_a, s->d_a[s->cur], 32, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaMemcpyAsync(out_b, s->d_b[s->cur], 32, cudaMemcpyDeviceToHost, c->stream));
  if (s->static_gens) {
    // G_hat = sum_k w[k] G_k: one MSM of the fold weights over the original generators
    const uint64_t n = s->n0;
    if (s->curve == 0) k_ipa_from_mont<FqCfg><<<sc_cdiv(n, 128), 128, 0, c->stream>>>((const Fe<FqCfg>*)s->d_w, n, (Fe<FqCfg>*)s->d_scal);
    else k_ipa_from_mont<FpCfg><<<sc_cdiv(n, 128), 128, 0, c->stream>>>((const Fe<FpCfg>*)s->d_w, n, (Fe<FpCfg>*)s->d_scal);
    REEF_LAUNCHED();
    MsmRunArgs ar;
    ar.plan = s->plan;
    ar.d_levels = s->d_levels_s;
    ar.n_bases = n + 1;
    ar.d_scalars = s->d_scal;
    ar.scalars_u32 = 0;
    ar.n = n;
    ar.w_begin = 0;
    ar.w_end = s->plan.W;
    ar.h_out_affine = out_g;
    ar.h_out_xyzz = nullptr;
    ar.h_extra_xyzz_mont = nullptr;
    ar.n_extra = 0;
    return msm_run(c, s->curve, ar);          // synchronises the stream
  }
  REEF_CUDA(cudaMemcpyAsync(out_g, s->d_G[s->cur], 64, cudaMemcpyDeviceToHost, c->stream));
  REEF_CUDA(cudaStreamSynchronize(c->stream));
  return REEF_OK;
}

void reef_ipa_free(reef_ipa* s) {
  if (!s) return;
  reef_ctx* c = s->ctx;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(s->d_buf);
  }
  delete s;
  ctx_release(c);
}

}  // extern "C"
