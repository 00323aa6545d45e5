// Poseidon (t = 5, R_F = 8, R_P = 56, x^5) over Fq for sm_100a -- permutation bodies.
//
// Replaces neptune 8.1.0's `Poseidon::hash` as reached from
//   /root/reference/src/backend/merkle_tree.rs:80-114   (new_parent)
//   /root/reference/src/backend/commitment.rs:495-510   (calc_d)
//   /root/reference/src/backend/r1cs.rs:2260-2311, r1cs_helper.rs:479-488 (Fiat-Shamir)
//
// Uses the optimised-but-equivalent schedule derived in tools/gen_poseidon_consts.py
// (lane-0-only partial-round constants, sparse partial-round matrices, one dense 4x4
// block after the partial rounds): ~1000 field multiplications instead of ~1900.
#pragma once
#include "fp.cuh"

namespace reef {

typedef Fe<FqCfg> Fq;

// All tables in Montgomery form.  Filled once per device by poseidon_upload_constants().
struct PoseidonTables {
  Fq rc_full[8][5];
  Fq mds[5][5];
  Fq post[4][4];
  // rescaled partial rounds (tools/gen_poseidon_consts.py):  u_r = w_r^5,
  //   w_{r+1} = u_r + sum_i beta[r][i] s_i(r) + kp[r+1],   s_i(r+1) = s_i(r) + D[r][i] u_r,
  //   after round 55: s_0 = lam_end * w_56.   dshift[r] = D[r-1] (dshift[0] = 0).
  Fq kp[57];
  Fq beta[56][4];
  Fq dshift[57][4];
  Fq lam_end;
};

// x^5
REEF_HD Fq quintic(const Fq& x) {
  Fq x2 = mont_sqr<FqCfg>(x);
  Fq x4 = mont_sqr<FqCfg>(x2);
  return mont_mul<FqCfg>(x4, x);
}

// Montgomery reduction of a sum of <= 8 products (result < 3p before the two subtractions).
REEF_HD Fq reduce_sum8(u32* T /*16*/) {
  Fq r;
  // mont_reduce ends with one conditional subtraction; sums of k products give < (k/4+1) p.
  mont_reduce<FqCfg>(r.v, T);
  cond_sub_p<FqCfg>(r.v);
  return r;
}

// out[j] = sum_i M[j][i] * s[i]  for a dense NxN block, lazily reduced.
template <int N, int LD>
REEF_HD void dense_mul(Fq* out, const Fq* M /*row-major, leading dim LD*/, const Fq* s) {
#pragma unroll 1
  for (int j = 0; j < N; j++) {
    u32 acc[16];
    mul_wide(acc, M[j * LD + 0].v, s[0].v);
#pragma unroll 1
    for (int i = 1; i < N; i++) {
      u32 t[16];
      mul_wide(t, M[j * LD + i].v, s[i].v);
      acc_add<16>(acc, t);
    }
    out[j] = reduce_sum8(acc);
  }
}

// One permutation, state in Montgomery form.  `K` may live in __constant__ or global memory.
REEF_HD void poseidon_permute(Fq* s /*5*/, const PoseidonTables& K) {
  Fq t[5];
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int i = 0; i < 5; i++) t[i] = quintic(fe_add<FqCfg>(s[i], K.rc_full[r][i]));
    dense_mul<5, 5>(s, &K.mds[0][0], t);
  }
  {
    Fq w = fe_add<FqCfg>(s[0], K.kp[0]);
#pragma unroll 1
    for (int r = 0; r < 56; r++) {
      Fq u = quintic(w);
      u32 acc[16];
      mul_wide(acc, K.beta[r][0].v, s[1].v);
#pragma unroll 1
      for (int i = 1; i < 4; i++) {
        u32 t2[16];
        mul_wide(t2, K.beta[r][i].v, s[1 + i].v);
        acc_add<16>(acc, t2);
      }
      w = fe_add<FqCfg>(fe_add<FqCfg>(u, reduce_sum8(acc)), K.kp[r + 1]);
#pragma unroll 1
      for (int i = 0; i < 4; i++) s[1 + i] = fe_add<FqCfg>(s[1 + i], mont_mul<FqCfg>(K.dshift[r + 1][i], u));
    }
    s[0] = mont_mul<FqCfg>(K.lam_end, w);
  }
  {
    Fq u[4];
    dense_mul<4, 4>(u, &K.post[0][0], s + 1);
#pragma unroll
    for (int i = 0; i < 4; i++) s[1 + i] = u[i];
  }
#pragma unroll 1
  for (int r = 4; r < 8; r++) {
#pragma unroll 1
    for (int i = 0; i < 5; i++) t[i] = quintic(fe_add<FqCfg>(s[i], K.rc_full[r][i]));
    dense_mul<5, 5>(s, &K.mds[0][0], t);
  }
}


#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------
// Warp-cooperative permutation for the latency-critical Fiat-Shamir path: lanes 0..4 of a
// warp hold one state element each (Montgomery form); all 32 lanes must call.  Critical
// path: 4 field multiplications per partial round, ~6 per full round (vs ~1000 sequential
// multiplications for one thread doing the whole permutation).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ Fq ldg_fq(const Fq* p) {
  Fq r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

__device__ __forceinline__ Fq shfl_fq(const Fq& x, int src) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, x.v[i], src);
  return r;
}

__device__ __forceinline__ Fq shfl_xor_fq(const Fq& x, int m) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, x.v[i], m);
  return r;
}

__device__ __forceinline__ Fq sel_fq(bool c, const Fq& a, const Fq& b) {
  Fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}

static __device__ __noinline__ void poseidon_full_round_warp5(Fq& s, int r, int li, const PoseidonTables* K) {
  Fq x = quintic(fe_add<FqCfg>(s, ldg_fq(&K->rc_full[r][li])));
  u32 acc[16];
#pragma unroll 1
  for (int i = 0; i < 5; i++) {
    Fq xi = shfl_fq(x, i);
    Fq m = ldg_fq(&K->mds[li][i]);
    if (i == 0) {
      mul_wide(acc, m.v, xi.v);
    } else {
      u32 t[16];
      mul_wide(t, m.v, xi.v);
      acc_add<16>(acc, t);
    }
  }
  s = reduce_sum8(acc);
}

// One rescaled partial round.  lane 0 carries w (the rescaled lane-0 state), lanes 1..4 carry
// s_i; `ub` = u of the previous round on every lane.  Three multiplications deep:
//   step 1  lane 0: w^2          lanes 1..4: s_i += dshift[r][i] * u_{r-1}
//   step 2  lane 0: w^4          lanes 1..4: p_i = beta[r][i] * s_i      (then summed onto lane 0)
//   step 3  lane 0: u = w^4 * w
//   lane 0: w <- u + sum p_i + kp[r+1]
static __device__ __noinline__ void poseidon_partial_round_warp5(Fq& s, Fq& ub, int r, int lane, int li,
                                                                 const PoseidonTables* K) {
  const bool l0 = (li == 0);
  const bool mid = (lane >= 1 && lane <= 4);
  const int ci = li > 0 ? li - 1 : 0;
  const Fq kd = ldg_fq(&K->dshift[r][ci]);
  const Fq kb = ldg_fq(&K->beta[r][ci]);
  const Fq kk = ldg_fq(&K->kp[r + 1]);
  const Fq w = s;
  Fq m1 = mont_mul<FqCfg>(sel_fq(l0, w, kd), sel_fq(l0, w, ub));
  Fq si = fe_add<FqCfg>(s, m1);                        // lanes 1..4: s_i(r)
  Fq m2 = mont_mul<FqCfg>(sel_fq(l0, m1, kb), sel_fq(l0, m1, si));
  // C = p_1 + p_2 + p_3 + p_4 on lane 0
  Fq v = sel_fq(mid, m2, fe_zero<FqCfg>());
  Fq v1 = fe_add<FqCfg>(v, shfl_fq(v, (lane + 1) & 31));
  Fq c = fe_add<FqCfg>(fe_add<FqCfg>(shfl_fq(v1, 1), shfl_fq(v1, 3)), kk);
  Fq u = mont_mul<FqCfg>(m2, w);                       // lane 0: w^5
  ub = shfl_fq(u, 0);
  s = sel_fq(l0, fe_add<FqCfg>(u, c), si);
}

static __device__ __noinline__ void poseidon_post_warp5(Fq& s, int li, const PoseidonTables* K) {
  const int row = li > 0 ? li - 1 : 0;
  u32 acc[16];
#pragma unroll 1
  for (int i = 1; i < 5; i++) {
    Fq si = shfl_fq(s, i);
    Fq m = ldg_fq(&K->post[row][i - 1]);
    if (i == 1) {
      mul_wide(acc, m.v, si.v);
    } else {
      u32 t[16];
      mul_wide(t, m.v, si.v);
      acc_add<16>(acc, t);
    }
  }
  Fq u = reduce_sum8(acc);
  s = sel_fq(li == 0, s, u);
}

// In/out: `s` = state element `lane` for lanes 0..4 (other lanes: don't care).
__device__ __forceinline__ void poseidon_permute_warp5(Fq& s, const PoseidonTables* K) {
  const int lane = threadIdx.x & 31;
  const int li = lane < 5 ? lane : 4;
  if (lane >= 5) s = fe_zero<FqCfg>();
#pragma unroll 1
  for (int r = 0; r < 4; r++) poseidon_full_round_warp5(s, r, li, K);
  {
    Fq ub = fe_zero<FqCfg>();
    if (lane == 0) s = fe_add<FqCfg>(s, ldg_fq(&K->kp[0]));
#pragma unroll 1
    for (int r = 0; r < 56; r++) poseidon_partial_round_warp5(s, ub, r, lane, li, K);
    // closing step: s_i(56) = s_i(55) + dshift[56][i] * u_55 ;  s_0 = lam_end * w_56
    const int ci = li > 0 ? li - 1 : 0;
    Fq m = mont_mul<FqCfg>(sel_fq(li == 0, ldg_fq(&K->lam_end), ldg_fq(&K->dshift[56][ci])), sel_fq(li == 0, s, ub));
    s = li == 0 ? m : fe_add<FqCfg>(s, m);
  }
  poseidon_post_warp5(s, li, K);
#pragma unroll 1
  for (int r = 4; r < 8; r++) poseidon_full_round_warp5(s, r, li, K);
}
#endif  // __CUDACC__

}  // namespace reef
