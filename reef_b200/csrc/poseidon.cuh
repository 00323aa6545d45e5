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
  Fq rc_part[56];
  Fq mds[5][5];
  Fq sp_row[56][5];
  Fq sp_col[56][4];
  Fq post[4][4];
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
#pragma unroll 1
  for (int r = 0; r < 56; r++) {
    Fq z0 = quintic(fe_add<FqCfg>(s[0], K.rc_part[r]));
    // new0 = row . (z0, s1..s4)
    u32 acc[16];
    mul_wide(acc, K.sp_row[r][0].v, z0.v);
#pragma unroll 1
    for (int i = 1; i < 5; i++) {
      u32 w[16];
      mul_wide(w, K.sp_row[r][i].v, s[i].v);
      acc_add<16>(acc, w);
    }
#pragma unroll 1
    for (int i = 1; i < 5; i++) s[i] = fe_add<FqCfg>(s[i], mont_mul<FqCfg>(K.sp_col[r][i - 1], z0));
    s[0] = reduce_sum8(acc);
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

static __device__ __noinline__ void poseidon_partial_round_warp5(Fq& s, int r, int lane, int li,
                                                          const PoseidonTables* K) {
  const bool l0 = (li == 0);
  const bool mid = (lane >= 1 && lane <= 4);
  Fq t = fe_add<FqCfg>(s, ldg_fq(&K->rc_part[r]));
  // step 1: lane 0 squares t; lanes 1..4 compute b_i * s_i (off the critical path)
  Fq opa = sel_fq(l0, t, ldg_fq(&K->sp_row[r][li]));
  Fq opb = sel_fq(l0, t, s);
  Fq m1 = mont_mul<FqCfg>(opa, opb);
  // sum of b_i * s_i over lanes 1..4, made available on every lane of the 8-lane group
  Fq v = sel_fq(mid, m1, fe_zero<FqCfg>());
  v = fe_add<FqCfg>(v, shfl_xor_fq(v, 1));
  v = fe_add<FqCfg>(v, shfl_xor_fq(v, 2));
  v = fe_add<FqCfg>(v, shfl_xor_fq(v, 4));
  // steps 2,3: lane 0 finishes t^5
  Fq t4 = mont_sqr<FqCfg>(m1);
  Fq z0 = mont_mul<FqCfg>(t4, t);
  z0 = shfl_fq(z0, 0);
  // step 4: lane 0: a * z0 ; lanes 1..4: d_i * z0
  Fq c = sel_fq(l0, ldg_fq(&K->sp_row[r][0]), ldg_fq(&K->sp_col[r][li > 0 ? li - 1 : 0]));
  Fq m4 = mont_mul<FqCfg>(c, z0);
  s = fe_add<FqCfg>(m4, sel_fq(l0, v, s));
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
#pragma unroll 1
  for (int r = 0; r < 56; r++) poseidon_partial_round_warp5(s, r, lane, li, K);
  poseidon_post_warp5(s, li, K);
#pragma unroll 1
  for (int r = 4; r < 8; r++) poseidon_full_round_warp5(s, r, li, K);
}
#endif  // __CUDACC__

}  // namespace reef
