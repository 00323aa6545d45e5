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
#include "fp29.cuh"

namespace reef {

typedef Fe<FqCfg> Fq;

// 29-bit-limb copies (Montgomery-261, fp29.cuh) of the tables the warp-cooperative
// Fiat-Shamir permutation uses; 9 limbs padded to 48 bytes for 128-bit loads.
struct alignas(16) F29s {
  u32 l[12];
};
struct Poseidon29Tables {
  F29s rc_full[8][5];
  F29s mds[5][5];
  F29s post[4][4];
  F29s kp[57];
  F29s beta[56][4];
  F29s emat[56][4];     // beta[r][i] * dshift[r][i]
  F29s dshift[57][4];
  F29s lam_end;
  F29s k266;            // 2^266 mod p as a plain integer: Montgomery-256 -> Montgomery-261
};

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
  Poseidon29Tables t29;
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

// ---- 29-bit-limb implementation (fp29.cuh): no carry flags, no conditional subtractions ----
__device__ __forceinline__ F29 ld29(const F29s* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  const uint4 a = __ldg(q), b = __ldg(q + 1);
  F29 r;
  r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
  r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
  r.l[8] = __ldg(&p->l[8]);
  return r;
}

__device__ __forceinline__ F29 shfl29(const F29& x, int src) {
  F29 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = __shfl_sync(0xffffffffu, x.l[i], src);
  return r;
}

__device__ __forceinline__ F29 sel29(bool c, const F29& a, const F29& b) {
  F29 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = c ? a.l[i] : b.l[i];
  return r;
}

// Full round on lanes 0..4 (state) with the MDS product spread over 25 lanes: lane j + 5 i
// computes mds[j][i] * x_i^5 with ONE multiplication, three lazy shuffle-adds bring row j back
// to lane j (instead of five sequential 81-product passes per state lane).
static __device__ __noinline__ void p29_full_round(F29& s, int r, int lane, const Poseidon29Tables* T) {
  const int li = lane < 5 ? lane : 4;
  const int mi = lane < 25 ? lane / 5 : 4, mj = lane < 25 ? lane % 5 : 4;
  const F29 m = ld29(&T->mds[mj][mi]);                         // issued ahead of the S-box
  const F29 x = f29_add_lazy(s, ld29(&T->rc_full[r][li]));     // < 2^30 per limb
  const F29 x2 = sqr29<FqCfg>(x);
  const F29 x4 = sqr29<FqCfg>(x2);
  const F29 x5 = mul29<FqCfg>(x4, x);
  const F29 xi = shfl29(x5, mi);
  const F29 t = mul29<FqCfg>(m, xi);
  const F29 s1 = f29_add_lazy(t, shfl29(t, (lane + 5) & 31));
  const F29 t20 = shfl29(t, (lane + 20) & 31);
  const F29 s2 = f29_add_lazy(s1, shfl29(s1, (lane + 10) & 31));
  s = f29_relax(f29_add_lazy(s2, t20));                        // 5 terms: limbs < 2^31.4, value < 10p
}

// One rescaled partial round, three multiplications deep, nothing else on the critical path.
//   lane 0 carries w_r; lanes 1..4 carry s_i(r-1) (relaxed, not reduced); ub = u_{r-1} everywhere.
//   slot 1  lane 0: w^2         lanes 1..4: beta[r][i] * s_i(r-1)
//   slot 2  lane 0: w^4         lanes 1..4: (beta[r][i] dshift[r][i]) * u_{r-1}
//           p_i = slot1 + slot2 = beta[r][i] * s_i(r)  -> summed onto lane 0 while slot 3 runs
//   slot 3  lane 0: u = w^4 w   lanes 1..4: dshift[r][i] * u_{r-1}   (s_i(r) = s_i(r-1) + that)
//   lane 0: w <- u + sum p_i + kp[r+1]
// The broadcast of u_r is only consumed in slot 2 of the next round.
static __device__ __noinline__ void p29_partial_round(F29& s, F29& ub, int r, int lane, int li,
                                                      const Poseidon29Tables* T) {
  const bool l0 = (li == 0);
  const bool mid = (lane >= 1 && lane <= 4);
  const int ci = li > 0 ? li - 1 : 0;
  const F29 kb = ld29(&T->beta[r][ci]);
  const F29 ke = ld29(&T->emat[r][ci]);
  const F29 kd = ld29(&T->dshift[r][ci]);
  const F29 kk = ld29(&T->kp[r + 1]);
  const F29 w = s;
  const F29 m1 = mul29<FqCfg>(sel29(l0, w, kb), w);
  const F29 m2 = mul29<FqCfg>(sel29(l0, m1, ke), sel29(l0, m1, ub));
  const F29 v = sel29(mid, f29_relax(f29_add_lazy(m1, m2)), f29_zero());
  const F29 v1 = f29_add_lazy(v, shfl29(v, (lane + 1) & 31));
  const F29 c = f29_add_lazy(f29_add_lazy(shfl29(v1, 1), shfl29(v1, 3)), kk);
  const F29 m3 = mul29<FqCfg>(sel29(l0, m2, kd), sel29(l0, w, ub));
  ub = shfl29(m3, 0);
  s = f29_relax(f29_add_lazy(m3, sel29(l0, c, w)));
}

// Dense 4x4 block after the partial rounds on state lanes 1..4, spread over lanes 1..16:
// lane 1 + jj + 4 ii computes post[jj][ii] * s_{ii+1}; two lazy shuffle-adds land row jj on lane 1 + jj.
static __device__ __noinline__ void p29_post(F29& s, int lane, const Poseidon29Tables* T) {
  const int k = (lane >= 1 && lane <= 16) ? lane - 1 : 15;
  const int jj = k & 3, ii = k >> 2;
  const F29 m = ld29(&T->post[jj][ii]);
  const F29 si = shfl29(s, ii + 1);
  const F29 t = mul29<FqCfg>(m, si);
  const F29 s1 = f29_add_lazy(t, shfl29(t, (lane + 4) & 31));
  const F29 s2 = f29_relax(f29_add_lazy(s1, shfl29(s1, (lane + 8) & 31)));
  s = sel29(lane == 0, s, s2);
}

// In/out: `s` = state element `lane` for lanes 0..4 (other lanes: don't care), Montgomery-256.
__device__ __forceinline__ void poseidon_permute_warp5(Fq& s, const PoseidonTables* K) {
  const int lane = threadIdx.x & 31;
  const int li = lane < 5 ? lane : 4;
  const Poseidon29Tables* T = &K->t29;
  {
    // warm L1 with this launch's constants: one 128-byte line per lane per step
    const char* base = reinterpret_cast<const char*>(T);
    for (uint32_t off = lane * 128u; off < (uint32_t)sizeof(Poseidon29Tables); off += 32u * 128u)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(base + off));
  }
  if (lane >= 5) s = fe_zero<FqCfg>();
  F29 x = f29_from_mont256<FqCfg>(s, ld29(&T->k266));
#pragma unroll 1
  for (int r = 0; r < 4; r++) p29_full_round(x, r, lane, T);
  {
    F29 ub = f29_zero();
    if (lane == 0) x = f29_add_lazy(x, ld29(&T->kp[0]));
#pragma unroll 1
    for (int r = 0; r < 56; r++) p29_partial_round(x, ub, r, lane, li, T);
    // closing step: s_i(56) = s_i(55) + dshift[56][i] * u_55 ;  s_0 = lam_end * w_56
    const int ci = li > 0 ? li - 1 : 0;
    const F29 m = mul29<FqCfg>(sel29(li == 0, ld29(&T->lam_end), ld29(&T->dshift[56][ci])), sel29(li == 0, x, ub));
    x = li == 0 ? m : f29_relax(f29_add_lazy(x, m));
  }
  p29_post(x, lane, T);
#pragma unroll 1
  for (int r = 4; r < 8; r++) p29_full_round(x, r, lane, T);
  s = f29_to_mont256<FqCfg>(x);
}

// ---------------------------------------------------------------------------------------
// Two-warp permutation for the transcript kernels (k_nl_begin, k_round, k_tail, sharded
// rounds): warps 0 and 1 of the CTA call it together (all 64 threads).
//   warp 0 ("A") owns the state (lanes 0..4, like poseidon_permute_warp5) and, during the 56
//   partial rounds, runs NOTHING but lane 0's chain  w -> w^2 -> w^4 -> w^5  (two sqr29 + one
//   mul29 per round: no selects, no shuffles, no constant loads);
//   warp 1 ("B") carries s_1..s_4: ONE multiplication per round computes all twelve side
//   products (lanes 0-3: beta*s, 4-7: (beta*D)*u, 8-11: D*u), sums sum_i beta_i s_i + kp and
//   hands it over through shared memory; it receives u_r the same way.
// One named barrier (id 1, 64 threads) per round, double-buffered exchange slots.
// A partial round costs A ~790 instructions instead of ~1020 (tools/bench_perm.cu).
// ---------------------------------------------------------------------------------------
#ifdef REEF_PERM_TIMING
__device__ long long reef_perm_timing[4];
#endif
struct PermPairShared {
  u32 u[2][12];
  u32 c[2][12];
  u32 s[4][12];
};

__device__ __forceinline__ void perm_pair_barrier() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

__device__ __forceinline__ void st29_shared(u32* dst, const F29& x) {
#pragma unroll
  for (int i = 0; i < 9; i++) dst[i] = x.l[i];
}
__device__ __forceinline__ F29 ld29_shared(const u32* src) {
  F29 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = src[i];
  return r;
}

static __device__ __noinline__ void p29_partial_rounds_A(F29& w, PermPairShared* sh, int lane) {
#pragma unroll 1
  for (int r = 0; r < 56; r++) {
    const F29 m1 = sqr29<FqCfg>(w);
    const F29 m2 = sqr29<FqCfg>(m1);
    const F29 u = mul29<FqCfg>(m2, w);
    if (lane == 0) st29_shared(sh->u[r & 1], u);
    perm_pair_barrier();
    const F29 c = ld29_shared(sh->c[r & 1]);   // relaxed by warp B: limbs < 2^29 + 8
    w = f29_add_lazy(u, c);                    // limbs < 2^30 + 2^8: still a valid sqr29 / mul29 operand
  }
}

static __device__ __noinline__ void p29_partial_rounds_B(PermPairShared* sh, int lane, const Poseidon29Tables* T) {
  const int g = lane < 12 ? (lane >> 2) : 2;      // 0: beta * s   1: (beta D) * u   2: D * u
  const int i = lane & 3;
  F29 sp = ld29_shared(sh->s[i]);                 // s_i(r-1), relaxed, not reduced (meaningful on lanes 0..3)
  F29 S = g == 0 ? sp : f29_zero();               // u_{-1} = 0
#pragma unroll 1
  for (int r = 0; r < 56; r++) {
    const F29s* kp = g == 0 ? &T->beta[r][i] : (g == 1 ? &T->emat[r][i] : &T->dshift[r][i]);
    const F29 k1 = ld29(kp);
    const F29 kk = ld29(&T->kp[r + 1]);
    const F29 t = mul29<FqCfg>(k1, S);
    const F29 te = shfl29(t, (lane + 4) & 31);
    const F29 td = shfl29(t, (lane + 8) & 31);
    const F29 p = f29_relax(f29_add_lazy(t, te));            // beta_i * s_i(r)          (lanes 0..3)
    sp = f29_relax(f29_add_lazy(sp, td));                     // s_i(r) = s_i(r-1) + D u  (lanes 0..3)
    F29 v = sel29(lane < 4, p, f29_zero());
    v = f29_add_lazy(v, shfl29(v, lane ^ 1));
    v = f29_add_lazy(v, shfl29(v, lane ^ 2));
    if (lane == 0) st29_shared(sh->c[r & 1], f29_relax(f29_add_lazy(v, kk)));   // 5 terms, relaxed here (off A's path)
    perm_pair_barrier();
    const F29 ub = ld29_shared(sh->u[r & 1]);
    S = g == 0 ? sp : ub;
  }
  // closing: s_i(56) = s_i(55) + dshift[56][i] * u_55 (S holds u_55 on lanes 8..11 only: reload)
  const F29 ub = ld29_shared(sh->u[1]);                       // r = 55 wrote slot 1
  const F29 t = mul29<FqCfg>(ld29(&T->dshift[56][i]), ub);
  sp = f29_relax(f29_add_lazy(sp, t));
  if (lane < 4) st29_shared(sh->s[lane], sp);
}

// warps 0 and 1 of the CTA, all lanes.  `s`: warp 0 lanes 0..4 = state (Montgomery-256), ignored on warp 1.
__device__ __forceinline__ void poseidon_permute_pair(Fq& s, const PoseidonTables* K) {
  __shared__ PermPairShared sh;
  const int lane = threadIdx.x & 31;
  const bool roleA = (threadIdx.x >> 5) == 0;
  const Poseidon29Tables* T = &K->t29;
  {
    const char* base = reinterpret_cast<const char*>(T);
    for (uint32_t off = (threadIdx.x & 63) * 128u; off < (uint32_t)sizeof(Poseidon29Tables); off += 64u * 128u)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(base + off));
  }
  if (roleA) {
    const int li = lane < 5 ? lane : 4;
    if (lane >= 5) s = fe_zero<FqCfg>();
    F29 x = f29_from_mont256<FqCfg>(s, ld29(&T->k266));
#pragma unroll 1
    for (int r = 0; r < 4; r++) p29_full_round(x, r, lane, T);
    if (lane >= 1 && lane <= 4) st29_shared(sh.s[lane - 1], x);
    F29 w = f29_add_lazy(x, ld29(&T->kp[0]));
    perm_pair_barrier();
#ifdef REEF_PERM_TIMING
    const long long tp0 = clock64();
#endif
    p29_partial_rounds_A(w, &sh, lane);
#ifdef REEF_PERM_TIMING
    if (lane == 0) reef_perm_timing[0] = clock64() - tp0;
#endif
    const F29 m = mul29<FqCfg>(ld29(&T->lam_end), w);
    perm_pair_barrier();
    x = lane == 0 ? m : ld29_shared(sh.s[(li > 0 ? li : 1) - 1]);
    p29_post(x, lane, T);
#pragma unroll 1
    for (int r = 4; r < 8; r++) p29_full_round(x, r, lane, T);
    s = f29_to_mont256<FqCfg>(x);
  } else {
    perm_pair_barrier();
    p29_partial_rounds_B(&sh, lane, T);
    perm_pair_barrier();
  }
}
#endif  // __CUDACC__

}  // namespace reef
