// Poseidon (t = 5, R_F = 8, R_P = 56, x^5) over Fq for the Fiat-Shamir transcript of the nlookup
// sum-check (r1cs.rs:2260-2311, r1cs_helper.rs:479-488): ONE permutation as fast as a 256-thread CTA
// can make it.  Same function as poseidon.cuh's permutations (and neptune 8.1.0's); the schedule is
// built for latency:
//
//   * every multiplication of the critical path is LANE-PARALLEL (lpmul.cuh): ~345 cycles instead of
//     620 (sqr29) / 810 (mul29) for one thread;
//   * partial rounds in the "Gamma" form (tools/lp_model.py):  u_r = w_r^5,  w_(r+1) = u_r + c_r,
//         c_r = C0_r + sum_(t<r) Gamma[r][t] u_t,   Gamma[r][t] = sum_i beta[r][i] D[t][i]
//     so that the chain warp (A) does nothing but w -> w^2 -> w^4 -> w^5 (+ c_r) and everything else
//     is a constant times an OLD u_t:
//       warp B    the newest term Gamma[r][r-1] u_(r-1), lane-parallel, one round of slack
//       warps C   two lanes per future round r (even / odd t) accumulate Gamma[r][t] u_t, t <= r-2, with
//                 one-thread mul29; four more lane pairs accumulate the state lanes 1..4 with the dense
//                 4x4 block already applied:  y_j = sum_i post[j][i] s_i(0) + sum_t PD[j][t] u_t
//     the warps talk through shared-memory words that carry their own valid tag in bit 31 (limbs are
//     < 2^31), so there is no barrier and no fence on the chain;
//   * full rounds: five warps, one state element each: three lane-parallel multiplications for x^5,
//     one 5-term lane-parallel multi-product for the MDS row, next round's constants merged into its
//     normalisation; one named barrier per round.
//
// The sponge state lives in shared memory between permutations as 5 x 10 lazy 29-bit limbs (plain
// residues): absorbing a canonical element is a limb-wise addition, no Montgomery conversions.
#pragma once
#include "lpmul.cuh"
#include "poseidon.cuh"

#ifndef REEF_LP_EXPERIMENT
#define REEF_LP_EXPERIMENT 0      // 1, 2: timing experiments of tools/ (wrong results on purpose)
#endif

namespace reef {

struct alignas(16) LpPad {
  u32 w[LP_PAD];
};

// Global-memory tables of the LP permutation (filled by poseidon_lp_tables_host)
struct PoseidonLpTables {
  u32 rcf[8][5][12];     // rc_full[r][i]: plain 29-bit limbs (words 9..11 = 0)
  u32 kp0[12];           // kp[0]
  u32 kp[57][12];        // kp[r], plain limbs
  u32 rc4[4][12];        // rc_full[4][1..4]
  LpPad mds[5][5];       // mds[j][i], plain, padded operand layout
  LpPad gam1[56];        // Gamma[r][r-1]  (entry 0 unused)
  LpPad pd55[4];         // PD[j][55]
  LpPad lam;             // lam_end
  LpPad r256;            // 2^256 mod p: plain value -> Montgomery-256 form
  F29s beta[56][4];      // Montgomery-261 form: mul29(plain, this) = plain product
  F29s post[4][4];
  F29s gam[56][56];      // Gamma[r][t], t <= r-2
  F29s pd[4][56];        // PD[j][t], t <= 54
};

void poseidon_lp_tables_host(PoseidonLpTables* t);

#if defined(__CUDACC__)

static constexpr int LP_PERM_THREADS = 256;
static constexpr u32 LP_TAG = 0x80000000u;

struct alignas(16) LpPermShared {
  u32 S[5][12];          // sponge state: 10 lazy limbs per element
  LpPad pad[5][3];       // operand pads of the five lane-parallel warps
  u32 hbuf[5][12];
  LpPad x5[2][5];        // S-box outputs of a full round (double-buffered), operand layout
  u32 U[4][12];          // u_t (10 limbs), tag = bit 31 = (t >> 2) & 1
  u32 Uend[12];          // u_55 once more, tag = permutation parity (its readers may arrive long before it exists)
  u32 Cs[2][12];         // c_r (10 limbs), tag = (r >> 1) & 1
  u32 S0[4][12];         // s_i(0), 9 limbs (prologue input)
  u32 C0[56][12];        // C0_r, 9 limbs, tag = permutation parity
  u32 Y0[4][12];
  u32 fin[2][56][12];    // [parity of t][r]: accumulated terms t <= r-2, 9 limbs, tag = permutation parity
  u32 yf[2][4][12];
  LpPad mds[5][5];       // copy of PoseidonLpTables::mds (lane-shifted reads every full round)
  u32 prog[4];           // per accumulator warp: number of u_t consumed
  long long dbg[10];      // warp 0's clock at the phase boundaries of the last permutation (test hook only)
};

__device__ __forceinline__ u32 lp_ldv(const u32* p) { return *(const volatile u32*)p; }
__device__ __forceinline__ void lp_stv(u32* p, u32 v) { *(volatile u32*)p = v; }
__device__ __forceinline__ void lp_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// one carry step on lanes 0..9 (limb 9 takes the carry of limb 8 and is not split)
__device__ __forceinline__ u32 lp_norm(u32 v, int lane) {
  return (lane >= 9 ? v : (v & M29)) + lp_up(lane >= 9 ? 0u : (v >> 29), 1, lane);
}

// column `lane` of a * b: a = ten limbs at a10 (16-byte aligned, every lane reads all of them; `mask`
// strips a tag bit), b = operand-layout array
__device__ __forceinline__ u64 lp_cols_a(const u32* a10, u32 mask, const u32* b_pad, int lane, u64 acc = 0) {
  u64 c0 = acc, c1 = 0, c2 = 0;
  const u32* bl = b_pad + LP_OFF + lane;
  const uint4 a03 = *reinterpret_cast<const uint4*>(a10);
  const uint4 a47 = *reinterpret_cast<const uint4*>(a10 + 4);
  const uint2 a89 = *reinterpret_cast<const uint2*>(a10 + 8);
  c0 = lp_mad(a03.x & mask, bl[0], c0);
  c1 = lp_mad(a03.y & mask, bl[-1], c1);
  c2 = lp_mad(a03.z & mask, bl[-2], c2);
  c0 = lp_mad(a03.w & mask, bl[-3], c0);
  c1 = lp_mad(a47.x & mask, bl[-4], c1);
  c2 = lp_mad(a47.y & mask, bl[-5], c2);
  c0 = lp_mad(a47.z & mask, bl[-6], c0);
  c1 = lp_mad(a47.w & mask, bl[-7], c1);
  c2 = lp_mad(a89.x & mask, bl[-8], c2);
  c0 = lp_mad(a89.y & mask, bl[-9], c0);
  return c0 + c1 + c2;
}

// the same with b already in registers: bq[l] = b_(lane - l)
__device__ __forceinline__ u64 lp_cols_reg(const u32* a10, const u32* bq, u32 mask = 0xffffffffu) {
  u64 c0 = 0, c1 = 0, c2 = 0;
  const uint4 a03 = *reinterpret_cast<const uint4*>(a10);
  const uint4 a47 = *reinterpret_cast<const uint4*>(a10 + 4);
  const uint2 a89 = *reinterpret_cast<const uint2*>(a10 + 8);
  c0 = lp_mad(a03.x & mask, bq[0], c0);
  c1 = lp_mad(a03.y & mask, bq[1], c1);
  c2 = lp_mad(a03.z & mask, bq[2], c2);
  c0 = lp_mad(a03.w & mask, bq[3], c0);
  c1 = lp_mad(a47.x & mask, bq[4], c1);
  c2 = lp_mad(a47.y & mask, bq[5], c2);
  c0 = lp_mad(a47.z & mask, bq[6], c0);
  c1 = lp_mad(a47.w & mask, bq[7], c1);
  c2 = lp_mad(a89.x & mask, bq[8], c2);
  c0 = lp_mad(a89.y & mask, bq[9], c0);
  return c0 + c1 + c2;
}

// lane-shifted register copy of an operand-layout constant: q[l] = limb (lane - l)
__device__ __forceinline__ void lp_load_shifted(u32* q /*10*/, const u32* b_pad, int lane) {
  const u32* b = b_pad + LP_OFF + lane;
#pragma unroll
  for (int l = 0; l < 10; l++) q[l] = b[-l];
}

// lane-shifted register copy of the five MDS constants of row i: q[t][l] = limb (lane - l) of mds[i][t].
// Loaded once per permutation; the eight full rounds then read only the broadcast S-box outputs from shared
// memory (the MDS phase starts on all five warps at once, right after their barrier: 65 shared-memory
// loads per lane per warp there were a queue, not a latency).
struct LpMdsRow {
  u32 q[5][10];
};

// lp_fold with two results from one fold: with and without the addend (w_(r+1) = u_r + c_r and u_r)
__device__ __forceinline__ u32 lp_fold2(u64 col, const LpLane& c, u32* h_buf, int lane, u32 addend, u32* out_plain) {
  u32 p0 = (u32)col & M29, p1 = (u32)(col >> 29) & M29, p2 = (u32)(col >> 58);
  const u32 limb = p0 + lp_up(p1, 1, lane) + lp_up(p2, 2, lane);
  __syncwarp();
  if (lane >= 9 && lane < 21) h_buf[lane - 9] = limb;
  __syncwarp();
  const uint4 h03 = *reinterpret_cast<const uint4*>(h_buf);
  const uint4 h47 = *reinterpret_cast<const uint4*>(h_buf + 4);
  const uint4 h8b = *reinterpret_cast<const uint4*>(h_buf + 8);
  u64 a0 = (u64)(lane < 9 ? limb : 0u), a1 = 0, a2 = 0;
  a0 = lp_mad(h03.x, c.K[0], a0);
  a1 = lp_mad(h03.y, c.K[1], a1);
  a2 = lp_mad(h03.z, c.K[2], a2);
  a0 = lp_mad(h03.w, c.K[3], a0);
  a1 = lp_mad(h47.x, c.K[4], a1);
  a2 = lp_mad(h47.y, c.K[5], a2);
  a0 = lp_mad(h47.z, c.K[6], a0);
  a1 = lp_mad(h47.w, c.K[7], a1);
  a2 = lp_mad(h8b.x, c.K[8], a2);
  a0 = lp_mad(h8b.y, c.K[9], a0);
  a1 = lp_mad(h8b.z, c.K[10], a1);
  a2 = lp_mad(h8b.w, c.K[11], a2);
  const u64 colU = a0 + a1 + a2;
  const u64 colW = colU + addend;
  const u32 u0 = (u32)colU & M29, u1 = (u32)(colU >> 29) & M29, u2 = (u32)(colU >> 58);
  const u32 w0 = (u32)colW & M29, w1 = (u32)(colW >> 29) & M29, w2 = (u32)(colW >> 58);
  *out_plain = u0 + lp_up(u1, 1, lane) + lp_up(u2, 2, lane);
  return w0 + lp_up(w1, 1, lane) + lp_up(w2, 2, lane);
}

__device__ __forceinline__ F29 lp_load9(const u32* p, u32 mask) {
  F29 r;
  const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 4);
  r.l[0] = a.x & mask; r.l[1] = a.y & mask; r.l[2] = a.z & mask; r.l[3] = a.w & mask;
  r.l[4] = b.x & mask; r.l[5] = b.y & mask; r.l[6] = b.z & mask; r.l[7] = b.w & mask;
  r.l[8] = p[8] & mask;
  return r;
}

// publish nine limbs (< 2^31) with a tag; one thread
__device__ __forceinline__ void lp_publish9(u32* dst, const F29& v, u32 tag) {
#pragma unroll
  for (int k = 0; k < 9; k++) lp_stv(dst + k, v.l[k] | tag);
}

// spin until the nine tagged words at p carry `tag`, return them untagged (every lane of the warp reads)
__device__ __forceinline__ F29 lp_wait9(const u32* p, u32 tag) {
  while ((lp_ldv(p + 8) & LP_TAG) != tag) __nanosleep(20);
  F29 r;
  bool ok;
  do {
    ok = true;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const u32 v = lp_ldv(p + k);
      ok = ok && ((v & LP_TAG) == tag);
      r.l[k] = v & ~LP_TAG;
    }
  } while (!ok);
  return r;
}

// lanes 0..n-1 of a warp wait for n tagged words (word `lane`), returns the untagged word (0 on other lanes)
__device__ __forceinline__ u32 lp_wait_lane(const u32* p, int n, u32 tag, int lane) {
  u32 v;
  do {
    v = lane < n ? lp_ldv(p + lane) : tag;
  } while (!__all_sync(0xffffffffu, (v & LP_TAG) == tag));
  return lane < n ? (v & ~LP_TAG) : 0u;
}

// ---------------------------------------------------------------------------------------
// one full round for warp i (state element i): x -> sum_t mds[i][t] x_t^5 (+ addend)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ u32 lp_full_round(LpPermShared* sh, const LpMdsRow& M, const LpLane& c, int i, int lane,
                                             u32 x, int buf, const u32* add_row /* [5][12] or nullptr */, bool first_only = false) {
  LpPad* p = sh->pad[i];
  u32* hb = sh->hbuf[i];
  // next round's constants ride on the normalisation (first_only: only state element 0 gets one, kp[0]);
  // they come from global memory: issue the load now, a whole S-box before its use
  const u32 add = (add_row && lane < 9 && (!first_only || i == 0)) ? add_row[(first_only ? 0 : i * 12) + lane] : 0u;
#if defined(REEF_LP_TIMING) && REEF_LP_TIMING == 2
  long long tmark = clock64();
#define LP_TF(k) do { const long long _t = clock64(); if (threadIdx.x == 0) sh->dbg[k] += _t - tmark; tmark = _t; } while (0)
#else
#define LP_TF(k) do { } while (0)
#endif
  lp_store(p[0].w, lane, x);
  __syncwarp();
  const u32 m2 = lp_mul<false>(p[0].w, p[0].w, c, hb, lane);
  lp_store(p[1].w, lane, m2);
  __syncwarp();
  const u32 m4 = lp_mul<false>(p[1].w, p[1].w, c, hb, lane);
  lp_store(p[2].w, lane, m4);
  __syncwarp();
  const u32 x5 = lp_mul<true>(p[2].w, p[0].w, c, hb, lane);
  lp_store(sh->x5[buf][i].w, lane, x5);
  LP_TF(5);
  lp_bar(2, 160);
  LP_TF(6);
  // all fifteen broadcast loads first (one latency), then the fifty products
  uint4 a03[5], a47[5];
  uint2 a89[5];
#pragma unroll
  for (int t = 0; t < 5; t++) {
    const u32* a10 = sh->x5[buf][t].w + LP_OFF;
    a03[t] = *reinterpret_cast<const uint4*>(a10);
    a47[t] = *reinterpret_cast<const uint4*>(a10 + 4);
    a89[t] = *reinterpret_cast<const uint2*>(a10 + 8);
  }
  u64 c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
#pragma unroll
  for (int t = 0; t < 5; t++) {
    c0 = lp_mad(a03[t].x, M.q[t][0], c0);
    c1 = lp_mad(a03[t].y, M.q[t][1], c1);
    c2 = lp_mad(a03[t].z, M.q[t][2], c2);
    c3 = lp_mad(a03[t].w, M.q[t][3], c3);
    c4 = lp_mad(a47[t].x, M.q[t][4], c4);
    c0 = lp_mad(a47[t].y, M.q[t][5], c0);
    c1 = lp_mad(a47[t].z, M.q[t][6], c1);
    c2 = lp_mad(a47[t].w, M.q[t][7], c2);
    c3 = lp_mad(a89[t].x, M.q[t][8], c3);
    c4 = lp_mad(a89[t].y, M.q[t][9], c4);
  }
  const u64 col = (c0 + c1) + (c2 + c3) + c4;
  LP_TF(7);
  const u32 res = lp_fold<false>(col, c, hb, lane, add);
  LP_TF(8);
  return res;
}

// ---------------------------------------------------------------------------------------
// partial rounds
// ---------------------------------------------------------------------------------------
// warp 0: the chain.  w = limb `lane` of w_0 (lanes 0..9); returns w_56.
#if defined(REEF_LP_TIMING) && REEF_LP_TIMING == 1
#define LP_T(k) do { const long long _t = clock64(); if (lane == 0) sh->dbg[k] += _t - tmark; tmark = _t; } while (0)
#else
#define LP_T(k) do { } while (0)
#endif
__device__ __forceinline__ u32 lp_partial_A(LpPermShared* sh, const LpLane& c, int lane, u32 w, u32 ptag) {
  LpPad* p = sh->pad[0];
  u32* hb = sh->hbuf[0];
#if defined(REEF_LP_TIMING) && REEF_LP_TIMING == 1
  long long tmark = clock64();
#endif
#pragma unroll 1
  for (int r = 0; r < 56; r++) {
    lp_store(p[0].w, lane, w);
    __syncwarp();
    const u32 m2 = lp_mul<false>(p[0].w, p[0].w, c, hb, lane);
    lp_store(p[1].w, lane, m2);
    __syncwarp();
    // Both things the end of this round waits for are normally there long before: read them now, off the
    // chain (c_r is published by warp B about a third into the round; the accumulator warps run a round ahead)
    const int wa = (r & 1) * 2;
    u32 pg0 = lp_ldv(&sh->prog[wa]), pg1 = lp_ldv(&sh->prog[wa + 1]);
    const u32 m4 = lp_mul<false>(p[1].w, p[1].w, c, hb, lane);
    lp_store(p[2].w, lane, m4);
    __syncwarp();
    LP_T(5);
    const u32 ctag = ((u32)(r >> 1) & 1u) << 31;
    u32 cw = lane < 10 ? lp_ldv(&sh->Cs[r & 1][lane]) : ctag;
    const u64 col = lp_cols(p[2].w, p[0].w, lane);
#if REEF_LP_EXPERIMENT == 1 || REEF_LP_EXPERIMENT == 2
    cw = 0;
#else
    while (!__all_sync(0xffffffffu, (cw & LP_TAG) == ctag)) cw = lane < 10 ? lp_ldv(&sh->Cs[r & 1][lane]) : ctag;
    cw = lane < 10 ? (cw & ~LP_TAG) : 0u;
#endif
    LP_T(6);
    u32 u;
    w = lp_fold2(col, c, hb, lane, cw, &u);
    LP_T(7);
#if REEF_LP_EXPERIMENT == 0
    if (r >= 4) {
      // slot r & 3 still holds u_(r-4): its readers are the accumulator warps of parity r & 1
      while (pg0 < (u32)(r - 3) || pg1 < (u32)(r - 3)) {
        pg0 = lp_ldv(&sh->prog[wa]);
        pg1 = lp_ldv(&sh->prog[wa + 1]);
      }
    }
#endif
    if (lane < 10) lp_stv(&sh->U[r & 3][lane], u | (((u32)(r >> 2) & 1u) << 31));
    if (r == 55 && lane < 10) lp_stv(&sh->Uend[lane], u | ptag);
    LP_T(8);
  }
  return w;
}

// warp 1: c_r = (accumulated terms t <= r-2) + Gamma[r][r-1] u_(r-1)
__device__ __forceinline__ void lp_partial_B(LpPermShared* sh, const PoseidonLpTables* T, const LpLane& c, int lane, u32 ptag) {
  u32* hb = sh->hbuf[1];
#pragma unroll 1
  for (int r = 0; r < 56; r++) {
    u64 col = 0;
    if (r >= 1) {
      u32 gq[10];
      lp_load_shifted(gq, T->gam1[r].w, lane);      // global memory: in flight while this warp waits for u_(r-1)
      const u32 ut = ((u32)((r - 1) >> 2) & 1u) << 31;
      const u32* us = sh->U[(r - 1) & 3];
      while ((lp_ldv(us + 9) & LP_TAG) != ut) {
      }
      (void)lp_wait_lane(us, 10, ut, lane);
      __syncwarp();
      col = lp_cols_reg(us, gq, ~LP_TAG);
    }
    const u32 fe = lp_wait_lane(sh->fin[0][r], 9, ptag, lane);
    const u32 fo = lp_wait_lane(sh->fin[1][r], 9, ptag, lane);
    const u32 out = lp_fold<false>(col, c, hb, lane, fe + fo);
    if (lane < 10) lp_stv(&sh->Cs[r & 1][lane], out | (((u32)(r >> 1) & 1u) << 31));
  }
}

// warps 2, 3 (even t) and 6, 7 (odd t): accumulator lanes (one-thread arithmetic).  Warps 4 and 5 share their
// schedulers with the chain warp A (0) and with B (1) and stay idle: a computing warp on A's scheduler
// costs the chain 50 % (tools/bench_lp.cu, modes 3-9), polling and computing warps elsewhere cost nothing.
__device__ __forceinline__ void lp_partial_C(LpPermShared* sh, const PoseidonLpTables* T, int warp, int lane, u32 ptag) {
  const int par = warp >> 2;                           // warps 2,3 -> 0; 6,7 -> 1
  const int pslot = par * 2 + (warp & 1);
  const int q = (warp & 1) * 32 + lane;                // 0..55: round r = q;  56..59: state lane j = q - 56
  const bool is_r = q < 56, is_y = q >= 56 && q < 60;
  const int r = is_r ? q : 55, j = is_y ? q - 56 : 0;
  F29 acc = f29_zero();
  if (par == 0) {
    if (is_r) acc = lp_wait9(sh->C0[r], ptag);
    else if (is_y) acc = lp_wait9(sh->Y0[j], ptag);
  }
  // last t of this parity with t <= r - 2 (rounds), t <= 54 (state lanes)
  int t_last = is_r ? r - 2 : 54;
  if ((t_last & 1) != par) t_last -= 1;
  if (is_r && t_last < 0) lp_publish9(sh->fin[par][r], acc, ptag);
#pragma unroll 1
  for (int t = par; t <= 54; t += 2) {
    const bool active = (is_r || is_y) && t <= t_last;
    const F29s* kp = is_y ? &T->pd[j][t] : &T->gam[active ? r : 55][t];
    const F29 k = ld29(kp);
    const u32 ut = ((u32)(t >> 2) & 1u) << 31;
    const u32* us = sh->U[t & 3];
    while ((lp_ldv(us + 9) & LP_TAG) != ut) __nanosleep(40);
    u32 u10[10];
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int i = 0; i < 10; i++) {
        const u32 v = lp_ldv(us + i);
        ok = ok && ((v & LP_TAG) == ut);
        u10[i] = v & ~LP_TAG;
      }
    } while (!ok);
    __syncwarp();
    if (lane == 0) lp_stv(&sh->prog[pslot], (u32)(t + 1));
    if (active) {
      const F29 uf = lp10_to_f29<0>(u10);
      acc = f29_relax(f29_add_lazy(acc, mul29<FqCfg>(uf, k)));
      if (is_r && t == t_last) lp_publish9(sh->fin[par][r], acc, ptag);
    }
  }
  if (is_y) lp_publish9(sh->yf[par][j], acc, ptag);
}

// prologue products (warps 1..7, 224 threads): C0_r = sum_i beta[r][i] s_i(0) + kp[r+1]; warp 7 also
// y_j(0) = sum_i post[j][i] s_i(0) + rc_full[4][j+1]
__device__ __forceinline__ void lp_prologue(LpPermShared* sh, const PoseidonLpTables* T, int warp, int lane, u32 ptag) {
  const int tau = (warp - 1) * 32 + lane;
  const int r = tau >> 2, i = tau & 3;
  const F29 s0 = lp_load9(sh->S0[i], 0xffffffffu);
  F29 t = mul29<FqCfg>(s0, ld29(&T->beta[r][i]));
  t = f29_add_lazy(t, shfl29(t, lane ^ 1));
  t = f29_add_lazy(t, shfl29(t, lane ^ 2));
  if (i == 0) {
    F29 kk;
#pragma unroll
    for (int k = 0; k < 9; k++) kk.l[k] = T->kp[r + 1][k];
    lp_publish9(sh->C0[r], f29_relax(f29_add_lazy(t, kk)), ptag);
  }
  if (warp == 7) {
    const int j = (lane >> 2) & 3;
    F29 y = mul29<FqCfg>(s0, ld29(&T->post[j][i]));
    y = f29_add_lazy(y, shfl29(y, lane ^ 1));
    y = f29_add_lazy(y, shfl29(y, lane ^ 2));
    if (i == 0 && lane < 16) {
      F29 kk;
#pragma unroll
      for (int k = 0; k < 9; k++) kk.l[k] = T->rc4[j][k];
      lp_publish9(sh->Y0[j], f29_relax(f29_add_lazy(y, kk)), ptag);
    }
  }
}

// Once per kernel, by all LP_PERM_THREADS threads, before the first permutation.
__device__ __forceinline__ void lp_perm_init(LpPermShared* sh, const PoseidonLpTables* T) {
  u32* w = reinterpret_cast<u32*>(sh);
  for (int i = threadIdx.x; i < (int)(sizeof(LpPermShared) / 4); i += LP_PERM_THREADS) w[i] = 0;
  __syncthreads();
  {
    const u32* src = &T->mds[0][0].w[0];
    u32* dst = &sh->mds[0][0].w[0];
    for (int i = threadIdx.x; i < 25 * LP_PAD; i += LP_PERM_THREADS) dst[i] = src[i];
  }
  // slots whose first use expects tag 0 start "invalid"
  if (threadIdx.x < 48) sh->U[threadIdx.x / 12][threadIdx.x % 12] = LP_TAG;
  if (threadIdx.x >= 64 && threadIdx.x < 88) sh->Cs[(threadIdx.x - 64) / 12][(threadIdx.x - 64) % 12] = LP_TAG;
  __syncthreads();
}

// One permutation of sh->S by the whole CTA (LP_PERM_THREADS threads, all must call).
// seq: 1, 2, 3, ... number of this permutation within the kernel (the same on every thread).
static __device__ __noinline__ void poseidon_permute_lp(LpPermShared* sh, const PoseidonLpTables* T, u32 seq) {
  // warp index through a broadcast shuffle: the compiler then KNOWS it is warp-uniform and keeps the role
  // branches below convergent (plain SHFL / no WARPSYNC+ENDCOLLECTIVE around every shuffle and __syncwarp)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5) & 7, 0), lane = threadIdx.x & 31;
  const u32 ptag = (seq & 1u) << 31;
  if (threadIdx.x < 4) sh->prog[threadIdx.x] = 0;      // read by warp 0 only, from round 4 on
  if (threadIdx.x == 0) sh->dbg[5] = sh->dbg[6] = sh->dbg[7] = sh->dbg[8] = 0;
  if (warp < 5) {
    const LpLane c = lp_lane_consts<0>(lane);
    const int i = warp;
    LpMdsRow M;
#pragma unroll
    for (int t = 0; t < 5; t++) {
      const u32* b = sh->mds[i][t].w + LP_OFF + lane;
#pragma unroll
      for (int l = 0; l < 10; l++) M.q[t][l] = b[-l];
    }
    // entry: state + first round constants, one carry step
    if (threadIdx.x == 0) sh->dbg[0] = clock64();
    u32 x = lane < 10 ? sh->S[i][lane] : 0u;
    if (lane < 9) x += T->rcf[0][i][lane];
    x = lp_norm(x, lane);
    int buf = 0;
#pragma unroll 1
    for (int r = 0; r < 3; r++, buf ^= 1) x = lp_full_round(sh, M, c, i, lane, x, buf, &T->rcf[r + 1][0][0]);
    // last full round before the partial rounds: only element 0 of the state gets a constant (kp[0])
    x = lp_full_round(sh, M, c, i, lane, x, buf, T->kp0, true);
    buf ^= 1;
    if (threadIdx.x == 0) sh->dbg[1] = clock64();
    if (i == 0) {
      x = lp_partial_A(sh, c, lane, x, ptag);                 // w_56
      if (threadIdx.x == 0) sh->dbg[2] = clock64();
      // x_0 = lam_end * w_56 + rc_full[4][0]
      u32 lq[10];
      lp_load_shifted(lq, T->lam.w, lane);
      const u32 rc40 = lane < 9 ? T->rcf[4][0][lane] : 0u;
      lp_store(sh->pad[0][0].w, lane, x);
      __syncwarp();
      x = lp_fold<false>(lp_cols_reg(sh->pad[0][0].w + LP_OFF, lq), c, sh->hbuf[0], lane, rc40);
    } else {
      // s_i(0) for the prologue (nine limbs), then this warp's role in the partial rounds
      const u32 s9 = lp_fold_b(x, c, lane);
      if (lane < 9) sh->S0[i - 1][lane] = s9;
      lp_bar(3, 224);
      lp_prologue(sh, T, warp, lane, ptag);
#if REEF_LP_EXPERIMENT == 1
      (void)0;                                   // timing experiment: the chain alone (results are wrong)
#elif REEF_LP_EXPERIMENT == 2
      if (i == 2 || i == 3) lp_partial_C(sh, T, warp, lane, ptag);   // accumulators run, B does not
#else
      if (i == 1) lp_partial_B(sh, T, c, lane, ptag);
      else if (i < 4) lp_partial_C(sh, T, warp, lane, ptag);
#endif
      // x_i = y_i + PD[i-1][55] u_55   (rc_full[4][i] is already in y)
      u32 pq[10];
      lp_load_shifted(pq, T->pd55[i - 1].w, lane);
      const u32 ut = ptag;
      const u32* us = sh->Uend;
      while ((lp_ldv(us + 9) & LP_TAG) != ut) __nanosleep(100);
      (void)lp_wait_lane(us, 10, ut, lane);
      __syncwarp();
      const u64 col = lp_cols_reg(us, pq, ~LP_TAG);
#if REEF_LP_EXPERIMENT == 1
      const u32 ye = 0, yo = 0;
#else
      const u32 ye = lp_wait_lane(sh->yf[0][i - 1], 9, ptag, lane);
      const u32 yo = lp_wait_lane(sh->yf[1][i - 1], 9, ptag, lane);
#endif
      x = lp_fold<false>(col, c, sh->hbuf[i], lane, ye + yo);
    }
#pragma unroll 1
    if (threadIdx.x == 0) sh->dbg[3] = clock64();
    for (int r = 4; r < 8; r++, buf ^= 1) x = lp_full_round(sh, M, c, i, lane, x, buf, r < 7 ? &T->rcf[r + 1][0][0] : nullptr);
    if (lane < 10) sh->S[i][lane] = x;
    if (threadIdx.x == 0) sh->dbg[4] = clock64();
  } else {
    lp_bar(3, 224);
    lp_prologue(sh, T, warp, lane, ptag);
#if REEF_LP_EXPERIMENT != 1
    if (warp >= 6) lp_partial_C(sh, T, warp, lane, ptag);
#endif
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// sponge helpers on sh->S (plain residues, lazy limbs)
// ---------------------------------------------------------------------------------------
// limb k (29 bits) of a canonical 256-bit integer given as 8 words
__device__ __forceinline__ u32 lp_limb_of(const u32* w, int k) {
  const int bit = 29 * k, i = bit >> 5, sh = bit & 31;
  u64 v = w[i];
  if (i + 1 < 8) v |= (u64)w[i + 1] << 32;
  v >>= sh;
  return k == 8 ? (u32)v : ((u32)v & M29);
}

// S[pos] += e (canonical).  Threads 0..8 of the CTA; callers separate it from a permutation with a barrier.
__device__ __forceinline__ void lp_absorb(LpPermShared* sh, int pos, const Fq& e) {
  if (threadIdx.x < 9) sh->S[pos][threadIdx.x] += lp_limb_of(e.v, threadIdx.x);
}

// canonical value of state element `pos` (any single thread)
__device__ __forceinline__ Fq lp_squeeze(const LpPermShared* sh, int pos) {
  u32 l[10];
#pragma unroll
  for (int k = 0; k < 10; k++) l[k] = sh->S[pos][k];
  const F29 t = lp10_to_f29<0>(l);
  return lp_to_canonical<FqCfg>(t.l);
}

#endif  // __CUDACC__

}  // namespace reef
