// Lane-parallel multiplication in the Pasta fields for the latency-critical Fiat-Shamir chain.
//
// One field multiplication spread over the lanes of a warp: lane k owns limb k of every value
// (29-bit limbs, PLAIN residue mod p, lazy) and column k of every product.  A lane issues
// 10 IMAD.WIDE for the product and 12 for the fold -- ~70 instructions in total instead of the
// ~260 (mul29) / ~300 (mont_mul) a single thread needs, which is what bounds a DEPENDENT chain of
// multiplications (r1cs.rs:2260-2311, r1cs_helper.rs:479-488: one Poseidon permutation per
// sum-check round, ~200 dependent multiplications each).
//
// Values inside a chain have TEN limbs: limbs 0..8 < 2^30 + 2^7 and limb 9 < 2^28 (the part above
// 2^261), value < 2^289.  Reduction uses the shape of both Pasta primes, p = 2^254 + c, c < 2^126:
//   fold A:  2^(261+29j) == K_j (mod p), K_j < p: the twelve high limbs H_j of a product are multiplied
//            by per-lane constants K_j[k] and added to the low columns  -> value < 2^289 again.
//            This is the only reduction a chain needs: the product of two such values has 21 limbs.
//   fold B:  2^261 == -c' (mod p), c' = 2^7 c < 2^133: folds limb 9 into limbs 0..8 (value < 2^262);
//            used where a value leaves the chain.  The offsets Z (== 64 p) keep every column >= 0.
// Carries move between lanes with shuffles (two 29-bit pieces + the 6-bit top of a 64-bit column).
// tools/lp_model.py is the exact lane-by-lane model of this file (bounds asserted, results checked
// against Python integers) and the generator of lp_consts.inc.
#pragma once
#include <cstdint>

#include "fp29.cuh"

namespace reef {

#include "lp_consts.inc"

// Operand storage: 48 words, limbs at [12..21], zeros elsewhere, so that lane k can read b_(k-i) as
// pad[12 + k - i] for every lane and every i < 10 without a bounds test.
static constexpr int LP_PAD = 48;
static constexpr int LP_OFF = 12;
static constexpr int LP_PAD_WORDS = LP_PAD, LP_OFF_WORDS = LP_OFF;

#if defined(__CUDACC__)

struct LpLane {      // per-lane constants (registers)
  u32 K[12];         // K_j[lane]   (0 for lanes >= 9)
  u32 cp;            // c'_lane     (0 for lanes >= 5)
  u64 Z;             // Z_lane      (0 for lanes >= 9)
};

template <int FIELD>   // 0 = Fq, 1 = Fp
__device__ __forceinline__ LpLane lp_lane_consts(int lane) {
  LpLane c;
  const int k = lane < 9 ? lane : 0;
#pragma unroll
  for (int j = 0; j < 12; j++) c.K[j] = lane < 9 ? (FIELD == 0 ? LP_K_FQ[j][k] : LP_K_FP[j][k]) : 0u;
  c.cp = lane < 5 ? (FIELD == 0 ? LP_CP_FQ[lane < 5 ? lane : 0] : LP_CP_FP[lane < 5 ? lane : 0]) : 0u;
  c.Z = lane < 9 ? (FIELD == 0 ? LP_Z_FQ[k] : LP_Z_FP[k]) : 0ull;
  return c;
}

// one IMAD.WIDE.U32: the compiler otherwise widens loop-invariant operands to 64 bits (3 instructions
// per product) and chains independent accumulators
__device__ __forceinline__ u64 lp_mad(u32 a, u32 b, u64 c) {
  u64 r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
  return r;
}

__device__ __forceinline__ u32 lp_up(u32 v, int d, int lane) {
  const u32 t = __shfl_up_sync(0xffffffffu, v, d);
  return lane >= d ? t : 0u;
}

// column `lane` of a * b (+ acc): a read as broadcast (a_pad[LP_OFF + i]), b lane-shifted
// (b_pad[LP_OFF + lane - i]); both arrays hold 10 limbs at LP_OFF and zeros around them
__device__ __forceinline__ u64 lp_cols(const u32* a_pad, const u32* b_pad, int lane, u64 acc = 0) {
  u64 c0 = acc, c1 = 0, c2 = 0;
  const u32* bl = b_pad + LP_OFF + lane;
  const uint4 a03 = *reinterpret_cast<const uint4*>(a_pad + LP_OFF);
  const uint4 a47 = *reinterpret_cast<const uint4*>(a_pad + LP_OFF + 4);
  const uint2 a89 = *reinterpret_cast<const uint2*>(a_pad + LP_OFF + 8);
  c0 = lp_mad(a03.x, bl[0], c0);
  c1 = lp_mad(a03.y, bl[-1], c1);
  c2 = lp_mad(a03.z, bl[-2], c2);
  c0 = lp_mad(a03.w, bl[-3], c0);
  c1 = lp_mad(a47.x, bl[-4], c1);
  c2 = lp_mad(a47.y, bl[-5], c2);
  c0 = lp_mad(a47.z, bl[-6], c0);
  c1 = lp_mad(a47.w, bl[-7], c1);
  c2 = lp_mad(a89.x, bl[-8], c2);
  c0 = lp_mad(a89.y, bl[-9], c0);
  return c0 + c1 + c2;
}

// Fold A: the 19 product columns held by lanes 0..18 (col = 0 on the other lanes) -> ten lazy limbs
// on lanes 0..9 (other lanes: unspecified).
//   h_buf: 12 words of shared memory private to this warp (16-byte aligned)
//   addend: per-lane addend merged into the normalisation (any u32 on lanes 0..8, 0 elsewhere)
template <bool TIGHT>
__device__ __forceinline__ u32 lp_fold(u64 col, const LpLane& c, u32* h_buf, int lane, u32 addend) {
  // N1: limbs 0..20
  u32 p0 = (u32)col & M29, p1 = (u32)(col >> 29) & M29, p2 = (u32)(col >> 58);
  const u32 limb = p0 + lp_up(p1, 1, lane) + lp_up(p2, 2, lane);
  // fold A
  __syncwarp();
  if (lane >= 9 && lane < 21) h_buf[lane - 9] = limb;
  __syncwarp();
  const uint4 h03 = *reinterpret_cast<const uint4*>(h_buf);
  const uint4 h47 = *reinterpret_cast<const uint4*>(h_buf + 4);
  const uint4 h8b = *reinterpret_cast<const uint4*>(h_buf + 8);
  u64 a0 = (u64)(lane < 9 ? limb : 0u) + addend, a1 = 0, a2 = 0;
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
  const u64 colA = a0 + a1 + a2;          // 0 on lanes >= 9
  // N2: limbs 0..9
  p0 = (u32)colA & M29;
  p1 = (u32)(colA >> 29) & M29;
  p2 = (u32)(colA >> 58);
  u32 out = p0 + lp_up(p1, 1, lane) + lp_up(p2, 2, lane);
  if (TIGHT) out = (lane >= 9 ? out : (out & M29)) + lp_up(lane >= 9 ? 0u : (out >> 29), 1, lane);
  return out;
}

// Variant of lp_fold that hands the carry pieces of the product columns to fold A through shared memory
// instead of normalising them with shuffles first (one dependent hop less):
//   pieces: 3 x 24 words, zero-initialised once (entries P1[0], P2[0], P2[1] are never written)
template <bool TIGHT>
__device__ __forceinline__ u32 lp_fold_sm(u64 col, const LpLane& c, u32* pieces, int lane, u32 addend) {
  u32* P0 = pieces;
  u32* P1 = pieces + 24;
  u32* P2 = pieces + 48;
  u32 p0 = (u32)col & M29, p1 = (u32)(col >> 29) & M29, p2 = (u32)(col >> 58);
  __syncwarp();
  if (lane < 19) {
    P0[lane] = p0;
    P1[lane + 1] = p1;
    P2[lane + 2] = p2;
  }
  __syncwarp();
  const int kk = lane < 9 ? lane : 0;
  const u32 own = p0 + P1[kk] + P2[kk];
  u32 H[12];
  // H_j = P0[9 + j] + P1[9 + j] + P2[9 + j]; the arrays are read as 16-byte vectors from index 8
  {
    const uint4 a0 = *reinterpret_cast<const uint4*>(P0 + 8), a1 = *reinterpret_cast<const uint4*>(P0 + 12), a2 = *reinterpret_cast<const uint4*>(P0 + 16), a3 = *reinterpret_cast<const uint4*>(P0 + 20);
    const uint4 b0 = *reinterpret_cast<const uint4*>(P1 + 8), b1 = *reinterpret_cast<const uint4*>(P1 + 12), b2 = *reinterpret_cast<const uint4*>(P1 + 16), b3 = *reinterpret_cast<const uint4*>(P1 + 20);
    const uint4 c0 = *reinterpret_cast<const uint4*>(P2 + 8), c1 = *reinterpret_cast<const uint4*>(P2 + 12), c2 = *reinterpret_cast<const uint4*>(P2 + 16), c3 = *reinterpret_cast<const uint4*>(P2 + 20);
    H[0] = a0.y + b0.y + c0.y; H[1] = a0.z + b0.z + c0.z; H[2] = a0.w + b0.w + c0.w;
    H[3] = a1.x + b1.x + c1.x; H[4] = a1.y + b1.y + c1.y; H[5] = a1.z + b1.z + c1.z; H[6] = a1.w + b1.w + c1.w;
    H[7] = a2.x + b2.x + c2.x; H[8] = a2.y + b2.y + c2.y; H[9] = a2.z + b2.z + c2.z; H[10] = a2.w + b2.w + c2.w;
    H[11] = a3.x + b3.x + c3.x;
  }
  u64 a0 = (u64)(lane < 9 ? own : 0u) + addend, a1 = 0, a2 = 0;
#pragma unroll
  for (int j = 0; j < 12; j += 3) {
    a0 = lp_mad(H[j], c.K[j], a0);
    a1 = lp_mad(H[j + 1], c.K[j + 1], a1);
    a2 = lp_mad(H[j + 2], c.K[j + 2], a2);
  }
  const u64 colA = a0 + a1 + a2;
  p0 = (u32)colA & M29;
  p1 = (u32)(colA >> 29) & M29;
  p2 = (u32)(colA >> 58);
  u32 out = p0 + lp_up(p1, 1, lane) + lp_up(p2, 2, lane);
  if (TIGHT) out = (lane >= 9 ? out : (out & M29)) + lp_up(lane >= 9 ? 0u : (out >> 29), 1, lane);
  return out;
}

// Fold B: ten lazy limbs (lanes 0..9) -> nine lazy limbs on lanes 0..8, limbs 0..7 < 2^30 + 2^8,
// limb 8 < 2^31, value < 2^262.
__device__ __forceinline__ u32 lp_fold_b(u32 limb, const LpLane& c, int lane) {
  const u32 h = __shfl_sync(0xffffffffu, limb, 9);
  const u64 colB = (u64)(lane < 9 ? limb : 0u) + c.Z - lp_mad(h, c.cp, 0ull);
  return (lane == 8 ? (u32)colB : ((u32)colB & M29)) + lp_up((u32)(colB >> 29), 1, lane);
}

// lanes 0..9 publish their limb as the next operand; the caller orders the write against the
// readers with __syncwarp (same warp) or a barrier / tag protocol (other warps)
__device__ __forceinline__ void lp_store(u32* pad, int lane, u32 limb) {
  if (lane < 10) pad[LP_OFF + lane] = limb;
}

__device__ __forceinline__ void lp_pad_clear(u32* pad, int lane) {
  pad[lane] = 0;
  if (lane < LP_PAD - 32) pad[32 + lane] = 0;
}

// a * b with both operands already in padded shared arrays
template <bool TIGHT>
__device__ __forceinline__ u32 lp_mul(const u32* a_pad, const u32* b_pad, const LpLane& c, u32* h_buf, int lane, u32 addend = 0) {
  return lp_fold<TIGHT>(lp_cols(a_pad, b_pad, lane), c, h_buf, lane, addend);
}

#endif  // __CUDACC__

// ---- single-thread helpers (host + device): conversions at the boundary of an LP computation ----

// ten lazy limbs -> nine (single thread; what lp_fold_b does across lanes): limbs 0..7 < 2^29,
// top limb < 2^31 (value < 2^262): a valid operand of mul29 / sqr29
template <int FIELD>
REEF_HD F29 lp10_to_f29(const u32* l /*10*/) {
  const u32 h = l[9];
  u64 carry = 0;
  F29 r;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const u64 z = FIELD == 0 ? lp_z_fq(k) : lp_z_fp(k);
    const u32 cpk = FIELD == 0 ? lp_cp_fq(k) : lp_cp_fp(k);
    const u64 v = (u64)l[k] + z - (u64)h * cpk + carry;
    if (k < 8) {
      r.l[k] = (u32)v & M29;
      carry = v >> 29;
    } else {
      r.l[k] = (u32)v;
    }
  }
  return r;
}

// lazy plain limbs (value < 2^262) -> canonical integer < p as 8 x 32-bit words
template <class C>
REEF_HD Fe<C> lp_to_canonical(const u32* limbs /*9*/) {
  F29 t;
#pragma unroll
  for (int k = 0; k < 9; k++) t.l[k] = limbs[k];
  f29_normalize(t);                       // limbs < 2^29, top limb < 2^30 + carry
  // x = q 2^254 + r, q < 2^8:  x == r - q c (mod p), made non-negative by adding p when q > 0
  const u32 q = t.l[8] >> 22;
  t.l[8] &= (1u << 22) - 1u;
  Fe<C> r;
  f29_to_words(r.v, t);                   // r < 2^254
  // qc = q * c (c = p - 2^254 < 2^126): 5 words
  u32 qc[8];
  u64 carry = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const u32 ci = i < 7 ? modulus_limb<C>(i) : (modulus_limb<C>(7) - 0x40000000u);
    const u64 v = (u64)q * ci + carry;
    qc[i] = (u32)v;
    carry = v >> 32;
  }
  Fe<C> s;
#pragma unroll
  for (int i = 0; i < 8; i++) s.v[i] = qc[i];
  return fe_sub<C>(r, s);                 // r, qc < p: modular subtraction is exact
}

}  // namespace reef
