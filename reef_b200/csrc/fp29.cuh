// Carry-free, low-latency Montgomery arithmetic for the Pasta fields on sm_100a:
// 9 limbs of 29 bits, Montgomery radix R' = 2^261, 64-bit column accumulators.
//
// Why a second representation next to fp.cuh (8 x 32-bit limbs, carry-flag chains):
// the Fiat-Shamir transcript of the nlookup sum-check (/root/reference/src/backend/r1cs.rs:2260-2311,
// r1cs_helper.rs:479-488) is ONE dependent chain of ~200 field multiplications per Poseidon
// permutation executed by a single warp, so what matters there is the latency of one
// multiplication, not chip throughput.  Measured on B200 (tools/bench_lat.cu, tools/bench_fp.cu):
//   * IMAD.WIDE.U32 with a 64-bit accumulator issues every 2 cycles per warp and a dependent
//     accumulate chain costs ~3 cycles per link; an add-with-carry link costs ~2.3 cycles;
//   * fp.cuh's mont_mul is a web of ~180 instructions serialised by the carry flag: 896 cycles.
// With 29-bit limbs every 29x29 product fits 58 bits, a column of 9 products (even of 45, for a
// 5-term matrix row) fits 64 bits, so the schoolbook product is 81 INDEPENDENT-column
// IMAD.WIDEs with no carry flag at all, the Montgomery factor of a column is m = -col mod 2^29
// (p == 1 mod 2^29), and p = 2^254 + c contributes 4 products and one shift per column.
// Values are kept "almost normalised" (limbs < 2^29 + 2^7, value < 8p): no conditional
// subtraction and no carry ripple inside a chain; canonicalisation happens once at the boundary.
#pragma once
#include "fp.cuh"

namespace reef {

static constexpr u32 M29 = (1u << 29) - 1u;

struct F29 {
  u32 l[9];
};

// bits [29k, 29k+29) of a 256-bit constant given as 8 x 32-bit limbs
REEF_HD constexpr u32 slice29(const u32* w, int k) {
  const int bit = 29 * k, i = bit >> 5, sh = bit & 31;
  const u64 lo = w[i], hi = (i + 1 < 8) ? w[i + 1] : 0;
  return (u32)(((lo | (hi << 32)) >> sh) & M29);
}

template <class C>
REEF_HD constexpr u32 p29(int k) {
  const u32 w[8] = {modulus_limb<C>(0), modulus_limb<C>(1), modulus_limb<C>(2), modulus_limb<C>(3),
                    modulus_limb<C>(4), modulus_limb<C>(5), modulus_limb<C>(6), modulus_limb<C>(7)};
  return slice29(w, k);
}

// 8 x 32 (any 256-bit integer) -> 9 x 29, exact
REEF_HD F29 f29_from_words(const u32* w) {
  F29 r;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const int bit = 29 * k, i = bit >> 5, sh = bit & 31;
    u32 v = w[i] >> sh;
    if (sh > 3 && i + 1 < 8) v |= w[i + 1] << (32 - sh);
    r.l[k] = (k == 8) ? v : (v & M29);
  }
  return r;
}

// exact carry ripple: limbs < 2^29 except the top one
REEF_HD void f29_normalize(F29& a) {
  u32 c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const u32 t = a.l[k] + c;
    a.l[k] = t & M29;
    c = t >> 29;
  }
  a.l[8] += c;
}

// 9 x 29 (normalised, value < 2^256) -> 8 x 32
REEF_HD void f29_to_words(u32* w, const F29& a) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int bit = 32 * i, k = bit / 29, sh = bit % 29;   // word i starts inside limb k at offset sh
    u64 v = (u64)a.l[k] >> sh;
    v |= (u64)a.l[k + 1] << (29 - sh);
    if (k + 2 < 9) v |= (u64)a.l[k + 2] << (58 - sh);
    w[i] = (u32)v;
  }
}

// one parallel carry step: limbs < 2^31.7 in -> limbs < 2^29 + 8 out (top limb unmasked)
REEF_HD F29 f29_relax(const F29& a) {
  F29 r;
  r.l[0] = a.l[0] & M29;
#pragma unroll
  for (int k = 1; k < 8; k++) r.l[k] = (a.l[k] & M29) + (a.l[k - 1] >> 29);
  r.l[8] = a.l[8] + (a.l[7] >> 29);
  return r;
}

// limb-wise sum, no carry handling (callers track the bound: limbs must stay < 2^30 for a
// multiplication operand, < 2^32 for f29_relax)
REEF_HD F29 f29_add_lazy(const F29& a, const F29& b) {
  F29 r;
#pragma unroll
  for (int k = 0; k < 9; k++) r.l[k] = a.l[k] + b.l[k];
  return r;
}

REEF_HD F29 f29_zero() {
  F29 r;
#pragma unroll
  for (int k = 0; k < 9; k++) r.l[k] = 0;
  return r;
}

// col[0..17) (+)= a * b, column k = sum_{i+j=k} a_i b_j; col[17] untouched
template <bool ACC>
REEF_HD void mul29_cols(u64* col, const F29& a, const F29& b) {
#pragma unroll
  for (int k = 0; k < 17; k++) {
    u64 acc = ACC ? col[k] : 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const int j = k - i;
      if (j >= 0 && j < 9) acc += (u64)a.l[i] * b.l[j];
    }
    col[k] = acc;
  }
}

// Montgomery reduction of 18 columns (col[17] must be 0 on entry; every col < 2^63.6):
// returns sum(col_k 2^(29k)) / 2^261 mod p, almost normalised, value < value_in / 2^261 + p.
template <class C>
REEF_HD F29 redc29(u64* col) {
  constexpr u32 P1 = p29<C>(1), P2 = p29<C>(2), P3 = p29<C>(3), P4 = p29<C>(4);
  static_assert(p29<C>(0) == 1 && p29<C>(5) == 0 && p29<C>(6) == 0 && p29<C>(7) == 0 && p29<C>(8) == (1u << 22),
                "p = 2^254 + c with c < 2^126 and p == 1 mod 2^29");
  u64 carry = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const u64 t = col[i] + carry;
    const u32 tl = (u32)t & M29;
    const u32 m = (0u - tl) & M29;               // t + m == 0 (mod 2^29)
    carry = (t >> 29) + (tl != 0u ? 1u : 0u);    // = (t + m) >> 29, without waiting for m
    col[i + 1] += (u64)m * P1;
    col[i + 2] += (u64)m * P2;
    col[i + 3] += (u64)m * P3;
    col[i + 4] += (u64)m * P4;
#ifdef __CUDA_ARCH__
    asm("mad.wide.u32 %0, %1, 4194304, %0;" : "+l"(col[i + 8]) : "r"(m));   // one IMAD.WIDE instead of 4 shift/add ops
#else
    col[i + 8] += (u64)m << 22;
#endif
  }
  col[9] += carry;
  // two parallel carry steps instead of a 9-deep ripple
  u64 x[9];
  x[0] = col[9] & M29;
#pragma unroll
  for (int k = 1; k < 8; k++) x[k] = (col[9 + k] & M29) + (col[8 + k] >> 29);
  x[8] = col[17] + (col[16] >> 29);
  F29 r;
  r.l[0] = (u32)x[0];
#pragma unroll
  for (int k = 1; k < 8; k++) r.l[k] = ((u32)x[k] & M29) + (u32)(x[k - 1] >> 29);
  r.l[8] = (u32)x[8] + (u32)(x[7] >> 29);
  return r;
}

// a * b / 2^261 mod p.  Operand limbs < 2^30 + 2^8, values < 2^7 p  ->  result almost normalised, < 2p.
template <class C>
REEF_HD F29 mul29(const F29& a, const F29& b) {
  u64 col[18];
  mul29_cols<false>(col, a, b);
  col[17] = 0;
  return redc29<C>(col);
}

// a * a / 2^261 mod p: 36 doubled cross products + 9 squares instead of 81 products.
// Operand limbs < 2^30 + 2^8 (a column then holds < 4 * 2^61.01 + 2^60.01 + reduction terms < 2^63.4).
template <class C>
REEF_HD F29 sqr29(const F29& a) {
  u32 d[9];
#pragma unroll
  for (int i = 0; i < 9; i++) d[i] = a.l[i] << 1;
  u64 col[18];
#pragma unroll
  for (int k = 0; k < 17; k++) {
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const int j = k - i;
      if (j > i && j < 9) acc += (u64)a.l[i] * d[j];
    }
    if ((k & 1) == 0) acc += (u64)a.l[k / 2] * a.l[k / 2];
    col[k] = acc;
  }
  col[17] = 0;
  return redc29<C>(col);
}

// ---- constants (as 29-bit limb slices of 8 x 32 tables) ---------------------------------
template <class C>
REEF_HD F29 f29_const_2_266() {   // 2^266 mod p: Montgomery-256 -> Montgomery-261
  // 2^266 = 2^256 * 2^10: ten modular doublings of the Montgomery one
  Fe<C> x = fe_one<C>();
  for (int i = 0; i < 10; i++) x = fe_dbl<C>(x);
  return f29_from_words(x.v);
}

// Montgomery-256 element (canonical integer x 2^256 mod p) -> Montgomery-261, almost normalised
template <class C>
REEF_HD F29 f29_from_mont256(const Fe<C>& a, const F29& k266) {
  return mul29<C>(f29_from_words(a.v), k266);
}

// Montgomery-261 (almost normalised, < 8p) -> Montgomery-256 canonical element
template <class C>
REEF_HD Fe<C> f29_to_mont256(const F29& a) {
  u32 rw[8];
#pragma unroll
  for (int i = 0; i < 8; i++) rw[i] = C::r(i);
  F29 t = mul29<C>(a, f29_from_words(rw));       // x 2^261 * 2^256 / 2^261
  f29_normalize(t);                              // value <= p
  Fe<C> r;
  f29_to_words(r.v, t);
  cond_sub_p<C>(r.v);
  return r;
}

// canonical integer (< p) -> Montgomery-261 via (2^261)^2 = 2^522 = 2^512 * 2^10
template <class C>
REEF_HD F29 f29_from_canonical(const Fe<C>& a) {
  Fe<C> x = fe_r2<C>();
  for (int i = 0; i < 10; i++) x = fe_dbl<C>(x);
  return mul29<C>(f29_from_words(a.v), f29_from_words(x.v));
}

// Montgomery-261 -> canonical integer (< p)
template <class C>
REEF_HD Fe<C> f29_to_canonical(const F29& a) {
  F29 one = f29_zero();
  one.l[0] = 1;
  F29 t = mul29<C>(a, one);
  f29_normalize(t);
  Fe<C> r;
  f29_to_words(r.v, t);
  cond_sub_p<C>(r.v);
  return r;
}

}  // namespace reef
