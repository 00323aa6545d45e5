// 255-bit prime-field arithmetic for the Pasta cycle (Fq = Pallas scalar, Fp = Pallas base)
// on sm_100a.  8 x 32-bit limbs, Montgomery form (R = 2^256) internally, canonical
// little-endian integers at every memory / ABI boundary.
//
// Replaces, for the hot path, what the reference gets from GMP (`rug::Integer` mul +
// `rem_floor`, /root/reference/src/backend/r1cs_helper.rs:457-503) and from
// `fil_pasta_curves` field types (/root/reference/src/backend/framework.rs:1-2).
//
// Design notes
//  * One IMAD.WIDE.U32[.X] per 32x32 partial product: `mad.lo.cc/madc.hi.cc` pairs in
//    even/odd column accumulators (carry stays in a predicate, no extra adds).
//  * Both Pasta primes are p = 2^254 + c with c < 2^126 and p == 1 (mod 2^32):
//    limbs {1, P1, P2, P3, 0, 0, 0, 2^30}.  Montgomery reduction therefore needs
//    m = -T[i] (no multiply) and only 3 real products per limb; the 2^254 term is a shift.
//  * Everything is __host__ __device__: the host build emulates the carry chains with
//    64-bit integers so the *structure* (column bookkeeping, reduction) is unit-tested
//    on CPU (tests/test_host_abi.py::test_shared_field_code_host_instantiation) before it ever runs on a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define REEF_HD __host__ __device__ __forceinline__
#define REEF_D __device__ __forceinline__
#else
#define REEF_HD inline
#define REEF_D inline
#endif

namespace reef {

typedef uint32_t u32;
typedef uint64_t u64;

#include "fp_consts.inc"

#ifdef __CUDA_ARCH__
#include "fp_asm.inc"
#endif

// ---------------------------------------------------------------------------------------
// carry-chain primitives (device: single PTX blocks; host: 64-bit emulation)
// ---------------------------------------------------------------------------------------
template <int N>
REEF_HD u32 acc_add(u32* acc, const u32* b) {
#ifdef __CUDA_ARCH__
  if constexpr (N == 4) return ptx_acc_add_4(acc, b);
  else if constexpr (N == 5) return ptx_acc_add_5(acc, b);
  else if constexpr (N == 8) return ptx_acc_add_8(acc, b);
  else if constexpr (N == 9) return ptx_acc_add_9(acc, b);
  else if constexpr (N == 16) return ptx_acc_add_16(acc, b);
  else { static_assert(N == 17, "unsupported chain length"); return ptx_acc_add_17(acc, b); }
#else
  u64 c = 0;
  for (int i = 0; i < N; i++) { c += (u64)acc[i] + b[i]; acc[i] = (u32)c; c >>= 32; }
  return (u32)c;
#endif
}

template <int N>
REEF_HD u32 acc_add_cin(u32* acc, const u32* b, u32 cin) {
#ifdef __CUDA_ARCH__
  if constexpr (N == 8) return ptx_acc_add_cin_8(acc, b, cin);
  else { static_assert(N == 9, "unsupported chain length"); return ptx_acc_add_cin_9(acc, b, cin); }
#else
  u64 c = cin;
  for (int i = 0; i < N; i++) { c += (u64)acc[i] + b[i]; acc[i] = (u32)c; c >>= 32; }
  return (u32)c;
#endif
}

template <int N>
REEF_HD u32 acc_sub(u32* acc, const u32* b) {  // returns borrow (0/1)
#ifdef __CUDA_ARCH__
  if constexpr (N == 8) return ptx_acc_sub_8(acc, b);
  else { static_assert(N == 9, "unsupported chain length"); return ptx_acc_sub_9(acc, b); }
#else
  u64 br = 0;
  for (int i = 0; i < N; i++) {
    u64 d = (u64)acc[i] - b[i] - br;
    acc[i] = (u32)d;
    br = (d >> 32) & 1;
  }
  return (u32)br;
#endif
}

REEF_HD u32 add3_8(u32* r, const u32* a, const u32* b) {
#ifdef __CUDA_ARCH__
  return ptx_add3_8(r, a, b);
#else
  u64 c = 0;
  for (int i = 0; i < 8; i++) { c += (u64)a[i] + b[i]; r[i] = (u32)c; c >>= 32; }
  return (u32)c;
#endif
}

REEF_HD u32 sub3_8(u32* r, const u32* a, const u32* b) {  // returns borrow
#ifdef __CUDA_ARCH__
  return ptx_sub3_8(r, a, b);
#else
  u64 br = 0;
  for (int i = 0; i < 8; i++) {
    u64 d = (u64)a[i] - b[i] - br;
    r[i] = (u32)d;
    br = (d >> 32) & 1;
  }
  return (u32)br;
#endif
}

// acc[0..8) += {x0,x1,x2,x3} * b, product k landing on limbs (2k, 2k+1); returns carry out.
REEF_HD u32 mad_row4(u32* acc, u32 x0, u32 x1, u32 x2, u32 x3, u32 b) {
#ifdef __CUDA_ARCH__
  return ptx_mad_row4(acc, x0, x1, x2, x3, b);
#else
  const u32 x[4] = {x0, x1, x2, x3};
  u64 c = 0;
  for (int k = 0; k < 4; k++) {
    u64 pr = (u64)x[k] * b;
    u64 lo = (u64)acc[2 * k] + (u32)pr + c;
    acc[2 * k] = (u32)lo;
    u64 hi = (u64)acc[2 * k + 1] + (pr >> 32) + (lo >> 32);
    acc[2 * k + 1] = (u32)hi;
    c = hi >> 32;
  }
  return (u32)c;
#endif
}

// (m1, m2) += x*y as a 64-bit pair, carry into m3.
REEF_HD void mad_mid(u32& m1, u32& m2, u32& m3, u32 x, u32 y) {
#ifdef __CUDA_ARCH__
  ptx_mad_mid(m1, m2, m3, x, y);
#else
  u64 pr = (u64)x * y;
  u64 lo = (u64)m1 + (u32)pr;
  m1 = (u32)lo;
  u64 hi = (u64)m2 + (pr >> 32) + (lo >> 32);
  m2 = (u32)hi;
  m3 += (u32)(hi >> 32);
#endif
}

// ---------------------------------------------------------------------------------------
// 8x8 -> 16 limb schoolbook product (64 IMAD.WIDE) and square
// ---------------------------------------------------------------------------------------
REEF_HD void mul_wide(u32* r /*16*/, const u32* a, const u32* b) {
  u32 ev[16], od[16];  // od[k] holds column k+1
#pragma unroll
  for (int i = 0; i < 16; i++) { ev[i] = 0; od[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    u32 c;
    c = mad_row4(ev + i, a[0], a[2], a[4], a[6], b[i]);
    ev[i + 8] += c;
    c = mad_row4(od + i, a[1], a[3], a[5], a[7], b[i]);
    od[i + 8] += c;
    c = mad_row4(od + i, a[0], a[2], a[4], a[6], b[i + 1]);
    od[i + 8] += c;
    c = mad_row4(ev + i + 2, a[1], a[3], a[5], a[7], b[i + 1]);
    if (i + 10 < 16) ev[i + 10] += c;
  }
  // r = ev + (od << 32)
  r[0] = ev[0];
#pragma unroll
  for (int i = 1; i < 16; i++) r[i] = ev[i];
  u32 t[16];
  t[0] = 0;
#pragma unroll
  for (int i = 1; i < 16; i++) t[i] = od[i - 1];
  acc_add<16>(r, t);
}

// ---------------------------------------------------------------------------------------
// field configuration
// ---------------------------------------------------------------------------------------
struct FqCfg {  // Pallas scalar field (circuit / MLE field).  r1cs_helper.rs:37-38
  static constexpr u32 P1 = 0x8c46eb21u, P2 = 0x0994a8ddu, P3 = 0x224698fcu;
  REEF_HD static constexpr u32 r(int i) { constexpr u32 k[8] = REEF_FQ_R; return k[i]; }
  REEF_HD static constexpr u32 r2(int i) { constexpr u32 k[8] = REEF_FQ_R2; return k[i]; }
  REEF_HD static constexpr u32 r3(int i) { constexpr u32 k[8] = REEF_FQ_R3; return k[i]; }
};
struct FpCfg {  // Pallas base field (coordinates of Pallas points; scalars of Vesta)
  static constexpr u32 P1 = 0x992d30edu, P2 = 0x094cf91bu, P3 = 0x224698fcu;
  REEF_HD static constexpr u32 r(int i) { constexpr u32 k[8] = REEF_FP_R; return k[i]; }
  REEF_HD static constexpr u32 r2(int i) { constexpr u32 k[8] = REEF_FP_R2; return k[i]; }
  REEF_HD static constexpr u32 r3(int i) { constexpr u32 k[8] = REEF_FP_R3; return k[i]; }
};

template <class C>
REEF_HD constexpr u32 modulus_limb(int i) {
  return i == 0 ? 1u : i == 1 ? C::P1 : i == 2 ? C::P2 : i == 3 ? C::P3 : i == 7 ? 0x40000000u : 0u;
}

template <class C>
struct alignas(32) Fe {
  u32 v[8];
};

// r = (x >= p) ? x - p : x        (x < 2p)
template <class C>
REEF_HD void cond_sub_p(u32* x) {
  u32 p[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = modulus_limb<C>(i);
  u32 borrow = sub3_8(d, x, p);
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : d[i];
}

// Montgomery reduction of a 16-limb T (T < p * 2^256): returns T / 2^256 mod p, canonical.
// `extra` = value of an optional 17th limb of the running sum that the caller already
// folded away (always 0 for plain products).
template <class C>
REEF_HD void mont_reduce(u32* out /*8*/, u32* T /*16, destroyed*/) {
  u32 m[8];
  u32 cA = 0;       // carry out of the 5-limb chain of the previous step (belongs to col i+4)
  u32 cS = 0;       // carry out of the early (m0 << 30) add into column 7
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i == 7) {   // column 7 must contain m0 << 30 before m7 is derived from it
      u64 s = (u64)T[7] + (u64)(m[0] << 30);
      T[7] = (u32)s;
      cS = (u32)(s >> 32);
    }
    u32 mi = 0u - T[i];
    m[i] = mi;
    u64 l1 = (u64)mi * C::P1;
    u64 l3 = (u64)mi * C::P3;
    u32 A[5];
    A[0] = mi;
    A[1] = (u32)l1;
    A[2] = (u32)(l1 >> 32);
    A[3] = (u32)l3;
    A[4] = (u32)(l3 >> 32);
    mad_mid(A[2], A[3], A[4], mi, C::P2);
    A[4] += cA;                       // < 2^30 + 1, cannot wrap
    cA = acc_add<5>(T + i, A);        // T[i] becomes 0
  }
  // pending: cA at column 12; (M << 254) at columns 7..15 (+ carry limb), with column 7's
  // low part already added (carry cS pending at column 8).
  u32 S[9];
  S[0] = 0;                           // column 7: only m0 << 30, already added
#pragma unroll
  for (int j = 1; j < 8; j++) S[j] = (m[j] << 30) | (m[j - 1] >> 2);
  S[8] = m[7] >> 2;
  u32 hi[9];
#pragma unroll
  for (int j = 0; j < 8; j++) hi[j] = T[8 + j];
  hi[8] = 0;
  // add S[1..8] at columns 8..15(+16) with carry-in cS
  acc_add_cin<8>(hi, S + 1, cS);      // carry out impossible: result < 2p < 2^256 (see header)
  // add cA at column 12
  u32 cvec[4] = {cA, 0, 0, 0};
  acc_add<4>(hi + 4, cvec);
#pragma unroll
  for (int j = 0; j < 8; j++) out[j] = hi[j];
  cond_sub_p<C>(out);
}

template <class C>
REEF_HD Fe<C> mont_mul(const Fe<C>& a, const Fe<C>& b) {
  u32 T[16];
  mul_wide(T, a.v, b.v);
  Fe<C> r;
  mont_reduce<C>(r.v, T);
  return r;
}

template <class C>
REEF_HD Fe<C> mont_sqr(const Fe<C>& a) { return mont_mul<C>(a, a); }

template <class C>
REEF_HD Fe<C> fe_add(const Fe<C>& a, const Fe<C>& b) {
  Fe<C> r;
  add3_8(r.v, a.v, b.v);  // < 2p < 2^256: no carry out
  cond_sub_p<C>(r.v);
  return r;
}

template <class C>
REEF_HD Fe<C> fe_sub(const Fe<C>& a, const Fe<C>& b) {
  Fe<C> r;
  u32 borrow = sub3_8(r.v, a.v, b.v);
  u32 p[8];
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = borrow ? modulus_limb<C>(i) : 0u;
  acc_add<8>(r.v, p);
  return r;
}

template <class C>
REEF_HD Fe<C> fe_neg(const Fe<C>& a) {
  Fe<C> z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.v[i] = 0;
  return fe_sub<C>(z, a);
}

template <class C>
REEF_HD Fe<C> fe_dbl(const Fe<C>& a) { return fe_add<C>(a, a); }

template <class C>
REEF_HD bool fe_is_zero(const Fe<C>& a) {
  u32 o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i];
  return o == 0;
}

template <class C>
REEF_HD bool fe_eq(const Fe<C>& a, const Fe<C>& b) {
  u32 o = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

template <class C>
REEF_HD Fe<C> fe_zero() {
  Fe<C> z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.v[i] = 0;
  return z;
}

template <class C>
REEF_HD Fe<C> fe_one() {  // Montgomery one
  Fe<C> z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.v[i] = C::r(i);
  return z;
}

template <class C>
REEF_HD Fe<C> fe_r2() {
  Fe<C> z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.v[i] = C::r2(i);
  return z;
}

// canonical integer (< p) -> Montgomery form
template <class C>
REEF_HD Fe<C> to_mont(const Fe<C>& canon) { return mont_mul<C>(canon, fe_r2<C>()); }

// Montgomery form -> canonical integer
template <class C>
REEF_HD Fe<C> from_mont(const Fe<C>& a) {
  u32 T[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { T[i] = a.v[i]; T[8 + i] = 0; }
  Fe<C> r;
  mont_reduce<C>(r.v, T);
  return r;
}

// small non-negative integer -> Montgomery form
template <class C>
REEF_HD Fe<C> fe_from_u64(u64 x) {
  Fe<C> c = fe_zero<C>();
  c.v[0] = (u32)x;
  c.v[1] = (u32)(x >> 32);
  return to_mont<C>(c);
}

// ---------------------------------------------------------------------------------------
// lazy accumulation: sum of 16-limb products in a 17-limb accumulator, one reduction at
// the end.  Capacity: 2^32 products of canonical operands.
// ---------------------------------------------------------------------------------------
struct Wide17 {
  u32 v[17];
};

REEF_HD void wide_zero(Wide17& w) {
#pragma unroll
  for (int i = 0; i < 17; i++) w.v[i] = 0;
}

// w += a * b   (plain integers, any 8-limb values)
REEF_HD void wide_mac(Wide17& w, const u32* a, const u32* b) {
  u32 t[17];
  mul_wide(t, a, b);
  t[16] = 0;
  acc_add<17>(w.v, t);
}

REEF_HD void wide_add(Wide17& w, const Wide17& o) { acc_add<17>(w.v, o.v); }

// w += t * b for a single-limb t (document codes): 8 products instead of 64.
REEF_HD void wide_mac_small(Wide17& w, u32 t, const u32* b) {
  u32 pr[17];
#pragma unroll
  for (int i = 0; i < 17; i++) pr[i] = 0;
  u32 od[9];
#pragma unroll
  for (int i = 0; i < 9; i++) od[i] = 0;
  mad_row4(pr, b[0], b[2], b[4], b[6], t);       // columns (0,1),(2,3),(4,5),(6,7): no carry out
  mad_row4(od, b[1], b[3], b[5], b[7], t);       // columns (1,2),...,(7,8)
  u32 sh[9];
  sh[0] = 0;
#pragma unroll
  for (int i = 1; i < 9; i++) sh[i] = od[i - 1];
  acc_add<9>(pr, sh);
  acc_add<17>(w.v, pr);
}

// (w / R) mod p, canonical, for ANY 17-limb w.  Used once per thread at the end of a lazy
// accumulation: fold limb 16 (2^512 == R2 mod p), squeeze the high half below p, then one
// Montgomery reduction.
template <class C>
REEF_HD Fe<C> wide_reduce_div_R(const Wide17& w) {
  u32 T[17];
#pragma unroll
  for (int i = 0; i < 17; i++) T[i] = w.v[i];
#pragma unroll
  for (int round = 0; round < 2; round++) {
    u32 top = T[16];
    T[16] = 0;
    u32 add[17];
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      c += (u64)top * C::r2(i);
      add[i] = (u32)c;
      c >>= 32;
    }
    add[8] = (u32)c;
#pragma unroll
    for (int i = 9; i < 17; i++) add[i] = 0;
    acc_add<17>(T, add);
  }
  // T < 2^512; make the high half < p so that T < p * 2^256 (precondition of mont_reduce).
  u32 hi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) hi[i] = T[8 + i];
#pragma unroll
  for (int k = 0; k < 4; k++) cond_sub_p<C>(hi);
#pragma unroll
  for (int i = 0; i < 8; i++) T[8 + i] = hi[i];
  Fe<C> r;
  mont_reduce<C>(r.v, T);
  return r;
}

// (w mod p) as a canonical integer (NOT divided by R).
template <class C>
REEF_HD Fe<C> wide_reduce_canonical(const Wide17& w) {
  return mont_mul<C>(wide_reduce_div_R<C>(w), fe_r2<C>());
}

// ---------------------------------------------------------------------------------------
// exponentiation helpers (used for inversion: a^(p-2); cold paths only)
// ---------------------------------------------------------------------------------------
template <class C>
REEF_HD Fe<C> fe_pow_pm2(const Fe<C>& a) {
  // exponent p - 2 = {0xffffffff (= 1 - 2 wraps), ...}: compute limbs of p-2 explicitly
  u32 e[8];
#pragma unroll
  for (int i = 0; i < 8; i++) e[i] = modulus_limb<C>(i);
  // p ends in ...00000001, so p - 2 borrows through limb 0
  e[0] = 0xffffffffu;
  e[1] = C::P1 - 1u;
  Fe<C> acc = fe_one<C>();
  for (int i = 7; i >= 0; i--) {
    for (int b = 31; b >= 0; b--) {
      acc = mont_sqr<C>(acc);
      if ((e[i] >> b) & 1u) acc = mont_mul<C>(acc, a);
    }
  }
  return acc;
}

// Modular inverse by the binary extended Euclidean algorithm on the CANONICAL integer
// (about 10x fewer instructions than a^(p-2)); returns 0 for 0.  Branchy: meant for the
// single-thread tails (final affine conversion), not for SIMT-wide use.
template <class C>
REEF_HD void inv_canonical(u32* out, const u32* a_in) {
  u32 p[8], u[8], v[8], x1[8], x2[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { p[i] = modulus_limb<C>(i); u[i] = a_in[i]; v[i] = p[i]; x1[i] = 0; x2[i] = 0; }
  x1[0] = 1;
  u32 nz = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) nz |= u[i];
  if (nz == 0) {
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = 0;
    return;
  }
  auto is_one = [](const u32* x) {
    u32 o = x[0] ^ 1u;
    for (int i = 1; i < 8; i++) o |= x[i];
    return o == 0;
  };
  auto shr1 = [](u32* x, u32 top) {
    for (int i = 0; i < 7; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[7] = (x[7] >> 1) | (top << 31);
  };
  auto halve_mod = [&](u32* x) {   // x <- x / 2 mod p
    u32 carry = 0;
    if (x[0] & 1u) carry = acc_add<8>(x, p);
    shr1(x, carry);
  };
  auto geq = [](const u32* x, const u32* y) {
    for (int i = 7; i >= 0; i--) {
      if (x[i] > y[i]) return true;
      if (x[i] < y[i]) return false;
    }
    return true;
  };
  auto sub_mod = [&](u32* x, const u32* y) {   // x <- x - y mod p
    if (acc_sub<8>(x, y)) acc_add<8>(x, p);
  };
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1u)) { shr1(u, 0); halve_mod(x1); }
    while (!(v[0] & 1u)) { shr1(v, 0); halve_mod(x2); }
    if (geq(u, v)) { acc_sub<8>(u, v); sub_mod(x1, x2); }
    else { acc_sub<8>(v, u); sub_mod(x2, x1); }
  }
  const bool uo = is_one(u);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = uo ? x1[i] : x2[i];
}

template <class C>
REEF_HD Fe<C> fe_r3() {
  Fe<C> z;
#pragma unroll
  for (int i = 0; i < 8; i++) z.v[i] = C::r3(i);
  return z;
}

// inverse of a Montgomery-form element, result in Montgomery form:
// inv(aR) = a^-1 R^-1 (as an integer);  mont_mul(., R^3) = a^-1 R^-1 R^3 R^-1 = a^-1 R.
template <class C>
REEF_HD Fe<C> fe_inv(const Fe<C>& a) {
  Fe<C> t;
  inv_canonical<C>(t.v, a.v);
  return mont_mul<C>(t, fe_r3<C>());
}


}  // namespace reef
