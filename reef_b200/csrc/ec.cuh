// Short-Weierstrass arithmetic for the Pasta curves (y^2 = x^3 + 5, a = 0) over Fe<C>.
//   Pallas: coordinates in Fp (FpCfg), scalars in Fq.   Vesta: coordinates in Fq, scalars in Fp.
//
// Replaces what the reference gets from `fil_pasta_curves` / `pasta-msm` inside nova-snark's
// `vartime_multiscalar_mul` (reached from /root/reference/src/backend/framework.rs:668-675,
// 695-698 and commitment.rs:187, 350-393).
//
// Accumulators are XYZZ (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0):
//   mixed add (XYZZ += affine)   8M + 2S      (madd-2008-s)
//   full add  (XYZZ += XYZZ)    12M + 2S      (add-2008-s)
//   double                       6M + 3S      (dbl-2008-s-1, a = 0)
// All coordinates are in Montgomery form.  __host__ __device__ so that the formulas are
// unit-tested on the CPU through include/reef_b200_testing.h.
#pragma once
#include "fp.cuh"

namespace reef {

template <class C>
struct Affine {   // infinity encoded as (0, 0) (not on the curve since b = 5 != 0)
  Fe<C> x, y;
};

template <class C>
struct XYZZ {
  Fe<C> x, y, zz, zzz;
};

template <class C>
REEF_HD bool affine_is_inf(const Affine<C>& p) { return fe_is_zero<C>(p.x) && fe_is_zero<C>(p.y); }

template <class C>
REEF_HD XYZZ<C> xyzz_inf() {
  XYZZ<C> r;
  r.x = fe_zero<C>();
  r.y = fe_zero<C>();
  r.zz = fe_zero<C>();
  r.zzz = fe_zero<C>();
  return r;
}

template <class C>
REEF_HD bool xyzz_is_inf(const XYZZ<C>& p) { return fe_is_zero<C>(p.zz); }

template <class C>
REEF_HD XYZZ<C> xyzz_from_affine(const Affine<C>& p) {
  XYZZ<C> r;
  if (affine_is_inf<C>(p)) return xyzz_inf<C>();
  r.x = p.x;
  r.y = p.y;
  r.zz = fe_one<C>();
  r.zzz = fe_one<C>();
  return r;
}

// 2 * (affine point)
template <class C>
REEF_HD XYZZ<C> xyzz_dbl_affine(const Affine<C>& p) {
  if (affine_is_inf<C>(p) || fe_is_zero<C>(p.y)) return xyzz_inf<C>();
  XYZZ<C> r;
  Fe<C> u = fe_dbl<C>(p.y);                  // U = 2 Y1
  Fe<C> v = mont_sqr<C>(u);                  // V = U^2
  Fe<C> w = mont_mul<C>(u, v);               // W = U V
  Fe<C> s = mont_mul<C>(p.x, v);             // S = X1 V
  Fe<C> xx = mont_sqr<C>(p.x);
  Fe<C> m = fe_add<C>(fe_dbl<C>(xx), xx);    // M = 3 X1^2
  r.x = fe_sub<C>(mont_sqr<C>(m), fe_dbl<C>(s));
  r.y = fe_sub<C>(mont_mul<C>(m, fe_sub<C>(s, r.x)), mont_mul<C>(w, p.y));
  r.zz = v;
  r.zzz = w;
  return r;
}

template <class C>
REEF_HD XYZZ<C> xyzz_dbl(const XYZZ<C>& p) {
  if (xyzz_is_inf<C>(p) || fe_is_zero<C>(p.y)) return xyzz_inf<C>();
  XYZZ<C> r;
  Fe<C> u = fe_dbl<C>(p.y);
  Fe<C> v = mont_sqr<C>(u);
  Fe<C> w = mont_mul<C>(u, v);
  Fe<C> s = mont_mul<C>(p.x, v);
  Fe<C> xx = mont_sqr<C>(p.x);
  Fe<C> m = fe_add<C>(fe_dbl<C>(xx), xx);
  r.x = fe_sub<C>(mont_sqr<C>(m), fe_dbl<C>(s));
  r.y = fe_sub<C>(mont_mul<C>(m, fe_sub<C>(s, r.x)), mont_mul<C>(w, p.y));
  r.zz = mont_mul<C>(v, p.zz);
  r.zzz = mont_mul<C>(w, p.zzz);
  return r;
}

// acc += (neg ? -q : q), q affine
template <class C>
REEF_HD void xyzz_add_affine(XYZZ<C>& acc, const Affine<C>& q, bool neg) {
  if (affine_is_inf<C>(q)) return;
  Fe<C> qy = neg ? fe_neg<C>(q.y) : q.y;
  if (xyzz_is_inf<C>(acc)) {
    acc.x = q.x;
    acc.y = qy;
    acc.zz = fe_one<C>();
    acc.zzz = fe_one<C>();
    return;
  }
  Fe<C> u2 = mont_mul<C>(q.x, acc.zz);
  Fe<C> s2 = mont_mul<C>(qy, acc.zzz);
  Fe<C> p = fe_sub<C>(u2, acc.x);
  Fe<C> r = fe_sub<C>(s2, acc.y);
  if (fe_is_zero<C>(p)) {
    if (fe_is_zero<C>(r)) {
      Affine<C> t;
      t.x = q.x;
      t.y = qy;
      acc = xyzz_dbl_affine<C>(t);
    } else {
      acc = xyzz_inf<C>();
    }
    return;
  }
  Fe<C> pp = mont_sqr<C>(p);
  Fe<C> ppp = mont_mul<C>(p, pp);
  Fe<C> qq = mont_mul<C>(acc.x, pp);
  Fe<C> x3 = fe_sub<C>(fe_sub<C>(mont_sqr<C>(r), ppp), fe_dbl<C>(qq));
  Fe<C> y3 = fe_sub<C>(mont_mul<C>(r, fe_sub<C>(qq, x3)), mont_mul<C>(acc.y, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = mont_mul<C>(acc.zz, pp);
  acc.zzz = mont_mul<C>(acc.zzz, ppp);
}

// acc += q, both XYZZ
template <class C>
REEF_HD void xyzz_add(XYZZ<C>& acc, const XYZZ<C>& q) {
  if (xyzz_is_inf<C>(q)) return;
  if (xyzz_is_inf<C>(acc)) {
    acc = q;
    return;
  }
  Fe<C> u1 = mont_mul<C>(acc.x, q.zz);
  Fe<C> u2 = mont_mul<C>(q.x, acc.zz);
  Fe<C> s1 = mont_mul<C>(acc.y, q.zzz);
  Fe<C> s2 = mont_mul<C>(q.y, acc.zzz);
  Fe<C> p = fe_sub<C>(u2, u1);
  Fe<C> r = fe_sub<C>(s2, s1);
  if (fe_is_zero<C>(p)) {
    if (fe_is_zero<C>(r)) acc = xyzz_dbl<C>(acc);
    else acc = xyzz_inf<C>();
    return;
  }
  Fe<C> pp = mont_sqr<C>(p);
  Fe<C> ppp = mont_mul<C>(p, pp);
  Fe<C> qq = mont_mul<C>(u1, pp);
  Fe<C> x3 = fe_sub<C>(fe_sub<C>(mont_sqr<C>(r), ppp), fe_dbl<C>(qq));
  Fe<C> y3 = fe_sub<C>(mont_mul<C>(r, fe_sub<C>(qq, x3)), mont_mul<C>(s1, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = mont_mul<C>(mont_mul<C>(acc.zz, q.zz), pp);
  acc.zzz = mont_mul<C>(mont_mul<C>(acc.zzz, q.zzz), ppp);
}

// XYZZ -> affine given inv = 1 / ZZZ:  1/Z = ZZ * inv,  1/ZZ = (1/Z)^2
template <class C>
REEF_HD Affine<C> xyzz_to_affine_with_inv(const XYZZ<C>& p, const Fe<C>& inv_zzz) {
  Affine<C> r;
  if (xyzz_is_inf<C>(p)) {
    r.x = fe_zero<C>();
    r.y = fe_zero<C>();
    return r;
  }
  Fe<C> iz = mont_mul<C>(p.zz, inv_zzz);
  Fe<C> izz = mont_sqr<C>(iz);
  r.x = mont_mul<C>(p.x, izz);
  r.y = mont_mul<C>(p.y, inv_zzz);
  return r;
}

template <class C>
REEF_HD Affine<C> xyzz_to_affine(const XYZZ<C>& p) {
  if (xyzz_is_inf<C>(p)) {
    Affine<C> r;
    r.x = fe_zero<C>();
    r.y = fe_zero<C>();
    return r;
  }
  return xyzz_to_affine_with_inv<C>(p, fe_inv<C>(p.zzz));
}

// k * p for a small non-negative integer k (double-and-add, MSB first)
template <class C>
REEF_HD XYZZ<C> xyzz_mul_small(const XYZZ<C>& p, u64 k) {
  XYZZ<C> r = xyzz_inf<C>();
  for (int b = 63; b >= 0; b--) {
    r = xyzz_dbl<C>(r);
    if ((k >> b) & 1) xyzz_add<C>(r, p);
  }
  return r;
}

}  // namespace reef
