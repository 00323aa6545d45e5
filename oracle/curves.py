"""Pallas / Vesta group law and a naive MSM (oracle; test infrastructure only).

The reference reaches curve arithmetic only through un-vendored crates
(`fil_pasta_curves 0.5.2`, `nova-snark`; /root/reference/Cargo.toml:12-14), e.g.
`hyrax_gen.commit` /root/reference/src/backend/commitment.rs:187 and
`RecursiveSNARK::prove_step` /root/reference/src/backend/framework.rs:668-675.
The curves are public standards: both are y^2 = x^3 + 5, Pallas over Fp with
group order Fq, Vesta over Fq with order Fp, generator (-1, 2) on each.  A
multi-scalar multiplication has ONE correct answer as an affine point, so the
naive double-and-add below is a complete oracle for the MSM kernels.
"""
from __future__ import annotations

from .fields import FP, FQ

B = 5
INF = None  # point at infinity


class Curve:
    def __init__(self, name, p, order):
        self.name, self.p, self.order = name, p, order
        self.gen = (p - 1, 2)

    def on_curve(self, P):
        if P is INF:
            return True
        x, y = P
        return (y * y - x * x * x - B) % self.p == 0

    def neg(self, P):
        return INF if P is INF else (P[0], (-P[1]) % self.p)

    def add(self, P, Q):
        p = self.p
        if P is INF:
            return Q
        if Q is INF:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return INF
            lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return (x3, (lam * (x1 - x3) - y1) % p)

    def mul(self, k, P):
        k %= self.order
        R = INF
        while k:
            if k & 1:
                R = self.add(R, P)
            P = self.add(P, P)
            k >>= 1
        return R

    def msm(self, scalars, points):
        R = INF
        for k, P in zip(scalars, points):
            R = self.add(R, self.mul(k, P))
        return R

    # Jacobian helpers: fast generation of many distinct bases k*G for fixtures.
    def multiples(self, n, start=1):
        """[start*G, (start+1)*G, ...] (n points) with one batched inversion."""
        p = self.p
        G = self.gen
        cur = self.mul(start, G)
        pts = []
        for _ in range(n):
            pts.append(cur)
            cur = self.add(cur, G)
        return pts


PALLAS = Curve("pallas", FP, FQ)
VESTA = Curve("vesta", FQ, FP)


def point_to_bytes(P) -> bytes:
    """C-ABI affine encoding: x (32 B LE) || y (32 B LE); infinity = 64 zero bytes
    ((0,0) is not on y^2 = x^3 + 5)."""
    if P is INF:
        return bytes(64)
    return int(P[0]).to_bytes(32, "little") + int(P[1]).to_bytes(32, "little")


def point_from_bytes(b: bytes):
    b = bytes(b)
    x = int.from_bytes(b[:32], "little")
    y = int.from_bytes(b[32:64], "little")
    return INF if x == 0 and y == 0 else (x, y)
