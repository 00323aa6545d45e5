#!/usr/bin/env python3
"""Exact model of the LANE-PARALLEL field multiplication used by the Fiat-Shamir permutation
(reef_b200/csrc/lpmul.cuh) and of the permutation schedule built on it (poseidon_lp.cuh).

Build/verification tool (not a product path, not the oracle): it restates, lane by lane and with
every 64-bit bound asserted, what the CUDA code does, so that the algorithm can be checked against
Python integers before it runs on a GPU, and it emits the constant tables the kernels use
(`python tools/lp_model.py --emit reef_b200/csrc/lp_consts.inc`).

Representation: x = sum_k x_k 2^(29 k), 9 limbs, PLAIN residue mod p (no Montgomery factor), lazy:
limbs 0..7 < 2^29.7, limb 8 < 2^30.7, value < 2^261.8.  One lane owns one limb / one column.

multi-product  sum_t a^(t) * b^(t)  (mod p):
  P   lane k: col_k = sum_t sum_i a^(t)_i * b^(t)_(k-i)                    (<= 9 n IMAD.WIDE per lane)
  N1  col = p0 + 2^29 p1 + 2^58 p2;  limb_k = p0_k + p1_(k-1) + p2_(k-2)   (two shuffles)
  A   H_j = limb_(9+j), j < 10;  lane k < 9: col'_k = limb_k + sum_j H_j * K_j[k],  K_j = 2^(261+29j) mod p
  N2  as N1 on col'_0..8 -> limb'_0..9 (limb'_9 = h < 2^28)
  B   2^261 = -c' (mod p), c' = 2^7 (p - 2^254) < 2^133:  col''_k = limb'_k + Z_k - h * c'_k (+ addend_k)
      Z = 64 p + telescoping offsets that keep every column non-negative
  N3  limb''_k = p0_k + p1_(k-1) (one shuffle), limb 8 keeps its carry
"""
import random
import sys

FQ = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
FP = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
M29 = (1 << 29) - 1
U64 = 1 << 64
NL = 9


def limbs_of(x, n=NL):
    """canonical 29-bit limbs (top limb takes the rest)"""
    out = [(x >> (29 * k)) & M29 for k in range(n - 1)]
    out.append(x >> (29 * (n - 1)))
    return out


def value_of(l):
    return sum(int(v) << (29 * k) for k, v in enumerate(l))


class Field:
    def __init__(self, p):
        self.p = p
        self.c = p - (1 << 254)
        assert 0 < self.c < (1 << 126) and p % (1 << 29) == 1
        self.cp = limbs_of(self.c << 7, 5)                         # c' = 2^7 c  (2^261 = -c' mod p)
        assert (self.c << 7) < (1 << 133) and self.cp[4] < (1 << 17)
        self.K = [limbs_of(pow(2, 261 + 29 * j, p)) for j in range(12)]
        # fold-B offsets: 64 p + (2^57 at column 0, 2^57 - 2^28 at columns 1..7, -2^28 at column 8) == 0 mod p
        p64 = limbs_of(64 * p)
        tel = [1 << 57] + [(1 << 57) - (1 << 28)] * 7 + [-(1 << 28)]
        assert sum(t << (29 * k) for k, t in enumerate(tel)) == 0
        self.Z = [p64[k] + tel[k] for k in range(9)]
        assert all(z >= 0 for z in self.Z) and value_of(self.Z) == 64 * p


def split3(col):
    assert 0 <= col < U64, col.bit_length()
    return col & M29, (col >> 29) & M29, col >> 58


def lp_multi(F, terms, addend=None, tight=False, stats=None):
    """terms: list of (a_limbs, b_limbs), 10 limbs each (limb 9 = the part above 2^261, < 2^28.7);
    returns the 10 lazy limbs of sum a*b (+ addend) mod p after fold A only (value < 2^289).
    addend: 9 or 10 non-negative column addends (< 2^32) merged into the normalisation.
    tight: one more carry step (limbs 0..8 <= 2^29 + 2), used before the 50-term MDS products."""
    # P: 19 columns
    col = [0] * 22
    for a, b in terms:
        a, b = list(a) + [0] * (10 - len(a)), list(b) + [0] * (10 - len(b))
        for k in range(19):
            for i in range(10):
                j = k - i
                if 0 <= j < 10:
                    col[k] += a[i] * b[j]
    assert all(c < U64 for c in col), max(c.bit_length() for c in col)
    # N1: limbs 0..20
    pc = [split3(c) for c in col]
    limb = [pc[k][0] + (pc[k - 1][1] if k >= 1 else 0) + (pc[k - 2][2] if k >= 2 else 0) for k in range(22)]
    assert limb[21] == 0
    # fold A: H_j = limb_(9+j), j < 12
    H = limb[9:21]
    ad = (list(addend) + [0] * 10)[:10] if addend else [0] * 10
    assert ad[9] == 0 or True
    colA = [limb[k] + sum(H[j] * F.K[j][k] for j in range(12)) + ad[k] for k in range(9)]
    assert all(c < U64 for c in colA), max(c.bit_length() for c in colA)
    # N2: limbs 0..9 (limb 9 = h)
    pa = [split3(c) for c in colA]
    out = [pa[k][0] + (pa[k - 1][1] if k >= 1 else 0) + (pa[k - 2][2] if k >= 2 else 0) for k in range(9)]
    out.append(pa[8][1] + pa[7][2] + ad[9])
    assert pa[8][2] == 0
    if tight:
        o2 = [(out[k] & M29) + ((out[k - 1] >> 29) if k >= 1 else 0) for k in range(9)]
        o2.append(out[9] + (out[8] >> 29))
        out = o2
        assert all(v <= (1 << 29) + 2 for v in out[:9])
    assert all(v < (1 << 30) + (1 << 7) for v in out[:9]) and out[9] < (1 << 28), [v.bit_length() for v in out]
    if stats is not None:
        stats["max_low"] = max(stats.get("max_low", 0), max(out[:9]))
        stats["max_top"] = max(stats.get("max_top", 0), out[9])
        stats["max_col"] = max(stats.get("max_col", 0), max(col), max(colA))
    want = (sum(value_of(a) * value_of(b) for a, b in terms) + (value_of(ad))) % F.p
    assert value_of(out) % F.p == want
    return out


def lp_fold_b(F, l10, stats=None):
    """10 lazy limbs -> 9 lazy limbs (value < 2^262): 2^261 h == -c' h, offsets Z keep columns >= 0."""
    h = l10[9]
    assert h < (1 << 28)
    colB = []
    for k in range(9):
        v = l10[k] + F.Z[k] - (h * F.cp[k] if k < 5 else 0)
        assert 0 <= v < (1 << 59), (k, v)
        colB.append(v)
    out = [(colB[k] & M29) + ((colB[k - 1] >> 29) if k >= 1 else 0) for k in range(8)]
    out.append(colB[8] + (colB[7] >> 29))
    assert all(v < (1 << 30) + (1 << 8) for v in out[:8]) and out[8] < (1 << 31)
    assert value_of(out) % F.p == value_of(l10) % F.p
    return out


def canon(F, l):
    return value_of(l) % F.p


# ---------------------------------------------------------------------------------------------
# permutation schedule (Gamma formulation of the rescaled partial rounds)
# ---------------------------------------------------------------------------------------------
def derive_lp_tables():
    """From tools/gen_poseidon_consts.derive():
         c_r  = w_(r+1) - u_r = C0_r + sum_(t<r) Gamma[r][t] u_t,   C0_r = sum_i beta[r][i] s_i(0) + kp[r+1]
         Gamma[r][t] = sum_i beta[r][i] D[t][i]
         state lanes 1..4 after the partial rounds AND the dense 4x4 block:
         y_j = sum_i post[j][i] s_i(0) + sum_t PD[j][t] u_t,        PD[j][t] = sum_i post[j][i] D[t][i]"""
    sys.path.insert(0, __file__.rsplit("/", 1)[0])
    import gen_poseidon_consts as G
    K = G.derive()
    P = FQ
    RP = 56
    Gam = [[sum(K["beta"][r][i] * K["D"][t][i] for i in range(4)) % P if t < r else 0 for t in range(RP)] for r in range(RP)]
    PD = [[sum(K["post"][j][i] * K["D"][t][i] for i in range(4)) % P for t in range(RP)] for j in range(4)]
    return K, Gam, PD


def permute_gamma(state, K, Gam, PD):
    """Plain-integer evaluation of the Gamma schedule (what the kernel computes, without the lane detail)."""
    P = FQ
    s = list(state)

    def full(s, r):
        s = [pow((x + K["rc_full"][r][i]) % P, 5, P) for i, x in enumerate(s)]
        return [sum(K["mds"][j][i] * s[i] for i in range(5)) % P for j in range(5)]

    for r in range(4):
        s = full(s, r)
    s0 = s[1:]
    C0 = [(sum(K["beta"][r][i] * s0[i] for i in range(4)) + K["kp"][r + 1]) % P for r in range(56)]
    y = [sum(K["post"][j][i] * s0[i] for i in range(4)) % P for j in range(4)]
    w = (s[0] + K["kp"][0]) % P
    us = []
    for r in range(56):
        u = pow(w, 5, P)
        c = (C0[r] + sum(Gam[r][t] * us[t] for t in range(r))) % P
        us.append(u)
        w = (u + c) % P
    y = [(y[j] + sum(PD[j][t] * us[t] for t in range(56))) % P for j in range(4)]
    s = [K["lam_end"] * w % P] + y
    for r in range(4, 8):
        s = full(s, r)
    return s


def permute_lp(F, state, K, Gam, PD, stats=None):
    """The same schedule with every multiplication of the critical path done by lp_multi on lazy limbs
    (full rounds: 3 LP squarings/multiplications + one 5-term LP multi-product per state element)."""
    L = limbs_of
    x = [L((state[i] + K["rc_full"][0][i]) % F.p) for i in range(5)]          # entry: state + first constants

    def sbox(v, tight):
        v2 = lp_multi(F, [(v, v)], stats=stats)
        v4 = lp_multi(F, [(v2, v2)], stats=stats)
        return lp_multi(F, [(v4, v)], tight=tight, stats=stats)

    def full(x, add_next):
        x5 = [sbox(v, True) for v in x]
        return [lp_multi(F, [(x5[i], L(K["mds"][j][i])) for i in range(5)], addend=L(add_next[j]) if add_next else None, stats=stats)
                for j in range(5)]

    for r in range(4):
        nxt = K["rc_full"][r + 1] if r < 3 else [K["kp"][0], 0, 0, 0, 0]
        x = full(x, nxt)
    w = x[0]
    s0 = [canon(F, v) for v in x[1:]]
    C0 = [(sum(K["beta"][r][i] * s0[i] for i in range(4)) + K["kp"][r + 1]) % F.p for r in range(56)]     # helper warps (mul29)
    y = [(sum(K["post"][j][i] * s0[i] for i in range(4)) + K["rc_full"][4][j + 1]) % F.p for j in range(4)]
    us = []
    for r in range(56):
        w2 = lp_multi(F, [(w, w)], stats=stats)
        w4 = lp_multi(F, [(w2, w2)], stats=stats)
        # c_r: accumulator lanes (terms t <= r-2, mul29) + warp B (term t = r-1, LP); delivered normalised
        c = (C0[r] + sum(Gam[r][t] * canon(F, us[t]) for t in range(r))) % F.p
        u = lp_multi(F, [(w4, w)], stats=stats)                 # published to the side warps as is (10 limbs)
        w = lp_multi(F, [(w4, w)], addend=L(c), stats=stats)    # the chain continues with u + c_r
        lp_fold_b(F, u)                                         # what an accumulator lane does before its mul29
        us.append(u)
    # end: x_0 = lam_end * w_56 + rc, x_j = y_j + PD[j][55] u_55 (+ rc, already in y)
    ycan = [(y[j] + sum(PD[j][t] * canon(F, us[t]) for t in range(55))) % F.p for j in range(4)]
    x = [lp_multi(F, [(w, L(K["lam_end"]))], addend=L(K["rc_full"][4][0]), stats=stats)]
    x += [lp_multi(F, [(us[55], L(PD[j][55]))], addend=L(ycan[j]), stats=stats) for j in range(4)]
    for r in range(4, 8):
        x = full(x, K["rc_full"][r + 1] if r < 7 else None)
    return [canon(F, v) for v in x]


def selftest(n=20):
    F = Field(FQ)
    rnd = random.Random(3)
    st = {}
    for _ in range(3000):
        a = [rnd.randrange(1 << 29) + rnd.randrange(1 << 29) for _ in range(9)] + [rnd.randrange(1 << 28)]
        b = [rnd.randrange(1 << 29) + rnd.randrange(1 << 29) for _ in range(9)] + [rnd.randrange(1 << 28)]
        lp_fold_b(F, lp_multi(F, [(a, b)], stats=st))
    # worst-case magnitudes of the lazy form
    top = [(1 << 30) + 127] * 9 + [(1 << 28) - 1]
    lp_multi(F, [(top, top)], addend=[(1 << 32) - 1] * 9, stats=st)
    lp_fold_b(F, top)
    lp_multi(F, [([(1 << 29) + 2] * 9 + [(1 << 28) - 1], limbs_of(FQ - 1))] * 5, addend=[(1 << 32) - 1] * 9, stats=st)
    K, Gam, PD = derive_lp_tables()
    import gen_poseidon_consts as G
    for _ in range(n):
        s = [rnd.randrange(FQ) for _ in range(5)]
        ref = G.permute_textbook(s, K)
        assert permute_gamma(s, K, Gam, PD) == ref
    for _ in range(2):
        s = [rnd.randrange(FQ) for _ in range(5)]
        assert permute_lp(F, s, K, Gam, PD, st) == G.permute_textbook(s, K)
    print("lp_model selftest OK; lazy bounds seen: limbs 0..8 < 2^%.2f, limb 9 < 2^%.2f, column < 2^%.2f"
          % tuple(__import__("math").log2(st[k]) for k in ("max_low", "max_top", "max_col")))


def emit(path):
    """lp_consts.inc: fold constants of both Pasta fields (Poseidon tables of the LP permutation are appended by
    emit_poseidon below when the permutation kernels are generated)."""
    out = ["// GENERATED by tools/lp_model.py --emit -- do not edit.",
           "// Lane-parallel multiplication constants (lpmul.cuh): K_j = 2^(261+29j) mod p as 9 x 29-bit limbs,",
           "// c' = 2^7 (p - 2^254) as 5 limbs, Z = 64 p + telescoping column offsets (u64 per column).",
           "#if defined(__CUDACC__)", "#define LP_CONST static __device__ __constant__", "#else", "#define LP_CONST static const", "#endif"]
    for name, p in (("FQ", FQ), ("FP", FP)):
        F = Field(p)
        out.append("LP_CONST uint32_t LP_K_%s[12][9] = {" % name)
        for j in range(12):
            out.append("  {" + ",".join("0x%08xu" % v for v in F.K[j]) + "},")
        out.append("};")
        out.append("LP_CONST uint32_t LP_CP_%s[5] = {" % name + ",".join("0x%08xu" % v for v in F.cp) + "};")
        out.append("LP_CONST unsigned long long LP_Z_%s[9] = {" % name + ",".join("0x%016xull" % v for v in F.Z) + "};")
        sel = lambda vals, fmt: " : ".join(["k == %d ? %s" % (i, fmt % v) for i, v in enumerate(vals[:-1])] + [fmt % vals[-1]])
        out.append("REEF_HD constexpr uint32_t lp_cp_%s(int k) { return %s; }" % (name.lower(), sel(F.cp + [0], "0x%08xu")))
        out.append("REEF_HD constexpr unsigned long long lp_z_%s(int k) { return %s; }" % (name.lower(), sel(F.Z, "0x%016xull")))
    out += emit_poseidon()
    out.append("#undef LP_CONST")
    open(path, "w").write("\n".join(out) + "\n")


def emit_poseidon():
    return []


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--emit":
        emit(sys.argv[2])
    else:
        selftest()
