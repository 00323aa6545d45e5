#!/usr/bin/env python3
"""Generate reef_b200/csrc/poseidon_consts.inc  (Poseidon t=5 over Fq, neptune-compatible).

Build-time generator for the PRODUCT constants; deliberately standalone (does not import
oracle/).  tests/test_poseidon_consts.py cross-checks the emitted tables against the
oracle's independent derivation and against the textbook permutation.

Emitted (all canonical integers, 8 x u32 little-endian limbs):
  RC_FULL[8][5]      round constants of the 8 full rounds (the first full round after the
                     partial rounds has the pushed-forward partial-round leftovers folded in)
  RC_PART[56]        lane-0 constant of each partial round
  MDS[5][5]          Cauchy matrix 1/(i + j + 5)   (symmetric)
  SP_ROW[56][5]      sparse matrix first row   (a, b_1..b_4):  new0 = a*z0 + sum b_i z_i
  SP_COL[56][4]      sparse matrix first column (d_1..d_4):    new_i = d_i*z0 + z_i
  POST[4][4]         dense block applied to lanes 1..4 after the last partial round

Derivation (column-vector convention; M symmetric so it equals neptune's row-vector one):
  textbook partial round:   y_{r+1} = M S(y_r) + c_{r+1},  S = x^5 on lane 0 only.
  Write y_r = B_r z_r + g_r with B_r = diag(1, Bh_r), (g_r)_0 = 0.  Then S(y_r) = B_r S(z_r) + g_r and
      M B_r = B_{r+1} Sp_r        (Sp_r sparse: first row/column + identity)
      M g_r + c_{r+1} = k_{r+1} e_0 + g_{r+1}
  so z_{r+1} = Sp_r S(z_r) + k_{r+1} e_0, starting from B_0 = I, z_0 = s + k_0 e_0, g_0 = c_0 - k_0 e_0,
  and after the last partial round  s = B_RP z + M g_{RP-1}  (the vector is folded into the next
  full round's constants).
"""
import sys

FQ = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
T, RF, RP = 5, 8, 56
P = FQ


def grain_constants():
    bits = []
    for width, val in ((2, 1), (4, 1), (12, 255), (12, T), (10, RF), (10, RP), (30, (1 << 30) - 1)):
        bits += [(val >> (width - 1 - i)) & 1 for i in range(width)]
    s = bits

    def clock():
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    for _ in range(160):
        clock()

    def bit():
        b = clock()
        while b == 0:
            clock()
            b = clock()
        return clock()

    rc = []
    while len(rc) < (RF + RP) * T:
        x = 0
        for _ in range(255):
            x = (x << 1) | bit()
        if x < P:
            rc.append(x)
    return rc


def matmul(A, B):
    return [[sum(A[i][k] * B[k][j] for k in range(len(B))) % P for j in range(len(B[0]))] for i in range(len(A))]


def matvec(A, v):
    return [sum(A[i][k] * v[k] for k in range(len(v))) % P for i in range(len(A))]


def inverse(A):
    n = len(A)
    M = [list(r) + [int(i == j) for j in range(n)] for i, r in enumerate(A)]
    for c in range(n):
        piv = next(r for r in range(c, n) if M[r][c] % P)
        M[c], M[piv] = M[piv], M[c]
        inv = pow(M[c][c], -1, P)
        M[c] = [x * inv % P for x in M[c]]
        for r in range(n):
            if r != c and M[r][c]:
                f = M[r][c]
                M[r] = [(x - f * y) % P for x, y in zip(M[r], M[c])]
    return [r[n:] for r in M]


def derive():
    rc = grain_constants()
    mds = [[pow(i + j + T, -1, P) for j in range(T)] for i in range(T)]
    c = [rc[r * T:(r + 1) * T] for r in range(RF + RP)]
    half = RF // 2
    cp = c[half:half + RP]                      # partial-round constants
    # constants: forward push
    k = [cp[0][0]]
    g = [0] + cp[0][1:]
    for r in range(1, RP):
        v = [(a + b) % P for a, b in zip(matvec(mds, g), cp[r])]
        k.append(v[0])
        g = [0] + v[1:]
    tail = matvec(mds, g)                       # folded into the next full round's constants
    # matrices: forward factorisation  M B_r = B_{r+1} Sp_r
    B = [[int(i == j) for j in range(T)] for i in range(T)]
    sp_row, sp_col = [], []
    for r in range(RP):
        N = matmul(mds, B)
        Nh = [row[1:] for row in N[1:]]
        w = [row[0] for row in N[1:]]
        d = matvec(inverse(Nh), w)
        sp_row.append([N[0][0]] + N[0][1:])
        sp_col.append(d)
        B = [[1] + [0] * (T - 1)] + [[0] + row for row in Nh]
    post = [row[1:] for row in B[1:]]
    rc_full = [list(c[r]) for r in range(half)] + [list(c[half + RP + r]) for r in range(half)]
    rc_full[half] = [(a + b) % P for a, b in zip(rc_full[half], tail)]
    # Rescaled lane 0 (y_r = s0 / lam_r, lam_0 = 1, lam_{r+1} = a_r lam_r^5): the multiplication by the
    # sparse matrix corner a_r leaves the critical path, which becomes w -> w^5 only:
    #     u_r = w_r^5,  w_{r+1} = u_r + sum_i beta[r][i] s_i(r) + kp[r+1],  s_i(r+1) = s_i(r) + D[r][i] u_r
    # and after the last round  s0 = lam_end * (u_55 + sum_i beta[55][i] s_i(55)).
    lam = [1]
    for r in range(RP):
        lam.append(sp_row[r][0] * pow(lam[r], 5, P) % P)
    kp = [k[r] * pow(lam[r], -1, P) % P for r in range(RP)] + [0]
    beta = [[sp_row[r][i] * pow(lam[r + 1], -1, P) % P for i in range(1, T)] for r in range(RP)]
    Dm = [[sp_col[r][i] * pow(lam[r], 5, P) % P for i in range(T - 1)] for r in range(RP)]
    return dict(rc=rc, mds=mds, rc_full=rc_full, rc_part=k, sp_row=sp_row, sp_col=sp_col, post=post,
                kp=kp, beta=beta, D=Dm, lam_end=lam[RP])


def permute_optimized(state, K):
    """Reference evaluation of the emitted tables (used by the self-check and the tests)."""
    s = list(state)
    half = RF // 2

    def full(s, r):
        s = [pow((x + K["rc_full"][r][i]) % P, 5, P) for i, x in enumerate(s)]
        return matvec(K["mds"], s)

    for r in range(half):
        s = full(s, r)
    for r in range(RP):
        z0 = pow((s[0] + K["rc_part"][r]) % P, 5, P)
        row, col = K["sp_row"][r], K["sp_col"][r]
        n0 = (row[0] * z0 + sum(row[i] * s[i] for i in range(1, T))) % P
        s = [n0] + [(col[i - 1] * z0 + s[i]) % P for i in range(1, T)]
    s = [s[0]] + matvec(K["post"], s[1:])
    for r in range(half, RF):
        s = full(s, r)
    return s


def permute_rescaled(state, K):
    """Evaluation through the rescaled partial-round recurrence (what the GPU runs)."""
    s = list(state)
    half = RF // 2

    def full(s, r):
        s = [pow((x + K["rc_full"][r][i]) % P, 5, P) for i, x in enumerate(s)]
        return matvec(K["mds"], s)

    for r in range(half):
        s = full(s, r)
    w = (s[0] + K["kp"][0]) % P
    rest = s[1:]
    for r in range(RP):
        u = pow(w, 5, P)
        c = sum(K["beta"][r][i] * rest[i] for i in range(T - 1)) % P
        w = (u + c + K["kp"][r + 1]) % P
        rest = [(rest[i] + K["D"][r][i] * u) % P for i in range(T - 1)]
    s = [K["lam_end"] * w % P] + matvec(K["post"], rest)
    for r in range(half, RF):
        s = full(s, r)
    return s


def permute_textbook(state, K):
    s = list(state)
    rc, mds = K["rc"], K["mds"]
    half = RF // 2
    for r in range(RF + RP):
        s = [(x + rc[r * T + i]) % P for i, x in enumerate(s)]
        if r < half or r >= half + RP:
            s = [pow(x, 5, P) for x in s]
        else:
            s[0] = pow(s[0], 5, P)
        s = matvec(mds, s)
    return s


def limbs(x):
    return "{" + ",".join("0x%08xu" % ((x >> (32 * i)) & 0xFFFFFFFF) for i in range(8)) + "}"


def emit(K, f):
    w = lambda s="": print(s, file=f)
    w("// GENERATED by tools/gen_poseidon_consts.py -- do not edit.")
    w("// Poseidon over Fq (Pallas scalar field), width 5 (arity 4), R_F = 8, R_P = 56, x^5,")
    w("// neptune 8.1.0 `Strength::Standard` parameters.  Canonical integers, u32 LE limbs.")
    w("#define REEF_POSEIDON_T 5")
    w("#define REEF_POSEIDON_RF 8")
    w("#define REEF_POSEIDON_RP 56")

    def table(name, rows):
        flat = [x for row in rows for x in (row if isinstance(row, (list, tuple)) else [row])]
        w("static const uint32_t %s[%d][8] = {" % (name, len(flat)))
        for x in flat:
            w("  %s," % limbs(x))
        w("};")

    table("REEF_POSEIDON_RC_FULL", K["rc_full"])
    table("REEF_POSEIDON_RC_PART", K["rc_part"])
    table("REEF_POSEIDON_MDS", K["mds"])
    table("REEF_POSEIDON_SP_ROW", K["sp_row"])
    table("REEF_POSEIDON_SP_COL", K["sp_col"])
    table("REEF_POSEIDON_POST", K["post"])
    table("REEF_POSEIDON_KP", K["kp"])
    table("REEF_POSEIDON_BETA", K["beta"])
    table("REEF_POSEIDON_D", K["D"])
    table("REEF_POSEIDON_LAM_END", [K["lam_end"]])


if __name__ == "__main__":
    import random
    K = derive()
    rnd = random.Random(7)
    for _ in range(5):
        st = [rnd.randrange(P) for _ in range(T)]
        assert permute_optimized(st, K) == permute_textbook(st, K), "optimised form != textbook form"
        assert permute_rescaled(st, K) == permute_textbook(st, K), "rescaled form != textbook form"
    out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
    emit(K, out)
