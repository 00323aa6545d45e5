"""Composed provers and verifiers behind `CompressedSNARK::prove` and `prove_consistency` (oracle; test
infrastructure only -- never imported by reef_b200/).

PARITY UNPINNED, like oracle/spartan.py: nova-snark (git dependency on github.com/sga001/Nova with no pinned
revision, /root/reference/Cargo.toml:12) is not under /root/reference, so what is restated is the PUBLISHED
upstream construction, anchored on the in-tree call sites:
  * RelaxedR1CSSNARK::prove / verify (framework.rs:695-698, 818; S = spartan::RelaxedR1CSSNARK<G,
    ipa_pc::EvaluationEngine<G>>, framework.rs:5-8; also `cap_prove` / `cap_verify`, commitment.rs:261-268, 472):
    outer cubic sum-check of eq(tau,x) (Az Bz - u Cz - E)(x), inner quadratic sum-check of
    (A + r B + r^2 C)(r_x, y) z(y), openings of W and E by inner-product arguments.
  * InnerProductArgument::prove / verify (commitment.rs:371-393 through hyrax_pc; ipa_pc.rs upstream).
  * HyraxPC::prove_eval / verify_eval: LZ = L^T M, the row commitments combined with L, one IPA of length 2^right.
The Fiat-Shamir TRANSCRIPT of nova-snark lives in the same un-vendored crate: `Transcript` below is a
stand-in (SHA-256) with the same absorb / squeeze call pattern, so that prover and verifier agree here; the
GPU path (reef_b200/snark.py) leaves the transcript to the caller for exactly that reason.
What these functions DO pin is the algebra: a proof produced by the GPU driver must be accepted by the
verifier below, and must equal, byte for byte, the proof of the prover below under the same transcript.
"""
from __future__ import annotations

import hashlib

from . import spartan as S
from .curves import INF


class Transcript:
    """stand-in for nova-snark's transcript: absorb(label, bytes), squeeze(label) -> field element"""

    def __init__(self, label: bytes, modulus: int):
        self.state = hashlib.sha256(b"reef-b200-oracle-transcript" + label).digest()
        self.p = modulus
        self.round = 0

    def absorb(self, label: bytes, data: bytes):
        self.state = hashlib.sha256(self.state + len(label).to_bytes(4, "little") + label + len(data).to_bytes(8, "little") + data).digest()

    def absorb_scalars(self, label: bytes, xs):
        self.absorb(label, b"".join(int(x).to_bytes(32, "little") for x in xs))

    def absorb_point(self, label: bytes, P):
        self.absorb(label, bytes(64) if P is None else int(P[0]).to_bytes(32, "little") + int(P[1]).to_bytes(32, "little"))

    def squeeze(self, label: bytes) -> int:
        self.round += 1
        d0 = hashlib.sha256(self.state + b"\x00" + label + self.round.to_bytes(4, "little")).digest()
        d1 = hashlib.sha256(self.state + b"\x01" + label + self.round.to_bytes(4, "little")).digest()
        self.state = d0
        return int.from_bytes(d0 + d1, "little") % self.p


def eq_table(r, p):
    """eq(r, x) for x in {0,1}^k, r[0] <-> top index bit (the order bound_poly_var_top consumes)"""
    t = [1]
    for ri in r:
        t = [e * (1 - ri) % p for e in t] + [e * ri % p for e in t]
        # index = (previous index) + 2^j * bit: keep MSB-first by interleaving instead
    # the loop above makes r[0] the LOWEST bit; rebuild MSB-first
    k = len(r)
    out = [0] * (1 << k)
    for i, v in enumerate(t):
        j = int(format(i, "0%db" % k)[::-1], 2) if k else 0
        out[j] = v
    return out


def mle_eval(table, r, p):
    """value of the multilinear extension at r (r[0] binds the top variable)"""
    t = list(table)
    for ri in r:
        t = S.bound_top(t, ri, p)
    return t[0]


# ------------------------------------------------------------------------------------ inner-product argument
def inner(a, b, p):
    return sum(x * y for x, y in zip(a, b)) % p


def _weights(n, rs, p):
    """w[k] = prod_i (r_i if bit_i(k) else r_i^-1), bit_i = i-th index bit FROM THE TOP: the folded generator
    G'_j after len(rs) rounds of  G' = G_lo r^-1 + G_hi r  is  sum_{k = j mod m} w[k] G_k,  m = n / 2^len(rs)."""
    w = [1] * n
    k_bits = n.bit_length() - 1
    for i, r in enumerate(rs):
        ri = pow(r, -1, p)
        sh = k_bits - 1 - i
        for k in range(n):
            w[k] = w[k] * (r if (k >> sh) & 1 else ri) % p
    return w


def ipa_prove(curve, gens, gen_c, a, b, tr: Transcript, msm=None):
    """Returns (L_vec, R_vec, a_hat).  Convention in include/reef_b200.h (reef_ipa_*).
    msm(scalars, points) -> point: with it the generators are never folded explicitly (every L / R is one MSM over
    the ORIGINAL generators with the fold weights multiplied into the scalars) -- same points, C-speed oracle."""
    p = curve.order
    a, b, G = list(a), list(b), list(gens)
    n = len(a)
    Ls, Rs, rs = [], [], []
    while len(a) > 1:
        h = len(a) // 2
        cL, cR = inner(a[:h], b[h:], p), inner(a[h:], b[:h], p)
        if msm is None:
            L = curve.add(curve.msm(a[:h], G[h:]), curve.mul(cL, gen_c))
            R = curve.add(curve.msm(a[h:], G[:h]), curve.mul(cR, gen_c))
        else:
            w, m = _weights(n, rs, p), len(a)
            xl = [0] * h + a[:h]                 # pairs a_lo with the upper half of the folded generators
            xr = a[h:] + [0] * h
            L = curve.add(msm([xl[k % m] * w[k] % p for k in range(n)], gens), curve.mul(cL, gen_c))
            R = curve.add(msm([xr[k % m] * w[k] % p for k in range(n)], gens), curve.mul(cR, gen_c))
        tr.absorb_point(b"L", L)
        tr.absorb_point(b"R", R)
        r = tr.squeeze(b"r")
        ri = pow(r, -1, p)
        a = [(a[i] * r + a[h + i] * ri) % p for i in range(h)]
        b = [(b[i] * ri + b[h + i] * r) % p for i in range(h)]
        if msm is None:
            G = S.ipa_fold_bases(curve, G, ri, r)
        rs.append(r)
        Ls.append(L)
        Rs.append(R)
    return Ls, Rs, a[0]


def ipa_verify(curve, gens, gen_c, comm_a, b, c, proof, tr: Transcript, msm=None) -> bool:
    p = curve.order
    Ls, Rs, a_hat = proof
    P = curve.add(comm_a, curve.mul(c, gen_c))
    b, G = list(b), list(gens)
    n = len(G)
    if len(Ls) != (n.bit_length() - 1):
        return False
    rs = []
    for L, R in zip(Ls, Rs):
        h = len(b) // 2
        tr.absorb_point(b"L", L)
        tr.absorb_point(b"R", R)
        r = tr.squeeze(b"r")
        ri = pow(r, -1, p)
        P = curve.add(curve.add(curve.mul(r * r % p, L), P), curve.mul(ri * ri % p, R))
        b = [(b[i] * ri + b[h + i] * r) % p for i in range(h)]
        if msm is None:
            G = S.ipa_fold_bases(curve, G, ri, r)
        rs.append(r)
    G_hat = G[0] if msm is None else msm(_weights(n, rs, p), gens)
    return P == curve.add(curve.mul(a_hat, G_hat), curve.mul(a_hat * b[0] % p, gen_c))


# ------------------------------------------------------------------------------------ Hyrax prove_eval
def hyrax_commit(curve, gens, matrix, rows, cols):
    return [curve.msm(matrix[r * cols:(r + 1) * cols], gens[:cols]) for r in range(rows)]


def hyrax_prove_eval(curve, gens, gen_c, matrix, rows, cols, q, tr: Transcript, msm=None):
    """matrix: rows x cols row-major evaluations; q: log2(rows) + log2(cols) point (q[0] binds the top variable).
    Returns (value, (L_vec, R_vec, a_hat))."""
    p = curve.order
    kl = rows.bit_length() - 1
    Lv, Rv = eq_table(q[:kl], p), eq_table(q[kl:], p)
    LZ = [sum(Lv[i] * matrix[i * cols + j] for i in range(rows)) % p for j in range(cols)]
    v = inner(LZ, Rv, p)
    tr.absorb_scalars(b"v", [v])
    return v, ipa_prove(curve, gens[:cols], gen_c, LZ, Rv, tr, msm)


def hyrax_verify_eval(curve, gens, gen_c, comms, rows, cols, q, v, proof, tr: Transcript, msm=None) -> bool:
    p = curve.order
    kl = rows.bit_length() - 1
    Lv, Rv = eq_table(q[:kl], p), eq_table(q[kl:], p)
    comm_LZ = (msm or (lambda sc, pts: curve.msm(sc, pts)))(Lv, comms)
    tr.absorb_scalars(b"v", [v])
    return ipa_verify(curve, gens[:cols], gen_c, comm_LZ, Rv, v, proof, tr, msm)


# ------------------------------------------------------------------------------------ relaxed R1CS SNARK (Spartan)
class R1CSShape:
    """num_cons, num_vars powers of two; z = W (num_vars) ++ [u] ++ X ++ zeros (2 * num_vars entries).
    A, B, C: lists of (row, col, value)."""

    def __init__(self, num_cons, num_vars, num_io, A, B, C):
        assert num_cons & (num_cons - 1) == 0 and num_vars & (num_vars - 1) == 0 and num_io + 1 <= num_vars
        self.num_cons, self.num_vars, self.num_io = num_cons, num_vars, num_io
        self.A, self.B, self.C = A, B, C

    def csr(self, M, transpose=False):
        n = 2 * self.num_vars if transpose else self.num_cons
        rows = [[] for _ in range(n)]
        for r, c, v in M:
            (rows[c] if transpose else rows[r]).append((r if transpose else c, v))
        ptr, idx, val = [0], [], []
        for row in rows:
            for c, v in row:
                idx.append(c)
                val.append(v)
            ptr.append(len(idx))
        return ptr, idx, val

    def z(self, W, u, X):
        return list(W) + [u] + list(X) + [0] * (self.num_vars - 1 - len(X))

    def mul(self, M, z, p):
        out = [0] * self.num_cons
        for r, c, v in M:
            out[r] = (out[r] + v * z[c]) % p
        return out

    def is_sat(self, W, E, u, X, p):
        z = self.z(W, u, X)
        az, bz, cz = (self.mul(M, z, p) for M in (self.A, self.B, self.C))
        return all((a * b - u * c - e) % p == 0 for a, b, c, e in zip(az, bz, cz, E))


def _sumcheck_prove(tables, tr, p, label):
    """drives oracle.spartan round functions with transcript-derived challenges; returns (polys, r, finals, claim)"""
    kind = len(tables)
    tabs = [list(t) for t in tables]
    claim = sum(a * b for a, b in zip(*tabs)) % p if kind == 2 else sum(S.comb_cubic(*x) for x in zip(*tabs)) % p
    polys, rs = [], []
    while len(tabs[0]) > 1:
        ev = S.round_quad(*tabs, p) if kind == 2 else S.round_cubic(*tabs, p)
        full = [ev[0], (claim - ev[0]) % p] + list(ev[1:])
        tr.absorb_scalars(label, full)
        r = tr.squeeze(label)
        claim = S.interpolate_eval(full, r, p)
        tabs = [S.bound_top(t, r, p) for t in tabs]
        polys.append(full)
        rs.append(r)
    return polys, rs, [t[0] for t in tabs], claim


def _sumcheck_verify(claim, polys, tr, p, label, degree):
    rs = []
    for full in polys:
        if len(full) != degree + 1 or (full[0] + full[1]) % p != claim:
            return None, None
        tr.absorb_scalars(label, full)
        r = tr.squeeze(label)
        claim = S.interpolate_eval(full, r, p)
        rs.append(r)
    return claim, rs


def snark_prove(curve, shape: R1CSShape, gens, gen_c, comm_W, comm_E, W, E, u, X, tr: Transcript, msm=None):
    """RelaxedR1CSSNARK::prove.  gens: >= max(num_vars, num_cons) generators of `curve`."""
    p = curve.order
    z = shape.z(W, u, X)
    Az, Bz, Cz = (shape.mul(M, z, p) for M in (shape.A, shape.B, shape.C))
    tr.absorb_point(b"W", comm_W)
    tr.absorb_point(b"E", comm_E)
    tr.absorb_scalars(b"uX", [u] + list(X))
    k_x, k_y = shape.num_cons.bit_length() - 1, (2 * shape.num_vars).bit_length() - 1
    tau = [tr.squeeze(b"tau") for _ in range(k_x)]
    uCzE = [(u * c + e) % p for c, e in zip(Cz, E)]
    polys_o, r_x, fin_o, claim_o = _sumcheck_prove([eq_table(tau, p), Az, Bz, uCzE], tr, p, b"outer")
    claim_Az, claim_Bz = fin_o[1], fin_o[2]
    claim_Cz, claim_E = mle_eval(Cz, r_x, p), mle_eval(E, r_x, p)
    tr.absorb_scalars(b"claims", [claim_Az, claim_Bz, claim_Cz, claim_E])
    r = tr.squeeze(b"r")
    ex = eq_table(r_x, p)
    ABC = [0] * (2 * shape.num_vars)
    for M, coef in ((shape.A, 1), (shape.B, r), (shape.C, r * r % p)):
        for row, col, v in M:
            ABC[col] = (ABC[col] + coef * ex[row] % p * v) % p
    polys_i, r_y, fin_i, claim_i = _sumcheck_prove([ABC, z], tr, p, b"inner")
    eval_W = mle_eval(W, r_y[1:], p)
    tr.absorb_scalars(b"evals", [eval_W, claim_E])
    ipa_W = ipa_prove(curve, gens[:shape.num_vars], gen_c, W, eq_table(r_y[1:], p), tr, msm)
    ipa_E = ipa_prove(curve, gens[:shape.num_cons], gen_c, E, ex, tr, msm)
    return {"outer": polys_o, "claims": (claim_Az, claim_Bz, claim_Cz, claim_E), "inner": polys_i, "eval_W": eval_W,
            "ipa_W": ipa_W, "ipa_E": ipa_E}


def snark_verify(curve, shape: R1CSShape, gens, gen_c, comm_W, comm_E, u, X, proof, tr: Transcript, msm=None) -> bool:
    p = curve.order
    tr.absorb_point(b"W", comm_W)
    tr.absorb_point(b"E", comm_E)
    tr.absorb_scalars(b"uX", [u] + list(X))
    k_x = shape.num_cons.bit_length() - 1
    tau = [tr.squeeze(b"tau") for _ in range(k_x)]
    claim_o, r_x = _sumcheck_verify(0, proof["outer"], tr, p, b"outer", 3)
    if claim_o is None or len(r_x) != k_x:
        return False
    cAz, cBz, cCz, cE = proof["claims"]
    eq_tr = 1
    for t, x in zip(tau, r_x):
        eq_tr = eq_tr * ((t * x + (1 - t) * (1 - x)) % p) % p
    if claim_o != eq_tr * (cAz * cBz - u * cCz - cE) % p:
        return False
    tr.absorb_scalars(b"claims", [cAz, cBz, cCz, cE])
    r = tr.squeeze(b"r")
    claim_i, r_y = _sumcheck_verify((cAz + r * cBz + r * r * cCz) % p, proof["inner"], tr, p, b"inner", 2)
    if claim_i is None or len(r_y) != (2 * shape.num_vars).bit_length() - 1:
        return False
    ex, ey = eq_table(r_x, p), eq_table(r_y, p)
    ev = lambda M: sum(ex[row] * ey[col] % p * v for row, col, v in M) % p
    eval_ABC = (ev(shape.A) + r * ev(shape.B) + r * r % p * ev(shape.C)) % p
    eval_W = proof["eval_W"]
    uX = [u] + list(X) + [0] * (shape.num_vars - 1 - len(X))
    eval_X = mle_eval(uX, r_y[1:], p)
    eval_Z = ((1 - r_y[0]) * eval_W + r_y[0] * eval_X) % p
    if claim_i != eval_ABC * eval_Z % p:
        return False
    tr.absorb_scalars(b"evals", [eval_W, cE])
    if not ipa_verify(curve, gens[:shape.num_vars], gen_c, comm_W, eq_table(r_y[1:], p), eval_W, proof["ipa_W"], tr, msm):
        return False
    return ipa_verify(curve, gens[:shape.num_cons], gen_c, comm_E, ex, cE, proof["ipa_E"], tr, msm)


# ------------------------------------------------------------------------------------ the CAP circuit as an R1CS instance
def poseidon_h2_r1cs(v, salt, p):
    """R1CS of d = H2(v, salt) = calc_d (the `ConsistencyCircuit` of commitment.rs:538-622 proves exactly this hash
    with neptune's SpongeCircuit, ~300 constraints): three constraints per S-box (x^2, x^4, x^5), the round constants
    and MDS layers folded into the linear combinations, one last constraint binding the public digest.
    Returns (shape, W, X = [d]); u = 1 and E = 0 make it a relaxed instance.  z = W ++ [1] ++ [d] ++ 0..."""
    from . import poseidon as P
    t = 5
    rf, rp, rc, mds = P.constants(p, t)
    tag = P.io_pattern_tag([(P.ABSORB, 2), (P.SQUEEZE, 1)]) % p
    W = [v % p, salt % p]                      # witness variables 0, 1
    cons = []                                  # (lcA, lcB, lcC) with lc = {var: coeff}, var -1 = the constant one

    def lc_val(lc):
        return sum(c * (1 if k == -1 else W[k]) for k, c in lc.items()) % p

    def lc_add(x, y):
        out = dict(x)
        for k, c in y.items():
            out[k] = (out.get(k, 0) + c) % p
        return out

    def lc_scale(x, c):
        return {k: v * c % p for k, v in x.items()}

    def mul(x, y):
        W.append(lc_val(x) * lc_val(y) % p)
        cons.append((x, y, {len(W) - 1: 1}))
        return {len(W) - 1: 1}

    def sbox(x):
        x2 = mul(x, x)
        x4 = mul(x2, x2)
        return mul(x4, x)

    state = [{-1: tag}, {0: 1}, {1: 1}, {}, {}]
    k = 0
    for r in range(rf + rp):
        full = r < rf // 2 or r >= rf // 2 + rp
        state = [lc_add(x, {-1: rc[k + i]}) for i, x in enumerate(state)]
        k += t
        if full:
            state = [sbox(x) for x in state]
        else:
            state[0] = sbox(state[0])
        state = [{} if False else _lc_sum([lc_scale(state[i], mds[i][j]) for i in range(t)], p) for j in range(t)]
    d = lc_val(state[1])
    assert d == P.calc_d(v, salt)
    # bind the digest: (state[1]) * 1 = d_public
    num_w = len(W)
    num_vars = 1
    while num_vars < max(num_w, 2):
        num_vars <<= 1
    num_cons = 1
    while num_cons < len(cons) + 1:
        num_cons <<= 1
    one_col, d_col = num_vars, num_vars + 1

    def rows(idx):
        out = []
        for r, con in enumerate(cons):
            for var, c in con[idx].items():
                if c % p:
                    out.append((r, one_col if var == -1 else var, c % p))
        return out

    A, B, C = rows(0), rows(1), rows(2)
    r = len(cons)
    for var, c in state[1].items():
        if c % p:
            A.append((r, one_col if var == -1 else var, c % p))
    B.append((r, one_col, 1))
    C.append((r, d_col, 1))
    shape = R1CSShape(num_cons, num_vars, 1, A, B, C)
    Wp = W + [0] * (num_vars - num_w)
    return shape, Wp, [d]


def _lc_sum(lcs, p):
    out = {}
    for lc in lcs:
        for k, c in lc.items():
            out[k] = (out.get(k, 0) + c) % p
    return out


# ------------------------------------------------------------------------------------ NIFS (the fold of prove_step)
def nifs_ro_elements(pp_digest, U1, U2, comm_T, base_p):
    """absorb order of `NIFS::prove` (published upstream nova-snark nifs.rs / r1cs.rs `absorb_in_ro`; PARITY-UNPINNED):
    pp digest | U1: comm_W, comm_E as (x, y, is_infinity), u, each X entry as four 64-bit limbs | U2: comm_W, X | comm_T"""
    def pt(P):
        return [0, 0, 1] if P is None else [int(P[0]), int(P[1]), 0]
    e = [pp_digest % base_p] + pt(U1["comm_W"]) + pt(U1["comm_E"]) + [U1["u"] % base_p]
    for x in U1["X"]:
        e += [(x >> (64 * k)) & (2 ** 64 - 1) for k in range(4)]
    e += pt(U2["comm_W"]) + [x % base_p for x in U2["X"]] + pt(comm_T)
    return e


def nifs_prove(curve, shape: R1CSShape, gens, pp_digest, U1, W1, U2, W2, msm=None, num_challenge_bits=128):
    """Folds the fresh pair (U2, W2) (u = 1, E = 0) into the running relaxed pair (U1, W1): the cross term
    T = Az1 o Bz2 + Az2 o Bz1 - u1 Cz2 - u2 Cz1, comm_T, r = PoseidonRO(...), then X, u, comm_W, comm_E, W, E <- . + r .
    Reached once per curve from every prove_step (framework.rs:668-675)."""
    from .poseidon import poseidon_ro
    p, base_p = curve.order, curve.p
    msm = msm or curve.msm
    z1, z2 = shape.z(W1["W"], U1["u"], U1["X"]), shape.z(W2["W"], 1, U2["X"])
    a1, b1, c1 = (shape.mul(M, z1, p) for M in (shape.A, shape.B, shape.C))
    a2, b2, c2 = (shape.mul(M, z2, p) for M in (shape.A, shape.B, shape.C))
    T = [(x1 * y2 + x2 * y1 - U1["u"] * w2 - w1) % p for x1, y1, w1, x2, y2, w2 in zip(a1, b1, c1, a2, b2, c2)]
    comm_T = msm(T, gens[:len(T)])
    r = poseidon_ro(nifs_ro_elements(pp_digest, U1, U2, comm_T, base_p), base_p, p, num_challenge_bits)
    U = {"comm_W": curve.add(U1["comm_W"], curve.mul(r, U2["comm_W"])), "comm_E": curve.add(U1["comm_E"], curve.mul(r, comm_T)),
         "u": (U1["u"] + r) % p, "X": [(a + r * b) % p for a, b in zip(U1["X"], U2["X"])]}
    W = {"W": [(a + r * b) % p for a, b in zip(W1["W"], W2["W"])], "E": [(a + r * b) % p for a, b in zip(W1["E"], T)]}
    return comm_T, r, U, W
