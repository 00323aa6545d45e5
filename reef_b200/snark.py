"""Composed provers on the GPU: the host-side mirror of what the reference calls inside nova-snark once per proof.

  ipa_prove          InnerProductArgument::prove         (commitment.rs:371-393 through hyrax_pc; framework.rs:695-698)
  hyrax_prove_eval   HyraxPC::prove_eval                 (commitment.rs:371-393: `proof_dot_prod_prover`)
  snark_prove        RelaxedR1CSSNARK::prove             (framework.rs:695-698 `CompressedSNARK::prove`, one call per curve;
                                                          commitment.rs:261-268 `cap_prove` on the ConsistencyCircuit)

Every field / curve operation runs in libreef_b200 (sum-check sessions, sparse mat-vec, eq tables, IPA session with
its L / R multi-scalar multiplications and generator folds); this module only sequences the calls and talks to the
TRANSCRIPT, which stays with the caller: nova-snark -- and with it the exact transcript and message layout -- is a git
dependency without a pinned revision that is not under /root/reference (Cargo.toml:12), so the construction follows
the published upstream one and is parity-unpinned (tests pin it against oracle/snark.py: same proof bytes under the
same transcript, and the oracle's verifier accepts).  `tr` is any object with absorb_point(label, P),
absorb_scalars(label, xs) and squeeze(label) -> int.
"""
from __future__ import annotations

import ctypes as C

from ._lib import check, lib
from .backend import _CURVES, _FIELDS, _buf, _pack, _pt_bytes, _pt_from, _unpack

ORDER = {"pallas": 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001,
         "vesta": 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001}
SCALAR_FIELD = {"pallas": "fq", "vesta": "fp"}


def _points_bytes(points) -> bytes:
    return bytes(points) if isinstance(points, (bytes, bytearray)) else b"".join(_pt_bytes(P) for P in points)


def eq_table(ctx, field: str, r) -> list:
    """eq(r, .) with r[0] <-> top index bit, computed on the device"""
    out = C.create_string_buffer((1 << len(r)) * 32)
    check(lib.reef_eq_table(ctx._h, _FIELDS[field], _buf(_pack(r)), len(r), out))
    return _unpack(out.raw)


def axpy(ctx, field: str, a: int, x, y=None) -> list:
    out = C.create_string_buffer(len(x) * 32)
    check(lib.reef_vec_axpy(ctx._h, _FIELDS[field], _buf(_pack([a])), _buf(_pack(x)), _buf(_pack(y)) if y is not None else None,
                            len(x), out))
    return _unpack(out.raw)


class Ipa:
    """Device-resident IPA prover session (reef_ipa_*)."""

    def __init__(self, ctx, curve: str, gens, gen_c, a, b):
        """gens: points / bytes (folded on the device every round) or a registered `Bases` handle (static commitment key:
        never folded, every round is one two-row MSM over the precomputed window levels -- reef_ipa_begin_bases)"""
        h = C.c_void_p()
        ab = bytes(a) if isinstance(a, (bytes, bytearray)) else _pack(a)        # vectors as lists of ints or packed 32-byte LE values
        bb = bytes(b) if isinstance(b, (bytes, bytearray)) else _pack(b)
        self.n = len(ab) // 32
        if hasattr(gens, "_h") and hasattr(gens, "msm"):
            check(lib.reef_ipa_begin_bases(ctx._h, gens._h, _buf(_pt_bytes(gen_c)), _buf(ab), _buf(bb), self.n, C.byref(h)))
        else:
            check(lib.reef_ipa_begin(ctx._h, _CURVES[curve], _buf(_points_bytes(gens)[:64 * self.n]), _buf(_pt_bytes(gen_c)),
                                     _buf(ab), _buf(bb), self.n, C.byref(h)))
        self._h = h

    def round(self):
        L, R = C.create_string_buffer(64), C.create_string_buffer(64)
        check(lib.reef_ipa_round(self._h, L, R))
        return _pt_from(L.raw), _pt_from(R.raw)

    def fold(self, r: int, r_inv: int):
        check(lib.reef_ipa_fold(self._h, _buf(_pack([r])), _buf(_pack([r_inv]))))

    def finish(self):
        a, b, g = C.create_string_buffer(32), C.create_string_buffer(32), C.create_string_buffer(64)
        check(lib.reef_ipa_finish(self._h, a, b, g))
        return int.from_bytes(a.raw, "little"), int.from_bytes(b.raw, "little"), _pt_from(g.raw)

    def free(self):
        if self._h:
            lib.reef_ipa_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def ipa_prove(ctx, curve: str, gens, gen_c, a, b, tr):
    """(L_vec, R_vec, a_hat): log2(n) rounds of two MSMs + folds of a, b and the generators, all on the device."""
    p = ORDER[curve]
    s = Ipa(ctx, curve, gens, gen_c, a, b)
    Ls, Rs = [], []
    try:
        n = s.n
        while n > 1:
            L, R = s.round()
            tr.absorb_point(b"L", L)
            tr.absorb_point(b"R", R)
            r = tr.squeeze(b"r")
            s.fold(r, pow(r, -1, p))
            Ls.append(L)
            Rs.append(R)
            n //= 2
        a_hat, _, _ = s.finish()
    finally:
        s.free()
    return Ls, Rs, a_hat


def hyrax_prove_eval(ctx, table, rows: int, cols: int, gens, gen_c, q, tr):
    """Opening of the committed document polynomial at q (Pallas / Fq): LZ = L^T M on the device over the resident
    table (u32 document codes or field elements), v = <LZ, R>, one IPA of length cols."""
    p = ORDER["pallas"]
    kl = rows.bit_length() - 1
    Lv, Rv = eq_table(ctx, "fq", q[:kl]), eq_table(ctx, "fq", q[kl:])
    LZ = ctx.hyrax_lz(table, rows, cols, Lv)
    v = sum(x * y for x, y in zip(LZ, Rv)) % p
    tr.absorb_scalars(b"v", [v])
    return v, ipa_prove(ctx, "pallas", gens, gen_c, LZ, Rv, tr)


def _sumcheck(ctx, tables, field, tr, p, label, extra=None):
    """drives a device sum-check session with transcript challenges; `extra`: a second session bound with the same
    challenges (its round messages are discarded) to obtain more bound values.  Returns (polys, r, finals, extra finals)."""
    kind = len(tables)
    sc = ctx.sumcheck(tables, field)
    sx = ctx.sumcheck(extra, field) if extra else None
    try:
        if kind == 2:
            claim = sum(a * b for a, b in zip(*tables)) % p
        else:
            claim = sum(a * (b * c - d) for a, b, c, d in zip(*tables)) % p
        polys, rs = [], []
        r_prev = None
        n = len(tables[0])
        while n > 1:
            ev = sc.round(r_prev)
            if sx:
                sx.round(r_prev)
            full = [ev[0], (claim - ev[0]) % p] + list(ev[1:])
            tr.absorb_scalars(label, full)
            r = tr.squeeze(label)
            claim = _interp(full, r, p)
            polys.append(full)
            rs.append(r)
            r_prev = r
            n //= 2
        finals = sc.final(r_prev)
        xfin = sx.final(r_prev) if sx else None
    finally:
        sc.free()
        if sx:
            sx.free()
    return polys, rs, finals, xfin


def _interp(evals, r, p):
    total = 0
    for k, yk in enumerate(evals):
        num, den = 1, 1
        for j in range(len(evals)):
            if j != k:
                num = num * (r - j) % p
                den = den * (k - j) % p
        total += yk * num * pow(den, -1, p)
    return total % p


def _csr(entries, n_rows, transpose=False):
    rows = [[] for _ in range(n_rows)]
    for r, c, v in entries:
        if transpose:
            rows[c].append((r, v))
        else:
            rows[r].append((c, v))
    ptr, idx, val = [0], [], []
    for row in rows:
        for c, v in row:
            idx.append(c)
            val.append(v)
        ptr.append(len(idx))
    return ptr, idx, val


def snark_prove(ctx, curve: str, shape, gens, gen_c, comm_W, comm_E, W, E, u, X, tr):
    """RelaxedR1CSSNARK::prove on the device.  shape: num_cons, num_vars (powers of two), A / B / C as (row, col, value)
    lists; z = W ++ [u] ++ X ++ 0.. (2 * num_vars entries)."""
    p, field = ORDER[curve], SCALAR_FIELD[curve]
    nc, nv = shape.num_cons, shape.num_vars
    z = list(W) + [u] + list(X) + [0] * (nv - 1 - len(X))
    Az, Bz, Cz = (ctx.r1cs_spmv(*_csr(M, nc), z, field) for M in (shape.A, shape.B, shape.C))
    tr.absorb_point(b"W", comm_W)
    tr.absorb_point(b"E", comm_E)
    tr.absorb_scalars(b"uX", [u] + list(X))
    k_x = nc.bit_length() - 1
    tau = [tr.squeeze(b"tau") for _ in range(k_x)]
    uCzE = axpy(ctx, field, u, Cz, E)
    polys_o, r_x, fin_o, fin_x = _sumcheck(ctx, [eq_table(ctx, field, tau), Az, Bz, uCzE], field, tr, p, b"outer", extra=[Cz, E])
    claim_Az, claim_Bz, (claim_Cz, claim_E) = fin_o[1], fin_o[2], fin_x
    tr.absorb_scalars(b"claims", [claim_Az, claim_Bz, claim_Cz, claim_E])
    r = tr.squeeze(b"r")
    ex = eq_table(ctx, field, r_x)
    At, Bt, Ct = (ctx.r1cs_spmv(*_csr(M, 2 * nv, transpose=True), ex, field) for M in (shape.A, shape.B, shape.C))
    ABC = axpy(ctx, field, r * r % p, Ct, axpy(ctx, field, r, Bt, At))
    polys_i, r_y, _, _ = _sumcheck(ctx, [ABC, z], field, tr, p, b"inner")
    ey = eq_table(ctx, field, r_y[1:])
    eval_W = sum(a * b for a, b in zip(W, ey)) % p
    tr.absorb_scalars(b"evals", [eval_W, claim_E])
    ipa_W = ipa_prove(ctx, curve, gens, gen_c, W, ey, tr)
    ipa_E = ipa_prove(ctx, curve, gens, gen_c, E, ex, tr)
    return {"outer": polys_o, "claims": (claim_Az, claim_Bz, claim_Cz, claim_E), "inner": polys_i, "eval_W": eval_W,
            "ipa_W": ipa_W, "ipa_E": ipa_E}


# ------------------------------------------------------------------------------------ NIFS (the fold of prove_step)
def _limbs64(x: int):
    return [(x >> (64 * k)) & (2 ** 64 - 1) for k in range(4)]


def nifs_ro_elements(pp_digest: int, U1, U2, comm_T, base_p: int):
    """What `NIFS::prove` absorbs (nova-snark nifs.rs, published upstream; PARITY-UNPINNED): the public-parameter digest,
    the relaxed instance U1 = (comm_W, comm_E, u, X) -- points as (x, y, is_infinity), u as a base-field element, every X
    entry as four 64-bit limbs --, the fresh instance U2 = (comm_W, X) with X as base-field elements, then comm_T:
    24 elements for two public IOs (nova's NUM_FE_FOR_RO)."""
    def pt(P):
        return [0, 0, 1] if P is None else [int(P[0]), int(P[1]), 0]
    e = [pp_digest % base_p] + pt(U1["comm_W"]) + pt(U1["comm_E"]) + [U1["u"] % base_p]
    for x in U1["X"]:
        e += _limbs64(x)
    e += pt(U2["comm_W"]) + [x % base_p for x in U2["X"]] + pt(comm_T)
    return e


def nifs_prove(ctx, curve: str, shape, gens_bases, pp_digest: int, U1, W1, U2, W2, num_challenge_bits: int = 128):
    """`NIFS::prove` (nova-snark, once per curve inside every prove_step, framework.rs:668-675) on the device:
    six sparse products, the cross term T, commit(T) (MSM over the registered generators `gens_bases`), the Poseidon
    random-oracle challenge r, the folds W1 + r W2 / E1 + r T, and the two-term commitment folds.
    U1: dict(comm_W, comm_E, u, X); W1: dict(W, E); U2: dict(comm_W, X); W2: dict(W).
    Returns (comm_T, r, folded U, folded W)."""
    p, field = ORDER[curve], SCALAR_FIELD[curve]
    base_p = ORDER["vesta" if curve == "pallas" else "pallas"]
    nc = shape.num_cons
    z1 = shape.z(W1["W"], U1["u"], U1["X"])
    z2 = shape.z(W2["W"], 1, U2["X"])
    abc1 = [v for M in (shape.A, shape.B, shape.C) for v in ctx.r1cs_spmv(*_csr(M, nc), z1, field)]
    abc2 = [v for M in (shape.A, shape.B, shape.C) for v in ctx.r1cs_spmv(*_csr(M, nc), z2, field)]
    out = C.create_string_buffer(nc * 32)
    check(lib.reef_nova_cross_term(ctx._h, _FIELDS[field], _buf(_pack(abc1)), _buf(_pack(abc2)), _buf(_pack([U1["u"]])), _buf(_pack([1])), nc, out))
    T = _unpack(out.raw)
    comm_T = gens_bases.msm(T)
    # the random oracle hashes coordinates: elements of the curve's BASE field (= the other curve's scalar field)
    ro_field = "fp" if curve == "pallas" else "fq"
    r = ctx.poseidon_ro(nifs_ro_elements(pp_digest, U1, U2, comm_T, base_p), ro_field, num_challenge_bits)
    W = axpy(ctx, field, r, W2["W"], W1["W"])
    E = axpy(ctx, field, r, T, W1["E"])
    comm_W = _lincomb(ctx, curve, U1["comm_W"], U2["comm_W"], r)
    comm_E = _lincomb(ctx, curve, U1["comm_E"], comm_T, r)
    U = {"comm_W": comm_W, "comm_E": comm_E, "u": (U1["u"] + r) % p, "X": [(a + r * b) % p for a, b in zip(U1["X"], U2["X"])]}
    return comm_T, r, U, {"W": W, "E": E}


def _lincomb(ctx, curve, P1, P2, r):
    """P1 + r P2 on the device (a generic two-term MSM); None = the identity"""
    pts, sc = [], []
    if P1 is not None:
        pts.append(P1)
        sc.append(1)
    if P2 is not None:
        pts.append(P2)
        sc.append(r)
    return ctx.msm(curve, pts, sc) if pts else None
