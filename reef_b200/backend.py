"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

Names follow /root/reference/src/backend: `doc_transform` (framework.rs:978), `logmn`
(costs.rs:10), `MerkleCommitment` (merkle_tree.rs:10-192), `calc_d` (commitment.rs:495),
`wit_nlookup_gadget` (r1cs.rs:2177), `gen_eq_table` / `linear_mle_product` /
`prover_mle_partial_eval` / `verifier_mle_eval` (r1cs_helper.rs:441-641).
Values cross this layer as Python ints; on the wire they are 32-byte LE canonical.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import NlookupOut, ReefError, check, lib

TAG_NL, TAG_NLDOC, TAG_NLHYBRID = 0, 1, 2
_TAGS = {"nl": TAG_NL, "nldoc": TAG_NLDOC, "nlhybrid": TAG_NLHYBRID}
ABSORB_BIT = 1 << 31


def _pack(xs) -> bytes:
    return b"".join(int(x).to_bytes(32, "little") for x in xs)


def _unpack(b) -> list:
    b = bytes(b)
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def _buf(b: bytes):
    return C.create_string_buffer(b, len(b)) if len(b) else C.create_string_buffer(1)


def _u64(xs):
    if isinstance(xs, np.ndarray):
        return np.ascontiguousarray(xs.astype(np.uint64, copy=False))
    return np.ascontiguousarray(np.asarray(list(xs), dtype=np.uint64))


def _ops(pattern):
    """[('A'|'S', n), ...] -> u32 words (bit 31 = absorb)."""
    out = []
    for kind, n in pattern:
        if kind not in ("A", "S"):
            raise ValueError("pattern ops are ('A', n) or ('S', n)")
        out.append((ABSORB_BIT | n) if kind == "A" else n)
    return np.asarray(out, dtype=np.uint32)


# ------------------------------------------------------------------------- pure host helpers
def logmn(mn: int) -> int:
    return int(lib.reef_logmn(int(mn)))


def doc_transform(ab: str, doc: str) -> list:
    a = np.asarray([ord(c) for c in ab], dtype=np.uint32)
    d = np.asarray([ord(c) for c in doc], dtype=np.uint32)
    n = C.c_uint64(0)
    cap = 1 << (max(len(doc) + 2, 2).bit_length() + 1)
    out = np.zeros(cap, dtype=np.uint64)
    check(lib.reef_doc_transform(a.ctypes.data, len(a), d.ctypes.data if len(d) else None, len(d), out.ctypes.data,
                                 cap, C.byref(n)))
    return out[:n.value].tolist()


def combined_q(q, sc_l: int) -> list:
    qa = _u64(q)
    cap = (len(qa) * sc_l) // 254 + 2
    out = C.create_string_buffer(cap * 32)
    n = C.c_uint32(0)
    check(lib.reef_combined_q(qa.ctypes.data if len(qa) else None, len(qa), sc_l, out, cap, C.byref(n)))
    return _unpack(out.raw[:n.value * 32])


def io_pattern_tag(pattern, domain_separator: int = 0) -> int:
    ops = _ops(pattern)
    out = C.create_string_buffer(32)
    check(lib.reef_io_pattern_tag(ops.ctypes.data if len(ops) else None, len(ops), domain_separator, out))
    return int.from_bytes(out.raw, "little")


# ------------------------------------------------------------------------- device objects
class Context:
    """One libreef_b200 context (one CUDA stream) on one GPU."""

    def __init__(self, device: int = 0, latency_critical: bool = None):
        """latency_critical: True / False = highest / lowest CUDA stream priority (reef_init_prio); None = default stream"""
        h = C.c_void_p()
        if latency_critical is None:
            check(lib.reef_init(device, C.byref(h)))
        else:
            check(lib.reef_init_prio(device, int(latency_critical) if not isinstance(latency_critical, bool) else (1 if latency_critical else 0), C.byref(h)))
        self._h = h
        self.device = device

    def close(self):
        if self._h:
            lib.reef_shutdown(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib.reef_sync(self._h))

    @property
    def stream(self) -> int:
        return int(lib.reef_stream(self._h) or 0)

    # ---- Poseidon
    def poseidon_hash(self, rows, arity: int) -> list:
        """n one-shot hashes; rows = flat list of n*arity ints."""
        n = len(rows) // arity
        out = C.create_string_buffer(max(n, 1) * 32)
        check(lib.reef_poseidon_hash(self._h, _buf(_pack(rows)), arity, n, out))
        return _unpack(out.raw[:n * 32])

    def calc_d(self, v: int, salt: int) -> int:
        out = C.create_string_buffer(32)
        check(lib.reef_calc_d(self._h, _buf(_pack([v])), _buf(_pack([salt])), out))
        return int.from_bytes(out.raw, "little")

    def poseidon_sponge(self, pattern, elems, domain_separator: int = 0) -> list:
        ops = _ops(pattern)
        n_out = sum(n for k, n in pattern if k == "S")
        out = C.create_string_buffer(max(n_out, 1) * 32)
        check(lib.reef_poseidon_sponge(self._h, _buf(_pack(elems)), len(elems), ops.ctypes.data, len(ops),
                                       domain_separator, out, n_out))
        return _unpack(out.raw[:n_out * 32])

    def poseidon_ro(self, elems, base_field: str = "fp", num_bits: int = 256) -> int:
        """nova-snark `PoseidonRO` (arity 24): absorb `elems` (elements of base_field), squeeze(num_bits) into the other
        Pasta field (commitment.rs:190-198; the NIFS challenge of prove_step)."""
        out = C.create_string_buffer(32)
        check(lib.reef_poseidon_ro(self._h, _FIELDS[base_field], _buf(_pack(elems)), len(elems), num_bits, out))
        return int.from_bytes(out.raw, "little")

    def poseidon_ro_points(self, curve, points, num_bits: int = 256) -> int:
        """`absorb_in_ro` of every point (x, y, is_infinity), then squeeze(num_bits): doc_commit_hash (commitment.rs:190-198)"""
        raw = bytes(points) if isinstance(points, (bytes, bytearray)) else b"".join(_pt_bytes(P) for P in points)
        out = C.create_string_buffer(32)
        check(lib.reef_poseidon_ro_points(self._h, _CURVES[curve], _buf(raw), len(raw) // 64, num_bits, out))
        return int.from_bytes(out.raw, "little")

    # ---- tables
    def table(self, values) -> "Table":
        return Table(self, values=values)

    def table_u32(self, codes, async_upload: bool = False) -> "Table":
        """async_upload: reef_table_upload_u32_async (the caller keeps `codes` -- page-locked -- alive until the table is consumed)"""
        return Table(self, codes=codes, async_upload=async_upload)

    def table_hybrid(self, pub_table, fill: int, half_len: int, doc_codes) -> "Table":
        """The `--hybrid` merged table (r1cs.rs:2101-2112), expanded on the device."""
        a = np.ascontiguousarray(np.asarray(doc_codes, dtype=np.uint32))
        pub = _pack(pub_table) if not isinstance(pub_table, (bytes, bytearray)) else bytes(pub_table)
        h = C.c_void_p()
        check(lib.reef_table_hybrid_u32(self._h, _buf(pub), len(pub) // 32, _buf(_pack([fill])), half_len, a.ctypes.data, len(a),
                                        C.byref(h)))
        t = Table.__new__(Table)
        t.ctx, t._h = self, h
        return t

    # ---- MLE building blocks (reference-shaped)
    def gen_eq_table(self, rs, qs, last_q) -> list:
        ell = len(last_q)
        out = C.create_string_buffer((1 << ell) * 32)
        qa = _u64(qs)
        check(lib.reef_gen_eq_table(self._h, _buf(_pack(rs)), qa.ctypes.data if len(qa) else None, len(qa),
                                    _buf(_pack(last_q)), ell, out))
        return _unpack(out.raw)

    def linear_mle_product(self, table_t: "Table", table_eq: "Table", ell: int, i: int, sponge: "Sponge"):
        out = C.create_string_buffer(128)
        check(lib.reef_linear_mle_product(self._h, table_t._h, table_eq._h, ell, i, sponge._h, out))
        r, xsq, x, con = _unpack(out.raw)
        return r, xsq, x, con

    def verifier_mle_eval(self, table: "Table", q) -> int:
        out = C.create_string_buffer(32)
        check(lib.reef_verifier_mle_eval(self._h, table._h, _buf(_pack(q)), len(q), out))
        return int.from_bytes(out.raw, "little")

    def prover_mle_partial_eval(self, table: "Table", x):
        """x entries are ints, -1 marks the hole (at most one)."""
        holes = [k for k, v in enumerate(x) if v == -1]
        if len(holes) > 1:
            raise ValueError("at most one hole is supported")
        hole = holes[0] if holes else -1
        xs = [0 if v == -1 else v for v in x]
        oc, ok = C.create_string_buffer(32), C.create_string_buffer(32)
        check(lib.reef_prover_mle_partial_eval(self._h, table._h, _buf(_pack(xs)), len(xs), hole, oc, ok))
        return int.from_bytes(oc.raw, "little"), int.from_bytes(ok.raw, "little")

    def hyrax_lz(self, table: "Table", rows: int, cols: int, L) -> list:
        """LZ[j] = sum_i L[i] * M[i][j] over the rows x cols row-major view of `table`."""
        out = C.create_string_buffer(cols * 32)
        check(lib.reef_hyrax_lz(self._h, table._h, rows, cols, _buf(_pack(L)), out))
        return _unpack(out.raw)

    # ---- nlookup
    def wit_nlookup_gadget(self, table: "Table", q, v, running_q=None, running_v=None, tag="nl", doc_hash=None, witness=None, slots=None):
        """r1cs.rs:2177-2393.  Returns NlookupResult.  witness / slots: also scatter the outputs into an index-addressed
        witness buffer on the device (slots = dict(claim_r=, rounds=, last_claim=, next_claim=), missing = skip)."""
        m = len(q)
        assert m == len(v)
        ell_cap = 64
        cq_cap = (m * ell_cap) // 254 + 2
        bufs = dict(prev=C.create_string_buffer(32), cq=C.create_string_buffer(cq_cap * 32),
                    claim=C.create_string_buffer(32), rounds=C.create_string_buffer(ell_cap * 4 * 32),
                    last=C.create_string_buffer(32), nxt=C.create_string_buffer(32))
        o = NlookupOut()
        o.prev_running_claim = C.addressof(bufs["prev"])
        o.combined_q = C.addressof(bufs["cq"])
        o.combined_q_cap = cq_cap
        o.claim_r = C.addressof(bufs["claim"])
        o.rounds = C.addressof(bufs["rounds"])
        o.rounds_cap = ell_cap
        o.sc_last_claim = C.addressof(bufs["last"])
        o.next_running_claim = C.addressof(bufs["nxt"])
        qa = _u64(q)
        vb = _buf(_pack(v))
        pq = _buf(_pack(running_q)) if running_q is not None else None
        pv = _buf(_pack([running_v])) if running_v is not None else None
        dh = _buf(_pack([doc_hash])) if doc_hash is not None else None
        if witness is None:
            check(lib.reef_nlookup_prove(self._h, _TAGS[tag] if isinstance(tag, str) else tag, table._h,
                                         qa.ctypes.data if m else None, vb if m else None, m, pq, pv, dh, C.byref(o)))
        else:
            from ._lib import NlookupSlots
            none = 2 ** 64 - 1
            sl = NlookupSlots(*(int((slots or {}).get(k, none)) for k in ("claim_r", "rounds", "last_claim", "next_claim")))
            check(lib.reef_nlookup_prove_w(self._h, _TAGS[tag] if isinstance(tag, str) else tag, table._h,
                                           qa.ctypes.data if m else None, vb if m else None, m, pq, pv, dh, C.byref(o),
                                           witness._h, C.byref(sl)))
        ell = o.ell
        rounds = _unpack(bufs["rounds"].raw[:ell * 4 * 32])
        rounds = [tuple(rounds[4 * i:4 * i + 4]) for i in range(ell)]
        return NlookupResult(
            prev_running_claim=int.from_bytes(bufs["prev"].raw, "little"),
            combined_q=_unpack(bufs["cq"].raw[:o.num_cqs * 32]),
            claim_r=int.from_bytes(bufs["claim"].raw, "little"),
            rounds=rounds,
            sc_last_claim=int.from_bytes(bufs["last"].raw, "little"),
            next_running_claim=int.from_bytes(bufs["nxt"].raw, "little"),
            next_running_q=[r[0] for r in rounds],
        )

    # ---- multi-GPU: small-message exchange over NVLink peer memory (p2p.cu)
    def mailbox_create(self, world: int) -> bytes:
        h = C.create_string_buffer(64)
        check(lib.reef_mailbox_create(self._h, world, h))
        return h.raw

    def mailbox_ptr(self) -> int:
        return int(lib.reef_mailbox_ptr(self._h) or 0)

    def mailbox_connect(self, rank: int, world: int, handles: bytes):
        check(lib.reef_mailbox_connect(self._h, rank, world, _buf(bytes(handles))))

    def mailbox_connect_local(self, rank: int, world: int, ptrs):
        arr = (C.c_void_p * world)(*[C.c_void_p(p) for p in ptrs])
        check(lib.reef_mailbox_connect_local(self._h, rank, world, arr))

    def p2p_allgather(self, mine_dev_ptr: int, nbytes: int, out_dev_ptr: int):
        check(lib.reef_p2p_allgather(self._h, C.c_void_p(mine_dev_ptr), nbytes, C.c_void_p(out_dev_ptr)))

    def p2p_status(self):
        check(lib.reef_p2p_status(self._h))

    # ---- Spartan sweeps behind CompressedSNARK::prove (framework.rs:695-698), R1CS products, IPA fold
    def sumcheck(self, tables, field: str = "fq") -> "Sumcheck":
        """2 tables: prove_quad (A*B); 4 tables: prove_cubic_with_additive_term (A*(B*C-D))."""
        return Sumcheck(self, tables, field)

    def r1cs_spmv(self, row_ptr, col_idx, vals, z, field: str = "fq") -> list:
        rp = np.ascontiguousarray(np.asarray(row_ptr, dtype=np.uint64))
        ci = np.ascontiguousarray(np.asarray(col_idx, dtype=np.uint32))
        n_rows = len(rp) - 1
        out = C.create_string_buffer(max(n_rows, 1) * 32)
        check(lib.reef_r1cs_spmv(self._h, _FIELDS[field], rp.ctypes.data, ci.ctypes.data if len(ci) else None,
                                 _buf(_pack(vals)) if len(vals) else None, n_rows, len(z), _buf(_pack(z)), out))
        return _unpack(out.raw[:n_rows * 32])

    def ipa_fold_bases(self, curve, points, s_lo: int, s_hi: int) -> list:
        """out[i] = s_lo * G[i] + s_hi * G[i + n/2]  (nova ipa_pc `ck.fold`)."""
        n = len(points)
        out = C.create_string_buffer(max(n // 2, 1) * 64)
        check(lib.reef_ipa_fold_bases(self._h, _CURVES[curve], _buf(b"".join(_pt_bytes(P) for P in points)), n,
                                      _buf(_pack([s_lo])), _buf(_pack([s_hi])), out))
        return [_pt_from(out.raw[i * 64:(i + 1) * 64]) for i in range(n // 2)]

    # ---- Merkle
    def merkle(self, doc) -> "MerkleCommitment":
        return MerkleCommitment(self, doc)

    def merkle_raw(self, doc, levels_out=None):
        """MerkleCommitment::new with raw buffers: (root int, levels bytes: every level, leaf parents first) -- the
        whole tree comes back to the host, as the .cmt of --commit holds it (merkle_tree.rs:10-15).
        doc: uint64 numpy array is used in place; levels_out: optional caller-owned (page-locked) uint8 numpy buffer of
        reef_merkle_tree_elems(n) * 32 bytes that receives the levels instead of a fresh bytes object."""
        d = doc if isinstance(doc, np.ndarray) and doc.dtype == np.uint64 and doc.flags["C_CONTIGUOUS"] else _u64(doc)
        total = int(lib.reef_merkle_tree_elems(len(d)))
        sizes = np.zeros(64, dtype=np.uint64)
        nl = C.c_uint32(0)
        root = C.create_string_buffer(32)
        if levels_out is not None:
            assert levels_out.nbytes >= total * 32
            check(lib.reef_merkle_build(self._h, d.ctypes.data, len(d), levels_out.ctypes.data, sizes.ctypes.data, C.byref(nl), root))
            levels = levels_out
        else:
            levels = C.create_string_buffer(max(total, 1) * 32)
            check(lib.reef_merkle_build(self._h, d.ctypes.data, len(d), levels, sizes.ctypes.data, C.byref(nl), root))
        self.last_level_sizes, self.last_n_levels = sizes, int(nl.value)      # for the .cmt writer (reef_cmt_merkle_write)
        return int.from_bytes(root.raw, "little"), levels

    def merkle_root(self, doc) -> int:
        """Root only (power-of-two documents): the tree stays on the device."""
        d = _u64(doc)
        root = C.create_string_buffer(32)
        check(lib.reef_merkle_subtree(self._h, d.ctypes.data, len(d), 0, None, root))
        return int.from_bytes(root.raw, "little")

    # ---- MSM
    def bases(self, curve, points, scalar_bits: int = 0) -> "Bases":
        return Bases(self, curve, points, scalar_bits)

    def msm(self, curve, points, scalars):
        """One-shot convenience: register `points`, multiply, free."""
        b = Bases(self, curve, points)
        try:
            return b.msm(scalars)
        finally:
            b.free()


_FIELDS = {"fq": 0, "fp": 1}
_CURVES = {"pallas": 0, "vesta": 1}


class Witness:
    """(f1) index-addressed witness buffer on the device (reef_witness_*): the host writes plain values by index, the
    sum-check writes its outputs into their slots on the device, commit(W) reads it in place."""

    def __init__(self, ctx: "Context", n: int):
        self.ctx, self.n = ctx, n
        h = C.c_void_p()
        check(lib.reef_witness_create(ctx._h, n, C.byref(h)))
        self._h = h

    def set(self, idx, vals):
        i = _u64(idx)
        check(lib.reef_witness_set(self._h, i.ctypes.data, _buf(_pack(vals)), len(i)))

    def set_small(self, idx, vals):
        i, v = _u64(idx), _u64(vals)
        check(lib.reef_witness_set_u64(self._h, i.ctypes.data, v.ctypes.data, len(i)))

    def read(self, first: int = 0, k: int = None) -> list:
        k = self.n - first if k is None else k
        out = C.create_string_buffer(max(k, 1) * 32)
        check(lib.reef_witness_read(self._h, first, k, out))
        return _unpack(out.raw[:k * 32])

    @property
    def dev_ptr(self) -> int:
        return int(lib.reef_witness_dev(self._h) or 0)

    def free(self):
        if self._h:
            lib.reef_witness_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Sumcheck:
    """Device-resident sum-check session: round(None) -> evaluations of round 1; round(r_i) binds the
    top variable with r_i and returns the evaluations of round i+1; final(r_k) -> bound values."""

    def __init__(self, ctx: "Context", tables, field: str = "fq"):
        self.ctx, self.kind, self.n = ctx, len(tables), len(tables[0])
        bufs = [_buf(_pack(t)) for t in tables]
        arr = (C.c_void_p * self.kind)(*[C.cast(b, C.c_void_p) for b in bufs])
        h = C.c_void_p()
        check(lib.reef_sumcheck_begin(ctx._h, _FIELDS[field], self.kind, arr, self.n, C.byref(h)))
        self._h = h

    def round(self, r_prev=None) -> list:
        ne = 2 if self.kind == 2 else 3
        out = C.create_string_buffer(ne * 32)
        check(lib.reef_sumcheck_round(self._h, _buf(_pack([r_prev])) if r_prev is not None else None, out))
        return _unpack(out.raw)

    def final(self, r_last: int) -> list:
        out = C.create_string_buffer(self.kind * 32)
        check(lib.reef_sumcheck_final(self._h, _buf(_pack([r_last])), out))
        return _unpack(out.raw)

    def free(self):
        if self._h:
            lib.reef_sumcheck_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


@dataclass
class NlookupResult:
    prev_running_claim: int
    combined_q: list
    claim_r: int
    rounds: list            # per round: (sc_r, xsq, x, const)
    sc_last_claim: int
    next_running_claim: int
    next_running_q: list


class Table:
    def __init__(self, ctx: Context, values=None, codes=None, dev_ptr=None, n=None, is_u32=False, async_upload=False):
        self.ctx = ctx
        h = C.c_void_p()
        if values is not None and isinstance(values, np.ndarray):
            a = np.ascontiguousarray(values.view(np.uint8).reshape(-1))       # n x 32 bytes, LE canonical
            check(lib.reef_table_upload(ctx._h, a.ctypes.data, a.size // 32, C.byref(h)))
        elif values is not None:
            b = _pack(values)
            check(lib.reef_table_upload(ctx._h, _buf(b), len(values), C.byref(h)))
        elif codes is not None:
            a = np.ascontiguousarray(np.asarray(codes, dtype=np.uint32))
            if async_upload:
                self._keep = a                       # the upload may still be reading it when this returns
                check(lib.reef_table_upload_u32_async(ctx._h, a.ctypes.data, len(a), C.byref(h)))
            else:
                check(lib.reef_table_upload_u32(ctx._h, a.ctypes.data, len(a), C.byref(h)))
        else:
            check(lib.reef_table_wrap_dev(ctx._h, C.c_void_p(dev_ptr), n, 1 if is_u32 else 0, C.byref(h)))
        self._h = h

    def __len__(self):
        return int(lib.reef_table_len(self._h))

    def download(self, n: int) -> list:
        out = C.create_string_buffer(max(n, 1) * 32)
        check(lib.reef_table_download(self._h, out, n))
        return _unpack(out.raw[:n * 32])

    def free(self):
        if self._h:
            lib.reef_table_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Sponge:
    """SpongeAPI session (start/absorb/squeeze/finish), state resident on the device."""

    def __init__(self, ctx: Context, pattern, domain_separator: int = 0):
        ops = _ops(pattern)
        h = C.c_void_p()
        check(lib.reef_sponge_start(ctx._h, ops.ctypes.data, len(ops), domain_separator, C.byref(h)))
        self._h = h

    def absorb(self, elems):
        check(lib.reef_sponge_absorb(self._h, _buf(_pack(elems)), len(elems)))

    def squeeze(self, n: int) -> list:
        out = C.create_string_buffer(max(n, 1) * 32)
        check(lib.reef_sponge_squeeze(self._h, n, out))
        return _unpack(out.raw[:n * 32])

    def finish(self):
        h, self._h = self._h, None
        check(lib.reef_sponge_finish(h))


class MerkleCommitment:
    """merkle_tree.rs:10-192: `commitment`, `tree` (levels, leaf parents first), `doc`."""

    def __init__(self, ctx: Context, doc):
        self.doc = [int(c) for c in doc]
        d = _u64(self.doc)
        total = int(lib.reef_merkle_tree_elems(len(d)))
        levels = C.create_string_buffer(max(total, 1) * 32)
        sizes = np.zeros(64, dtype=np.uint64)
        nl = C.c_uint32(0)
        root = C.create_string_buffer(32)
        check(lib.reef_merkle_build(ctx._h, d.ctypes.data if len(d) else None, len(d), levels, sizes.ctypes.data,
                                    C.byref(nl), root))
        self._levels, self._sizes, self._nl, self._d = levels, sizes, nl.value, d
        self.commitment = int.from_bytes(root.raw, "little")
        flat = _unpack(levels.raw[:total * 32])
        self.tree, off = [], 0
        for k in range(nl.value):
            self.tree.append(flat[off:off + int(sizes[k])])
            off += int(sizes[k])

    @classmethod
    def build_sharded(cls, ctx: Context, doc_local, n_doc: int, rank: int, world: int, gather, full_tree: bool = False):
        """Multi-GPU build: this rank holds the contiguous leaves [rank * n_doc / world, ...).
        Returns (commitment, tree or None); `gather(bytes) -> list[bytes]` is the all-gather."""
        from .sharding import merkle_sharded
        d = _u64(doc_local)
        per = n_doc // world
        assert len(d) == per

        def build_subtree(a, b):
            lv = C.create_string_buffer(max(per - 1, 1) * 32)
            root = C.create_string_buffer(32)
            check(lib.reef_merkle_subtree(ctx._h, d.ctypes.data, per, a, lv if full_tree else None, root))
            levels, off, n = [], 0, per // 2
            while full_tree and n >= 1:
                levels.append([lv.raw[(off + i) * 32:(off + i + 1) * 32] for i in range(n)])
                off += n
                n //= 2
            return levels, root.raw

        def hash_top(roots):
            g = len(roots)
            lv = C.create_string_buffer(max(g - 1, 1) * 32)
            root = C.create_string_buffer(32)
            check(lib.reef_merkle_top(ctx._h, b"".join(roots), g, lv, root))
            levels, off, n = [], 0, g // 2
            while n >= 1:
                levels.append([lv.raw[(off + i) * 32:(off + i + 1) * 32] for i in range(n)])
                off += n
                n //= 2
            return levels, root.raw

        root, tree = merkle_sharded(build_subtree, hash_top, n_doc, rank, world, gather, full_tree)
        tree_int = [[int.from_bytes(x, "little") for x in lvl] for lvl in tree] if tree is not None else None
        return int.from_bytes(root, "little"), tree_int

    def path_wits(self, idx: int):
        nl = self._nl
        lr = np.zeros(nl, dtype=np.uint8)
        hi = np.zeros(nl, dtype=np.uint8)
        oi = np.zeros(nl, dtype=np.uint64)
        op = C.create_string_buffer(nl * 32)
        check(lib.reef_merkle_path_wits(self._d.ctypes.data, len(self._d), self._levels, self._sizes.ctypes.data, nl,
                                        idx, lr.ctypes.data, hi.ctypes.data, oi.ctypes.data, op))
        opp = _unpack(op.raw)
        return [(bool(lr[k]), int(oi[k]) if hi[k] else None, opp[k]) for k in range(nl)]

    def make_wits(self, lookups):
        return [self.path_wits(q) for q in lookups]


CURVES = {"pallas": 0, "vesta": 1}


def _pt_bytes(P) -> bytes:
    if P is None:
        return bytes(64)
    return int(P[0]).to_bytes(32, "little") + int(P[1]).to_bytes(32, "little")


def _pt_from(b: bytes):
    x, y = int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")
    return None if x == 0 and y == 0 else (x, y)


class Bases:
    """Registered generators (Pedersen `CommitmentGens`): window levels precomputed, HBM-resident."""

    def __init__(self, ctx: Context, curve, points, scalar_bits: int = 0):
        self.ctx = ctx
        self.curve = CURVES[curve] if isinstance(curve, str) else int(curve)
        self.n = len(points)
        raw = b"".join(_pt_bytes(P) for P in points) if not isinstance(points, (bytes, bytearray)) else bytes(points)
        if isinstance(points, (bytes, bytearray)):
            self.n = len(raw) // 64
        h = C.c_void_p()
        check(lib.reef_bases_register(ctx._h, self.curve, _buf(raw), self.n, scalar_bits, C.byref(h)))
        self._h = h

    @property
    def windows(self) -> int:
        return int(lib.reef_bases_windows(self._h))

    @property
    def window_bits(self) -> int:
        return int(lib.reef_bases_window_bits(self._h))

    def msm(self, scalars):
        out = C.create_string_buffer(64)
        if isinstance(scalars, (bytes, bytearray)):
            raw, n = bytes(scalars), len(scalars) // 32
        else:
            raw, n = _pack(scalars), len(scalars)
        check(lib.reef_msm(self.ctx._h, self._h, _buf(raw), n, out))
        return _pt_from(out.raw)

    def msm_u32(self, scalars):
        a = np.ascontiguousarray(np.asarray(scalars, dtype=np.uint32))
        out = C.create_string_buffer(64)
        check(lib.reef_msm_u32(self.ctx._h, self._h, a.ctypes.data if len(a) else None, len(a), out))
        return _pt_from(out.raw)

    def msm_rows(self, matrix, rows: int, cols: int, entry_bits: int = 0, blinds=None):
        """Hyrax `commit`: one commitment per matrix row (commitment.rs:187).  `matrix` is a numpy
        uint32 array (document codes) or a flat list of field elements, row-major."""
        out = C.create_string_buffer(rows * 64)
        bl = _buf(_pack(blinds)) if blinds is not None else None
        if isinstance(matrix, np.ndarray):
            a = np.ascontiguousarray(matrix.astype(np.uint32).reshape(-1))
            check(lib.reef_msm_rows_u32(self.ctx._h, self._h, a.ctypes.data, rows, cols, entry_bits, bl, out))
        else:
            check(lib.reef_msm_rows(self.ctx._h, self._h, _buf(_pack(matrix)), rows, cols, bl, out))
        return [_pt_from(out.raw[i * 64:(i + 1) * 64]) for i in range(rows)]

    def doc_commit(self, codes: "np.ndarray", rows: int, cols: int, entry_bits: int, blinds):
        """`NLDocCommitment::new` arithmetic with injected blinds (commitment.rs:133-212): (row commitments, doc_commit_hash)."""
        a = np.ascontiguousarray(np.asarray(codes, dtype=np.uint32).reshape(-1))
        out, h = C.create_string_buffer(rows * 64), C.create_string_buffer(32)
        bl = bytes(blinds) if isinstance(blinds, (bytes, bytearray)) else _pack(blinds)
        check(lib.reef_doc_commit_u32(self.ctx._h, self._h, a.ctypes.data, rows, cols, entry_bits, _buf(bl), out, h))
        return out.raw, int.from_bytes(h.raw, "little")

    def msm_dev(self, dev_ptr: int, n: int):
        out = C.create_string_buffer(64)
        check(lib.reef_msm_dev(self.ctx._h, self._h, C.c_void_p(dev_ptr), n, out))
        return _pt_from(out.raw)

    def msm_sharded_dev(self, dev_ptr: int, n: int):
        """Window-sharded MSM across the ranks of the context's mailbox world: partial MSM, 128-byte
        all-gather over NVLink peer memory and combine in one stream-ordered call; same result on every rank."""
        out = C.create_string_buffer(64)
        check(lib.reef_msm_sharded_dev(self.ctx._h, self._h, C.c_void_p(dev_ptr), n, out))
        return _pt_from(out.raw)

    def commit_rows_sharded(self, matrix: "np.ndarray", rows: int, cols: int, entry_bits: int, blinds, rank: int, world: int, gather):
        """Hyrax `commit` with the rows split over `world` ranks; returns all `rows` points."""
        from .sharding import hyrax_commit_sharded
        m = np.ascontiguousarray(matrix.astype(np.uint32).reshape(rows, cols))

        def mine(r0, r1):
            out = C.create_string_buffer((r1 - r0) * 64)
            bl = _buf(_pack(blinds[r0:r1])) if blinds is not None else None
            a = np.ascontiguousarray(m[r0:r1].reshape(-1))
            check(lib.reef_msm_rows_u32(self.ctx._h, self._h, a.ctypes.data, r1 - r0, cols, entry_bits, bl, out))
            return [out.raw[i * 64:(i + 1) * 64] for i in range(r1 - r0)]

        return [_pt_from(p) for p in hyrax_commit_sharded(mine, rows, rank, world, gather)]

    def msm_partial_dev(self, dev_ptr: int, n: int, w_begin: int, w_end: int) -> bytes:
        out = C.create_string_buffer(128)
        check(lib.reef_msm_partial_dev(self.ctx._h, self._h, C.c_void_p(dev_ptr), n, w_begin, w_end, out))
        return out.raw

    def combine(self, partials: bytes):
        out = C.create_string_buffer(64)
        check(lib.reef_msm_combine(self.ctx._h, self.curve, _buf(partials), len(partials) // 128, out))
        return _pt_from(out.raw)

    def free(self):
        if self._h:
            lib.reef_bases_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ShardedNlookup:
    """One rank of the multi-GPU sum-check (protocol in oracle/sharded.py).

    `local_table` holds T[j*world + rank].  `gather(dev_ptr_in, nbytes, dev_ptr_out)` must fill
    `dev_ptr_out` with the world x nbytes rank-major all-gather of `dev_ptr_in`, ordered after
    the context's stream (NCCL `all_gather_into_tensor` under `torch.cuda.stream(ExternalStream)`).
    """

    def __init__(self, ctx: Context, local_table: "Table", rank: int, world: int, q, v, prev_q, prev_v, tag="nl",
                 doc_hash=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        m = len(q)
        self._bufs = dict(prev=C.create_string_buffer(32), cq=C.create_string_buffer(((m * 64) // 254 + 2) * 32),
                          claim=C.create_string_buffer(32), rounds=C.create_string_buffer(64 * 4 * 32),
                          last=C.create_string_buffer(32), nxt=C.create_string_buffer(32))
        o = NlookupOut()
        o.prev_running_claim = C.addressof(self._bufs["prev"])
        o.combined_q = C.addressof(self._bufs["cq"])
        o.combined_q_cap = (m * 64) // 254 + 2
        o.claim_r = C.addressof(self._bufs["claim"])
        o.rounds = C.addressof(self._bufs["rounds"])
        o.rounds_cap = 64
        o.sc_last_claim = C.addressof(self._bufs["last"])
        o.next_running_claim = C.addressof(self._bufs["nxt"])
        self._o = o
        qa = _u64(q)
        self._keep = (qa, _buf(_pack(v)), _buf(_pack(prev_q)) if prev_q is not None else None,
                      _buf(_pack([prev_v])) if prev_v is not None else None,
                      _buf(_pack([doc_hash])) if doc_hash is not None else None)
        h = C.c_void_p()
        check(lib.reef_nl_shard_begin(ctx._h, _TAGS[tag] if isinstance(tag, str) else tag, local_table._h, rank, world,
                                      qa.ctypes.data if m else None, self._keep[1] if m else None, m, self._keep[2],
                                      self._keep[3], self._keep[4], C.byref(o), C.byref(h)))
        self._h = h
        self.ell = o.ell
        self.ell_local = o.ell - (world.bit_length() - 1)

    def round_local(self, out_dev_ptr: int):
        check(lib.reef_nl_shard_round_local(self._h, C.c_void_p(out_dev_ptr)))

    def round_finish(self, all_triples_dev_ptr: int):
        check(lib.reef_nl_shard_round_finish(self._h, C.c_void_p(all_triples_dev_ptr)))

    def export(self, out_dev_ptr: int):
        check(lib.reef_nl_shard_export(self._h, C.c_void_p(out_dev_ptr)))

    def finish(self, all_pairs_dev_ptr: int) -> NlookupResult:
        check(lib.reef_nl_shard_finish(self._h, C.c_void_p(all_pairs_dev_ptr), C.byref(self._o)))
        b, ell = self._bufs, self.ell
        rounds = _unpack(b["rounds"].raw[:ell * 4 * 32])
        rounds = [tuple(rounds[4 * i:4 * i + 4]) for i in range(ell)]
        return NlookupResult(prev_running_claim=int.from_bytes(b["prev"].raw, "little"),
                             combined_q=_unpack(b["cq"].raw[:self._o.num_cqs * 32]),
                             claim_r=int.from_bytes(b["claim"].raw, "little"), rounds=rounds,
                             sc_last_claim=int.from_bytes(b["last"].raw, "little"),
                             next_running_claim=int.from_bytes(b["nxt"].raw, "little"),
                             next_running_q=[r[0] for r in rounds])

    def _result(self) -> NlookupResult:
        b, ell = self._bufs, self.ell
        rounds = _unpack(b["rounds"].raw[:ell * 4 * 32])
        rounds = [tuple(rounds[4 * i:4 * i + 4]) for i in range(ell)]
        return NlookupResult(prev_running_claim=int.from_bytes(b["prev"].raw, "little"),
                             combined_q=_unpack(b["cq"].raw[:self._o.num_cqs * 32]),
                             claim_r=int.from_bytes(b["claim"].raw, "little"), rounds=rounds,
                             sc_last_claim=int.from_bytes(b["last"].raw, "little"),
                             next_running_claim=int.from_bytes(b["nxt"].raw, "little"),
                             next_running_q=[r[0] for r in rounds])

    def enqueue_p2p(self):
        """All local rounds with the exchange fused into the kernels (peer mailboxes); returns without
        waiting for the device."""
        for _ in range(self.ell_local):
            check(lib.reef_nl_shard_round_p2p(self._h))

    def finish_p2p(self) -> NlookupResult:
        check(lib.reef_nl_shard_finish_p2p(self._h, C.byref(self._o)))
        return self._result()

    def run_p2p(self) -> NlookupResult:
        self.enqueue_p2p()
        return self.finish_p2p()

    def run(self, gather, scratch_dev_ptr: int) -> NlookupResult:
        """scratch: device buffer of (world + 1) * 96 bytes owned by the caller."""
        mine, allp = scratch_dev_ptr, scratch_dev_ptr + 96
        for _ in range(self.ell_local):
            self.round_local(mine)
            gather(mine, 96, allp)
            self.round_finish(allp)
        self.export(mine)
        gather(mine, 64, allp)
        return self.finish(allp)

    def free(self):
        if self._h:
            lib.reef_nl_shard_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
