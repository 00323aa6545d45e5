"""ctypes front-end of oracle/c/libreef_oracle.so (CPU restatement in C; test infrastructure only).

Same algorithms as oracle/*.py in the reference's own shape, fast enough for full-size parity
checks and for the CPU baseline of bench.py."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .nlookup import combined_qs, logmn, nlookup_pattern

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "libreef_oracle.so")


def build():
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")], stdout=subprocess.DEVNULL)


def _load():
    if not os.path.exists(_SO):
        build()
    lib = C.CDLL(_SO)
    lib.oracle_nlookup.restype = C.c_int
    lib.oracle_merkle.restype = C.c_int
    lib.oracle_msm.restype = C.c_int
    lib.oracle_max_threads.restype = C.c_int
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def max_threads() -> int:
    """Host threads the CPU arm may use: the cores this process is allowed to run on.  Not
    omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 into every rank."""
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or int(lib().oracle_max_threads()))


def _pack(xs):
    return b"".join(int(x).to_bytes(32, "little") for x in xs)


def _unpack(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def poseidon_hash(rows, arity):
    n = len(rows) // arity
    out = C.create_string_buffer(max(n, 1) * 32)
    lib().oracle_poseidon_hash(_pack(rows), C.c_uint32(arity), C.c_uint64(n), out)
    return _unpack(out.raw[:n * 32])


def merkle(doc, threads=1):
    d = np.ascontiguousarray(np.asarray(doc, dtype=np.uint64))
    total, n = 0, len(d)
    k = (n + 1) // 2
    total += k
    while k > 1:
        k = (k + 1) // 2
        total += k
    out = C.create_string_buffer(total * 32)
    sizes = np.zeros(64, dtype=np.uint64)
    nl = lib().oracle_merkle(C.c_void_p(d.ctypes.data), C.c_uint64(n), out, C.c_void_p(sizes.ctypes.data), C.c_int(threads))
    flat = _unpack(out.raw)
    tree, off = [], 0
    for i in range(nl):
        tree.append(flat[off:off + int(sizes[i])])
        off += int(sizes[i])
    return tree


def ops_words(pattern):
    return np.asarray([((1 << 31) | n) if k == "A" else n for k, n in pattern], dtype=np.uint32)


def nlookup_raw(table_arr, is_u32, n, q, query_bytes, n_query, ops, prev_q_bytes, ell):
    """Thin call used by bench.py (inputs pre-marshalled so that the timing is the C code)."""
    qa = np.ascontiguousarray(np.asarray(q, dtype=np.uint64))
    o_claim, o_rounds = C.create_string_buffer(32), C.create_string_buffer(ell * 128)
    o_last, o_next = C.create_string_buffer(32), C.create_string_buffer(32)
    rc = lib().oracle_nlookup(C.c_void_p(table_arr.ctypes.data), C.c_int(is_u32), C.c_uint64(n),
                              C.c_void_p(qa.ctypes.data if len(qa) else None), C.c_uint32(len(qa)), query_bytes,
                              C.c_uint32(n_query), C.c_void_p(ops.ctypes.data), C.c_uint32(len(ops)), prev_q_bytes,
                              o_claim, o_rounds, o_last, o_next)
    assert rc == ell, rc
    return o_claim.raw, o_rounds.raw, o_last.raw, o_next.raw


def wit_nlookup_gadget(table, q, v, running_q=None, running_v=None, tag="nl", doc_hash=0, u32=False):
    """Same contract as oracle.nlookup.wit_nlookup_gadget, evaluated by the C port."""
    n = len(table)
    ell = logmn(n)
    prev_q = list(running_q) if running_q is not None else [0] * ell
    prev_v = running_v if running_v is not None else int(table[0])
    cqs = combined_qs(list(q), ell)
    pattern = nlookup_pattern(tag, len(v), ell, len(cqs))
    query = ([] if tag == "nl" else [doc_hash]) + cqs + [int(x) for x in v] + prev_q + [prev_v]
    if u32:
        arr = np.ascontiguousarray(np.asarray(table, dtype=np.uint32))
    else:
        arr = np.frombuffer(_pack(table), dtype=np.uint8)
    claim, rounds, last, nxt = nlookup_raw(arr, 1 if u32 else 0, n, q, _pack(query), len(query), ops_words(pattern),
                                           _pack(prev_q), ell)
    r = _unpack(rounds)
    rounds = [tuple(r[4 * i:4 * i + 4]) for i in range(ell)]
    return {"claim_r": int.from_bytes(claim, "little"), "combined_q": cqs, "rounds": rounds,
            "sc_last_claim": int.from_bytes(last, "little"), "next_running_claim": int.from_bytes(nxt, "little"),
            "next_running_q": [x[0] for x in rounds], "prev_running_claim": prev_v}


def msm(curve, points, scalars, threads=1):
    cid = {"pallas": 0, "vesta": 1}[curve]
    pb = b"".join((bytes(64) if P is None else int(P[0]).to_bytes(32, "little") + int(P[1]).to_bytes(32, "little"))
                  for P in points) if not isinstance(points, (bytes, bytearray)) else bytes(points)
    sb = _pack(scalars) if not isinstance(scalars, (bytes, bytearray)) else bytes(scalars)
    out = C.create_string_buffer(64)
    lib().oracle_msm(C.c_int(cid), pb, sb, C.c_uint64(len(sb) // 32), out, C.c_int(threads))
    x, y = int.from_bytes(out.raw[:32], "little"), int.from_bytes(out.raw[32:], "little")
    return None if x == 0 and y == 0 else (x, y)


def poseidon_ro(elems, base_field: str, scalar_p: int, num_bits: int = 256) -> int:
    """oracle.poseidon.poseidon_ro evaluated by the C port (base_field: "fq" | "fp")."""
    out = C.create_string_buffer(32)
    data = _pack(elems) if not isinstance(elems, (bytes, bytearray)) else bytes(elems)
    rc = lib().oracle_poseidon_ro(C.c_int({"fq": 0, "fp": 1}[base_field]), data, C.c_uint64(len(data) // 32), out)
    assert rc == 0, rc
    return (int.from_bytes(out.raw, "little") & ((1 << num_bits) - 1)) % scalar_p
