"""CPU tier: the C-ABI library loads, exports every symbol the headers declare, its host-side
logic matches the oracle, the shared __host__ __device__ arithmetic is exact, and every compute
entry refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import random
import re

import pytest

import reef_b200
from oracle import poseidon as P
from oracle.fields import FP, FQ
from oracle.nlookup import ASCII_AB, DNA_AB
from oracle import nlookup as ON
from reef_b200._lib import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    names = set()
    for hdr in ("reef_b200.h", "reef_b200_testing.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(reef_[a-z0-9_]+)\s*\(", src))
    assert len(names) > 30
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in include/ but not exported by libreef_b200.so"


def _op(field, o, a, b=0):
    out = C.create_string_buffer(32)
    lib.reef_hosttest_field_op(field, o, a.to_bytes(32, "little"), b.to_bytes(32, "little"), out)
    return int.from_bytes(out.raw, "little")


@pytest.mark.parametrize("field,p", [(0, FQ), (1, FP)])
def test_shared_field_code_host_instantiation(field, p):
    rnd = random.Random(field + 1)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 254), (1 << 254) - 1, (1 << 128) - 1]
    pairs = [(a, b) for a in edge for b in edge] + [(rnd.randrange(p), rnd.randrange(p)) for _ in range(1500)]
    for a, b in pairs:
        assert _op(field, 0, a, b) == a * b % p
        assert _op(field, 1, a, b) == (a + b) % p
        assert _op(field, 2, a, b) == (a - b) % p
        assert _op(field, 4, a, b) == (a * b + a * a + b * b) % p
        assert _op(field, 6, a, b) == (a & 0xFFFFFFFF) * b % p
        assert _op(field, 7, a % p) == a * a % p
        assert _op(field, 8, a % p, b % p) == a * b % p                     # 29-bit limb Montgomery (fp29.cuh)
        assert _op(field, 9, a % p, b % p) == ((a + b) * (2 * a + b) + (a + b) ** 2) % p   # mul29 + sqr29, lazy/relaxed operands
    for _ in range(10):
        a = rnd.randrange(1, p)
        assert _op(field, 3, a) == pow(a, -1, p)
    big = (1 << 256) - 1                              # lazy accumulator must survive the 17th limb
    assert _op(field, 5, big, big) == 40000 * big * big % p


def test_mul_wide_is_exact():
    rnd = random.Random(9)
    out = C.create_string_buffer(64)
    for t in range(500):
        a, b = rnd.randrange(1 << 256), rnd.randrange(1 << 256)
        if t == 0:
            a = b = (1 << 256) - 1
        lib.reef_hosttest_mul_wide(a.to_bytes(32, "little"), b.to_bytes(32, "little"), out)
        assert int.from_bytes(out.raw, "little") == a * b


def test_optimised_poseidon_schedule_equals_textbook():
    rnd = random.Random(2)
    for _ in range(8):
        st = [rnd.randrange(FQ) for _ in range(5)]
        o = C.create_string_buffer(160)
        lib.reef_hosttest_poseidon_permute(b"".join(x.to_bytes(32, "little") for x in st), o)
        assert [int.from_bytes(o.raw[i * 32:(i + 1) * 32], "little") for i in range(5)] == P.permute(st)


def test_generated_constants_match_oracle():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_poseidon_consts as G
    K = G.derive()
    rf, rp, rc, mds = P.constants()
    assert (rf, rp) == (G.RF, G.RP) and list(rc) == K["rc"] and [list(r) for r in mds] == K["mds"]
    import io
    buf = io.StringIO()
    G.emit(K, buf)
    assert buf.getvalue() == open(os.path.join(ROOT, "reef_b200", "csrc", "poseidon_consts.inc")).read()


def test_host_helpers_match_oracle():
    for mn in [1, 2, 3, 4, 5, 11, 16, 17, 1 << 16, (1 << 16) + 2, (1 << 23) + 1, (1 << 24) - 1]:
        assert reef_b200.logmn(mn) == ON.logmn(mn)
    assert reef_b200.doc_transform(ASCII_AB, "aaaaaaaab") == ON.doc_transform(ASCII_AB, "aaaaaaaab")
    assert reef_b200.doc_transform(DNA_AB, "ACGTTGCA" * 5) == ON.doc_transform(DNA_AB, "ACGTTGCA" * 5)
    assert reef_b200.doc_transform(ASCII_AB, "") == ON.doc_transform(ASCII_AB, "")
    with pytest.raises(reef_b200.ReefError) as e:
        reef_b200.doc_transform(DNA_AB, "ACGX")
    assert e.value.code == 3
    rnd = random.Random(4)
    for (m, l) in [(1, 1), (1, 3), (3, 6), (15, 17), (4, 21), (30, 17), (31, 17), (64, 23), (0, 5)]:
        q = [rnd.randrange(1 << l) for _ in range(m)]
        assert reef_b200.combined_q(q, l) == ON.combined_qs(q, l), (m, l)
    pats = [[], [("A", 2), ("S", 2)], [("A", 4), ("S", 1)], [("A", 24), ("S", 1)] + [("A", 3), ("S", 1)] * 17]
    for p in pats:
        for ds in (0, 1, 123):
            assert reef_b200.io_pattern_tag(p, ds) == P.io_pattern_tag(p, ds)


def test_no_cpu_fallback():
    """Without a usable GPU every compute entry must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tier")
    with pytest.raises(reef_b200.ReefError) as e:
        reef_b200.Context(0)
    assert e.value.code == 2 and "no CPU fallback" in e.value.msg
