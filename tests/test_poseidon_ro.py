"""PoseidonRO (nova-snark, arity 24; commitment.rs:190-198 doc_commit_hash and the NIFS challenge).

CPU tier: the oracle's three statements of it (Python, C port, the library's host instantiation of
the shared field code) agree, and the constants the library derives equal the oracle's.  GPU tier:
the CUDA sponge (through the C ABI) equals the oracle bit for bit, on both Pasta fields, for ragged
lengths, for points with identities, and composed with the Hyrax rows as `reef_doc_commit_u32`.
PARITY UNPINNED against the reference binary (nova-snark is not under /root/reference)."""
import ctypes as C
import random

import numpy as np
import pytest

import reef_b200
import workloads as WL
from oracle import cport
from oracle import poseidon as P
from oracle.fields import FP, FQ
from reef_b200._lib import lib

FIELDS = [("fq", 0, FQ, FP), ("fp", 1, FP, FQ)]


def _pack(xs):
    return b"".join(int(x).to_bytes(32, "little") for x in xs)


def test_round_numbers_of_the_wide_instance():
    assert P.round_numbers(25) == (8, 59)          # what poseidon_ro.cu and oracle/c hard-code


@pytest.mark.parametrize("name,fid,bp,sp", FIELDS)
def test_library_constants_equal_the_oracle(name, fid, bp, sp):
    rf, rp, rc, mds = P.constants(bp, 25)
    rcb, mdsb = C.create_string_buffer(67 * 25 * 32), C.create_string_buffer(625 * 32)
    assert lib.reef_hosttest_poseidon_ro_constants(fid, rcb, mdsb) == 0
    assert rcb.raw == _pack(rc)
    assert mdsb.raw == _pack(x for row in mds for x in row)


@pytest.mark.parametrize("name,fid,bp,sp", FIELDS)
def test_optimised_schedule_derived_on_the_host_equals_the_textbook_rounds(name, fid, bp, sp):
    """sparse partial rounds + rescaled lane 0 for width 25 (what k_poseidon_ro_fast runs): derived in C++ at first use and
    cross-checked there against the textbook permutation; 0 would silently select the slower textbook kernel"""
    assert lib.reef_hosttest_poseidon_ro_fast_ok(fid) == 1


@pytest.mark.parametrize("name,fid,bp,sp", FIELDS)
def test_host_instantiation_and_c_port_equal_the_python_oracle(name, fid, bp, sp):
    rnd = random.Random(fid)
    for n in (1, 2, 23, 24, 25, 48, 50):
        e = [rnd.randrange(bp) for _ in range(n)]
        exp = P.poseidon_ro(e, bp, sp, 256)
        out = C.create_string_buffer(32)
        assert lib.reef_hosttest_poseidon_ro(fid, _pack(e), n, out) == 0
        assert int.from_bytes(out.raw, "little") % sp == exp
        assert cport.poseidon_ro(e, name, sp) == exp
    e = [bp - 1] * 24 + [0]
    assert cport.poseidon_ro(e, name, sp, 250) == P.poseidon_ro(e, bp, sp, 250)


def test_no_gpu_means_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(reef_b200.ReefError):
        reef_b200.Context(0)


# ----------------------------------------------------------------------------------------- GPU tier
@pytest.mark.gpu
@pytest.mark.parametrize("name,fid,bp,sp", FIELDS)
def test_gpu_poseidon_ro_matches_the_oracle(ctx, name, fid, bp, sp):
    rnd = random.Random(10 + fid)
    for n in (1, 3, 24, 25, 47, 48, 49, 24 * 9 + 5):
        e = [rnd.randrange(bp) for _ in range(n)]
        assert ctx.poseidon_ro(e, name, 256) == cport.poseidon_ro(e, name, sp, 256), n
    e = [bp - 1, 0, 1] * 11
    for bits in (128, 250, 255, 256):
        assert ctx.poseidon_ro(e, name, bits) == P.poseidon_ro(e, bp, sp, bits)
    with pytest.raises(reef_b200.ReefError):
        ctx.poseidon_ro([bp], name)                 # not canonical


@pytest.mark.gpu
def test_gpu_ro_over_points_with_identities(ctx):
    pts = WL.generators("pallas", 40)
    pts = pts[:64 * 7] + bytes(64) + pts[64 * 8:64 * 39] + bytes(64)
    elems = []
    for k in range(40):
        x, y = int.from_bytes(pts[64 * k:64 * k + 32], "little"), int.from_bytes(pts[64 * k + 32:64 * k + 64], "little")
        elems += list(P.point_coordinates(None if x == 0 and y == 0 else (x, y)))
    assert ctx.poseidon_ro_points("pallas", pts) == cport.poseidon_ro(elems, "fp", FQ)
    vp = WL.generators("vesta", 9)
    ve = []
    for k in range(9):
        ve += [int.from_bytes(vp[64 * k:64 * k + 32], "little"), int.from_bytes(vp[64 * k + 32:64 * k + 64], "little"), 0]
    assert ctx.poseidon_ro_points("vesta", vp, 250) == cport.poseidon_ro(ve, "fq", FP, 250)


@pytest.mark.gpu
@pytest.mark.parametrize("doc_log,bits", [(9, 8), (17, 8)])
def test_gpu_doc_commit_is_hyrax_rows_plus_ro(ctx, doc_log, bits):
    """NLDocCommitment::new with injected blinds (commitment.rs:133-212): rows == reef_msm_rows, hash == oracle RO over them"""
    rnd = random.Random(doc_log)
    rows, cols = WL.hyrax_dims(doc_log)
    codes = np.random.default_rng(doc_log).integers(0, 1 << bits, size=rows * cols, dtype=np.uint32)
    codes[cols:2 * cols] = 0                       # an all-zero row: with a zero blind its commitment is the identity
    blinds = [rnd.randrange(FQ) for _ in range(rows)]
    blinds[1] = 0
    gens = WL.generators("pallas", cols + 1)
    b = reef_b200.Bases(ctx, "pallas", gens, 255)
    got_rows, got_hash = b.doc_commit(codes, rows, cols, bits, blinds)
    ref_rows = b.msm_rows(codes.reshape(rows, cols), rows, cols, entry_bits=bits, blinds=blinds)
    assert got_rows == b"".join(reef_b200.backend._pt_bytes(Q) for Q in ref_rows)
    assert ref_rows[1] is None
    for r in (0, rows - 1):                        # spot rows against the oracle's MSM
        sc = [int(x) for x in codes[r * cols:(r + 1) * cols]] + [blinds[r]]
        assert ref_rows[r] == cport.msm("pallas", gens, sc, threads=cport.max_threads())
    elems = []
    for Q in ref_rows:
        elems += list(P.point_coordinates(Q))
    assert got_hash == cport.poseidon_ro(elems, "fp", FQ, 256)
    b.free()
