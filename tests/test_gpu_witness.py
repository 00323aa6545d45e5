"""GPU tier: (f1) the index-addressed witness buffer -- values written by index, sum-check outputs scattered into
their slots on the device, commit(W) taken over the buffer in place -- and (f2) the shared generator-level cache.
Reference seams: nova.rs:868-1399 / framework.rs:561-572 (string-keyed wires), framework.rs:668-675 (commit(W)),
framework.rs:297-303, 770 (the verifier derives the same commitment key again)."""
import ctypes as C
import random

import numpy as np
import pytest

import reef_b200
import workloads as WL
from oracle import cport
from oracle.fields import FQ
from oracle.nlookup import wit_nlookup_gadget
from reef_b200._lib import ReefError, check, lib

pytestmark = pytest.mark.gpu


def test_witness_buffer_round_trip_and_bounds(ctx):
    rnd = random.Random(1)
    w = reef_b200.Witness(ctx, 1000)
    assert w.read(0, 1000) == [0] * 1000
    idx = [5, 6, 7, 100, 999, 0]
    vals = [rnd.randrange(FQ) for _ in idx]
    w.set(idx, vals)
    w.set_small([10, 11, 500], [1, 0, 2 ** 64 - 1])
    got = w.read()
    exp = [0] * 1000
    for i, v in zip(idx, vals):
        exp[i] = v
    exp[10], exp[11], exp[500] = 1, 0, 2 ** 64 - 1
    assert got == exp
    with pytest.raises(ReefError) as e:
        w.set([1000], [1])
    assert e.value.code == 3                     # index out of bounds, as a Vec index would panic
    with pytest.raises(ReefError):
        w.set([1], [FQ])                         # not canonical
    w.free()


@pytest.mark.parametrize("ell,u32", [(12, True), (17, True), (9, False)])
def test_nlookup_writes_its_outputs_into_the_witness_on_the_device(ctx, ell, u32):
    rnd = random.Random(ell)
    n = 1 << ell
    tab = [rnd.randrange(131) for _ in range(n)] if u32 else [rnd.randrange(FQ) for _ in range(n)]
    t = ctx.table_u32(np.asarray(tab, dtype=np.uint32)) if u32 else ctx.table(tab)
    q = [rnd.randrange(n) for _ in range(3)]
    v = [tab[i] for i in q]
    tag = "nldoc" if u32 else "nl"
    dh = 77 if u32 else None
    nw = 1 << 10
    w = reef_b200.Witness(ctx, nw)
    slots = dict(claim_r=3, rounds=40, last_claim=1, next_claim=nw - 1)
    res = ctx.wit_nlookup_gadget(t, q, v, None, None, tag, dh, witness=w, slots=slots)
    if u32:
        exp = cport.wit_nlookup_gadget(tab, q, v, None, None, tag, dh, u32=True)
    else:
        exp = wit_nlookup_gadget(tab, q, v, None, None, tag, 0, fast=True)
    assert res.rounds == exp["rounds"] and res.claim_r == exp["claim_r"]
    got = w.read()
    assert got[3] == exp["claim_r"] and got[1] == exp["sc_last_claim"] and got[nw - 1] == exp["next_running_claim"]
    flat = [x for r in exp["rounds"] for x in r]
    assert got[40:40 + 4 * ell] == flat
    untouched = set(range(nw)) - {3, 1, nw - 1} - set(range(40, 40 + 4 * ell))
    assert all(got[i] == 0 for i in untouched)
    # skipping a slot leaves it alone; a slot range that does not fit is the reference's index panic
    w2 = reef_b200.Witness(ctx, 4 * ell + 2)
    ctx.wit_nlookup_gadget(t, q, v, None, None, tag, dh, witness=w2, slots=dict(rounds=1))
    g2 = w2.read()
    assert g2[1:1 + 4 * ell] == flat and g2[0] == 0 and g2[-1] == 0
    with pytest.raises(ReefError) as e:
        ctx.wit_nlookup_gadget(t, q, v, None, None, tag, dh, witness=w2, slots=dict(rounds=3))
    assert e.value.code == 3
    w.free()
    w2.free()
    t.free()


def test_commit_w_over_the_witness_buffer_in_place(ctx):
    """commit(W) = MSM over the device-resident witness (framework.rs:668-675): no host copy of W"""
    rnd = random.Random(4)
    n = 1 << 12
    gens = WL.generators("pallas", n)
    w = reef_b200.Witness(ctx, n)
    idx = list(range(0, n, 3))
    vals = [rnd.randrange(FQ) for _ in idx]
    w.set(idx, vals)
    w.set_small(list(range(1, n, 3)), [rnd.randrange(2) for _ in range(1, n, 3)])
    b = reef_b200.Bases(ctx, "pallas", gens)
    got = b.msm_dev(w.dev_ptr, n)
    assert got == cport.msm("pallas", gens, w.read(), threads=cport.max_threads())
    b.free()
    w.free()


def test_w_and_t_commit_as_two_rows_of_one_msm_over_device_memory(ctx):
    """reef_msm_rows_dev: commit(W) and commit(T) of a fold (same commitment key, framework.rs:668-675) as ONE row-batched
    MSM over scalars that already live on the device (here: one witness-sized buffer holding W then T)"""
    rnd = random.Random(9)
    n = 1 << 11
    gens = WL.generators("vesta", n)
    from oracle.fields import FP
    w = reef_b200.Witness(ctx, 2 * n)
    W = [rnd.randrange(2) if rnd.random() < 0.8 else rnd.randrange(FQ) for _ in range(n)]      # witness-like: mostly bits
    T = [rnd.randrange(FQ) for _ in range(n)]
    w.set(list(range(2 * n)), W + T)
    b = reef_b200.Bases(ctx, "vesta", gens)
    out = C.create_string_buffer(128)
    check(lib.reef_msm_rows_dev(ctx._h, b._h, C.c_void_p(w.dev_ptr), 2, n, out))
    pt = reef_b200.backend._pt_from
    assert pt(out.raw[:64]) == cport.msm("vesta", gens, W, threads=cport.max_threads())
    assert pt(out.raw[64:]) == cport.msm("vesta", gens, T, threads=cport.max_threads())
    with pytest.raises(ReefError):
        lib_rc = lib.reef_msm_rows_dev(ctx._h, b._h, C.c_void_p(w.dev_ptr), 2, n + 1, out)      # more columns than generators
        check(lib_rc)
    b.free()
    w.free()


def test_generator_levels_are_shared_between_contexts(ctx):
    hits, entries = C.c_uint64(), C.c_uint64()
    gens = WL.generators("vesta", 1 << 10)
    check(lib.reef_bases_cache_stats(C.byref(hits), C.byref(entries)))
    h0, e0 = hits.value, entries.value
    other = reef_b200.Context(0)
    b1 = reef_b200.Bases(ctx, "vesta", gens)
    b2 = reef_b200.Bases(other, "vesta", gens)             # same key on another context / stream: a cache hit
    b3 = reef_b200.Bases(other, "vesta", gens, 32)          # other scalar width: its own levels
    check(lib.reef_bases_cache_stats(C.byref(hits), C.byref(entries)))
    assert hits.value == h0 + 1 and entries.value == e0 + 2
    sc = [random.Random(8).randrange(1 << 250) for _ in range(1 << 10)]
    exp = cport.msm("vesta", gens, sc, threads=cport.max_threads())
    assert b1.msm(sc) == exp and b2.msm(sc) == exp
    b1.free()                                               # the shared levels survive their first owner
    assert b2.msm(sc) == exp
    assert b3.msm_u32(np.arange(1 << 10, dtype=np.uint32)) == cport.msm("vesta", gens, list(range(1 << 10)), threads=cport.max_threads())
    b2.free()
    b3.free()
    check(lib.reef_bases_cache_stats(C.byref(hits), C.byref(entries)))
    assert entries.value == e0
    other.close()


def test_async_table_upload_overlaps_the_first_absorb_and_changes_nothing(ctx):
    """reef_table_upload_u32_async: the copy is queued on the context's copy stream, the sum-check's first absorb runs
    beside it, the first sweep (and every other consumer) waits for it; results identical to the synchronous upload"""
    import torch
    rnd = random.Random(12)
    for ell in (9, 16):
        n = 1 << ell
        pinned = torch.from_numpy(np.random.default_rng(ell).integers(0, 131, size=n, dtype=np.uint32).astype(np.int32)).pin_memory()
        codes = pinned.numpy().view(np.uint32)
        q = [rnd.randrange(n) for _ in range(3)]
        v = [int(codes[i]) for i in q]
        exp = cport.wit_nlookup_gadget(list(map(int, codes)), q, v, None, None, "nldoc", 9, u32=True)
        for _ in range(3):                                   # buffers come back from the context's cache of freed tables
            t = ctx.table_u32(codes, async_upload=True)
            got = ctx.wit_nlookup_gadget(t, q, v, None, None, "nldoc", 9)
            assert got.rounds == exp["rounds"] and got.next_running_claim == exp["next_running_claim"]
            t.free()
        t = ctx.table_u32(codes, async_upload=True)          # a consumer that is not the sum-check
        x = [rnd.randrange(FQ) for _ in range(ell)]
        t2 = ctx.table_u32(codes)
        assert ctx.verifier_mle_eval(t, x) == ctx.verifier_mle_eval(t2, x)
        t.free()
        t2.free()
