"""GPU parity (through the C ABI): Pippenger MSM on Pallas/Vesta against the naive oracle.
An MSM has exactly one correct affine answer, so equality here is bit-exact parity."""
import random

import numpy as np
import pytest

import reef_b200
from oracle.curves import PALLAS, VESTA
from oracle.fields import FQ

pytestmark = pytest.mark.gpu
CUR = {"pallas": PALLAS, "vesta": VESTA}


@pytest.mark.parametrize("curve", ["pallas", "vesta"])
@pytest.mark.parametrize("n", [1, 2, 33, 200])
def test_msm_small_vs_naive(ctx, curve, n):
    cv = CUR[curve]
    rnd = random.Random(n)
    pts = [cv.mul(rnd.randrange(1, cv.order), cv.gen) for _ in range(n)]
    sc = [rnd.randrange(cv.order) for _ in range(n)]
    sc[0] = cv.order - 1
    if n > 2:
        sc[1], sc[2] = 0, 1
    assert ctx.msm(curve, pts, sc) == cv.msm(sc, pts)


def test_msm_degenerate_inputs(ctx):
    cv = PALLAS
    G = cv.gen
    pts = [G, G, cv.neg(G), None, cv.mul(7, G)]
    b = ctx.bases("pallas", pts)
    assert b.msm([0, 0, 0, 0, 0]) is None                       # all-zero scalars
    assert b.msm([5, 5, 10, 123, 0]) is None                    # cancels to infinity
    assert b.msm([1, 1, 0, 99, 0]) == cv.mul(2, G)              # equal points: doubling inside a bucket
    assert b.msm([3]) == cv.mul(3, G)                           # fewer scalars than bases
    assert b.msm([]) is None
    with pytest.raises(reef_b200.ReefError) as e:
        b.msm([1] * 6)                                          # more scalars than generators
    assert e.value.code == 3
    with pytest.raises(reef_b200.ReefError) as e:
        ctx.bases("pallas", [(1, 1)])                           # not on the curve
    assert e.value.code == 1


def _multiples(cv, n):
    return cv.multiples(n)


@pytest.mark.parametrize("curve,n", [("pallas", 1 << 12), ("vesta", 3000), ("pallas", 1 << 15)])
def test_msm_known_discrete_logs(ctx, curve, n):
    """bases = k*G for k = 1..n, so  sum s_k * (k G) = (sum s_k k mod r) G."""
    cv = CUR[curve]
    rnd = random.Random(n)
    pts = _multiples(cv, n)
    b = ctx.bases(curve, pts)
    for trial in range(2):
        sc = [rnd.randrange(cv.order) for _ in range(n)]
        exp = cv.mul(sum(s * (k + 1) for k, s in enumerate(sc)) % cv.order, cv.gen)
        assert b.msm(sc) == exp
    # adversarial histogram: every scalar equal (Reef's all-'a' document) -> one bucket per window
    s = rnd.randrange(cv.order)
    assert b.msm([s] * n) == cv.mul(s * (n * (n + 1) // 2) % cv.order, cv.gen)
    # small scalars through the u32 entry (document codes / witness bits)
    small = [rnd.randrange(131) for _ in range(n)]
    assert b.msm_u32(small) == cv.mul(sum(s * (k + 1) for k, s in enumerate(small)) % cv.order, cv.gen)
    # window-sharded partial sums + combine == whole MSM (the multi-GPU path on one GPU)
    import torch
    sc = [rnd.randrange(cv.order) for _ in range(n)]
    raw = b"".join(int(x).to_bytes(32, "little") for x in sc)
    dev = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    W = b.windows
    cuts = [0, W // 3, W // 2, W]
    parts = b"".join(b.msm_partial_dev(dev.data_ptr(), n, cuts[i], cuts[i + 1]) for i in range(3))
    exp = cv.mul(sum(s * (k + 1) for k, s in enumerate(sc)) % cv.order, cv.gen)
    assert b.combine(parts) == exp
    assert b.msm_dev(dev.data_ptr(), n) == exp
    b.free()


def test_hyrax_row_commit_small_vs_naive(ctx):
    """commitment.rs:187 `hyrax_gen.commit`: row r = sum_j M[r][j] G_j + blind_r H."""
    cv = PALLAS
    rnd = random.Random(5)
    rows, cols = 4, 8
    gens = [cv.mul(rnd.randrange(1, cv.order), cv.gen) for _ in range(cols + 1)]
    b = ctx.bases("pallas", gens)
    M = np.asarray([[rnd.randrange(131) for _ in range(cols)] for _ in range(rows)], dtype=np.uint32)
    M[1, :] = 0                                                     # an all-zero row commits to blind*H only
    blinds = [rnd.randrange(cv.order) for _ in range(rows)]
    got = b.msm_rows(M, rows, cols, 8, blinds)
    for r in range(rows):
        assert got[r] == cv.msm([int(x) for x in M[r]] + [blinds[r]], gens), r
    got = b.msm_rows(M, rows, cols, 8, None)
    assert got[1] is None
    assert got[2] == cv.msm([int(x) for x in M[2]], gens[:cols])
    # field-element matrix entries (full width)
    Mf = [rnd.randrange(cv.order) for _ in range(rows * cols)]
    got = b.msm_rows(Mf, rows, cols, 0, blinds)
    for r in range(rows):
        assert got[r] == cv.msm(Mf[r * cols:(r + 1) * cols] + [blinds[r]], gens), r


@pytest.mark.parametrize("ell,bits", [(17, 8), (15, 21)])
def test_hyrax_row_commit_document_shape(ctx, ell, bits):
    """cfg-2 shape: N = 2^17 document codes as a 256 x 512 matrix (left = ell/2, right = ell - left),
    generators k*G so each row has a closed-form answer; includes the all-'a' adversarial rows."""
    cv = PALLAS
    rnd = random.Random(ell)
    left = ell // 2
    rows, cols = 1 << left, 1 << (ell - left)
    gens = cv.multiples(cols + 1)
    b = ctx.bases("pallas", gens)
    M = np.random.default_rng(ell).integers(0, 1 << bits, size=(rows, cols), dtype=np.uint32)
    M[0, :] = 97                                                    # 'a' * cols
    M[1, :] = 0
    blinds = [rnd.randrange(cv.order) for _ in range(rows)]
    got = b.msm_rows(M, rows, cols, bits, blinds)
    w = np.arange(1, cols + 1, dtype=object)
    for r in range(rows):                                           # every row: warp-per-row scaling / inversion kernels
        k = (int((M[r].astype(object) * w).sum()) + blinds[r] * (cols + 1)) % cv.order
        assert got[r] == cv.mul(k, cv.gen), r


def test_msm_throughput_shapes_known_discrete_logs(ctx):
    """n = 2^18 takes the large-instance code paths (16-entry first pass, 4-lane combine groups, bucket sums
    ahead of the bit-decomposition tree): same closed form, plus the all-equal-scalar histogram and a
    Vesta run."""
    import numpy as np
    import torch
    n = 1 << 18
    for curve in ("pallas", "vesta"):
        cv = CUR[curve]
        pts = cv.multiples(n)
        b = ctx.bases(curve, b"".join(int(P[0]).to_bytes(32, "little") + int(P[1]).to_bytes(32, "little") for P in pts))
        assert b.window_bits >= 14
        raw = np.random.default_rng(18).integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
        raw[:, 3] &= (1 << 61) - 1                                    # < 2^253 < both group orders
        words = raw.astype(object)
        sc = [int(w[0]) | (int(w[1]) << 64) | (int(w[2]) << 128) | (int(w[3]) << 192) for w in words]
        dev = torch.from_numpy(raw.view(np.int64)).cuda()
        exp = cv.mul(sum(s * (k + 1) for k, s in enumerate(sc)) % cv.order, cv.gen)
        assert b.msm_dev(dev.data_ptr(), n) == exp
        if curve == "pallas":
            s = sc[7]
            eq = np.tile(raw[7], (n, 1))
            dev2 = torch.from_numpy(np.ascontiguousarray(eq).view(np.int64)).cuda()
            assert b.msm_dev(dev2.data_ptr(), n) == cv.mul(s * (n * (n + 1) // 2) % cv.order, cv.gen)
        b.free()


def test_background_context_keeps_off_the_reserved_sms_and_stays_exact():
    """reef_init_prio(.., 0, ..) with REEF_RESERVE_SMS=12: the background context's stream lives in a green-context
    partition that leaves 12 SMs to the latency-critical contexts; results are unchanged -- single MSM, the W / T pair as
    two rows, u32 rows with blinds"""
    import numpy as np
    import workloads as WL
    from oracle import cport
    import os
    hi = reef_b200.Context(0, latency_critical=True)
    bg_half = reef_b200.Context(0, latency_critical=False)      # default: an ordinary low-priority stream
    os.environ["REEF_RESERVE_SMS"] = "12"                       # opt-in: green-context partition of 136 SMs
    try:
        bg = reef_b200.Context(0, latency_critical=False)
    finally:
        del os.environ["REEF_RESERVE_SMS"]
    from reef_b200._lib import lib
    assert lib.reef_ctx_sm_count(hi._h) == lib.reef_ctx_sm_count(bg_half._h)
    assert 0 < lib.reef_ctx_sm_count(bg._h) <= lib.reef_ctx_sm_count(hi._h)
    print("background partition:", lib.reef_ctx_sm_count(bg._h), "of", lib.reef_ctx_sm_count(hi._h), "SMs")
    try:
        n = 1 << 13
        gens = WL.generators("pallas", n + 1)
        rnd = random.Random(21)
        sc = [rnd.randrange(1 << 253) for _ in range(2 * n)]
        for c in (bg, hi, bg_half):
            b = reef_b200.Bases(c, "pallas", gens)
            assert b.msm(sc[:n]) == cport.msm("pallas", gens[:64 * n], sc[:n], threads=cport.max_threads())
            rows = b.msm_rows(sc, 2, n)
            assert rows[0] == cport.msm("pallas", gens[:64 * n], sc[:n], threads=cport.max_threads())
            assert rows[1] == cport.msm("pallas", gens[:64 * n], sc[n:], threads=cport.max_threads())
            codes = np.random.default_rng(3).integers(0, 256, size=(16, 512), dtype=np.uint32)
            blinds = [rnd.randrange(FQ) for _ in range(16)]
            hb = reef_b200.Bases(c, "pallas", gens[:64 * 513], 255)
            got = hb.msm_rows(codes, 16, 512, entry_bits=8, blinds=blinds)
            for r in (0, 15):
                assert got[r] == cport.msm("pallas", gens[:64 * 513], [int(x) for x in codes[r]] + [blinds[r]], threads=cport.max_threads())
            hb.free()
            b.free()
    finally:
        bg.close()
        hi.close()
        bg_half.close()
