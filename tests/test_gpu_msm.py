"""GPU parity (through the C ABI): Pippenger MSM on Pallas/Vesta against the naive oracle.
An MSM has exactly one correct affine answer, so equality here is bit-exact parity."""
import random

import numpy as np
import pytest

import reef_b200
from oracle.curves import PALLAS, VESTA

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
