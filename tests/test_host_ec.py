"""CPU tier: the curve formulas of ec.cuh (host instantiation of the shared code) against the
affine-law oracle, including the special cases (doubling through add, inverse points, infinity)."""
import ctypes as C
import random

import pytest

from oracle.curves import PALLAS, VESTA, point_to_bytes, point_from_bytes
from reef_b200._lib import lib


def _op(curve_id, op, P, Q=None, k=None):
    out = C.create_string_buffer(64)
    q = point_to_bytes(Q) if k is None else int(k).to_bytes(8, "little") + bytes(56)
    lib.reef_hosttest_ec_op(curve_id, op, point_to_bytes(P), q, out)
    return point_from_bytes(out.raw)


@pytest.mark.parametrize("cid,curve", [(0, PALLAS), (1, VESTA)])
def test_group_law(cid, curve):
    rnd = random.Random(cid)
    G = curve.gen
    pts = [curve.mul(rnd.randrange(1, curve.order), G) for _ in range(12)]
    for i in range(0, 12, 2):
        P, Q = pts[i], pts[i + 1]
        assert _op(cid, 0, P, Q) == curve.add(P, Q)                 # full add
        assert _op(cid, 1, P, Q) == curve.add(P, Q)                 # mixed add
        assert _op(cid, 3, P, Q) == curve.add(P, curve.neg(Q))      # mixed add, negated
        assert _op(cid, 2, P) == curve.add(P, P)                    # double
        assert _op(cid, 0, P, P) == curve.add(P, P)                 # add detects P == Q
        assert _op(cid, 1, P, P) == curve.add(P, P)
        assert _op(cid, 0, P, curve.neg(P)) is None                 # P + (-P) = infinity
        assert _op(cid, 3, P, P) is None
        assert _op(cid, 0, P, None) == P and _op(cid, 1, P, None) == P
        assert _op(cid, 0, None, Q) == Q and _op(cid, 1, None, Q) == Q
        k = rnd.randrange(1 << 20)
        assert _op(cid, 4, P, k=k) == curve.mul(k, P)
    assert _op(cid, 4, G, k=0) is None
    assert _op(cid, 2, None) is None
