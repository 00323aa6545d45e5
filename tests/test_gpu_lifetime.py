"""Handle lifetime through the C ABI: a context may be shut down while tables, generators, sponge
and sum-check sessions created from it are still alive; the children can then only be freed, in
any order (round-1 bug: reef_table_free after reef_shutdown dereferenced a freed context)."""
import gc

import pytest

import reef_b200
from reef_b200._lib import ReefError
from oracle.curves import PALLAS
from oracle.fields import FQ

pytestmark = pytest.mark.gpu


def _children(c):
    t = c.table([1, 2, 3, 4, 5, 6, 7, 8])
    u = c.table_u32([3, 1, 4, 1, 5, 9, 2, 6])
    b = c.bases("pallas", PALLAS.multiples(8))
    sp = reef_b200.Sponge(c, [("A", 2), ("S", 1)])
    sc = c.sumcheck([[1, 2, 3, 4], [5, 6, 7, 8]])
    sn = reef_b200.ShardedNlookup(c, t, 0, 1, [1], [2], [0, 0, 0], 1, "nl")
    return t, u, b, sp, sc, sn


@pytest.mark.parametrize("order", ["forward", "reverse"])
def test_shutdown_with_live_children(order):
    c = reef_b200.Context(0)
    t, u, b, sp, sc, sn = _children(c)
    assert b.msm([1] * 8) == PALLAS.msm([1] * 8, PALLAS.multiples(8))
    c.close()                                   # reef_shutdown with six live children
    with pytest.raises(ReefError):
        t.download(8)                           # children of a closed context can only be freed
    with pytest.raises(ReefError):
        sc.round(None)
    with pytest.raises(ReefError):
        sp.absorb([1, 2])
    frees = [t.free, u.free, b.free, sp.finish, sc.free, sn.free]
    if order == "reverse":
        frees.reverse()
    for f in frees:
        try:
            f()
        except ReefError:
            pass                                # sponge finish reports the unfinished IOPattern; it still frees
    # a fresh context on the same device works after all of that
    c2 = reef_b200.Context(0)
    assert c2.calc_d(5, 7) < FQ
    c2.close()


def test_finalizer_order_is_irrelevant():
    """What smoke() did in round 1: close the context, let Python finalise the table afterwards."""
    c = reef_b200.Context(0)
    t = c.table_u32([1, 2, 3, 4])
    r = c.wit_nlookup_gadget(t, [1], [2], None, None, "nldoc", 7)
    assert len(r.rounds) == 2
    c.close()
    c.close()                                   # idempotent
    del t
    gc.collect()


def test_round2_handles_outlive_their_context_too():
    """witness buffers, IPA sessions (both kinds) and cached generator levels: freed after reef_shutdown, any order"""
    from reef_b200 import snark as G
    c = reef_b200.Context(0)
    w = reef_b200.Witness(c, 64)
    w.set_small([0, 1], [1, 2])
    pts = PALLAS.multiples(9)
    b = c.bases("pallas", pts[:8], 255)
    b2 = c.bases("pallas", pts[:8], 255)                    # second handle on the same cached levels
    s1 = G.Ipa(c, "pallas", b, pts[8], [1, 2, 3, 4, 5, 6, 7, 8], [8, 7, 6, 5, 4, 3, 2, 1])
    s2 = G.Ipa(c, "pallas", pts[:8], pts[8], [1, 2, 3, 4, 5, 6, 7, 8], [8, 7, 6, 5, 4, 3, 2, 1])
    assert s1.round() == s2.round()
    c.close()
    with pytest.raises(ReefError):
        w.read(0, 2)
    with pytest.raises(ReefError):
        s1.round()
    for f in (b.free, s2.free, w.free, s1.free, b2.free):
        f()
    c2 = reef_b200.Context(0)
    b3 = c2.bases("pallas", pts[:8], 255)                   # the levels were released with their last handle: registered anew
    assert b3.msm([1] * 8) == PALLAS.msm([1] * 8, pts[:8])
    b3.free()
    c2.close()
