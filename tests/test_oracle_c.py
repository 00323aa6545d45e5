"""Pin the C restatement (oracle/c) against the Python oracle (itself pinned by the reference's
KATs in test_oracle_kats.py).  CPU only."""
import random

import pytest

from oracle import cport
from oracle import poseidon as P
from oracle.curves import PALLAS, VESTA
from oracle.fields import FQ
from oracle.merkle import MerkleCommitment
from oracle.nlookup import wit_nlookup_gadget


def test_c_poseidon_matches_python():
    rnd = random.Random(1)
    for arity in (2, 4):
        rows = [rnd.randrange(FQ) for _ in range(arity * 20)]
        got = cport.poseidon_hash(rows, arity)
        assert got == [P.hash_once(rows[i * arity:(i + 1) * arity]) for i in range(20)]


@pytest.mark.parametrize("doc", [[2, 3, 4, 5, 6, 7, 8], [5], list(range(37))])
def test_c_merkle_matches_python(doc):
    assert cport.merkle(doc, threads=2) == MerkleCommitment(doc).tree


@pytest.mark.parametrize("ell,m,tag,u32", [(1, 1, "nl", False), (4, 3, "nldoc", True), (7, 5, "nlhybrid", False),
                                           (9, 0, "nl", False), (11, 16, "nldoc", True)])
def test_c_nlookup_matches_python(ell, m, tag, u32):
    rnd = random.Random(ell)
    n = 1 << ell
    table = [rnd.randrange(131 if u32 else FQ) for _ in range(n)]
    q = [rnd.randrange(n) for _ in range(m)]
    v = [table[i] for i in q]
    rq = [rnd.randrange(FQ) for _ in range(ell)]
    rv = rnd.randrange(FQ)
    exp = wit_nlookup_gadget(table, q, v, rq, rv, tag, 99, fast=True)
    got = cport.wit_nlookup_gadget(table, q, v, rq, rv, tag, 99, u32=u32)
    for k in ("claim_r", "rounds", "sc_last_claim", "next_running_claim"):
        assert got[k] == exp[k], k
    exp = wit_nlookup_gadget(table, q, v, None, None, tag, 99, fast=True)
    got = cport.wit_nlookup_gadget(table, q, v, None, None, tag, 99, u32=u32)
    assert got["rounds"] == exp["rounds"] and got["next_running_claim"] == exp["next_running_claim"]


@pytest.mark.parametrize("curve,cv", [("pallas", PALLAS), ("vesta", VESTA)])
def test_c_msm_matches_python(curve, cv):
    rnd = random.Random(3)
    n = 70
    pts = [cv.mul(rnd.randrange(1, cv.order), cv.gen) for _ in range(n)]
    pts[5] = None
    sc = [rnd.randrange(cv.order) for _ in range(n)]
    sc[0], sc[1] = 0, cv.order - 1
    assert cport.msm(curve, pts, sc, threads=2) == cv.msm(sc, pts)
    assert cport.msm(curve, pts[:1], [0]) is None


def test_c_fast_poseidon_schedule_equals_textbook():
    rnd = random.Random(8)
    rows = [rnd.randrange(FQ) for _ in range(4 * 16)]
    slow = cport.poseidon_hash(rows, 4)
    cport.lib().oracle_set_fast_poseidon(1)
    try:
        assert cport.poseidon_hash(rows, 4) == slow
        assert cport.merkle(list(range(21)), threads=1) == MerkleCommitment(list(range(21))).tree
    finally:
        cport.lib().oracle_set_fast_poseidon(0)
