"""Pin the oracle against every known-answer test the reference holds for this path
 and against upstream neptune KATs.  CPU only."""
import random

import pytest

from oracle import poseidon as P
from oracle.fields import BLS_FR, FQ
from oracle.merkle import MerkleCommitment
from oracle.mle import (gen_eq_table, gen_eq_table_fast, linear_mle_product, mle_eval_fast,
                        prover_mle_partial_eval, verifier_mle_eval)
from oracle.nlookup import combined_qs, doc_transform, logmn, wit_nlookup_gadget, ASCII_AB, DNA_AB


# ----------------------------------------------------------------------------- r1cs.rs:2517-2578
def test_mle_partial_reference_kat():
    table = [1, 3, 8, 2, 9, 5, 13, 4]
    for x1 in (0, 1, -1):
        for x2 in (0, 1, -1):
            for x3 in (0, 1, -1):
                coeff, con = prover_mle_partial_eval(table, [x1, x2, x3], list(range(8)), True, None)
                holes = (x1 == -1) + (x2 == -1) + (x3 == -1)
                if holes == 1:
                    if x1 == -1:
                        assert (coeff + con) % FQ == table[4 + x2 * 2 + x3] and con == table[x2 * 2 + x3]
                    elif x2 == -1:
                        assert (coeff + con) % FQ == table[x1 * 4 + 2 + x3] and con == table[x1 * 4 + x3]
                    else:
                        assert (coeff + con) % FQ == table[x1 * 4 + x2 * 2 + 1] and con == table[x1 * 4 + x2 * 2]
                elif holes == 0:
                    assert con == table[x1 * 4 + x2 * 2 + x3]


# ----------------------------------------------------------------------------- r1cs.rs:2411-2515
def test_mle_linear_basic_reference_kat():
    evals = [2, 3, 5, 7, 9, 13, 17, 19]
    table = list(evals)
    qs = [2, 1, 7]
    last_q = [2, 3, 5]
    claims = [3, 9, 27, 81]
    term = sum(evals[qs[i]] * claims[i] for i in range(3))
    eq_a = gen_eq_table(claims, qs, list(reversed(last_q)))
    _, running_v = prover_mle_partial_eval(evals, last_q, list(range(8)), True, None)
    term += running_v * claims[3]
    claim = sum(t * e for t, e in zip(evals, eq_a))
    assert term % FQ == claim % FQ
    sponge = P.Sponge()
    sponge.start([("A", 3), ("S", 1)] * 3)
    sc_rs = []
    for i in range(1, 4):
        r_i, xsq, x, con = linear_mle_product(evals, eq_a, 3, i, sponge)
        assert claim % FQ == (2 * con + x + xsq) % FQ
        claim = (xsq * r_i * r_i + x * r_i + con) % FQ
        sc_rs.append(r_i)
    _, next_running_v = prover_mle_partial_eval(table, sc_rs, list(range(8)), True, None)
    _, eq_term = prover_mle_partial_eval(claims, sc_rs, qs, False, last_q)
    assert claim == eq_term * next_running_v % FQ
    sponge.finish()
    # the fully folded tables are the evaluations
    assert evals[0] == next_running_v and claim == evals[0] * eq_a[0] % FQ


# ----------------------------------------------------------------------------- merkle_tree.rs:209-257
@pytest.mark.parametrize("doc", [[2, 3, 4, 5, 6, 7, 8], [5], [1, 2], [9, 8, 7], list(range(1, 18))])
def test_make_mt_reference_kat(doc):
    mc = MerkleCommitment(doc)
    for q in range(len(doc)):
        w = mc.path_wits(q)
        l, oi, o = w[0]
        h = P.hash_once([q, doc[q], oi, o] if l else [oi, o, q, doc[q]])
        for (l, _, o) in w[1:]:
            h = P.hash_once([h, o] if l else [o, h])
        assert h == mc.commitment


# ----------------------------------------------------------------------------- neptune upstream KATs
def _limbs(l):
    return sum(x << (64 * i) for i, x in enumerate(l))


@pytest.mark.parametrize("arity,expected", [
    (2, [0x2e203c369a02e7ff, 0xa6fba9339d05a69d, 0x739e0fd902efe161, 0x396508d75e76a56b]),
    (4, [0x019814ff6662075d, 0xfb6b4605bf1327ec, 0x00db3c6579229399, 0x58a54b10a9e5848a]),
])
def test_neptune_hash_values_bls12_381(arity, expected):
    """neptune `poseidon::test::hash_values` (Strength::Standard, HashType::MerkleTree,
    preimage 0..arity): the generic algorithm (Grain constants with sbox flag 1, Cauchy MDS,
    round numbers, output lane 1) must reproduce the published digests."""
    t = arity + 1
    state = [(1 << arity) - 1] + list(range(arity))
    assert P.permute(state, BLS_FR, t)[1] == _limbs(expected)


def test_neptune_iopattern_tag_values():
    """neptune `sponge::api::test::test_tag_values`."""
    assert P.io_pattern_tag([], 0) == 0
    assert P.io_pattern_tag([], 123) == 340282366920938463463374607431768191899
    assert P.io_pattern_tag([("A", 2), ("S", 2)], 0) == 340282366920938463463374607090318361668
    # runs of the same kind are merged
    assert P.io_pattern_tag([("A", 1), ("A", 1), ("S", 2)], 0) == P.io_pattern_tag([("A", 2), ("S", 2)], 0)


def test_round_numbers_match_in_tree_constraint_counts():
    # costs.rs:132: 288 = (8*5 + 56) * 3 constraints per width-5 permutation
    rf, rp = P.round_numbers(5)
    assert (rf, rp) == (8, 56) and (rf * 5 + rp) * 3 == 288
    assert P.round_numbers(3) == (8, 55)


# ----------------------------------------------------------------------------- host logic
def test_logmn_f32_semantics():
    assert logmn(1) == 1 and logmn(2) == 1 and logmn(3) == 2 and logmn(11) == 4
    assert logmn(1 << 20) == 20 and logmn((1 << 20) + 2) == 21
    assert logmn((1 << 23) + 1) == 23          # f32 rounding quirk (costs.rs:13)
    # consequence: a document of exactly 2^22 (or 2^23) characters cannot be padded -- the reference's
    # `base.pow(logmn(len)) - len` underflows (framework.rs:1007); 2^21 + 2 is still rounded up correctly
    assert logmn((1 << 21) + 2) == 22 and logmn((1 << 22) + 2) == 22 and logmn((1 << 22) + 66) == 23


def test_doc_transform_ascii_dna():
    u = doc_transform(ASCII_AB, "aaaaaaaab")
    assert len(u) == 16 and u[:9] == [97] * 8 + [98]
    assert u[9] == 130 and u[10] == 129 and u[11:] == [0] * 5       # EOF = |ab|+2, EPSILON = |ab|+1
    u = doc_transform(DNA_AB, "ACGT")
    assert u == [0, 1, 2, 3, 6, 5, 0, 0]
    with pytest.raises(ValueError):
        doc_transform(DNA_AB, "ACGX")
    # char 26 inside an ascii document maps to EOF's code (insert overwrote the entry)
    assert doc_transform(ASCII_AB, chr(26))[0] == 130


def test_combined_q_quirks():
    assert combined_qs([5], 3) == [0b01]                 # bits MSB-first 1,0,(1 skipped at flush)
    # 15 lookups * 17 bits = 255 bits -> 2 limbs, second pass repeats the first 254 bits
    q = [random.Random(3).randrange(1 << 17) for _ in range(15)]
    c = combined_qs(q, 17)
    assert len(c) == 2 and c[0] == c[1] and c[0] < 1 << 254


def test_fast_variants_equal_reference_shape():
    rnd = random.Random(5)
    for ell in (1, 2, 5, 7):
        n = 1 << ell
        m = rnd.randrange(0, 5)
        rs = [rnd.randrange(FQ) for _ in range(m + 1)]
        qs = [rnd.randrange(n) for _ in range(m)]
        lq = [rnd.randrange(FQ) for _ in range(ell)]
        assert gen_eq_table(rs, qs, lq) == gen_eq_table_fast(rs, qs, lq)
        tab = [rnd.randrange(FQ) for _ in range(n)]
        x = [rnd.randrange(FQ) for _ in range(ell)]
        assert verifier_mle_eval(tab, x) == mle_eval_fast(tab, x)


@pytest.mark.parametrize("tag", ["nl", "nldoc", "nlhybrid"])
def test_nlookup_identities(tag):
    """Claim identity of the nlookup (SURVEY A.6 i) and sum-check consistency per round."""
    rnd = random.Random(11)
    ell, m = 6, 3
    n = 1 << ell
    table = [rnd.randrange(200) for _ in range(n)]
    q = [rnd.randrange(n) for _ in range(m)]
    v = [table[i] for i in q]
    prev_q = [rnd.randrange(FQ) for _ in range(ell)]
    prev_v = mle_eval_fast(table, prev_q)
    r = wit_nlookup_gadget(table, q, v, prev_q, prev_v, tag, doc_hash=12345)
    rs = [pow(r["claim_r"], k + 1, FQ) for k in range(m + 1)]
    claim = (sum(rs[k] * v[k] for k in range(m)) + rs[m] * prev_v) % FQ
    for (sc_r, xsq, x, con) in r["rounds"]:
        assert claim == (2 * con + x + xsq) % FQ
        claim = (xsq * sc_r * sc_r + x * sc_r + con) % FQ
    assert claim == r["sc_last_claim"]
    assert r["next_running_claim"] == mle_eval_fast(table, r["next_running_q"]) == r["folded_t0"]
    assert r["sc_last_claim"] == r["folded_t0"] * r["folded_eq0"] % FQ
    fast = wit_nlookup_gadget(table, q, v, prev_q, prev_v, tag, doc_hash=12345, fast=True)
    assert fast == r
