"""GPU parity (through the C ABI) for the Spartan-side sweeps behind CompressedSNARK::prove
(framework.rs:695-698), the R1CS mat-vec of the folding step and the IPA generator fold.
Bit-exact against oracle/spartan.py -- which restates the PUBLISHED upstream nova-snark algorithm
(the fork Reef builds with is not vendored: parity unpinned, see the oracle's header)."""
import random

import pytest

import reef_b200
from oracle import spartan as S
from oracle.curves import PALLAS, VESTA
from oracle.fields import FP, FQ

pytestmark = pytest.mark.gpu
MOD = {"fq": FQ, "fp": FP}


def _run(ctx, tabs, ch, field):
    p = MOD[field]
    claim0, rounds, finals, last = S.prove(tabs, ch, p)
    sc = ctx.sumcheck(tabs, field)
    got_rounds = []
    r_prev = None
    for r in ch:
        got_rounds.append(sc.round(r_prev))
        r_prev = r
    got_finals = sc.final(r_prev)
    sc.free()
    exp_rounds = [[rd[0]] + rd[2:] for rd in rounds]        # eval_1 = claim - eval_0 is the caller's
    assert got_rounds == exp_rounds
    assert got_finals == finals
    return last, finals


@pytest.mark.parametrize("field", ["fq", "fp"])
@pytest.mark.parametrize("kind", [2, 4])
@pytest.mark.parametrize("k", [1, 2, 5, 9, 13])
def test_sumcheck_rounds_match_oracle(ctx, field, kind, k):
    p = MOD[field]
    rnd = random.Random(1000 * kind + k)
    n = 1 << k
    tabs = [[rnd.randrange(p) for _ in range(n)] for _ in range(kind)]
    tabs[0][0], tabs[-1][n - 1] = p - 1, 0                  # extremes
    ch = [rnd.randrange(p) for _ in range(k)]
    last, finals = _run(ctx, tabs, ch, field)
    # the sum-check's own closing identity: last claim == comb(bound values)
    assert last == (finals[0] * finals[1] if kind == 2 else S.comb_cubic(*finals)) % p


def test_sumcheck_outer_shape_eq_az_bz_cz(ctx):
    """The outer Spartan sum-check on a satisfied instance: A = eq(tau, .), D = Az*Bz => claim 0
    and every round polynomial sums to the running claim."""
    p = FQ
    rnd = random.Random(5)
    k = 10
    n = 1 << k
    tau = [rnd.randrange(p) for _ in range(k)]
    eq = [1]
    for t in tau:                                           # MSB-first tensor product
        eq = [e * (1 - t) % p for e in eq] + [e * t % p for e in eq]
    az = [rnd.randrange(p) for _ in range(n)]
    bz = [rnd.randrange(p) for _ in range(n)]
    cz = [a * b % p for a, b in zip(az, bz)]
    ch = [rnd.randrange(p) for _ in range(k)]
    claim0, rounds, finals, last = S.prove([eq, az, bz, cz], ch, p)
    assert claim0 == 0
    _run(ctx, [eq, az, bz, cz], ch, "fq")


def test_sumcheck_argument_checks(ctx):
    with pytest.raises(reef_b200.ReefError) as e:
        ctx.sumcheck([[1, 2, 3], [4, 5, 6]])                # not a power of two
    assert e.value.code == 3
    with pytest.raises(reef_b200.ReefError):
        ctx.sumcheck([[1, 2], [FQ, 1]])                     # non-canonical element
    sc = ctx.sumcheck([[1, 2], [3, 4]])
    with pytest.raises(reef_b200.ReefError):
        sc.round(5)                                         # the first round takes no challenge
    assert sc.round(None) == [3, (2 * 2 - 1) * (2 * 4 - 3) % FQ]
    with pytest.raises(reef_b200.ReefError):
        sc.round(None)                                      # later rounds need one
    assert sc.final(7) == [(1 + 7 * (2 - 1)) % FQ, (3 + 7 * (4 - 3)) % FQ]
    sc.free()


@pytest.mark.parametrize("field", ["fq", "fp"])
def test_r1cs_spmv_matches_oracle(ctx, field):
    p = MOD[field]
    rnd = random.Random(11)
    n_rows, n_cols = 300, 257
    row_ptr, col, vals = [0], [], []
    for r in range(n_rows):
        nnz = [0, 1, 3, 40, 70][r % 5]                      # empty rows, short rows, rows longer than a warp
        for _ in range(nnz):
            col.append(rnd.randrange(n_cols))
            vals.append(rnd.choice([1, p - 1, rnd.randrange(p)]))
        row_ptr.append(len(col))
    z = [rnd.randrange(p) for _ in range(n_cols)]
    assert ctx.r1cs_spmv(row_ptr, col, vals, z, field) == S.spmv(row_ptr, col, vals, z, p)
    with pytest.raises(reef_b200.ReefError) as e:
        ctx.r1cs_spmv([0, 1], [n_cols], [1], z, field)      # column index out of bounds
    assert e.value.code == 3


@pytest.mark.parametrize("curve,cv", [("pallas", PALLAS), ("vesta", VESTA)])
def test_ipa_fold_bases_matches_oracle(ctx, curve, cv):
    rnd = random.Random(21)
    n = 16
    G = [cv.mul(rnd.randrange(1, cv.order), cv.gen) for _ in range(n)]
    G[3] = None                                             # infinity among the generators
    G[4], G[12] = G[5], G[5]                                # equal halves: doubling inside the joint ladder
    r = rnd.randrange(1, cv.order)
    r_inv = pow(r, -1, cv.order)
    assert ctx.ipa_fold_bases(curve, G, r_inv, r) == S.ipa_fold_bases(cv, G, r_inv, r)
    assert ctx.ipa_fold_bases(curve, G, 0, 1) == G[n // 2:]
    assert ctx.ipa_fold_bases(curve, G, cv.order - 1, 0) == [cv.neg(P) for P in G[:n // 2]]
