"""GPU tier: the composed provers (reef_b200/snark.py: every field / curve operation in libreef_b200, transcript on the
host) against oracle/snark.py -- the proofs must be identical under the same transcript and the oracle's verifiers must
accept them: the inner-product argument, Hyrax `prove_eval` on the resident document table, and the relaxed-R1CS
SNARK on both curves, including the CAP circuit (Poseidon H2 = calc_d, commitment.rs:538-622) as an R1CS instance."""
import random

import numpy as np
import pytest

import workloads as W
from oracle import cport, snark as N
from oracle.curves import PALLAS, VESTA
from oracle.fields import FP, FQ
from reef_b200 import snark as G

pytestmark = pytest.mark.gpu


def cmsm(curve):
    name = "pallas" if curve is PALLAS else "vesta"
    return lambda sc, pts: cport.msm(name, pts, sc, threads=cport.max_threads())


def _gens(name, n):
    raw = W.generators(name, n)
    return [(int.from_bytes(raw[i * 64:i * 64 + 32], "little"), int.from_bytes(raw[i * 64 + 32:i * 64 + 64], "little")) for i in range(n)]


@pytest.mark.parametrize("name,n", [("pallas", 2), ("pallas", 64), ("vesta", 256), ("pallas", 2048)])
def test_ipa_matches_oracle_and_verifies(ctx, name, n):
    curve = PALLAS if name == "pallas" else VESTA
    p = curve.order
    rnd = random.Random(n)
    gens = _gens(name, n)
    gen_c = curve.mul(rnd.randrange(p), gens[0])
    a = [rnd.randrange(p) for _ in range(n)]
    b = [rnd.randrange(p) for _ in range(n)]
    got = G.ipa_prove(ctx, name, gens, gen_c, a, b, N.Transcript(b"ipa", p))
    exp = N.ipa_prove(curve, gens, gen_c, a, b, N.Transcript(b"ipa", p), cmsm(curve))
    assert got == exp
    comm, c = cmsm(curve)(a, gens), N.inner(a, b, p)
    assert N.ipa_verify(curve, gens, gen_c, comm, b, c, got, N.Transcript(b"ipa", p), cmsm(curve))


@pytest.mark.parametrize("name,n", [("pallas", 1), ("pallas", 2), ("vesta", 64), ("pallas", 4096)])
def test_ipa_over_registered_generators_gives_the_same_proof(ctx, name, n):
    """reef_ipa_begin_bases: the generators are never folded (every round = one two-row MSM over the registered window
    levels, the fold lives in per-generator weights) -- same L, R, a_hat as the folding session and the oracle, and the
    final folded generator / b_hat agree with the folding session"""
    import reef_b200
    curve = PALLAS if name == "pallas" else VESTA
    p = curve.order
    rnd = random.Random(1000 + n)
    gens = _gens(name, n + 3)                          # more registered generators than the argument uses
    gen_c = curve.mul(rnd.randrange(p), gens[0])
    a = [rnd.randrange(p) for _ in range(n)]
    b = [rnd.randrange(p) for _ in range(n)]
    bases = reef_b200.Bases(ctx, name, gens, 255)
    try:
        got = G.ipa_prove(ctx, name, bases, gen_c, a, b, N.Transcript(b"ipa", p))
        exp = N.ipa_prove(curve, gens[:n], gen_c, a, b, N.Transcript(b"ipa", p), cmsm(curve))
        assert got == exp
        comm, c = cmsm(curve)(a, gens[:n]), N.inner(a, b, p)
        assert N.ipa_verify(curve, gens[:n], gen_c, comm, b, c, got, N.Transcript(b"ipa", p), cmsm(curve))
        # finish(): a_hat, b_hat, G_hat of both session kinds agree
        fins = []
        for g in (bases, gens[:n]):
            s = G.Ipa(ctx, name, g, gen_c, a, b)
            tr = N.Transcript(b"ipa", p)
            m = n
            while m > 1:
                L, R = s.round()
                tr.absorb_point(b"L", L)
                tr.absorb_point(b"R", R)
                r = tr.squeeze(b"r")
                s.fold(r, pow(r, -1, p))
                m //= 2
            fins.append(s.finish())
            s.free()
        assert fins[0] == fins[1]
    finally:
        bases.free()


def test_hyrax_prove_eval_on_the_document_table(ctx):
    """configs[1] shape: 2^17 document codes as a 256 x 512 matrix, commitment by rows, opening at a random point"""
    ab, cps = W.document("cfg2")
    udoc = W.encode(ab, cps)
    ell = W.logmn(len(udoc))
    rows, cols = W.hyrax_dims(ell)
    gens = _gens("pallas", cols)
    rnd = random.Random(17)
    gen_c = PALLAS.mul(rnd.randrange(FQ), gens[0])
    q = [rnd.randrange(FQ) for _ in range(ell)]
    t = ctx.table_u32(udoc)
    b = ctx.bases("pallas", gens, 255)
    try:
        comms = b.msm_rows(udoc.reshape(rows, cols), rows, cols, entry_bits=8)
        v, proof = G.hyrax_prove_eval(ctx, t, rows, cols, gens, gen_c, q, N.Transcript(b"hy", FQ))
        assert v == ctx.verifier_mle_eval(t, q)
        assert N.hyrax_verify_eval(PALLAS, gens, gen_c, comms, rows, cols, q, v, proof, N.Transcript(b"hy", FQ), cmsm(PALLAS))
    finally:
        b.free()
        t.free()


def _random_instance(p, rnd, num_cons, num_vars):
    Wt = [rnd.randrange(p) for _ in range(num_vars)]
    u, X = rnd.randrange(p), [rnd.randrange(p)]
    shape = N.R1CSShape(num_cons, num_vars, 1, [], [], [])
    z = shape.z(Wt, u, X)
    A, B, Cm = [], [], []
    for r in range(num_cons):
        for M in (A, B, Cm):
            for _ in range(3):
                M.append((r, rnd.randrange(num_vars + 2), rnd.randrange(p)))
    shape.A, shape.B, shape.C = A, B, Cm
    az, bz, cz = (shape.mul(M, z, p) for M in (A, B, Cm))
    E = [(a * b - u * c) % p for a, b, c in zip(az, bz, cz)]
    return shape, Wt, E, u, X


@pytest.mark.parametrize("name", ["pallas", "vesta"])
def test_relaxed_r1cs_snark_matches_oracle_and_verifies(ctx, name):
    curve = PALLAS if name == "pallas" else VESTA
    p = curve.order
    rnd = random.Random(6)
    shape, Wt, E, u, X = _random_instance(p, rnd, 64, 32)
    gens = _gens(name, 64)
    gen_c = curve.mul(424242, gens[0])
    cW, cE = cmsm(curve)(Wt, gens[:32]), cmsm(curve)(E, gens)
    got = G.snark_prove(ctx, name, shape, gens, gen_c, cW, cE, Wt, E, u, X, N.Transcript(b"snark", p))
    exp = N.snark_prove(curve, shape, gens, gen_c, cW, cE, Wt, E, u, X, N.Transcript(b"snark", p), cmsm(curve))
    assert got == exp
    assert N.snark_verify(curve, shape, gens, gen_c, cW, cE, u, X, got, N.Transcript(b"snark", p), cmsm(curve))


def test_cap_circuit_snark(ctx):
    """cap_prove's statement (d = H2(v, salt)) as a relaxed-R1CS instance (u = 1, E = 0), proved on the GPU"""
    shape, Wt, X = N.poseidon_h2_r1cs(31337, 271828, FQ)
    gens = _gens("pallas", 512)
    gen_c = PALLAS.mul(99, gens[0])
    E = [0] * shape.num_cons
    cW, cE = cmsm(PALLAS)(Wt, gens), None
    got = G.snark_prove(ctx, "pallas", shape, gens, gen_c, cW, cE, Wt, E, 1, X, N.Transcript(b"cap", FQ))
    assert N.snark_verify(PALLAS, shape, gens, gen_c, cW, cE, 1, X, got, N.Transcript(b"cap", FQ), cmsm(PALLAS))
    assert not N.snark_verify(PALLAS, shape, gens, gen_c, cW, cE, 1, [(X[0] + 1) % FQ], got, N.Transcript(b"cap", FQ), cmsm(PALLAS))
