"""CPU tier: the composed provers / verifiers of oracle/snark.py are self-consistent -- an honest proof verifies, a
tampered one does not, the C-speed IPA formulation (no explicit generator folding) yields the same proof as the
naive one, and the CAP circuit R1CS (Poseidon H2 = calc_d) is satisfied by its witness."""
import random

import pytest

from oracle import cport, snark as N
from oracle.curves import PALLAS, VESTA
from oracle.fields import FP, FQ
from oracle.poseidon import calc_d


def cmsm(curve):
    name = "pallas" if curve is PALLAS else "vesta"
    return lambda sc, pts: cport.msm(name, pts, sc, threads=cport.max_threads())


@pytest.mark.parametrize("curve", [PALLAS, VESTA])
def test_ipa_round_trip_and_fast_form(curve):
    p = curve.order
    rnd = random.Random(4)
    n = 16
    gens = curve.multiples(n)
    gen_c = curve.mul(987654321, gens[0])
    a = [rnd.randrange(p) for _ in range(n)]
    b = [rnd.randrange(p) for _ in range(n)]
    proof = N.ipa_prove(curve, gens, gen_c, a, b, N.Transcript(b"ipa", p))
    fast = N.ipa_prove(curve, gens, gen_c, a, b, N.Transcript(b"ipa", p), cmsm(curve))
    assert proof == fast
    comm, c = curve.msm(a, gens), N.inner(a, b, p)
    assert N.ipa_verify(curve, gens, gen_c, comm, b, c, proof, N.Transcript(b"ipa", p))
    assert N.ipa_verify(curve, gens, gen_c, comm, b, c, proof, N.Transcript(b"ipa", p), cmsm(curve))
    assert not N.ipa_verify(curve, gens, gen_c, comm, b, (c + 1) % p, proof, N.Transcript(b"ipa", p))
    bad = (proof[0], proof[1], (proof[2] + 1) % p)
    assert not N.ipa_verify(curve, gens, gen_c, comm, b, c, bad, N.Transcript(b"ipa", p))


def test_hyrax_prove_eval_round_trip():
    curve, p = PALLAS, FQ
    rnd = random.Random(5)
    rows, cols = 8, 16
    gens = curve.multiples(cols)
    gen_c = curve.mul(31337, gens[0])
    M = [rnd.randrange(131) for _ in range(rows * cols)]
    comms = N.hyrax_commit(curve, gens, M, rows, cols)
    q = [rnd.randrange(p) for _ in range(7)]
    v, proof = N.hyrax_prove_eval(curve, gens, gen_c, M, rows, cols, q, N.Transcript(b"hy", p), cmsm(curve))
    assert v == N.mle_eval(M, q, p)
    assert N.hyrax_verify_eval(curve, gens, gen_c, comms, rows, cols, q, v, proof, N.Transcript(b"hy", p), cmsm(curve))
    assert not N.hyrax_verify_eval(curve, gens, gen_c, comms, rows, cols, q, (v + 1) % p, proof, N.Transcript(b"hy", p), cmsm(curve))


def _random_instance(p, rnd, num_cons=16, num_vars=16):
    """a satisfied relaxed R1CS instance with random sparse rows (u and E non-trivial)"""
    W = [rnd.randrange(p) for _ in range(num_vars)]
    u, X = rnd.randrange(p), [rnd.randrange(p)]
    shape = N.R1CSShape(num_cons, num_vars, 1, [], [], [])
    z = shape.z(W, u, X)
    A, B, C = [], [], []
    for r in range(num_cons):
        for M in (A, B, C):
            for _ in range(3):
                M.append((r, rnd.randrange(num_vars + 2), rnd.randrange(p)))
    shape.A, shape.B, shape.C = A, B, C
    az, bz, cz = (shape.mul(M, z, p) for M in (A, B, C))
    E = [(a * b - u * c) % p for a, b, c in zip(az, bz, cz)]
    assert shape.is_sat(W, E, u, X, p)
    return shape, W, E, u, X


@pytest.mark.parametrize("curve", [PALLAS, VESTA])
def test_relaxed_r1cs_snark_round_trip(curve):
    p = curve.order
    rnd = random.Random(6)
    shape, W, E, u, X = _random_instance(p, rnd)
    gens = curve.multiples(16)
    gen_c = curve.mul(424242, gens[0])
    cW, cE = curve.msm(W, gens), curve.msm(E, gens)
    proof = N.snark_prove(curve, shape, gens, gen_c, cW, cE, W, E, u, X, N.Transcript(b"snark", p), cmsm(curve))
    assert N.snark_verify(curve, shape, gens, gen_c, cW, cE, u, X, proof, N.Transcript(b"snark", p), cmsm(curve))
    assert not N.snark_verify(curve, shape, gens, gen_c, cW, cE, u, [(X[0] + 1) % p], proof, N.Transcript(b"snark", p), cmsm(curve))
    W2 = list(W)
    W2[3] = (W2[3] + 1) % p
    bad = N.snark_prove(curve, shape, gens, gen_c, curve.msm(W2, gens), cE, W2, E, u, X, N.Transcript(b"snark", p), cmsm(curve))
    assert not N.snark_verify(curve, shape, gens, gen_c, curve.msm(W2, gens), cE, u, X, bad, N.Transcript(b"snark", p), cmsm(curve))


def test_cap_circuit_r1cs_is_calc_d():
    shape, W, X = N.poseidon_h2_r1cs(12345, 67890, FQ)
    assert X == [calc_d(12345, 67890)]
    assert shape.num_cons == 512 and 288 < len(shape.C) <= 290          # 96 S-boxes x 3 + the digest binding
    assert shape.is_sat(W, [0] * shape.num_cons, 1, X, FQ)
    W[5] = (W[5] + 1) % FQ
    assert not shape.is_sat(W, [0] * shape.num_cons, 1, X, FQ)


def test_oracle_nifs_fold_keeps_the_relaxed_instance_satisfied():
    """algebra pin of oracle.snark.nifs_prove (the checker of tests/test_gpu_nifs.py): two chained folds; the folded
    pair satisfies the relaxed R1CS and its commitments open to the folded vectors; the challenge has 128 bits"""
    import random
    from oracle import snark as N
    from oracle.curves import PALLAS
    curve, p = PALLAS, PALLAS.order
    rnd = random.Random(11)
    nc, nv, num_io = 4, 8, 2
    A = [(i, rnd.randrange(nv - nc), rnd.randrange(1, p)) for i in range(nc)] + [(i, nv + 1, 3) for i in range(nc)]
    B = [(i, rnd.randrange(nv - nc), rnd.randrange(1, p)) for i in range(nc)] + [(i, nv, 1) for i in range(nc)]
    Cm = [(i, nv - nc + i, 1) for i in range(nc)]
    shape = N.R1CSShape(nc, nv, num_io, A, B, Cm)
    gens = curve.multiples(nv)

    def fresh():
        X = [rnd.randrange(p) for _ in range(num_io)]
        z = [rnd.randrange(p) for _ in range(nv - nc)] + [0] * nc + [1] + X + [0] * (nv - 1 - num_io)
        for i in range(nc):
            z[nv - nc + i] = sum(v * z[c] for r, c, v in A if r == i) * sum(v * z[c] for r, c, v in B if r == i) % p
        Wv = z[:nv]
        assert shape.is_sat(Wv, [0] * nc, 1, X, p)
        return {"comm_W": curve.msm(Wv, gens), "X": X}, {"W": Wv}

    U1, W1 = fresh()
    U1, W1 = dict(U1, comm_E=None, u=1), dict(W1, E=[0] * nc)
    for _ in range(2):
        U2, W2 = fresh()
        comm_T, r, U1, W1 = N.nifs_prove(curve, shape, gens, 12345, U1, W1, U2, W2)
        assert 0 < r < (1 << 128)
        assert shape.is_sat(W1["W"], W1["E"], U1["u"], U1["X"], p)
        assert U1["comm_W"] == curve.msm(W1["W"], gens) and U1["comm_E"] == curve.msm(W1["E"], gens[:nc])
