"""GPU tier: `NIFS::prove` -- the fold inside every prove_step (framework.rs:668-675) -- composed from the library's
sparse products, cross-term kernel, MSM, PoseidonRO and vector folds (reef_b200/snark.py) against oracle/snark.py: same
comm_T, same challenge, same folded instance and witness, and the folded pair satisfies the relaxed R1CS with its
commitments (the algebra nova's verifier relies on).  PARITY-UNPINNED against the reference binary (nova-snark is not
under /root/reference); chained twice so the second fold starts from a genuinely relaxed instance (u != 1, E != 0)."""
import random

import pytest

import reef_b200
import workloads as WL
from oracle import cport, snark as N
from oracle.curves import PALLAS, VESTA
from reef_b200 import snark as G

pytestmark = pytest.mark.gpu


def _gens(name, n):
    raw = WL.generators(name, n)
    return [(int.from_bytes(raw[i * 64:i * 64 + 32], "little"), int.from_bytes(raw[i * 64 + 32:i * 64 + 64], "little")) for i in range(n)]


def _random_r1cs(rnd, p, nc, nv, num_io, n_wit):
    """a satisfiable shape: row i constrains  (sum a z) * (sum b z) = z[out_i]  with out_i a witness slot filled accordingly"""
    A, B, Cm = [], [], []
    cols = list(range(n_wit)) + [nv] + [nv + 1 + k for k in range(num_io)]

    def witness(X):
        z = [0] * (2 * nv)
        z[nv] = 1
        for k, x in enumerate(X):
            z[nv + 1 + k] = x
        for j in range(n_wit):
            z[j] = rnd.randrange(p)
        return z
    rows = []
    for i in range(nc):
        ra = [(rnd.choice(cols), rnd.randrange(1, p)) for _ in range(3)]
        rb = [(rnd.choice(cols), rnd.randrange(1, p)) for _ in range(2)]
        rows.append((ra, rb, n_wit + i))
        A += [(i, c, v) for c, v in ra]
        B += [(i, c, v) for c, v in rb]
        Cm.append((i, n_wit + i, 1))
    shape = N.R1CSShape(nc, nv, num_io, A, B, Cm)

    def solve(X):
        z = witness(X)
        for ra, rb, o in rows:
            z[o] = sum(v * z[c] for c, v in ra) * sum(v * z[c] for c, v in rb) % p
        return z[:nv]
    return shape, solve


@pytest.mark.parametrize("name", ["pallas", "vesta"])
def test_nifs_fold_matches_oracle_and_stays_satisfied(ctx, name):
    curve = PALLAS if name == "pallas" else VESTA
    p = curve.order
    rnd = random.Random(5 if name == "pallas" else 6)
    nc, nv, num_io = 64, 128, 2
    shape, solve = _random_r1cs(rnd, p, nc, nv, num_io, n_wit=nv - nc)
    gens = _gens(name, nv)
    msm = lambda sc, pts: cport.msm(name, pts, sc, threads=cport.max_threads())
    bases = reef_b200.Bases(ctx, name, gens[:nc])        # commit(T) over the first num_cons generators
    pp_digest = rnd.randrange(p)

    def fresh():
        X = [rnd.randrange(p) for _ in range(num_io)]
        Wv = solve(X)
        assert shape.is_sat(Wv, [0] * nc, 1, X, p)
        return {"comm_W": msm(Wv, gens), "X": X}, {"W": Wv}

    U2, W2 = fresh()
    U1, W1 = fresh()
    U1 = dict(U1, comm_E=None, u=1)
    W1 = dict(W1, E=[0] * nc)
    for step in range(2):
        got = G.nifs_prove(ctx, name, shape, bases, pp_digest, U1, W1, U2, W2)
        exp = N.nifs_prove(curve, shape, gens, pp_digest, U1, W1, U2, W2, msm)
        assert got[0] == exp[0], "comm_T"
        assert got[1] == exp[1] and got[1] < (1 << 128), "challenge"
        assert got[2] == exp[2], "folded instance"
        assert got[3] == exp[3], "folded witness"
        U1, W1 = got[2], got[3]
        # the folded pair is a satisfying relaxed instance whose commitments open to the folded vectors
        assert shape.is_sat(W1["W"], W1["E"], U1["u"], U1["X"], p)
        assert U1["comm_W"] == msm(W1["W"], gens) and U1["comm_E"] == msm(W1["E"], gens[:nc])
        U2, W2 = fresh()
    bases.free()
