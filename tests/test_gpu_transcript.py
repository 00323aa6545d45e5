"""The lane-parallel transcript permutation (reef_b200/csrc/poseidon_lp.cuh) alone: one 256-thread
CTA must compute exactly neptune's Poseidon permutation (oracle/poseidon.py), also when it is chained."""
import ctypes as C
import random

import pytest

from oracle import poseidon as P
from oracle.fields import FQ
from reef_b200._lib import check, lib

pytestmark = pytest.mark.gpu


def _lp(ctx, state, n):
    inp = b"".join(int(x).to_bytes(32, "little") for x in state)
    out = C.create_string_buffer(160)
    cyc = (C.c_uint64 * 9)()
    check(lib.reef_gputest_poseidon_permute_lp(ctx._h, inp, n, out, cyc))
    return [int.from_bytes(out.raw[i * 32:(i + 1) * 32], "little") for i in range(5)], list(cyc)


def test_single_permutation_matches_the_oracle(ctx):
    rnd = random.Random(11)
    for st in ([0, 0, 0, 0, 0], [1, 2, 3, 4, 5], [FQ - 1] * 5, [rnd.randrange(FQ) for _ in range(5)], [rnd.randrange(FQ) for _ in range(5)]):
        got, _ = _lp(ctx, st, 1)
        assert got == P.permute(list(st))


def test_chained_permutations_and_latency(ctx):
    rnd = random.Random(12)
    st = [rnd.randrange(FQ) for _ in range(5)]
    exp = list(st)
    for _ in range(7):
        exp = P.permute(exp)
    got, cyc = _lp(ctx, st, 7)
    assert got == exp
    print(f"lane-parallel permutation: {cyc[0]} SM cycles per permutation; last one: first full rounds {cyc[1]}, "
          f"partial rounds {cyc[2]}, end {cyc[3]}, last full rounds {cyc[4]}; chain phases (only with -DREEF_LP_TIMING): {cyc[5:9]}")
    assert 0 < cyc[0] < 200000
