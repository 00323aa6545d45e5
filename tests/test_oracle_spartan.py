"""CPU tier: the Spartan-side oracle (oracle/spartan.py) is self-consistent -- every round
polynomial sums to the running claim, interpolation at the challenge gives the next claim, the
final claim equals comb(bound values).  (Upstream nova-snark is not vendored: there is no golden
vector to pin against; see the oracle's header.)"""
import random

import pytest

from oracle import spartan as S
from oracle.curves import PALLAS
from oracle.fields import FP, FQ


@pytest.mark.parametrize("p", [FQ, FP])
@pytest.mark.parametrize("kind", [2, 4])
def test_sumcheck_identities(p, kind):
    rnd = random.Random(kind)
    for k in (1, 2, 3, 6):
        n = 1 << k
        tabs = [[rnd.randrange(p) for _ in range(n)] for _ in range(kind)]
        ch = [rnd.randrange(p) for _ in range(k)]
        claim0, rounds, finals, last = S.prove(tabs, ch, p)
        assert last == (finals[0] * finals[1] if kind == 2 else S.comb_cubic(*finals)) % p
        # eval_1 recomputed independently on the upper half
        if kind == 2:
            e1 = sum(a * b for a, b in zip(tabs[0][n // 2:], tabs[1][n // 2:])) % p
        else:
            e1 = sum(S.comb_cubic(*x) for x in zip(*[t[n // 2:] for t in tabs])) % p
        assert rounds[0][1] == e1
        # bound tables evaluate the multilinear extension: binding all variables == MLE at the point
        t = tabs[0]
        for r in ch:
            t = S.bound_top(t, r, p)
        mle = 0
        for i, v in enumerate(tabs[0]):
            w = 1
            for j, r in enumerate(ch):
                bit = (i >> (k - 1 - j)) & 1
                w = w * (r if bit else (1 - r)) % p
            mle = (mle + v * w) % p
        assert t[0] == mle


def test_ipa_fold_is_bilinear():
    cv = PALLAS
    G = [cv.mul(k + 2, cv.gen) for k in range(8)]
    a, b = 12345, 67890
    out = S.ipa_fold_bases(cv, G, a, b)
    assert out[1] == cv.mul(a * 3 + b * 7, cv.gen)          # G[1] = 3G, G[5] = 7G
