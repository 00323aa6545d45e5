"""GPU parity (through the C ABI): Poseidon hashes, sponge sessions, Merkle commitment."""
import random

import pytest

import reef_b200
from oracle import poseidon as P
from oracle.fields import FQ
from oracle.merkle import MerkleCommitment as OracleMerkle

pytestmark = pytest.mark.gpu


def test_hash_batch_arity_2_and_4(ctx):
    rnd = random.Random(1)
    for arity in (2, 4):
        n = 300
        rows = [rnd.randrange(FQ) for _ in range(n * arity)]
        rows[:arity] = [0] * arity
        rows[arity:2 * arity] = [FQ - 1] * arity
        got = ctx.poseidon_hash(rows, arity)
        for i in range(n):
            assert got[i] == P.hash_once(rows[i * arity:(i + 1) * arity]), (arity, i)
    assert ctx.poseidon_hash([], 2) == []


def test_hash_batch_thread_per_hash_path(ctx):
    """>= 148*256 hashes take the thread-per-hash kernel; spot-check against the oracle."""
    import numpy as np
    n, arity = 40000, 4
    raw = np.random.default_rng(5).integers(0, 1 << 62, size=(n * arity, 4), dtype=np.uint64)
    raw[:, 3] &= (1 << 61) - 1
    rows = [int.from_bytes(r.tobytes(), "little") for r in raw]
    got = ctx.poseidon_hash(rows, arity)
    for i in list(range(0, n, 997)) + [n - 1]:
        assert got[i] == P.hash_once(rows[i * arity:(i + 1) * arity]), i


def test_calc_d(ctx):
    # commitment.rs:495-510
    rnd = random.Random(2)
    for _ in range(5):
        v, s = rnd.randrange(FQ), rnd.randrange(FQ)
        assert ctx.calc_d(v, s) == P.calc_d(v, s)


def test_non_canonical_input_rejected(ctx):
    with pytest.raises(reef_b200.ReefError) as e:
        ctx.poseidon_hash([FQ, 0], 2)
    assert e.value.code == 1


@pytest.mark.parametrize("pattern", [
    [("A", 2), ("S", 1)],
    [("A", 4), ("S", 1)],
    [("A", 9), ("S", 1)],
    [("A", 24), ("S", 1)] + [("A", 3), ("S", 1)] * 4,
    [("A", 1), ("S", 6), ("A", 5), ("S", 2)],
])
def test_sponge_one_shot_and_incremental(ctx, pattern):
    rnd = random.Random(3)
    n_in = sum(n for k, n in pattern if k == "A")
    elems = [rnd.randrange(FQ) for _ in range(n_in)]
    sp = P.Sponge()
    sp.start(pattern)
    exp, pos = [], 0
    for k, n in pattern:
        if k == "A":
            sp.absorb(elems[pos:pos + n])
            pos += n
        else:
            exp += sp.squeeze(n)
    sp.finish()
    assert ctx.poseidon_sponge(pattern, elems) == exp
    s = reef_b200.Sponge(ctx, pattern)
    got, pos = [], 0
    for k, n in pattern:
        if k == "A":
            s.absorb(elems[pos:pos + n])
            pos += n
        else:
            got += s.squeeze(n)
    s.finish()
    assert got == exp


def test_sponge_pattern_mismatch_is_an_assert(ctx):
    s = reef_b200.Sponge(ctx, [("A", 2), ("S", 1)])
    with pytest.raises(reef_b200.ReefError) as e:
        s.absorb([1, 2, 3])
    assert e.value.code == 3
    s.absorb([1, 2])
    with pytest.raises(reef_b200.ReefError) as e:
        s.finish()                      # squeeze never happened: ParameterUsageMismatch
    assert e.value.code == 3


@pytest.mark.parametrize("doc", [
    [2, 3, 4, 5, 6, 7, 8],              # the reference's make_mt document (merkle_tree.rs:213)
    [7], [1, 2], [9, 8, 7], list(range(100, 133)), [3] * 64,
])
def test_merkle_small_matches_oracle_and_paths_recompute_root(ctx, doc):
    mc = ctx.merkle(doc)
    om = OracleMerkle(doc)
    assert mc.commitment == om.commitment
    assert mc.tree == om.tree
    for q in range(len(doc)):
        w = mc.path_wits(q)
        assert w == om.path_wits(q)
        l, oi, o = w[0]
        h = P.hash_once([q, doc[q], oi, o] if l else [oi, o, q, doc[q]])
        for (l, _, o) in w[1:]:
            h = P.hash_once([h, o] if l else [o, h])
        assert h == mc.commitment       # merkle_tree.rs:255
    with pytest.raises(reef_b200.ReefError) as e:
        mc.path_wits(len(doc))
    assert e.value.code == 3


def test_merkle_large_spot_checked(ctx):
    """2^16-leaf tree (thread-per-hash levels AND warp-per-hash levels): every level is checked
    by re-hashing random parent nodes with the oracle, and random paths recompute the root."""
    rnd = random.Random(4)
    n = (1 << 16) + 3                   # odd sizes at several levels
    doc = [rnd.randrange(6) for _ in range(n)]
    mc = ctx.merkle(doc)
    assert [len(l) for l in mc.tree][:3] == [(n + 1) // 2, ((n + 1) // 2 + 1) // 2, (((n + 1) // 2 + 1) // 2 + 1) // 2]
    for _ in range(40):
        k = rnd.randrange(len(mc.tree[0]))
        r = [2 * k + 1, doc[2 * k + 1]] if 2 * k + 1 < n else [0, 0]
        assert mc.tree[0][k] == P.hash_once([2 * k, doc[2 * k]] + r)
    for lvl in range(1, len(mc.tree)):
        prev = mc.tree[lvl - 1]
        for _ in range(6):
            k = rnd.randrange(len(mc.tree[lvl]))
            right = prev[2 * k + 1] if 2 * k + 1 < len(prev) else 0
            assert mc.tree[lvl][k] == P.hash_once([prev[2 * k], right])
    assert len(mc.tree[-1]) == 1 and mc.tree[-1][0] == mc.commitment
    for q in (0, 1, n - 1, n - 2, rnd.randrange(n)):
        w = mc.path_wits(q)
        l, oi, o = w[0]
        h = P.hash_once([q, doc[q], oi, o] if l else [oi, o, q, doc[q]])
        for (l, _, o) in w[1:]:
            h = P.hash_once([h, o] if l else [o, h])
        assert h == mc.commitment
