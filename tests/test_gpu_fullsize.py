"""GPU parity AT BASELINE.json's FULL SIZES, bit for bit against the C restatement of the reference's
algorithm (oracle/c, reference-shaped O(N*ell) schedule): the nldoc sum-check of the 2^20-char
documents (ell = 21, u32 codes), the `nlhybrid` sum-check over the 2^22-entry merged table of
configs[3] (r1cs.rs:2101-2136), the Merkle commitment of configs[2] (2^21 leaves,
merkle_tree.rs:25-114) and the Hyrax document commitment at 1024 x 2048 (commitment.rs:187)."""
import random

import numpy as np
import pytest

import reef_b200
import workloads as W
from oracle import cport
from oracle.fields import FQ

pytestmark = pytest.mark.gpu


def _same(got, exp):
    assert got.claim_r == exp["claim_r"]
    assert got.rounds == exp["rounds"]
    assert got.sc_last_claim == exp["sc_last_claim"]
    assert got.next_running_claim == exp["next_running_claim"]
    assert got.combined_q == exp["combined_q"]


@pytest.mark.parametrize("cfg", ["target", "cfg4"])
def test_nldoc_ell21_two_chained_folds(ctx, cfg):
    ab, cps = W.document(cfg)
    udoc = W.encode(ab, cps)
    assert len(udoc) == 1 << 21
    rnd = random.Random(5)
    t = ctx.table_u32(udoc)
    doc_hash = rnd.randrange(FQ)
    rq = rv = None
    for fold in range(2):
        q = [rnd.randrange(len(cps) + 2) for _ in range(3)]
        v = [int(udoc[i]) for i in q]
        got = ctx.wit_nlookup_gadget(t, q, v, rq, rv, "nldoc", doc_hash)
        exp = cport.wit_nlookup_gadget(udoc, q, v, rq, rv, "nldoc", doc_hash, u32=True)
        _same(got, exp)
        rq, rv = got.next_running_q, got.next_running_claim
    t.free()


def test_nlhybrid_ell22_merged_table(ctx):
    """configs[3]: one table = T (padded with its fill value to 2^21) ++ document (2^21), tag nlhybrid."""
    ab, cps = W.document("cfg4")
    udoc = W.encode(ab, cps)
    half = len(udoc)
    rnd = random.Random(6)
    n_T = 1 << 9
    T = sorted(rnd.randrange(1 << 100) for _ in range(n_T))
    fill = rnd.randrange(1 << 100)
    tab = np.zeros((2 * half, 4), dtype=np.uint64)
    for i, x in enumerate(T):
        tab[i] = [(x >> (64 * k)) & (2 ** 64 - 1) for k in range(4)]
    tab[n_T:half] = [(fill >> (64 * k)) & (2 ** 64 - 1) for k in range(4)]
    tab[half:, 0] = udoc
    t = ctx.table_hybrid(T, fill, half, udoc)          # expanded on the device (reef_table_hybrid_u32)
    doc_hash = rnd.randrange(FQ)
    raw = np.frombuffer(tab.tobytes(), dtype=np.uint8)
    rq = rv = None
    for fold in range(2):
        q = [rnd.randrange(n_T) for _ in range(2)] + [half + rnd.randrange(len(cps) + 2) for _ in range(2)]
        v = [int.from_bytes(tab[i].tobytes(), "little") for i in q]
        got = ctx.wit_nlookup_gadget(t, q, v, rq, rv, "nlhybrid", doc_hash)
        exp = _cport_nlookup_raw(raw, 2 * half, q, v, rq, rv, "nlhybrid", doc_hash, first=int.from_bytes(tab[0].tobytes(), "little"))
        _same(got, exp)
        rq, rv = got.next_running_q, got.next_running_claim
    t.free()


def _cport_nlookup_raw(raw_u8, n, q, v, running_q, running_v, tag, doc_hash, first):
    """cport.wit_nlookup_gadget on a table that is already 32-byte LE rows (no Python big-int list)."""
    from oracle.nlookup import combined_qs, logmn, nlookup_pattern
    ell = logmn(n)
    prev_q = list(running_q) if running_q is not None else [0] * ell
    prev_v = running_v if running_v is not None else first
    cqs = combined_qs(list(q), ell)
    pattern = nlookup_pattern(tag, len(v), ell, len(cqs))
    query = ([] if tag == "nl" else [doc_hash]) + cqs + [int(x) for x in v] + prev_q + [prev_v]
    pk = lambda xs: b"".join(int(x).to_bytes(32, "little") for x in xs)
    claim, rounds, last, nxt = cport.nlookup_raw(raw_u8, 0, n, q, pk(query), len(query), cport.ops_words(pattern), pk(prev_q), ell)
    r = [int.from_bytes(rounds[i:i + 32], "little") for i in range(0, len(rounds), 32)]
    rounds = [tuple(r[4 * i:4 * i + 4]) for i in range(ell)]
    return {"claim_r": int.from_bytes(claim, "little"), "combined_q": cqs, "rounds": rounds,
            "sc_last_claim": int.from_bytes(last, "little"), "next_running_claim": int.from_bytes(nxt, "little")}


def test_merkle_2p21_leaves(ctx):
    """configs[2]: the whole tree (2 097 151 permutations), every level compared."""
    ab, cps = W.document("cfg3")
    udoc = W.encode(ab, cps)
    assert len(udoc) == 1 << 21
    mc = ctx.merkle(udoc)
    cport.lib().oracle_set_fast_poseidon(1)     # same digests (tests/test_oracle_c.py), half the multiplications
    try:
        exp = cport.merkle(udoc, threads=cport.max_threads())
    finally:
        cport.lib().oracle_set_fast_poseidon(0)
    assert len(mc.tree) == len(exp) == 21
    for lvl in range(21):
        assert mc.tree[lvl] == exp[lvl], f"level {lvl}"
    assert mc.commitment == exp[-1][0]
    # path witnesses recompute the root (make_mt, merkle_tree.rs:209-257) at a few positions
    for idx in (0, 4101, len(cps), len(cps) + 1, (1 << 21) - 1):
        w = mc.path_wits(idx)
        leaf_l, leaf_r = (idx, int(udoc[idx])), (w[0][1], w[0][2])
        if not w[0][0]:
            leaf_l, leaf_r = leaf_r, leaf_l
        node = cport.poseidon_hash([leaf_l[0], leaf_l[1], leaf_r[0], leaf_r[1]], 4)[0]
        for lr, _, opp in w[1:]:
            node = cport.poseidon_hash([node, opp] if lr else [opp, node], 2)[0]
        assert node == mc.commitment


@pytest.mark.parametrize("shape", [(1024, 2048, "cfg4", 8), (256, 512, "cfg2", 8)])
def test_hyrax_commit_full_shape(ctx, shape):
    rows, cols, cfg, bits = shape
    ab, cps = W.document(cfg)
    udoc = W.encode(ab, cps)
    assert (rows, cols) == W.hyrax_dims(W.logmn(len(udoc)))
    gens = W.generators("pallas", cols + 1)
    b = ctx.bases("pallas", gens, 255)
    rnd = random.Random(7)
    blinds = [rnd.randrange(FQ) for _ in range(rows)]
    got = b.msm_rows(udoc.reshape(rows, cols), rows, cols, entry_bits=bits, blinds=blinds)
    for r in sorted({0, 1, rows // 2, rows - 1} | {rnd.randrange(rows) for _ in range(6)}):
        sc = [int(x) for x in udoc[r * cols:(r + 1) * cols]] + [blinds[r]]
        assert got[r] == cport.msm("pallas", gens, sc, threads=cport.max_threads()), f"row {r}"
    b.free()


def test_hybrid_table_layout_and_document_repeat_quirk(ctx):
    """r1cs.rs:2105-2112: T, its fill value up to half_len, then the zero-padded document REPEATED until the
    table is 2 * half_len long (happens when the padded document is shorter than the public half)."""
    rnd = random.Random(8)
    T = [rnd.randrange(FQ) for _ in range(5)]
    fill = rnd.randrange(FQ)
    doc = [rnd.randrange(131) for _ in range(5)]            # pads to 8, half_len 16: two copies
    t = ctx.table_hybrid(T, fill, 16, doc)
    exp = T + [fill] * 11 + (doc + [0] * 3) * 2
    assert t.download(32) == exp
    t.free()
