"""GPU parity (through the C ABI): MLE sweeps and the nlookup sum-check, bit-exact against the
oracle; the first tests replay the reference's own unit tests (r1cs.rs:2411-2578)."""
import random

import pytest

import reef_b200
from oracle import poseidon as P
from oracle.fields import FQ
from oracle.mle import gen_eq_table, mle_eval_fast, prover_mle_partial_eval
from oracle.nlookup import wit_nlookup_gadget

pytestmark = pytest.mark.gpu


def test_mle_partial(ctx):
    """r1cs.rs:2517-2578, same table, same assertions, evaluated on the GPU."""
    table = [1, 3, 8, 2, 9, 5, 13, 4]
    t = ctx.table(table)
    for x1 in (0, 1, -1):
        for x2 in (0, 1, -1):
            for x3 in (0, 1, -1):
                holes = (x1 == -1) + (x2 == -1) + (x3 == -1)
                if holes > 1:
                    continue
                coeff, con = ctx.prover_mle_partial_eval(t, [x1, x2, x3])
                assert (coeff, con) == prover_mle_partial_eval(table, [x1, x2, x3], list(range(8)), True, None)
                if holes == 1:
                    if x1 == -1:
                        assert (coeff + con) % FQ == table[4 + x2 * 2 + x3] and con == table[x2 * 2 + x3]
                    elif x2 == -1:
                        assert (coeff + con) % FQ == table[x1 * 4 + 2 + x3] and con == table[x1 * 4 + x3]
                    else:
                        assert (coeff + con) % FQ == table[x1 * 4 + x2 * 2 + 1] and con == table[x1 * 4 + x2 * 2]
                else:
                    assert con == table[x1 * 4 + x2 * 2 + x3]


def test_mle_linear_basic(ctx):
    """r1cs.rs:2411-2515: gen_eq_table + 3 x linear_mle_product with a real sponge session."""
    evals = [2, 3, 5, 7, 9, 13, 17, 19]
    qs, last_q, claims = [2, 1, 7], [2, 3, 5], [3, 9, 27, 81]
    eq_a = ctx.gen_eq_table(claims, qs, list(reversed(last_q)))
    assert eq_a == gen_eq_table(claims, qs, list(reversed(last_q)))
    t_tab, e_tab = ctx.table(evals), ctx.table(eq_a)
    running_v = ctx.verifier_mle_eval(t_tab, last_q)
    term = (sum(evals[qs[i]] * claims[i] for i in range(3)) + running_v * claims[3]) % FQ
    claim = sum(t * e for t, e in zip(evals, eq_a)) % FQ
    assert term == claim
    pattern = [("A", 3), ("S", 1)] * 3
    sponge = reef_b200.Sponge(ctx, pattern)
    osp = P.Sponge()
    osp.start(pattern)
    sc_rs = []
    for i in range(1, 4):
        r_i, xsq, x, con = ctx.linear_mle_product(t_tab, e_tab, 3, i, sponge)
        assert claim == (2 * con + x + xsq) % FQ
        osp.absorb([con, x, xsq])
        assert r_i == osp.squeeze(1)[0]
        claim = (xsq * r_i * r_i + x * r_i + con) % FQ
        sc_rs.append(r_i)
    sponge.finish()
    fresh = ctx.table(evals)
    next_running_v = ctx.verifier_mle_eval(fresh, sc_rs)
    _, eq_term = prover_mle_partial_eval(claims, sc_rs, qs, False, last_q)
    assert claim == eq_term * next_running_v % FQ
    # fully folded tables (SURVEY A.6 ii)
    assert t_tab.download(1)[0] == next_running_v
    assert claim == next_running_v * e_tab.download(1)[0] % FQ


@pytest.mark.parametrize("ell,m", [(1, 1), (3, 0), (5, 4), (9, 3), (12, 7)])
def test_gen_eq_table_and_eval(ctx, ell, m):
    rnd = random.Random(ell * 31 + m)
    n = 1 << ell
    rs = [rnd.randrange(FQ) for _ in range(m + 1)]
    qs = [rnd.randrange(n) for _ in range(m)]
    if m >= 2:
        qs[1] = qs[0]                   # duplicate lookups accumulate
    lq = [rnd.randrange(FQ) for _ in range(ell)]
    from oracle.mle import gen_eq_table_fast
    assert ctx.gen_eq_table(rs, qs, lq) == gen_eq_table_fast(rs, qs, lq)
    tab = [rnd.randrange(FQ) for _ in range(n)]
    x = [rnd.randrange(FQ) for _ in range(ell)]
    assert ctx.verifier_mle_eval(ctx.table(tab), x) == mle_eval_fast(tab, x)


def _check_nlookup(ctx, table, q, tag, u32, prev=True, seed=0, doc_hash=0xABCDEF, n_steps=2):
    rnd = random.Random(seed)
    ell = reef_b200.logmn(len(table))
    v = [table[i] for i in q]
    t = ctx.table_u32(table) if u32 else ctx.table(table)
    rq = [rnd.randrange(FQ) for _ in range(ell)] if prev else None
    rv = mle_eval_fast(table, rq) if prev else None
    for step in range(n_steps):                       # chained steps: running claim feeds the next
        got = ctx.wit_nlookup_gadget(t, q, v, rq, rv, tag, doc_hash if tag != "nl" else None)
        exp = wit_nlookup_gadget(table, q, v, rq, rv, tag, doc_hash, fast=True)
        assert got.claim_r == exp["claim_r"], "claim_r"
        assert got.combined_q == exp["combined_q"]
        assert got.prev_running_claim == exp["prev_running_claim"]
        for i, (g, e) in enumerate(zip(got.rounds, exp["rounds"])):
            assert g == e, f"round {i + 1} of {ell}: (sc_r, xsq, x, const)"
        assert len(got.rounds) == ell
        assert got.sc_last_claim == exp["sc_last_claim"]
        assert got.next_running_claim == exp["next_running_claim"]
        rq, rv = got.next_running_q, got.next_running_claim
    t.free()


@pytest.mark.parametrize("ell", [1, 2, 4, 7, 10, 11, 12, 13, 15])
@pytest.mark.parametrize("tag", ["nl", "nldoc"])
def test_nlookup_field_tables(ctx, ell, tag):
    rnd = random.Random(100 + ell)
    n = 1 << ell
    table = [rnd.randrange(FQ) for _ in range(n)]
    m = [1, 3, 5][ell % 3]
    q = [rnd.randrange(n) for _ in range(m)]
    if m >= 3:
        q[2] = q[0]                                   # repeated lookup
        q[1] = n - 1
    _check_nlookup(ctx, table, q, tag, u32=False, prev=(ell % 2 == 0), seed=ell)


@pytest.mark.parametrize("ell", [4, 10, 11, 12, 14, 17])
def test_nlookup_document_codes_u32(ctx, ell):
    """nldoc over a u32 document table (cfg-2 shape at ell = 17)."""
    rnd = random.Random(200 + ell)
    n = 1 << ell
    table = [rnd.randrange(131) for _ in range(n - 2)] + [130, 129]
    q = [rnd.randrange(n) for _ in range(4)] + [0, n - 1]
    _check_nlookup(ctx, table, q, "nldoc", u32=True, prev=True, seed=ell)
    _check_nlookup(ctx, table, q, "nlhybrid", u32=True, prev=False, seed=ell, n_steps=1)


def test_nlookup_first_step_defaults_and_no_lookups(ctx):
    rnd = random.Random(7)
    table = [rnd.randrange(FQ) for _ in range(64)]
    _check_nlookup(ctx, table, [], "nl", u32=False, prev=False)       # m = 0
    _check_nlookup(ctx, table, [5], "nl", u32=False, prev=False)      # prev_v = table[0]


def test_nlookup_many_lookups_two_combined_q_limbs(ctx):
    rnd = random.Random(8)
    n = 1 << 11
    table = [rnd.randrange(FQ) for _ in range(n)]
    q = [rnd.randrange(n) for _ in range(40)]                         # 440 bits -> 2 limbs
    _check_nlookup(ctx, table, q, "nl", u32=False, prev=True, n_steps=1)


def test_nlookup_rejects_what_the_reference_asserts(ctx):
    t = ctx.table([1, 2, 3])                                          # not a power of two
    with pytest.raises(reef_b200.ReefError) as e:
        ctx.wit_nlookup_gadget(t, [0], [1], tag="nl")
    assert e.value.code == 3
    r = ctx.wit_nlookup_gadget(t, [2], [3], tag="nldoc", doc_hash=1)  # nldoc pads with zeros
    exp = wit_nlookup_gadget([1, 2, 3], [2], [3], None, None, "nldoc", 1)
    assert r.rounds == exp["rounds"] and r.next_running_claim == exp["next_running_claim"]
    t4 = ctx.table([1, 2, 3, 4])
    with pytest.raises(reef_b200.ReefError) as e:
        ctx.wit_nlookup_gadget(t4, [4], [1], tag="nl")                # index out of bounds
    assert e.value.code == 3


@pytest.mark.parametrize("ell", [21, 23])
def test_nlookup_full_size_properties(ctx, ell):
    """cfg-3/4 size (N = 2^21) and cfg-5 size (N = 2^23, the deepest sweep shape: 32 pairs per lane),
    u32 document: too big for the Python oracle, so check the
    size-independent properties: claim identity, per-round g(0)+g(1), last claim, and that the
    next running claim is the MLE of the table at the challenges (recomputed independently by
    the GPU fold path AND, for the claim identity, from plain lookups)."""
    rnd = random.Random(9)
    n = 1 << ell
    import numpy as np
    codes = np.random.default_rng(21).integers(0, 131, size=n, dtype=np.uint32)
    t = ctx.table_u32(codes)
    q = [rnd.randrange(n) for _ in range(6)]
    v = [int(codes[i]) for i in q]
    got = ctx.wit_nlookup_gadget(t, q, v, None, None, "nldoc", 77)
    m = len(q)
    rs = [pow(got.claim_r, k + 1, FQ) for k in range(m + 1)]
    claim = (sum(rs[k] * v[k] for k in range(m)) + rs[m] * int(codes[0])) % FQ   # prev_q = 0 => T~(0) = T[0]
    for (sc_r, xsq, x, con) in got.rounds:
        assert claim == (2 * con + x + xsq) % FQ
        claim = (xsq * sc_r * sc_r + x * sc_r + con) % FQ
    assert claim == got.sc_last_claim
    assert got.next_running_claim == ctx.verifier_mle_eval(t, got.next_running_q)
    # eq(next_q) side of the last claim: last_claim = T~(r) * EQ~(r)
    eq_r = 0
    for k in range(m):
        term = rs[k]
        for j, rj in enumerate(got.next_running_q):
            bit = (q[k] >> (ell - 1 - j)) & 1
            term = term * (rj if bit else (1 - rj)) % FQ
        eq_r = (eq_r + term) % FQ
    term = rs[m]
    for rj in got.next_running_q:
        term = term * (1 - rj) % FQ                                   # eq(0, r)
    eq_r = (eq_r + term) % FQ
    assert got.sc_last_claim == got.next_running_claim * eq_r % FQ
    t.free()


@pytest.mark.parametrize("ell,u32", [(6, True), (9, False), (13, True), (17, True)])
def test_hyrax_lz_matvec(ctx, ell, u32):
    """LZ = L^T M with L = eq(q_left) (Hyrax prove_eval); also  <LZ, eq(q_right)> = doc~(q)."""
    rnd = random.Random(ell)
    left = ell // 2
    rows, cols = 1 << left, 1 << (ell - left)
    n = rows * cols
    tab = [rnd.randrange(131 if u32 else FQ) for _ in range(n)]
    t = ctx.table_u32(tab) if u32 else ctx.table(tab)
    q = [rnd.randrange(FQ) for _ in range(ell)]
    L = ctx.gen_eq_table([1], [], list(reversed(q[:left])))          # index bit (left-1-k) <-> q[k]
    R = ctx.gen_eq_table([1], [], list(reversed(q[left:])))
    lz = ctx.hyrax_lz(t, rows, cols, L)
    if ell <= 13:
        for j in [0, 1, cols // 2, cols - 1]:
            assert lz[j] == sum(L[i] * tab[i * cols + j] for i in range(rows)) % FQ
    assert sum(a * b for a, b in zip(lz, R)) % FQ == mle_eval_fast(tab, q)
    assert ctx.verifier_mle_eval(t, q) == mle_eval_fast(tab, q)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("ell,u32,tag", [(5, False, "nl"), (11, True, "nldoc"), (13, False, "nlhybrid"), (15, True, "nldoc")])
def test_sharded_sumcheck_all_ranks_on_one_gpu(ctx, world, ell, u32, tag):
    """the table is sharded by low index bits over `world` ranks; here all ranks run
    lock-step on one GPU and the all-gather is a device buffer every rank writes its slot of.
    Every rank must reproduce the unsharded oracle bit for bit."""
    import torch
    rnd = random.Random(ell * 100 + world)
    n = 1 << ell
    table = [rnd.randrange(131 if u32 else FQ) for _ in range(n)]
    q = [rnd.randrange(n) for _ in range(5)] + [0, n - 1]
    v = [table[i] for i in q]
    prev_q = [rnd.randrange(FQ) for _ in range(ell)]
    prev_v = mle_eval_fast(table, prev_q)
    exp = wit_nlookup_gadget(table, q, v, prev_q, prev_v, tag, 31337, fast=True)
    tabs = []
    for g in range(world):
        shard = table[g::world]
        tabs.append(ctx.table_u32(shard) if u32 else ctx.table(shard))
    ranks = [reef_b200.ShardedNlookup(ctx, tabs[g], g, world, q, v, prev_q, prev_v, tag, 31337 if tag != "nl" else None)
             for g in range(world)]
    buf = torch.zeros(world * 96, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for _ in range(ranks[0].ell_local):
        for g, r in enumerate(ranks):
            r.round_local(buf.data_ptr() + g * 96)
        for r in ranks:
            r.round_finish(buf.data_ptr())
    pairs = torch.zeros(world * 64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for g, r in enumerate(ranks):
        r.export(pairs.data_ptr() + g * 64)
    for r in ranks:
        got = r.finish(pairs.data_ptr())
        assert got.claim_r == exp["claim_r"]
        assert got.rounds == exp["rounds"]
        assert got.sc_last_claim == exp["sc_last_claim"]
        assert got.next_running_claim == exp["next_running_claim"]
    for r in ranks:
        r.free()


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("ell,u32,tag", [(11, True, "nldoc"), (13, False, "nlhybrid")])
def test_sharded_sumcheck_p2p_mailbox_exchange(world, ell, u32, tag, fused):
    """The per-round exchange done by the library's own P2P mailbox kernel (p2p.cu) instead of a
    collective: one context (= one stream) per rank on this GPU, mailboxes connected by device
    pointer; all ranks are enqueued without any host wait and must reproduce the oracle."""
    import torch
    rnd = random.Random(ell * 10 + world)
    n = 1 << ell
    table = [rnd.randrange(131 if u32 else FQ) for _ in range(n)]
    q = [rnd.randrange(n) for _ in range(4)] + [n - 1]
    v = [table[i] for i in q]
    prev_q = [rnd.randrange(FQ) for _ in range(ell)]
    prev_v = mle_eval_fast(table, prev_q)
    exp = wit_nlookup_gadget(table, q, v, prev_q, prev_v, tag, 4242, fast=True)
    ctxs = [reef_b200.Context(0) for _ in range(world)]
    tabs, ranks = [], []
    try:
        for c in ctxs:
            c.mailbox_create(world)
        ptrs = [c.mailbox_ptr() for c in ctxs]
        for g, c in enumerate(ctxs):
            c.mailbox_connect_local(g, world, ptrs)
        tabs = [ctxs[g].table_u32(table[g::world]) if u32 else ctxs[g].table(table[g::world]) for g in range(world)]
        ranks = [reef_b200.ShardedNlookup(ctxs[g], tabs[g], g, world, q, v, prev_q, prev_v, tag, 4242 if tag != "nl" else None)
                 for g in range(world)]
        bufs = [torch.zeros((world + 1) * 96, dtype=torch.uint8, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        got = [None] * world
        if fused:                                              # exchange inside the round kernels themselves
            for _ in range(ranks[0].ell_local):
                for r in ranks:
                    check_rc = reef_b200.lib.reef_nl_shard_round_p2p(r._h)
                    assert check_rc == 0, reef_b200.lib.reef_last_error()
            # export + final rounds: rank g's finish waits for every peer's export, so all exports are
            # enqueued by worker threads before any result is read back
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=world) as ex:
                got = list(ex.map(lambda r: r.finish_p2p(), ranks))
            ranks_done = True
        else:
            ranks_done = False
        for _ in range(0 if ranks_done else ranks[0].ell_local):
            for g, r in enumerate(ranks):                      # nothing below waits on the host
                mine, allp = bufs[g].data_ptr(), bufs[g].data_ptr() + 96
                r.round_local(mine)
                ctxs[g].p2p_allgather(mine, 96, allp)
                r.round_finish(allp)
        for g, r in enumerate(ranks):
            if ranks_done:
                break
            mine, allp = bufs[g].data_ptr(), bufs[g].data_ptr() + 96
            r.export(mine)
            ctxs[g].p2p_allgather(mine, 64, allp)
        for g, r in enumerate(ranks):
            if not ranks_done:
                got[g] = r.finish(bufs[g].data_ptr() + 96)
            ctxs[g].p2p_status()
        for g in range(world):
            assert got[g].claim_r == exp["claim_r"]
            assert got[g].rounds == exp["rounds"]
            assert got[g].sc_last_claim == exp["sc_last_claim"]
            assert got[g].next_running_claim == exp["next_running_claim"]
    finally:
        for r in ranks:
            r.free()
        for t in tabs:
            t.free()
        for c in ctxs:
            c.close()
