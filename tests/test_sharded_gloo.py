"""Multi-rank sum-check protocol on CPU: (1) the sharded protocol equals the
unsharded oracle for every world size; (2) the exchange pattern runs under torch.distributed
(gloo, world_size 2) with real all-gathers between two processes."""
import os
import random
import socket

import pytest

from oracle.fields import FQ
from oracle.mle import mle_eval_fast
from oracle.nlookup import wit_nlookup_gadget
from oracle.sharded import ShardedNlookupRank, run_single_process, shard_table


def _case(ell, m, seed, small=False):
    rnd = random.Random(seed)
    n = 1 << ell
    table = [rnd.randrange(131 if small else FQ) for _ in range(n)]
    q = [rnd.randrange(n) for _ in range(m)]
    v = [table[i] for i in q]
    prev_q = [rnd.randrange(FQ) for _ in range(ell)]
    prev_v = mle_eval_fast(table, prev_q)
    return table, q, v, prev_q, prev_v


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("ell,m,tag", [(3, 2, "nl"), (6, 5, "nldoc"), (9, 3, "nlhybrid")])
def test_sharded_protocol_equals_unsharded_oracle(world, ell, m, tag):
    table, q, v, prev_q, prev_v = _case(ell, m, ell * 10 + world)
    exp = wit_nlookup_gadget(table, q, v, prev_q, prev_v, tag, 4242, fast=True)
    outs = run_single_process(table, world, q, v, prev_q, prev_v, tag, 4242)
    for o in outs:
        assert o["claim_r"] == exp["claim_r"]
        assert o["rounds"] == exp["rounds"]
        assert o["sc_last_claim"] == exp["sc_last_claim"]
        assert o["next_running_claim"] == exp["next_running_claim"]


def _worker(rank, world, port, ell, m, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table, q, v, prev_q, prev_v = _case(ell, m, 77, small=True)
    node = ShardedNlookupRank(shard_table(table, rank, world), rank, world, q, v, prev_q, prev_v, "nldoc", 99)

    def exchange(vals):
        # field elements travel as 32-byte little-endian rows of a uint8 tensor (what NCCL moves too)
        mine = torch.tensor(list(b"".join(int(x).to_bytes(32, "little") for x in vals)), dtype=torch.uint8)
        outs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(outs, mine)
        return [[int.from_bytes(bytes(o[i * 32:(i + 1) * 32].tolist()), "little") for i in range(len(vals))] for o in outs]

    out = node.run(exchange)
    exp = wit_nlookup_gadget(table, q, v, prev_q, prev_v, "nldoc", 99, fast=True)
    ok = (out["rounds"] == exp["rounds"] and out["claim_r"] == exp["claim_r"]
          and out["next_running_claim"] == exp["next_running_claim"] and out["sc_last_claim"] == exp["sc_last_claim"])
    ret[rank] = ok
    dist.destroy_process_group()


def test_sharded_protocol_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, 4, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert ret.get(0) is True and ret.get(1) is True


def test_msm_window_partition_covers_all_windows():
    """bench.py shards an MSM by Pippenger windows: [W*g/G, W*(g+1)/G) must tile [0, W)."""
    for W in (16, 20, 24, 32, 52):
        for G in (1, 2, 4, 8):
            cuts = [(W * g // G, W * (g + 1) // G) for g in range(G)]
            assert cuts[0][0] == 0 and cuts[-1][1] == W
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(G - 1))
