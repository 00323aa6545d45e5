"""CPU tier: the partition / reassembly logic of the multi-GPU Hyrax and Merkle splits
(reef_b200/sharding.py) with oracle arithmetic as the compute step -- single process for
every world size, and with real all-gathers between two gloo processes."""
import os
import random
import socket

import pytest

from oracle.curves import PALLAS
from oracle.merkle import MerkleCommitment, new_parent
from reef_b200.sharding import hyrax_commit_sharded, merkle_sharded, row_range

le = lambda x: int(x).to_bytes(32, "little")


def _oracle_subtree(doc):
    def build(a, b):
        level = [new_parent((i, doc[i]), (i + 1, doc[i + 1])) for i in range(a, b, 2)]
        levels = [[le(x) for x in level]]
        while len(level) > 1:
            level = [new_parent((None, level[i]), (None, level[i + 1])) for i in range(0, len(level), 2)]
            levels.append([le(x) for x in level])
        return levels, levels[-1][0]
    return build


def _oracle_top(roots):
    level = [int.from_bytes(r, "little") for r in roots]
    levels = []
    while len(level) > 1:
        level = [new_parent((None, level[i]), (None, level[i + 1])) for i in range(0, len(level), 2)]
        levels.append([le(x) for x in level])
    return levels, levels[-1][0]


def _points_bytes(P):
    return bytes(64) if P is None else le(P[0]) + le(P[1])


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_merkle_split_equals_the_whole_tree(world):
    rnd = random.Random(world)
    doc = [rnd.randrange(7) for _ in range(64)]
    exp = MerkleCommitment(doc)
    # run the ranks one after the other with a recorded all-gather
    posts = {}
    outs = []
    for phase in (0, 1):                 # phase 0 records every rank's posts, phase 1 replays them as the gather
        outs = []
        for g in range(world):
            k = [0]

            def gather(b, g=g, k=k):
                posts.setdefault(k[0], {})[g] = b
                res = [posts[k[0]].get(r, b) for r in range(world)]
                k[0] += 1
                return res
            outs.append(merkle_sharded(_oracle_subtree(doc), _oracle_top, len(doc), g, world, gather, True))
    for root, tree in outs:
        assert int.from_bytes(root, "little") == exp.commitment
        assert [[int.from_bytes(x, "little") for x in lvl] for lvl in tree] == exp.tree


@pytest.mark.parametrize("world", [1, 2, 4])
def test_hyrax_row_split(world):
    rnd = random.Random(3)
    rows, cols = 8, 4
    gens = PALLAS.multiples(cols)
    M = [[rnd.randrange(200) for _ in range(cols)] for _ in range(rows)]
    exp = [_points_bytes(PALLAS.msm(M[r], gens)) for r in range(rows)]
    posts = {}
    for phase in (0, 1):
        outs = []
        for g in range(world):
            def gather(b, g=g):
                posts[g] = b
                return [posts.get(r, b) for r in range(world)]
            outs.append(hyrax_commit_sharded(lambda a, b: exp[a:b], rows, g, world, gather))
    assert all(o == exp for o in outs)
    assert row_range(rows, world - 1, world)[1] == rows


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def gather(b):
        mine = torch.tensor(list(b), dtype=torch.uint8)
        outs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(outs, mine)
        return [bytes(o.tolist()) for o in outs]

    rnd = random.Random(5)
    doc = [rnd.randrange(5) for _ in range(32)]
    root, tree = merkle_sharded(_oracle_subtree(doc), _oracle_top, len(doc), rank, world, gather, True)
    exp = MerkleCommitment(doc)
    ok = int.from_bytes(root, "little") == exp.commitment and [[int.from_bytes(x, "little") for x in l] for l in tree] == exp.tree
    rows, cols = 4, 3
    gens = PALLAS.multiples(cols)
    M = [[rnd.randrange(100) for _ in range(cols)] for _ in range(rows)]
    pts = [_points_bytes(PALLAS.msm(M[r], gens)) for r in range(rows)]
    ok = ok and hyrax_commit_sharded(lambda a, b: pts[a:b], rows, rank, world, gather) == pts
    ret[rank] = ok
    dist.destroy_process_group()


def test_splits_under_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(180)
            assert p.exitcode == 0
        assert ret.get(0) is True and ret.get(1) is True
