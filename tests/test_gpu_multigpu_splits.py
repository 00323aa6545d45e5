"""GPU tier, all ranks on one GPU (one context = one stream per rank, mailboxes connected by device
pointer): the window-sharded MSM with its 128-byte all-gather done by the library's own mailbox kernel,
the Merkle subtree / top split and the Hyrax row split against the un-sharded results."""
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import reef_b200
import workloads as W
from oracle.curves import PALLAS, VESTA
from oracle.merkle import MerkleCommitment as OracleMerkle

pytestmark = pytest.mark.gpu


def _world(world):
    ctxs = [reef_b200.Context(0) for _ in range(world)]
    for c in ctxs:
        c.mailbox_create(world)
    ptrs = [c.mailbox_ptr() for c in ctxs]
    for g, c in enumerate(ctxs):
        c.mailbox_connect_local(g, world, ptrs)
    return ctxs


# world <= 4 here: with all ranks on ONE device their streams share its 8 hardware work queues
# (CUDA_DEVICE_MAX_CONNECTIONS), and a rank queued behind a peer's waiting exchange kernel can never post;
# on a real box every rank has its own device (bench.py --gpus 8 runs this path on 8 GPUs)
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("curve,n", [("pallas", 1 << 10), ("vesta", 3000)])
def test_msm_window_sharded_with_mailbox_gather(world, curve, n):
    import torch
    cv = PALLAS if curve == "pallas" else VESTA
    rnd = random.Random(n + world)
    gens = W.generators(curve, n)
    sc = [rnd.randrange(cv.order) for _ in range(n)]
    raw = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in sc), dtype=np.uint8)
    ctxs = _world(world)
    bases = []
    try:
        bases = [c.bases(curve, gens) for c in ctxs]
        # un-sharded on every context first: the reference result, and on ONE GPU it also sizes each context's
        # scratch up front (a cudaMalloc while a peer's exchange kernel spins would dead-lock this device)
        whole = [b.msm(sc) for b in bases][0]
        dev = torch.from_numpy(raw.copy()).cuda()
        torch.cuda.synchronize()
        with ThreadPoolExecutor(max_workers=world) as ex:      # every rank's call waits for its peers' partials
            got = list(ex.map(lambda b: b.msm_sharded_dev(dev.data_ptr(), n), bases))
        assert all(g == whole for g in got)
        # closed form: sum k_i * (i+1) G
        k = sum(s * (i + 1) for i, s in enumerate(sc)) % cv.order
        assert whole == cv.mul(k, cv.multiples(1)[0])
    finally:
        for b in bases:
            b.free()
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("world", [1, 2, 8])
def test_merkle_subtrees_and_top(ctx, world):
    rnd = random.Random(world)
    n = 1 << 9
    doc = [rnd.randrange(131) for _ in range(n)]
    exp = OracleMerkle(doc)
    posts = {}
    outs = []
    for phase in (0, 1):
        outs = []
        for g in range(world):
            k = [0]

            def gather(b, g=g, k=k):
                posts.setdefault(k[0], {})[g] = b
                res = [posts[k[0]].get(r, b) for r in range(world)]
                k[0] += 1
                return res
            per = n // world
            outs.append(reef_b200.MerkleCommitment.build_sharded(ctx, doc[g * per:(g + 1) * per], n, g, world, gather, full_tree=True))
    for root, tree in outs:
        assert root == exp.commitment
        assert tree == exp.tree


@pytest.mark.parametrize("world", [2, 4])
def test_hyrax_rows_split(ctx, world):
    rnd = random.Random(9)
    rows, cols = 16, 64
    m = np.asarray([rnd.randrange(131) for _ in range(rows * cols)], dtype=np.uint32).reshape(rows, cols)
    gens = W.generators("pallas", cols + 1)
    blinds = [rnd.randrange(PALLAS.order) for _ in range(rows)]
    b = ctx.bases("pallas", gens, 255)
    try:
        whole = b.msm_rows(m, rows, cols, entry_bits=8, blinds=blinds)
        posts = {}
        for phase in (0, 1):
            outs = []
            for g in range(world):
                def gather(x, g=g):
                    posts[g] = x
                    return [posts.get(r, x) for r in range(world)]
                outs.append(b.commit_rows_sharded(m, rows, cols, 8, blinds, g, world, gather))
        assert all(o == whole for o in outs)
    finally:
        b.free()
