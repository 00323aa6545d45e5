"""Host-side logic of the multi-GPU splits that need one real exchange:

  * Hyrax document commitment (commitment.rs:187): the 2^left rows are independent MSMs over the same
    generators -> rank g commits rows [g R/G, (g+1) R/G), one all-gather of 64-byte points.
  * Merkle commitment (merkle_tree.rs:25-78): contiguous leaf ranges are independent subtrees (the leaf
    hash takes the GLOBAL index) -> one all-gather of the G subtree roots (plus the level slices when the
    whole tree has to be materialised for `path_wits`), the top log2 G levels are replicated.

The compute steps are passed in as callables so that the same partition / reassembly code runs against
libreef_b200 on GPUs (backend.py, bench.py) and against the oracle under gloo in the CPU test tier.
`gather(payload: bytes) -> list[bytes]` is the all-gather (rank-major), e.g. NCCL `all_gather` of uint8."""
from __future__ import annotations


def row_range(rows: int, rank: int, world: int):
    if rows % world:
        raise ValueError("the number of matrix rows must be a multiple of the world size")
    per = rows // world
    return rank * per, (rank + 1) * per


def hyrax_commit_sharded(commit_rows, rows: int, rank: int, world: int, gather):
    """commit_rows(r0, r1) -> list of (r1 - r0) points as 64-byte strings.  Returns all `rows` commitments."""
    r0, r1 = row_range(rows, rank, world)
    mine = commit_rows(r0, r1)
    if len(mine) != r1 - r0 or any(len(p) != 64 for p in mine):
        raise ValueError("commit_rows must return one 64-byte point per row")
    parts = gather(b"".join(mine))
    out = []
    for g, blob in enumerate(parts):
        a, b = row_range(rows, g, world)
        if len(blob) != 64 * (b - a):
            raise ValueError(f"rank {g} sent {len(blob)} bytes for {b - a} rows")
        out.extend(blob[i * 64:(i + 1) * 64] for i in range(b - a))
    return out


def merkle_sharded(build_subtree, hash_top, n_doc: int, rank: int, world: int, gather, full_tree: bool = True):
    """build_subtree(leaf_begin, leaf_end) -> (levels: list of lists of 32-byte nodes, leaf parents first, root: bytes)
    hash_top(roots: list[bytes]) -> (levels of G/2, G/4, .. 1 nodes, root)
    Returns (root, tree) with tree = every level of the whole tree (merkle_tree.rs `tree`) or None."""
    if world & (world - 1) or n_doc % (2 * world):
        raise ValueError("world must be a power of two and the padded document a multiple of 2 * world leaves")
    per = n_doc // world
    levels, root = build_subtree(rank * per, (rank + 1) * per)
    roots = gather(root)
    top_levels, top_root = hash_top(list(roots)) if world > 1 else ([], roots[0])
    if not full_tree:
        return top_root, None
    n_sub_levels = len(levels)
    tree = []
    for lvl in range(n_sub_levels):
        blobs = gather(b"".join(levels[lvl]))
        nodes = []
        for blob in blobs:
            nodes.extend(blob[i:i + 32] for i in range(0, len(blob), 32))
        tree.append(nodes)
    tree.extend(top_levels)
    return top_root, tree
