"""Poseidon Merkle commitment over (idx, char) leaves (oracle; test infrastructure only).

Follows /root/reference/src/backend/merkle_tree.rs:25-192.
"""
from __future__ import annotations

from .poseidon import hash_once


def new_parent(left, right):
    """merkle_tree.rs:80-114.  left = (idx|None, c); right = (idx|None, c) | None."""
    li, lc = left
    if li is not None and right is not None and right[0] is not None:
        query = [li, lc, right[0], right[1]]
    elif li is not None and right is None:
        query = [li, lc, 0, 0]
    elif li is None and right is not None and right[0] is None:
        query = [lc, right[1]]
    elif li is None and right is None:
        query = [lc, 0]
    else:
        raise ValueError("not a correctly formatted leaf or parent")
    return hash_once(query)


class MerkleCommitment:
    """merkle_tree.rs:10-78: fields commitment, tree (levels, leaf parents first), doc."""

    def __init__(self, doc):
        self.doc = [int(c) for c in doc]
        tree = []
        level = []
        i = 0
        while i < len(doc):
            left = (i, doc[i])
            right = (i + 1, doc[i + 1]) if i + 1 < len(doc) else None
            level.append(new_parent(left, right))
            i += 2
        tree.append(list(level))
        while len(level) > 1:
            prev, level = level, []
            i = 0
            while i < len(prev):
                right = (None, prev[i + 1]) if i + 1 < len(prev) else None
                level.append(new_parent((None, prev[i]), right))
                i += 2
            tree.append(list(level))
        self.tree = tree
        self.commitment = level[0]

    def path_wits(self, idx):
        """merkle_tree.rs:128-191: list of (l_or_r, opposite_idx|None, opposite)."""
        assert idx < len(self.doc)
        wits = []
        if idx % 2 == 0:
            if idx + 1 >= len(self.doc):
                wits.append((True, 0, 0))
            else:
                wits.append((True, idx + 1, self.doc[idx + 1]))
        else:
            wits.append((False, idx - 1, self.doc[idx - 1]))
        quo = idx // 2
        for h in range(len(self.tree) - 1):
            if quo % 2 == 0:
                if quo + 1 >= len(self.tree[h]):
                    wits.append((True, None, 0))
                else:
                    wits.append((True, None, self.tree[h][quo + 1]))
            else:
                wits.append((False, None, self.tree[h][quo - 1]))
            quo //= 2
        return wits

    def make_wits(self, lookups):
        """merkle_tree.rs:116-126."""
        return [self.path_wits(q) for q in lookups]
