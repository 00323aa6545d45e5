#!/usr/bin/env python3
"""Generates tests/golden/hotpath_v1.json: fixed inputs and the outputs the oracle (oracle/*.py) gives for
them, for every piece of the hot path.

WHAT THESE VECTORS PIN.  The reference (Rust, un-vendored crates, no toolchain in this image) cannot be
run here, so these are NOT outputs of the reference binary: they freeze the oracle -- itself pinned by the
reference's own unit tests and by upstream neptune known answers in tests/test_oracle_kats.py -- so that
(i) any later edit of the oracle that changes a result is caught on CPU, and (ii) the GPU tier compares
the CUDA path against committed bytes, not only against a recomputation.

Run from the repo root:  python tools/gen_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import poseidon as P                                    # noqa: E402
from oracle import spartan as S                                     # noqa: E402
from oracle.curves import PALLAS, VESTA                             # noqa: E402
from oracle.fields import FP, FQ                                    # noqa: E402
from oracle.merkle import MerkleCommitment                          # noqa: E402
from oracle.mle import mle_eval_fast                                # noqa: E402
from oracle.nlookup import ASCII_AB, DNA_AB, doc_transform, wit_nlookup_gadget   # noqa: E402


def hx(x):
    return None if x is None else hex(x)


def pt(P_):
    return None if P_ is None else [hex(P_[0]), hex(P_[1])]


def main():
    rnd = random.Random(20261017)
    g = {"_about": "see tools/gen_golden.py: oracle outputs for fixed inputs (oracle pinned by the reference's KATs; "
                   "NOT outputs of the reference binary)", "version": 1}

    # ---- Poseidon (neptune U4 / Standard over Fq): permutation, one-shot hashes, calc_d, tags, a sponge session
    st = [rnd.randrange(FQ) for _ in range(5)]
    g["poseidon"] = {
        "permute_in": [hx(x) for x in st], "permute_out": [hx(x) for x in P.permute(list(st))],
        "hash2_in": [hx(3), hx(FQ - 1)], "hash2_out": hx(P.hash_once([3, FQ - 1])),
        "hash4_in": [hx(x) for x in (1, 2, 3, 4)], "hash4_out": hx(P.hash_once([1, 2, 3, 4])),
        "calc_d": {"v": hx(5), "salt": hx(7), "out": hx(P.calc_d(5, 7))},
        "tag_a2s1": hx(P.io_pattern_tag([("A", 2), ("S", 1)])), "tag_a4s1": hx(P.io_pattern_tag([("A", 4), ("S", 1)])),
    }
    pat = [("A", 6), ("S", 1), ("A", 3), ("S", 2)]
    elems = [rnd.randrange(FQ) for _ in range(9)]
    sp = P.Sponge()
    sp.start(pat)
    out = []
    sp.absorb(elems[:6]); out += sp.squeeze(1); sp.absorb(elems[6:]); out += sp.squeeze(2)
    sp.finish()
    g["poseidon"]["sponge"] = {"pattern": pat, "in": [hx(x) for x in elems], "out": [hx(x) for x in out]}

    # ---- Merkle commitment (merkle_tree.rs): the reference's make_mt document and a ragged DNA document
    g["merkle"] = []
    for doc in ([2, 3, 4, 5, 6, 7, 8], doc_transform(DNA_AB, "ACGTTGCAACGTA")):
        mc = MerkleCommitment(doc)
        g["merkle"].append({"doc": doc, "root": hx(mc.commitment), "level_sizes": [len(l) for l in mc.tree],
                            "level0": [hx(x) for x in mc.tree[0]],
                            "path_wits_idx3": [[bool(a), b, hx(c)] for a, b, c in mc.path_wits(3)]})

    # ---- nlookup sum-check (r1cs.rs:2177-2393): cfg-1 document, all three tags, first step and a chained step
    udoc = doc_transform(ASCII_AB, "aaaaaaaab")
    cases = []
    for tag, table, q in (("nldoc", udoc, [8, 9, 10]), ("nl", [rnd.randrange(1 << 40) for _ in range(8)], [1, 6]),
                          ("nlhybrid", [rnd.randrange(FQ) for _ in range(32)], [0, 17, 31, 5])):
        v = [table[i] for i in q]
        dh = 0 if tag == "nl" else 0x1234567
        r1 = wit_nlookup_gadget(table, q, v, None, None, tag, dh)
        q2 = list(reversed(q))
        v2 = [table[i] for i in q2]
        prev_q = [r[0] for r in r1["rounds"]]
        r2 = wit_nlookup_gadget(table, q2, v2, prev_q, r1["next_running_claim"], tag, dh)
        def pack(r):
            return {"claim_r": hx(r["claim_r"]), "rounds": [[hx(x) for x in rd] for rd in r["rounds"]],
                    "sc_last_claim": hx(r["sc_last_claim"]), "next_running_claim": hx(r["next_running_claim"]),
                    "combined_q": [hx(x) for x in r["combined_q"]]}
        cases.append({"tag": tag, "table": [hx(x) for x in table], "doc_hash": hx(dh), "q": q, "step1": pack(r1),
                      "q2": q2, "step2": pack(r2)})
    g["nlookup"] = cases
    t = [rnd.randrange(FQ) for _ in range(16)]
    x = [rnd.randrange(FQ) for _ in range(4)]
    g["mle_eval"] = {"table": [hx(v) for v in t], "x": [hx(v) for v in x], "out": hx(mle_eval_fast(t, x))}

    # ---- MSM (commit(W) / commit(T), framework.rs:668-675): both curves, edge scalars included
    g["msm"] = []
    for name, cv in (("pallas", PALLAS), ("vesta", VESTA)):
        pts = [cv.mul(rnd.randrange(1, cv.order), cv.gen) for _ in range(12)]
        sc = [rnd.randrange(cv.order) for _ in range(12)]
        sc[0], sc[1], sc[2] = 0, 1, cv.order - 1
        g["msm"].append({"curve": name, "points": [pt(p) for p in pts], "scalars": [hx(s) for s in sc], "out": pt(cv.msm(sc, pts))})

    # ---- Spartan-side rounds (upstream nova-snark algorithm; parity unpinned, see oracle/spartan.py)
    g["sumcheck"] = []
    for field, p in (("fq", FQ), ("fp", FP)):
        for kind in (2, 4):
            tabs = [[rnd.randrange(p) for _ in range(8)] for _ in range(kind)]
            ch = [rnd.randrange(p) for _ in range(3)]
            claim0, rounds, finals, last = S.prove(tabs, ch, p)
            g["sumcheck"].append({"field": field, "kind": kind, "tables": [[hx(v) for v in tb] for tb in tabs],
                                  "challenges": [hx(c) for c in ch], "claim": hx(claim0),
                                  "rounds": [[hx(v) for v in rd] for rd in rounds], "finals": [hx(v) for v in finals]})

    path = os.path.join(ROOT, "tests", "golden", "hotpath_v1.json")
    with open(path, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
