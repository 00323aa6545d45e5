"""nlookup witness driver, doc encoding and helpers (oracle; test infrastructure only).

Follows /root/reference/src/backend/r1cs.rs:2177-2393 (`R1CS::wit_nlookup_gadget`),
/root/reference/src/backend/costs.rs:10-15 (`logmn`) and
/root/reference/src/backend/framework.rs:978-1011 (`doc_transform`).
"""
from __future__ import annotations

import math

import numpy as np

from .fields import FQ
from .mle import gen_eq_table, gen_eq_table_fast, linear_mle_product, prover_mle_partial_eval, mle_eval_fast
from .poseidon import ABSORB, SQUEEZE, Sponge


def logmn(mn: int) -> int:
    """costs.rs:10-15: `(mn as f32).log2().ceil()`, with the mn == 1 special case."""
    if mn == 1:
        return 1
    return int(math.ceil(float(np.log2(np.float32(mn)))))


ASCII_AB = "".join(chr(c) for c in range(128))
DNA_AB = "ACGT"


def doc_transform(ab: str, doc: str):
    """framework.rs:978-1011.  Later inserts overwrite earlier ones (FxHashMap::insert)."""
    num_ab = {}
    i = 0
    for c in ab:
        num_ab[c] = i
        i += 1
    num_ab[None] = i + 1        # EPSILON
    num_ab[chr(26)] = i + 2     # EOF (overwrites ab's own entry for char 26 if present)
    udoc = []
    for c in doc:
        if c not in num_ab:
            raise ValueError("Character in document that's not in alphabet")
        udoc.append(num_ab[c])
    udoc.append(num_ab[chr(26)])
    udoc.append(num_ab[None])
    ext = (1 << logmn(len(udoc))) - len(udoc)
    if ext < 0:
        raise OverflowError("attempt to subtract with overflow")
    udoc.extend([0] * ext)
    return udoc


def combined_qs(q, sc_l):
    """r1cs.rs:2208-2243: the bit-packing loop, quirks included (the flush-triggering bit is
    skipped; the outer `while` may re-run the whole pass)."""
    num_vs = len(q)
    num_cqs = int(math.ceil((num_vs * sc_l) / 254.0))
    out = []
    cq = 0
    while cq < num_cqs:
        combined_q, next_slot = 0, 1
        for i in range(num_vs):
            qjs = [(q[i] >> j) & 1 for j in range(sc_l)]
            j = 0
            for qj in reversed(qjs):
                if (i * sc_l) + j >= 254 * (cq + 1) or (i == num_vs - 1 and j == sc_l - 1):
                    cq += 1
                    out.append(combined_q)
                    combined_q, next_slot = 0, 1
                else:
                    combined_q += qj * next_slot
                    next_slot *= 2
                j += 1
    assert num_cqs == len(out), "assert_eq!(num_cqs, combined_qs.len())"
    return out


def nlookup_pattern(tag: str, num_vs: int, sc_l: int, num_cqs: int):
    """r1cs.rs:2263-2282."""
    if tag == "nl":
        first = num_vs + sc_l + 1 + num_cqs
    elif tag in ("nldoc", "nlhybrid"):
        first = num_vs + sc_l + 2 + num_cqs
    else:
        raise ValueError("weird tag")
    pat = [(ABSORB, first), (SQUEEZE, 1)]
    for _ in range(sc_l):
        pat += [(ABSORB, 3), (SQUEEZE, 1)]
    return pat


def wit_nlookup_gadget(table, q, v, running_q=None, running_v=None, tag="nl", doc_hash=0, fast=False):
    """r1cs.rs:2177-2393.  Returns a dict with every value the reference writes to `wits`
    (keyed like the reference's wire names minus the `{id}_` prefix) plus
    next_running_q / next_running_v.

    fast=True swaps gen_eq_table / prover_mle_partial_eval for their O(N) equivalents
    (identical values) so that large parity sizes finish in seconds.
    """
    table = [int(t) for t in table]
    sc_l = logmn(len(table))
    num_vs = len(v)
    assert num_vs == len(q)
    prev_q = list(running_q) if running_q is not None else [0] * sc_l
    prev_v = running_v if running_v is not None else table[0]
    assert len(prev_q) == sc_l

    cqs = combined_qs(q, sc_l)
    pattern = nlookup_pattern(tag, num_vs, sc_l, len(cqs))
    sponge = Sponge()
    sponge.start(pattern)
    query = [] if tag == "nl" else [doc_hash]
    query += cqs
    query += [x % FQ for x in v]
    query += [x % FQ for x in prev_q]
    query.append(prev_v % FQ)
    sponge.absorb(query)
    claim_r = sponge.squeeze(1)[0]

    rs = [claim_r]
    for i in range(1, len(q) + 1):
        rs.append(rs[i - 1] * claim_r % FQ)
    eq_fn = gen_eq_table_fast if fast else gen_eq_table
    eq_table = eq_fn(rs, q, list(reversed(prev_q)))
    sc_table = list(table)
    if tag == "nldoc":
        sc_table += [0] * ((1 << logmn(len(table))) - len(table))
    assert len(sc_table) == 1 << sc_l, "assert_eq!(table_t.len(), base.pow(ell))"

    rounds = []
    sc_rs = []
    sc_r = g_xsq = g_x = g_const = 0
    for i in range(1, sc_l + 1):
        prev_g_r = (g_xsq * sc_r * sc_r + g_x * sc_r + g_const) % FQ
        sc_r, g_xsq, g_x, g_const = linear_mle_product(sc_table, eq_table, sc_l, i, sponge)
        if i > 1:
            assert prev_g_r == (g_xsq + g_x + 2 * g_const) % FQ
        rounds.append((sc_r, g_xsq, g_x, g_const))
        sc_rs.append(sc_r)
    sponge.finish()

    last_claim = (g_xsq * sc_r * sc_r + g_x * sc_r + g_const) % FQ
    if fast:
        next_v = mle_eval_fast(table, sc_rs)
    else:
        _, next_v = prover_mle_partial_eval(table, sc_rs, list(range(len(table))), True, None)
    return {
        "prev_running_claim": prev_v % FQ,
        "combined_q": cqs,
        "claim_r": claim_r,
        "rounds": rounds,               # per round i=1..ell: (sc_r_i, sc_g_i_xsq, sc_g_i_x, sc_g_i_const)
        "sc_last_claim": last_claim,
        "next_running_claim": next_v,
        "next_running_q": sc_rs,
        "folded_t0": sc_table[0],
        "folded_eq0": eq_table[0],
    }
