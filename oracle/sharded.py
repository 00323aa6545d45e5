"""Multi-GPU sharding of the nlookup sum-check, stated with oracle arithmetic (test infrastructure).

This is the PROTOCOL reef_b200 runs across G = 2^gamma ranks, written with
Python ints so that (a) it can be compared with the unsharded oracle and (b) the exchange
pattern can be exercised on CPU with gloo (tests/test_sharded_gloo.py).

Rank g owns the table entries whose low gamma index bits equal g:  T_g[j] = T[j*G + g].
The MSB-first pairs (b, b + pow) of the reference (r1cs_helper.rs:457-461) then stay rank-local
for the first ell - gamma rounds.  Per round every rank contributes its local
(const, g(1), xsq); the contributions are all-gathered and summed, and EVERY rank runs the same
Poseidon transcript, so all ranks derive the same challenge.  After ell - gamma rounds each rank
holds one folded (T, EQ) pair; those G pairs are all-gathered and the last gamma rounds are
finished redundantly on every rank.
"""
from __future__ import annotations

from .fields import FQ
from .nlookup import combined_qs, logmn, nlookup_pattern
from .poseidon import Sponge


def shard_table(table, rank, world):
    return [table[j * world + rank] for j in range(len(table) // world)]


class ShardedNlookupRank:
    """State of one rank.  `exchange(list_of_values) -> list of per-rank lists` is the all-gather."""

    def __init__(self, local_table, rank, world, q, v, prev_q, prev_v, tag="nl", doc_hash=0):
        self.rank, self.world = rank, world
        self.gamma = world.bit_length() - 1
        assert 1 << self.gamma == world
        self.T = [int(x) % FQ for x in local_table]
        n_loc = len(self.T)
        self.ell = logmn(n_loc * world)
        self.ell_loc = self.ell - self.gamma
        assert n_loc == 1 << self.ell_loc
        m = len(q)
        cqs = combined_qs(list(q), self.ell)
        self.sponge = Sponge()
        self.sponge.start(nlookup_pattern(tag, m, self.ell, len(cqs)))
        query = ([] if tag == "nl" else [doc_hash]) + cqs + [x % FQ for x in v] + [x % FQ for x in prev_q] + [prev_v % FQ]
        self.sponge.absorb(query)
        self.claim_r = self.sponge.squeeze(1)[0]
        rs = [pow(self.claim_r, k + 1, FQ) for k in range(m + 1)]
        lq = list(reversed(prev_q))                      # bit t of the GLOBAL index <-> lq[t]
        sel = lambda bit, x: x % FQ if bit else (1 - x) % FQ
        c_g = 1
        for t in range(self.gamma):
            c_g = c_g * sel((rank >> t) & 1, lq[t]) % FQ
        # local eq table: dense tensor part over the remaining bits, plus the lookups this rank owns
        self.E = []
        for j in range(n_loc):
            term = rs[m] * c_g % FQ
            for t in range(self.ell_loc):
                term = term * sel((j >> t) & 1, lq[self.gamma + t]) % FQ
            self.E.append(term)
        for k in range(m):
            if q[k] % world == rank:
                self.E[q[k] // world] = (self.E[q[k] // world] + rs[k]) % FQ
        self.rounds = []

    def local_coeffs(self):
        """(const, g(1), xsq) of the current round over this rank's pairs."""
        half = len(self.T) // 2
        con = g1 = xsq = 0
        for b in range(half):
            t0, t1, e0, e1 = self.T[b], self.T[b + half], self.E[b], self.E[b + half]
            con += t0 * e0
            g1 += t1 * e1
            xsq += (t1 - t0) * (e1 - e0)
        return [con % FQ, g1 % FQ, xsq % FQ]

    def finish_round(self, all_triples):
        con = sum(t[0] for t in all_triples) % FQ
        g1 = sum(t[1] for t in all_triples) % FQ
        xsq = sum(t[2] for t in all_triples) % FQ
        x = (g1 - con - xsq) % FQ
        self.sponge.absorb([con, x, xsq])
        r = self.sponge.squeeze(1)[0]
        self.rounds.append((r, xsq, x, con))
        half = len(self.T) // 2
        if half >= 1:
            self.T = [(self.T[b] + r * (self.T[b + half] - self.T[b])) % FQ for b in range(half)]
            self.E = [(self.E[b] + r * (self.E[b + half] - self.E[b])) % FQ for b in range(half)]
        return r

    def run(self, exchange):
        for _ in range(self.ell_loc):
            self.finish_round(exchange(self.local_coeffs()))
        pairs = exchange([self.T[0], self.E[0]])         # rank g's folded pair sits at index g
        self.T = [p[0] for p in pairs]
        self.E = [p[1] for p in pairs]
        for _ in range(self.gamma):
            self.finish_round([self.local_coeffs()])     # identical on every rank, no exchange
        self.sponge.finish()
        r, xsq, x, con = self.rounds[-1]
        return {"claim_r": self.claim_r, "rounds": self.rounds,
                "sc_last_claim": (xsq * r * r + x * r + con) % FQ,
                "next_running_claim": self.T[0], "next_running_q": [t[0] for t in self.rounds]}


def run_single_process(table, world, q, v, prev_q, prev_v, tag="nl", doc_hash=0):
    """All ranks in one process (lock-step), the all-gather is a Python list."""
    ranks = [ShardedNlookupRank(shard_table(table, g, world), g, world, q, v, prev_q, prev_v, tag, doc_hash)
             for g in range(world)]
    gamma = world.bit_length() - 1
    for _ in range(ranks[0].ell_loc):
        triples = [r.local_coeffs() for r in ranks]
        for r in ranks:
            r.finish_round(triples)
    pairs = [[r.T[0], r.E[0]] for r in ranks]
    outs = []
    for r in ranks:
        r.T = [p[0] for p in pairs]
        r.E = [p[1] for p in pairs]
        for _ in range(gamma):
            r.finish_round([r.local_coeffs()])
        rr, xsq, x, con = r.rounds[-1]
        outs.append({"claim_r": r.claim_r, "rounds": r.rounds, "sc_last_claim": (xsq * rr * rr + x * rr + con) % FQ,
                     "next_running_claim": r.T[0], "next_running_q": [t[0] for t in r.rounds]})
    return outs
