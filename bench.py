#!/usr/bin/env python3
"""bench.py -- Reef prover hot path on B200 (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--workload target] [--impl reef|reference]

A "step" is ONE PASS OF THE PROVER HOT PATH over one synthetic document: for each Nova fold of the
`--prove` run the reference would do (framework.rs:405-625 / 642-754), by mode of the workload:
    nldoc   nl sum-check over T ("nl", r1cs.rs:2137-2161) + nldoc sum-check over the committed document,
            calc_d of the previous and the next document claim (framework.rs:517-526)
    hybrid  ONE nlhybrid sum-check over the merged table T ++ document (r1cs.rs:2101-2136), 2 calc_d
    merkle  nl sum-check over T + Merkle path witnesses of the document lookups (r1cs.rs:2088-2100,
            merkle_tree.rs:116-192)
  and the prove_step commitments commit(W), commit(T) on Pallas and on Vesta (framework.rs:668-675).
metric  = NFA steps/s proved = doc_len / time of one pass        (BASELINE.json)
workload = `target` by default: configs[1] at the 2^20-char document the north_star target names; the
          other BASELINE configs (cfg2..cfg5) are timed in the same run and reported under "also",
          each with its commit phase (Hyrax rows / Merkle tree), each verified against the oracle
value   = inputs resident in HBM when the timed region starts    (device timed, CUDA events)
e2e     = the same pass through the C ABI with HOST buffers: document/table upload, scalar
          upload and result read-back inside the timed region
--impl reference = the CPU restatement of the reference's algorithm (oracle/c, all host cores).

Multi-GPU (torchrun), weak scaling: ONE document of base_len * G characters.  Its sum-check is sharded by
low index bits; the 96 bytes per rank per round are exchanged by the round kernels themselves through
peer mailboxes over NVLink (P2P stores + system-scope release/acquire; no NCCL call on that path), every
rank runs the same transcript.  The fold commitments (latency-bound at 2^14..2^15 terms) are distributed
whole, round-robin over the ranks.  The line also carries, measured on the same ranks: the same document
on ONE GPU (`same_doc_1gpu`), a 2^20-term MSM sharded by Pippenger windows with its 128-byte all-gather
done by a kernel of the library (`msm_sharded`), and the commit phase split over the ranks (`commit`).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import random
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import workloads as WL                                  # numpy-only input generators shared by both arms and the tests

FQ, FP = WL.FQ, WL.FP
ASCII_AB = "".join(chr(c) for c in range(128))        # config.rs: the `ascii` alphabet
curve_multiples = WL.curve_multiples

# name: document, length, mode, Nova folds, lookups per fold, log2 |T|, primary / secondary MSM sizes, entry bits
# of the document codes.  Fold / lookup counts and circuit sizes are SYNTHETIC (the frontend and CirC that
# would fix them are out of scope); they are stated in `desc`.
WORKLOADS = {
    "target": dict(doc="target", doc_len=1 << 20, mode="nldoc", steps=2, m=2, t_log=6, n_pri=1 << 15, n_sec=1 << 14, bits=8,
                   desc="north_star target = configs[1] at 2^20 chars: ascii doc (2^20-1 x 'a' + 'b'), re '.*b', --prove: "
                        "per Nova fold nl(T=2^6) + nldoc(N=2^21,u32) sum-checks, 2 calc_d, commit(W),commit(T) on "
                        "Pallas(2^15) and Vesta(2^14); 2 folds"),
    "cfg2": dict(doc="cfg2", doc_len=1 << 16, mode="nldoc", steps=2, m=2, t_log=6, n_pri=1 << 15, n_sec=1 << 14, bits=8,
                 desc="configs[1]: ascii 2^16-char doc (65535 x 'a' + 'b'), re '.*b', --prove: per Nova fold "
                      "nl(T=2^6) + nldoc(N=2^17,u32) sum-checks, 2 calc_d, commit(W),commit(T) on "
                      "Pallas(2^15) and Vesta(2^14); 2 folds"),
    "cfg3": dict(doc="cfg3", doc_len=1 << 20, mode="merkle", steps=3, m=4, t_log=8, n_pri=1 << 16, n_sec=1 << 14, bits=8,
                 desc="configs[2]: dna 2^20-char doc, re '(A|C|G|T){4}TATA.*', --merkle --prove: per Nova fold nl(T=2^8) sum-check + "
                      "4 Merkle path witnesses (21 levels), commit(W),commit(T) on Pallas(2^16) and Vesta(2^14); 3 folds; "
                      "commit phase = Poseidon tree over 2^21 leaves"),
    "cfg4": dict(doc="cfg4", doc_len=1 << 20, mode="hybrid", steps=2, m=4, t_log=9, n_pri=1 << 16, n_sec=1 << 14, bits=8,
                 desc="configs[3]: ascii 2^20-char doc, re 'hello.*world', --hybrid --projections (projection resolves to None): per "
                      "Nova fold ONE nlhybrid sum-check over the merged 2^22 table (T=2^9 padded to 2^21 ++ document), 2 calc_d, "
                      "commit(W),commit(T) on Pallas(2^16) and Vesta(2^14); 2 folds; commit phase = Hyrax 1024 x 2048"),
    "cfg5": dict(doc="cfg5", doc_len=1 << 22, mode="nldoc", steps=1, m=4, t_log=8, n_pri=1 << 16, n_sec=1 << 14, bits=21,
                 desc="configs[4]: utf8 2^22-char doc (code points up to 0x2FFF, EOF/EPSILON codes 21 bits), re '.*(foo|bar|baz).*', "
                      "--prove: nl(T=2^8) + nldoc(N=2^23,u32) sum-checks, 2 calc_d, commit(W),commit(T) on Pallas(2^16) and "
                      "Vesta(2^14); 1 fold; commit phase = Hyrax 2048 x 4096"),
}


TRACE = [] if os.environ.get("REEF_BENCH_TRACE") == "1" else None      # host timestamps of a pass (debugging aid)


class ParityError(AssertionError):
    """The GPU path produced a result that differs from the oracle."""


def le32(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


def pack(xs) -> bytes:
    return b"".join(le32(x) for x in xs)


# --------------------------------------------------------------------------------------------
# synthetic workload (deterministic; the same bytes feed the GPU arm and the reference arm)
# --------------------------------------------------------------------------------------------
def make_workload(name: str, seed_shift: int = 0, world: int = 1):
    """Inputs of one pass; no import of reef_b200 or oracle/ (both arms call this)."""
    w = dict(WORKLOADS[name])
    w["name"] = name
    rnd = random.Random(1234 + seed_shift)
    # weak scaling: one document of base_len * G characters.  The reference's f32 `logmn` (costs.rs:10-15)
    # mis-rounds 2^22+2 and 2^23+2, so its doc_transform panics on documents of exactly 2^22 / 2^23
    # characters (framework.rs:1007): 64 more characters put the length where logmn is exact; the padded
    # table (2^21 entries per GPU) is unchanged.
    w["doc_len"] = WL.safe_len(w["doc_len"] * world)
    doc_len = w["doc_len"]
    ab, cps = WL.document(w["doc"], doc_len, seed_shift)
    w["udoc"] = np.ascontiguousarray(WL.encode(ab, cps))
    tl = w["t_log"]
    w["T"] = sorted(rnd.randrange(1 << 100) for _ in range(1 << tl))
    w["T_bytes"] = pack(w["T"])
    w["fill"] = rnd.randrange(1 << 100)                       # calc_fill value of the padded T (r1cs.rs:363-388)
    S, m = w["steps"], w["m"]
    n_doc_tab = len(w["udoc"])
    if w["mode"] == "hybrid":
        # hybrid_q = q ++ (doc_q + half_len) (r1cs.rs:2114-2117): m/2 lookups into T, m/2 into the document
        half = max(n_doc_tab, 1 << tl)
        w["half"] = half
        w["q_hyb"] = [[rnd.randrange(1 << tl) for _ in range(m // 2)] + [half + rnd.randrange(doc_len + 2) for _ in range(m - m // 2)]
                      for _ in range(S)]
    else:
        w["q_nl"] = [[rnd.randrange(1 << tl) for _ in range(m)] for _ in range(S)]
        w["q_doc"] = [[rnd.randrange(doc_len + 2) for _ in range(m)] for _ in range(S)]
    w["doc_hash"] = rnd.randrange(FQ)
    w["salt"] = rnd.randrange(FQ)

    def scalars(n, witness_like):
        rs = np.random.default_rng(rnd.randrange(1 << 30))
        raw = rs.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
        raw[:, 3] &= (1 << 61) - 1                                    # < 2^253 < both group orders
        if witness_like:                                              # R1CS witness: mostly bits / small values
            small = rs.random(n) < 0.85
            raw[small, 1:] = 0
            raw[small, 0] = rs.integers(0, 1 << 16, size=int(small.sum()), dtype=np.uint64)
            bits = rs.random(n) < 0.5
            raw[small & bits, 0] &= 1
        return np.ascontiguousarray(raw)

    w["bases_pri"] = WL.generators("pallas", w["n_pri"])     # Pallas: over Fp  (disk-cached k*G, build/gens/)
    w["bases_sec"] = WL.generators("vesta", w["n_sec"])      # Vesta: over Fq
    w["sc"] = [dict(Wp=scalars(w["n_pri"], True), Tp=scalars(w["n_pri"], False),
                    Ws=scalars(w["n_sec"], True), Ts=scalars(w["n_sec"], False)) for _ in range(S)]
    return w


def hybrid_value(w, i):
    """entry i of the merged table (r1cs.rs:2105-2112)"""
    half, nT = w["half"], len(w["T"])
    if i < half:
        return w["T"][i] if i < nT else w["fill"]
    return int(w["udoc"][(i - half) % len(w["udoc"])])


def wit_bytes(path):
    """one MerkleWit list (merkle_tree.rs:128-191) as bytes for the comparison"""
    return b"".join((1 if lr else 0).to_bytes(1, "little") + (2 ** 64 - 1 if oi is None else int(oi)).to_bytes(8, "little") + le32(opp)
                    for lr, oi, opp in path)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class GpuPass:
    """One prove pass through libreef_b200 with raw buffers (no Python big-int work inside)."""

    collect = None

    def __init__(self, ctxs, w, rank=0, world=1, dist=None):
        """ctxs: dict of libreef_b200 contexts (one CUDA stream each).  The reference runs the sum-checks on its
        solver thread and the fold commitments on its proving thread (framework.rs:98-110); here each of the
        independent chains of a fold has its own context/stream and host thread."""
        import reef_b200
        import torch
        from concurrent.futures import ThreadPoolExecutor
        self.rb, self.torch, self.ctxs, self.w = reef_b200, torch, ctxs, w
        self.lib, self.check = reef_b200.lib, reef_b200._lib.check
        self.rank, self.world, self.dist = rank, world, dist
        self.mode = w["mode"]
        # generators are static per PublicParams: registered once per context that commits with them (commit(T)
        # has its own context so that it does not queue behind commit(W) of the same curve)
        self.bases = {"Wp": reef_b200.Bases(ctxs["pri"], "pallas", w["bases_pri"]), "Ws": reef_b200.Bases(ctxs["sec"], "vesta", w["bases_sec"]),
                      "Tp": reef_b200.Bases(ctxs["pri2"], "pallas", w["bases_pri"]), "Ts": reef_b200.Bases(ctxs["sec2"], "vesta", w["bases_sec"])}
        self.ell_doc = reef_b200.logmn(len(w["udoc"]))
        self.ell_T = w["t_log"]
        self.pool = {k: ThreadPoolExecutor(max_workers=1) for k in ctxs}
        self._pins = []
        self.h_doc = self._pin(self._doc_shard())
        self.h_T = self._pin(np.frombuffer(w["T_bytes"], dtype=np.uint8))
        for s in w["sc"]:
            for k in list(s):
                s[k] = self._pin(s[k])
        self.pair_host = [{"p": self._pin(np.concatenate([s["Wp"], s["Tp"]])), "s": self._pin(np.concatenate([s["Ws"], s["Ts"]]))} for s in w["sc"]]
        self.pairs = world < 4          # W and T of a fold as the two rows of ONE row-batched MSM per curve
        self.tree = None
        self._mk_out()

    def _pin(self, arr):
        t = self.torch.from_numpy(np.array(arr, copy=True)).pin_memory()
        self._pins.append(t)
        return t.numpy()

    def _mk_out(self):
        from reef_b200._lib import NlookupOut
        self.o, self.bufs = {}, {}
        for key, ell in (("nl", self.ell_T), ("nldoc", self.ell_doc), ("nlhybrid", self.ell_doc + 1)):
            b = dict(prev=C.create_string_buffer(32), cq=C.create_string_buffer(64 * 32), claim=C.create_string_buffer(32),
                     rounds=C.create_string_buffer((ell + 2) * 128), last=C.create_string_buffer(32), nxt=C.create_string_buffer(32))
            o = NlookupOut()
            o.prev_running_claim = C.addressof(b["prev"]); o.combined_q = C.addressof(b["cq"]); o.combined_q_cap = 64
            o.claim_r = C.addressof(b["claim"]); o.rounds = C.addressof(b["rounds"]); o.rounds_cap = ell + 2
            o.sc_last_claim = C.addressof(b["last"]); o.next_running_claim = C.addressof(b["nxt"])
            self.o[key], self.bufs[key] = o, b

    # -- residency ----------------------------------------------------------------------
    def make_resident(self):
        w, t = self.w, self.torch
        if self.mode == "hybrid":
            self.hyb_tab = self.ctxs["doc"].table_hybrid(w["T_bytes"], w["fill"], w["half"], w["udoc"])
        elif self.mode == "nldoc":
            self.doc_tab = self.ctxs["doc"].table_u32(self._doc_shard())
        if self.mode != "hybrid":
            self.T_tab = self.rb.Table(self.ctxs["nl"], values=w["T"])
        if self.mode == "merkle":
            # the tree is the product of --commit (read back from the .cmt by --prove, main.rs:48-51): built once
            self.tree = self.ctxs["doc"].merkle(w["udoc"])
        if self.world > 1:
            self._connect_mailboxes()
        self.sc_dev = [{k: t.from_numpy(v.view(np.int64)).cuda() for k, v in s.items()} for s in w["sc"]]
        # commit(W) and commit(T) of a fold use the same commitment key: resident as the two rows of one matrix
        self.pair_dev = [{"p": t.cat([d["Wp"], d["Tp"]]).contiguous(), "s": t.cat([d["Ws"], d["Ts"]]).contiguous()} for d in self.sc_dev]
        t.cuda.synchronize()

    def _doc_shard(self):
        return np.ascontiguousarray(self.w["udoc"][self.rank::self.world]) if self.world > 1 else self.w["udoc"]

    def _connect_mailboxes(self):
        """Exchanges over NVLink peer memory (reef_b200/csrc/p2p.cu): every rank maps every peer's mailbox with
        CUDA IPC once; NCCL only carries the 64-byte handles.  One mailbox per context that exchanges."""
        t = self.torch
        for key in ("doc", "pri"):
            ctx = self.ctxs[key]
            mine = t.frombuffer(bytearray(ctx.mailbox_create(self.world)), dtype=t.uint8).cuda()
            allh = t.empty(self.world * 64, dtype=t.uint8, device="cuda")
            self.dist.all_gather_into_tensor(allh, mine)
            ctx.mailbox_connect(self.rank, self.world, bytes(allh.cpu().numpy().tobytes()))
        self.dist.barrier()

    # -- pieces of a fold ------------------------------------------------------------------
    def _calc_d(self, v):
        d = C.create_string_buffer(32)
        self.check(self.lib.reef_calc_d(self.ctxs["aux"]._h, v, self.salt, d))
        return d.raw

    def _nlookup(self, key, tab, q_arr, v_bytes, prev, first_v):
        o, b = self.o[key], self.bufs[key]
        tag = {"nl": 0, "nldoc": 1, "nlhybrid": 2}[key]
        pq = prev[0] if prev else None
        pv = prev[1] if prev else None
        ctx = self.ctxs["nl" if key == "nl" else "doc"]
        self.check(self.lib.reef_nlookup_prove(ctx._h, tag, tab._h, q_arr.ctypes.data, v_bytes, len(q_arr), pq, pv,
                                              self.dh if tag else None, C.byref(o)))
        ell = o.ell
        rounds = b["rounds"].raw
        next_q = b"".join(rounds[i * 128:i * 128 + 32] for i in range(ell))
        nxt = b["nxt"].raw
        if key != "nl" and os.environ.get("REEF_BENCH_SKIP_CALCD") != "1":
            # calc_d of the previous and of the next running claim (framework.rs:517-553): inputs of the step
            # circuit only, the next fold's sum-check does not wait for them
            self.d_futs.append(self.pool["aux"].submit(self._calc_d, pv if prev else first_v))
            self.d_futs.append(self.pool["aux"].submit(self._calc_d, nxt))
        if self.collect is not None:
            self.collect.append((key, b["claim"].raw, rounds[:ell * 128], b["last"].raw, nxt))
        return next_q, nxt

    def _nlookup_sharded(self, tab, q_list, v_list, prev):
        w = self.w
        ell = self.ell_doc
        if prev is None:
            pq, pv = [0] * ell, int(w["udoc"][0])
        else:
            pq = [int.from_bytes(prev[0][i * 32:(i + 1) * 32], "little") for i in range(ell)]
            pv = int.from_bytes(prev[1], "little")
        sn = self.rb.ShardedNlookup(self.ctxs["doc"], tab, self.rank, self.world, q_list, v_list, pq, pv, "nldoc", w["doc_hash"])
        res = sn.run_p2p()          # exchange fused into the round kernels (peer mailboxes over NVLink)
        sn.free()
        nxt = le32(res.next_running_claim)
        self.d_futs.append(self.pool["aux"].submit(self._calc_d, le32(pv)))
        self.d_futs.append(self.pool["aux"].submit(self._calc_d, nxt))
        if self.collect is not None:
            self.collect.append(("nldoc", le32(res.claim_r), b"".join(pack(r) for r in res.rounds), le32(res.sc_last_claim), nxt))
        return pack(res.next_running_q), nxt

    def _merkle_wits(self, lookups):
        """mc.make_wits(lookups) (framework.rs:576-580): host-side walk of the committed tree, as in the reference"""
        return b"".join(wit_bytes(self.tree.path_wits(int(q))) for q in lookups)

    def _msm_local(self, bases, dev_tensor, host_arr, n, resident):
        out = C.create_string_buffer(64)
        ctx = bases.ctx
        if resident:
            self.check(self.lib.reef_msm_dev(ctx._h, bases._h, C.c_void_p(dev_tensor.data_ptr()), n, out))
        else:
            self.check(self.lib.reef_msm(ctx._h, bases._h, host_arr.ctypes.data, n, out))
        return out.raw

    def _msm_pair(self, bases, dev_tensor, host_arr, n, resident):
        """commit(W), commit(T) of one fold over the same key: one row-batched MSM (2 x n), 2 x 64 bytes back"""
        out = C.create_string_buffer(128)
        ctx = bases.ctx
        if TRACE is not None:
            TRACE.append(("msm_start", n, time.perf_counter()))
        if resident:
            self.check(self.lib.reef_msm_rows_dev(ctx._h, bases._h, C.c_void_p(dev_tensor.data_ptr()), 2, n, out))
        else:
            self.check(self.lib.reef_msm_rows(ctx._h, bases._h, host_arr.ctypes.data, 2, n, None, out))
        if TRACE is not None:
            TRACE.append(("msm_end", n, time.perf_counter()))
        return out.raw

    def _exchange_points(self, outs, owners):
        """One all-gather per pass: every rank ends up with all commitments (64 B each, 128 B per W/T pair)."""
        t = self.torch
        k, sz = len(outs), (128 if self.pairs else 64)
        mine = bytearray(k * sz)
        for i, o in enumerate(outs):
            if o is not None:
                mine[i * sz:(i + 1) * sz] = o
        buf = t.frombuffer(mine, dtype=t.uint8).cuda()
        allb = t.empty(self.world * k * sz, dtype=t.uint8, device="cuda")
        self.dist.all_gather_into_tensor(allb, buf)
        host = allb.cpu().numpy().tobytes()
        return [outs[i] if owners[i] is None else host[(owners[i] * k + i) * sz:(owners[i] * k + i + 1) * sz] for i in range(k)]

    def run_collect(self, resident: bool = True):
        """One untimed pass that keeps EVERY output in cpu_pass(collect=...)'s layout for the bit-for-bit comparison."""
        self.collect = []
        try:
            _, outs, ds, wits = self.run(resident)
            got = {"sumchecks": list(self.collect), "d": list(ds), "msm": list(outs), "wits": wits}
        finally:
            self.collect = None
        return got

    def run(self, resident: bool):
        """One pass.  resident=False: every input crosses PCIe inside the call (e2e leg).
        Dependencies kept: sum-check of fold i+1 needs the running claim of fold i; the fold
        commitments of fold i are issued after the sum-checks of fold i (their witness)."""
        w = self.w
        self.dh = le32(w["doc_hash"])
        self.salt = le32(w["salt"])
        doc_tab = T_tab = hyb_tab = None
        if resident:
            doc_tab, T_tab, hyb_tab = getattr(self, "doc_tab", None), getattr(self, "T_tab", None), getattr(self, "hyb_tab", None)
        else:
            if self.mode == "hybrid":
                hyb_tab = self.ctxs["doc"].table_hybrid(w["T_bytes"], w["fill"], w["half"], self.h_doc)      # H2D of T + document codes
            else:
                f2 = self.pool["nl"].submit(self._upload_T)
                if self.mode == "nldoc":
                    # H2D of the document codes from page-locked memory, queued on the context's copy stream: the first absorb
                    # of the sum-check (which never reads the table) overlaps it, the first sweep waits for it
                    doc_tab = self.pool["doc"].submit(lambda: self.ctxs["doc"].table_u32(self.h_doc, async_upload=os.environ.get("REEF_BENCH_ASYNC_UPLOAD", "1") != "0")).result()
                T_tab = f2.result()
        import threading
        msm_futs, owners = [], []
        self.d_futs = []
        wits = []
        first_doc = le32(int(w["udoc"][0]))
        S = w["steps"]
        nl_done = [threading.Event() for _ in range(S)]

        def commitments(s):
            """the fold commitments of fold s: issued once both sum-checks of the fold are done (their witness)"""
            sc, scd = w["sc"][s], (self.sc_dev[s] if resident else None)
            if self.pairs:
                # the commitments of every fold but the last run beside the next fold's sum-checks (background contexts);
                # the last fold's have nothing to hide behind and use the second pair of contexts
                last = s == S - 1 and os.environ.get("REEF_BENCH_LAST_FOLD_CTX", "1") != "0"
                jobs = [("Tp" if last else "Wp", "pri2" if last else "pri", w["n_pri"], "p"), ("Ts" if last else "Ws", "sec2" if last else "sec", w["n_sec"], "s")]
            else:
                jobs = [("Wp", "pri", w["n_pri"], None), ("Ws", "sec", w["n_sec"], None), ("Tp", "pri2", w["n_pri"], None), ("Ts", "sec2", w["n_sec"], None)]
            for key, pool, n, pair in jobs:
                owner = None if self.world == 1 else len(owners) % self.world
                skip = os.environ.get("REEF_BENCH_SKIP_MSM")          # interference experiments only (never a bench value)
                if skip == "1" or (skip == "last" and s == S - 1) or (skip == "first" and s < S - 1):
                    msm_futs.append(None)
                    owners.append(owner)
                    continue
                if owner is None or owner == self.rank:
                    if pair:
                        msm_futs.append(self.pool[pool].submit(self._msm_pair, self.bases[key], self.pair_dev[s][pair] if resident else None,
                                                               self.pair_host[s][pair], n, resident))
                    else:
                        msm_futs.append(self.pool[pool].submit(self._msm_local, self.bases[key], scd[key] if resident else None, sc[key], n, resident))
                else:
                    msm_futs.append(None)
                owners.append(owner)

        # Two host threads, as in the reference (solver thread / proving thread, framework.rs:98-110): each runs ITS chain of
        # sum-checks fold after fold without handing control back in between (fold i+1 needs only the running claim of
        # fold i of the same table); the document chain issues the commitments of a fold as soon as both of its
        # sum-checks are done.
        def nl_chain():
            prev = None
            for s in range(S):
                qn = self.q_nl[s]
                prev = self._nlookup("nl", T_tab, qn[0], qn[1], prev, None)
                nl_done[s].set()
            return prev

        def doc_chain():
            prev = None
            for s in range(S):
                if self.mode == "hybrid":
                    qh = self.q_hyb[s]
                    prev = self._nlookup("nlhybrid", hyb_tab, qh[0], qh[1], prev, le32(w["T"][0]))
                else:
                    if self.mode == "merkle":
                        wits.append(self._merkle_wits(w["q_doc"][s]))
                    elif self.world > 1:
                        prev = self._nlookup_sharded(doc_tab, w["q_doc"][s], [int(w["udoc"][i]) for i in w["q_doc"][s]], prev)
                    else:
                        qd = self.q_doc[s]
                        prev = self._nlookup("nldoc", doc_tab, qd[0], qd[1], prev, first_doc)
                    nl_done[s].wait()
                if TRACE is not None:
                    TRACE.append(("sumchecks_done", s, time.perf_counter()))
                commitments(s)
            return prev

        if os.environ.get("REEF_BENCH_SKIP_NL") == "1":      # experiment only (never a bench value): the document chain alone
            for e in nl_done:
                e.set()
            f_nl = None
        else:
            f_nl = self.pool["nl"].submit(nl_chain) if self.mode != "hybrid" else None
        prev_doc = self.pool["doc"].submit(doc_chain).result()
        prev_nl = f_nl.result() if f_nl is not None else None
        outs = [f.result() if f is not None else None for f in msm_futs]
        ds = [f.result() for f in self.d_futs]
        if TRACE is not None:
            TRACE.append(("pass_end", 0, time.perf_counter()))
        if self.world > 1:
            outs = self._exchange_points(outs, owners)
        outs = [o[i:i + 64] for o in outs if o is not None for i in range(0, len(o), 64)]       # one 64-byte point per commitment
        if not resident:
            for t in (doc_tab, T_tab, hyb_tab):
                if t is not None:
                    t.free()
        return (prev_nl, prev_doc), outs, ds, wits

    def _upload_T(self):
        h = C.c_void_p()
        self.check(self.lib.reef_table_upload(self.ctxs["nl"]._h, self.h_T.ctypes.data, len(self.w["T"]), C.byref(h)))
        t = self.rb.Table.__new__(self.rb.Table)
        t.ctx, t._h = self.ctxs["nl"], h
        return t

    def prepare_queries(self):
        w = self.w
        if self.mode == "hybrid":
            self.q_hyb = [(np.asarray(q, dtype=np.uint64), pack(hybrid_value(w, i) for i in q)) for q in w["q_hyb"]]
        else:
            self.q_nl = [(np.asarray(q, dtype=np.uint64), pack(w["T"][i] for i in q)) for q in w["q_nl"]]
            self.q_doc = [(np.asarray(q, dtype=np.uint64), pack(int(w["udoc"][i]) for i in q)) for q in w["q_doc"]]

    def bytes_per_step(self):
        w = self.w
        h2d = (w["udoc"].nbytes if self.mode != "merkle" else 0) + len(w["T_bytes"])
        d2h = 0
        ells = {"nldoc": (self.ell_T, self.ell_doc), "hybrid": (self.ell_doc + 1,), "merkle": (self.ell_T,)}[self.mode]
        for s in range(w["steps"]):
            h2d += sum(v.nbytes for v in w["sc"][s].values())
            for ell in ells:
                h2d += (w["m"] + ell + 3) * 32 + ell * 32 + w["m"] * 8
                d2h += 8192
            d2h += 2 * 32 + 4 * 64                                   # calc_d x2, 4 points
        return h2d, d2h

    # -- commit phase (--commit: run_committer, framework.rs:62-79) -----------------------------------
    def commit_setup(self):
        if self.mode == "merkle":
            # host buffers of the commit phase, page-locked like every other timed input / output: the document as the
            # u64 codes reef_merkle_build takes, and room for the whole tree (the .cmt stores it)
            self.h_doc64 = self._pin(self.w["udoc"].astype(np.uint64))
            total = int(self.lib.reef_merkle_tree_elems(len(self.h_doc64)))
            self.h_tree = self._pin(np.zeros(total * 32, dtype=np.uint8))
            return
        rows, cols = WL.hyrax_dims(self.ell_doc)
        rnd = random.Random(99)
        self.hy_blind_ints = [rnd.randrange(FQ) for _ in range(rows)]
        self.hy_blinds = pack(self.hy_blind_ints)
        self.hy_gens = WL.generators("pallas", cols + 1)
        self.hy_bases = self.rb.Bases(self.ctxs["pri"], "pallas", self.hy_gens, 255)

    def commit(self):
        """Hyrax rows (commitment.rs:187, blinds injected) or the Merkle tree (merkle_tree.rs:25-78) of the document,
        from HOST codes.  Returns the commitment bytes (rows x 64 B | root)."""
        w = self.w
        if self.mode == "merkle":
            return le32(self.ctxs["doc"].merkle_raw(self.h_doc64, self.h_tree)[0])       # the whole tree comes back (it is what the .cmt stores)
        rows, cols = WL.hyrax_dims(self.ell_doc)
        out, h = C.create_string_buffer(rows * 64), C.create_string_buffer(32)
        # NLDocCommitment::new (commitment.rs:133-212): Hyrax rows + doc_commit_hash = PoseidonRO over the rows
        self.check(self.lib.reef_doc_commit_u32(self.ctxs["pri"]._h, self.hy_bases._h, self.h_doc.ctypes.data, rows, cols, w["bits"],
                                                self.hy_blinds, out, h))
        return out.raw + h.raw


def sample_clocks_start():
    """nvidia-smi sampler (20 ms period) running across the value and the profiled legs; the samples
    are later restricted to the wall-clock windows of the timed regions."""
    path = tempfile.mktemp(suffix=".csv")
    try:
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=timestamp,index,clocks.sm,clocks.max.sm,power.draw,"
                              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                              "--format=csv,noheader,nounits", "-lms", "20"], stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except Exception:
        return None, path
    return p, path


def sample_clocks_stop(p, path, device=0, windows=()):
    import datetime
    if p is None:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.05)
    p.terminate()
    try:
        p.wait(timeout=5)
    except Exception:
        p.kill()
    rows, mx = [], None
    for line in open(path):
        f = [x.strip() for x in line.split(",")]
        if len(f) < 9 or not f[1].isdigit() or int(f[1]) != device:
            continue
        try:
            ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            sm = float(f[2])
            mx = float(f[3])
        except ValueError:
            continue
        rs = [name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9])
              if val.lower().startswith("active")]
        rows.append((ts, sm, rs))
    try:
        os.unlink(path)
    except OSError:
        pass
    inside = [r for r in rows if any(a - 0.02 <= r[0] <= b + 0.02 for a, b in windows)]
    scope = "timed regions"
    if not inside:                      # timed regions shorter than the sampling period: the whole measured run
        inside, scope = rows, "whole measurement (warm-up + timed steps)"
    reasons = sorted({x for r in inside for x in r[2]})
    return {"sm_mhz": statistics.median([r[1] for r in inside]) if inside else None, "sm_max_mhz": mx, "reasons": reasons,
            "samples": len(inside), "scope": scope}


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            peaks["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"])
            peaks["source"] = "MEASURED_PEAKS.json (burst copy)"
        except Exception:
            pass
    pm = os.path.join(ROOT, "profiles", "peak_modmul.json")
    peaks["modmul_per_s"] = float(json.load(open(pm))["modmul_per_s"]) if os.path.exists(pm) else 69.0e9
    return peaks


def run_reef(args):
    import torch
    import reef_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev = local if world > 1 else 0
    # one context (= one CUDA stream + one host thread) per independent chain of a fold
    # the sum-check contexts are latency-critical (highest stream priority), the commitment contexts are not.  On one GPU
    # the background contexts also get the two opt-in scheduling aids of the library: a green-context
    # partition that leaves 12 SMs to the Fiat-Shamir kernels, and accumulation grids at two CTAs per SM (-1.4 % per pass);
    # REEF_RESERVE_SMS=0 / REEF_MSM_POLITE=0 in the environment turn them off (Nsight Compute's launch list needs the first off)
    if world == 1:
        os.environ.setdefault("REEF_RESERVE_SMS", "12")
        os.environ.setdefault("REEF_MSM_POLITE", "1")
    prio = os.environ.get("REEF_BENCH_PRIO", "1") != "0"
    level = {"nl": 1, "doc": 1, "pri": 0, "sec": 0, "pri2": 2, "sec2": 2}      # 1 highest, 2 middle, 0 lowest (background)
    ctxs = {k: reef_b200.Context(dev, level.get(k, 0) if prio else None) for k in ("nl", "doc", "pri", "sec", "pri2", "sec2", "aux")}
    w = make_workload(args.workload, seed_shift=0, world=world)
    gp = GpuPass(ctxs, w, rank, world, dist)
    gp.prepare_queries()
    gp.make_resident()
    streams = {k: torch.cuda.ExternalStream(c.stream) for k, c in ctxs.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2
    clock_windows = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_fn(fn, steps, warmup):
        """device time of `steps` calls of fn() (each drains every library stream), max over ranks"""
        barrier()
        for _ in range(warmup):
            flush.fill_(1)
            fn()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = {k: torch.cuda.Event(enable_timing=True) for k in streams}
        barrier()
        e0.record(streams["doc"])               # every stream is idle here (barrier above)
        t0, w0 = time.perf_counter(), time.time()
        for _ in range(steps):
            flush.fill_(1)                      # L2 flush between steps; fn() returns only after all streams drained
            torch.cuda.current_stream().synchronize()
            fn()
        for k in streams:
            e1[k].record(streams[k])
        barrier()
        wall = time.perf_counter() - t0
        clock_windows.append((w0, time.time()))
        ms = max(e0.elapsed_time(e1[k]) for k in streams)
        if world > 1:
            tt = torch.tensor([ms], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, wall * 1e3

    def timed(gpx, resident, steps, warmup, profile):
        barrier()
        for _ in range(warmup):                  # warm-up outside the counted / profiled region
            flush.fill_(1)
            gpx.run(resident)
        barrier()
        if profile:
            for c in ctxs.values():
                reef_b200._lib.check(reef_b200.lib.reef_profile_enable(c._h, 1))
        launches0 = int(reef_b200.lib.reef_launch_count())
        ms, wall = timed_fn(lambda: gpx.run(resident), steps, 0)
        launches = int(reef_b200.lib.reef_launch_count()) - launches0
        prof = None
        if profile:
            n = 9
            prof = [(0, 0, 0.0)] * n
            for c in ctxs.values():
                cnt, units, pms = (C.c_uint64 * n)(), (C.c_uint64 * n)(), (C.c_double * n)()
                reef_b200._lib.check(reef_b200.lib.reef_profile_read(c._h, n, cnt, units, pms))
                reef_b200._lib.check(reef_b200.lib.reef_profile_enable(c._h, 0))
                prof = [(p[0] + int(cnt[i]), p[1] + int(units[i]), p[2] + float(pms[i])) for i, p in enumerate(prof)]
        return ms, wall, launches, prof

    def verify(gpx, wx, what):
        """Untimed: the whole pass on this workload compared bit for bit with the CPU restatement of the reference's
        algorithm (oracle/c); every rank against rank 0's CPU run of the SAME document."""
        exp = [None]
        cpu_line = None
        if rank == 0:
            exp[0] = {}
            cpu_line = cpu_baseline(wx, None, exp[0])
        if world > 1:
            dist.broadcast_object_list(exp, src=0)
        msg = compare_outputs(gpx.run_collect(True), exp[0], f"{what}, rank {rank}")      # ParityError is loud
        compare_outputs(gpx.run_collect(False), exp[0], f"{what}, rank {rank}, host-buffer (e2e) path")
        if world > 1:
            msg += f"; every one of the {world} ranks checked its own copy of the results"
        return msg, cpu_line

    K, Wm = args.steps, args.warmup
    verified, cpu_line = (None, None) if args.no_cpu_baseline else verify(gp, w, args.workload)
    clk = sample_clocks_start() if rank == 0 else (None, None)
    ms, wall_ms, launches, _ = timed(gp, True, K, Wm, False)
    _, _, _, prof = timed(gp, True, K, Wm, True)
    clocks = sample_clocks_stop(*clk, device=dev, windows=clock_windows) if rank == 0 else None
    e2e_ms, _, _, _ = timed(gp, False, K, Wm, False)
    commit = commit_phase(gp, w, timed_fn, K, Wm, args.no_cpu_baseline) if world == 1 and not args.no_commit else None

    # ---- the other BASELINE configs (N = 1): same pass + their commit phase, each verified
    also = []
    if world == 1:
        for name in [x for x in args.also.split(",") if x and x != args.workload]:
            w2 = make_workload(name)
            gp2 = GpuPass(ctxs, w2, rank, world, dist)
            gp2.prepare_queries()
            gp2.make_resident()
            ver2, cpu2 = (None, None) if args.no_cpu_baseline else verify(gp2, w2, name)
            ms2, _, launches2, _ = timed(gp2, True, K, Wm, False)
            e2e2, _, _, _ = timed(gp2, False, K, Wm, False)
            h2d2, d2h2 = gp2.bytes_per_step()
            entry = {"workload": name + ": " + w2["desc"], "value": round(w2["doc_len"] / (ms2 / K / 1e3), 1),
                     "ms_per_step": round(ms2 / K, 4), "gpu_launches": launches2, "verified": ver2, "cpu_baseline": cpu2,
                     "e2e": {"value": round(w2["doc_len"] / (e2e2 / K / 1e3), 1), "unit": "NFA steps/s", "ms_per_step": round(e2e2 / K, 4),
                             "h2d_bytes_per_step": h2d2, "d2h_bytes_per_step": d2h2}}
            entry["commit"] = commit_phase(gp2, w2, timed_fn, K, Wm, args.no_cpu_baseline)
            also.append(entry)
            for t in ("doc_tab", "T_tab", "hyb_tab"):
                if getattr(gp2, t, None) is not None:
                    getattr(gp2, t).free()
            for b in gp2.bases.values():
                b.free()
            del gp2
            torch.cuda.empty_cache()

    # ---- multi-GPU: what sharding buys, measured on the same ranks
    multi = multi_gpu_extras(args, ctxs, gp, w, rank, world, dist, timed_fn, K, Wm, ms) if world > 1 else None
    transcript = transcript_object(ctxs["aux"], w, gp, prof, K, clocks) if rank == 0 else None
    openings = opening_phases(ctxs, w, gp, K, args.no_cpu_baseline) if world == 1 and not args.no_openings else None

    if TRACE is not None and rank == 0:
        last = [i for i, t in enumerate(TRACE) if t[0] == "pass_end"]
        seg = TRACE[last[-2] + 1:last[-1] + 1] if len(last) >= 2 else TRACE
        t0 = seg[0][2]
        print("TRACE (last pass, us):", [(a, b, round((t - t0) * 1e6)) for a, b, t in seg], file=sys.stderr)
    value = w["doc_len"] / (ms / K / 1e3)
    e2e_value = w["doc_len"] / (e2e_ms / K / 1e3)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    names = ["sweep_first", "sweep_fold", "round_transcript", "tail", "nl_setup", "msm_sort", "msm_accum", "msm_reduce", "poseidon"]
    total_prof = sum(p[2] for p in prof) or 1.0
    shares = {n: round(p[2] / total_prof, 4) for n, p in zip(names, prof)}
    kernel_ms = {n: round(p[2] / K, 4) for n, p in zip(names, prof)}
    classes = {"transcript": shares["round_transcript"] + shares["tail"] + shares["nl_setup"],
               "msm": shares["msm_sort"] + shares["msm_accum"] + shares["msm_reduce"],
               "sweep": shares["sweep_first"] + shares["sweep_fold"], "poseidon_batch": shares["poseidon"]}
    dominant = max(classes, key=classes.get)
    # MLE sweep: the only HBM-streaming kernel of the pass.  Algorithmic bytes of the reference-shaped fused schedule
    # (2 tables x 32 B): round-1 pass reads 2 L elements (64 L bytes); a fold+accumulate pass over an input
    # of length L reads 2 L and writes L elements (96 L bytes).
    sweep_bytes = 64.0 * prof[0][1] + 96.0 * prof[1][1]
    sweep_ms = prof[0][2] + prof[1][2]
    achieved = sweep_bytes / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else 0.0
    # integer-issue view of the same launches: a u32 first pass is ~0.25 modmul-equivalents per element, a fold pass one
    # modmul (r * diff) + one 256x256 multiply-accumulate (~0.7) per OUTPUT element
    sweep_modmul = 0.25 * prof[0][1] + 1.7 * (prof[1][1] / 2.0)
    roof = {"bound": "hbm", "kernel": "k_sweep (MLE fold+accumulate passes)", "achieved": round(achieved, 1),
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": None,
            "peak_source": peaks["source"], "launches": prof[0][0] + prof[1][0],
            "alg_bytes_per_launch": round(sweep_bytes / max(1, prof[0][0] + prof[1][0])),
            "avg_launch_us": round(1e3 * sweep_ms / max(1, prof[0][0] + prof[1][0]), 2),
            "note": "GB/s is ALGORITHMIC-equivalent (the document streams as 4-byte codes and the eq table is never materialised, "
                    "so real DRAM traffic is ~6x lower: see traffic); the binding resource of this kernel is integer issue slots",
            "int_issue_frac": round(sweep_modmul / (sweep_ms / 1e3) / peaks["modmul_per_s"], 4) if sweep_ms else None,
            "share_of_gpu_time": round(classes["sweep"], 4),
            "time_dominant_class": dominant, "class_shares": {k: round(v, 4) for k, v in classes.items()}}
    if world == 1:
        iso = isolated_sweep(ctxs["doc"], gp, w, peaks["hbm_gbs"])
        if iso:
            roof["isolated"] = iso
    tr = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    if os.path.exists(tr):
        try:
            tinfo = json.load(open(tr)).get(args.workload)
            if tinfo:
                roof["traffic"] = tinfo["dram_bytes_per_launch"]           # ncu --set full, DRAM read+write per launch
                roof["traffic_source"] = tinfo["source"]
        except Exception:
            pass
    # MSM: ops_alg = 10 n W + 14 W 2^c modmuls, against the measured modmul peak
    msm_ms = prof[5][2] + prof[6][2] + prof[7][2]
    ops = 0.0
    n_terms = 0
    for key, n in (("Wp", w["n_pri"]), ("Ws", w["n_sec"])):
        Wn, c = gp.bases[key].windows, gp.bases[key].window_bits
        ops += 2 * w["steps"] * K * (10.0 * n * Wn + 14.0 * Wn * (1 << c)) / max(1, world)
        n_terms += 2 * w["steps"] * K * n / max(1, world)
    msm = {"bound": "int-alu (255-bit modmul)", "mops": round(n_terms / (msm_ms / 1e3) / 1e6, 2) if msm_ms else None,
           "achieved_modmul_per_s": round(ops / (msm_ms / 1e3), 0) if msm_ms else None, "peak_modmul_per_s": peaks["modmul_per_s"],
           "frac": round(ops / (msm_ms / 1e3) / peaks["modmul_per_s"], 4) if msm_ms else None,
           "share_of_gpu_time": round(classes["msm"], 4),
           "note": "fold commitments of 2^14-2^16 terms are latency pipelines (~50 dependent curve operations); `large` is the same "
                   "kernels at a throughput size",
           "peak_source": "tools/bench_fp.cu on this pool (profiles/peak_modmul.json)"}
    if world == 1 and args.msm_large_log2:
        msm["large"] = msm_large(ctxs["doc"], args.msm_large_log2, peaks["modmul_per_s"])     # a context with the whole chip (not a background one)
    h2d, d2h = gp.bytes_per_step()
    out = {
        "metric": "NFA steps/s proved (= doc_len / hot-path prove time)", "value": round(value, 1), "unit": "NFA steps/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": round(ms / K, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 / u29x10 limbs (255-bit prime fields Fq/Fp, exact integer)",
        "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "l2": "256 MiB buffer written between steps (L2 flush)",
                   "timing": "CUDA events: start on an idle stream, end = latest of the library streams; max over ranks",
                   "streams": "contexts/streams: nl sum-check | nldoc sum-check (both highest stream priority) | fold commitments Pallas | Vesta | calc_d: "
                              "commit(W) and commit(T) of a fold share the commitment key and run as the two rows of ONE row-batched MSM per curve "
                              "(reef_msm_rows_dev; at N >= 4 GPUs four separate MSMs round-robin over the ranks); fold i+1 sum-checks overlap fold i "
                              "commitments; calc_d does not gate the next fold",
                   "scheduling": (f"background (commitment) contexts: REEF_RESERVE_SMS={os.environ.get('REEF_RESERVE_SMS', '0')} "
                                  f"(green-context SM partition), REEF_MSM_POLITE={os.environ.get('REEF_MSM_POLITE', '0')} (two CTAs per SM)"),
                   "registration": "commitment keys (static per PublicParams) are registered once per context BEFORE the timed region "
                                   "(k_precompute of all window levels: ~13 ms and 50 MiB per 2^15-point key); contexts that register the "
                                   "same key share one set of levels through the content-keyed cache (reef_bases_cache_stats)",
                   "verified": verified,
                   "parallelism": (f"1 document of {w['doc_len']} chars: nldoc sum-check sharded by low index bits x{world} "
                                   f"(96 bytes per rank per round; the round kernels themselves store them into the peers' mailboxes over NVLink and "
                                   f"acquire the peers' -- no NCCL call, no extra launch); fold commitments (2^14-2^15 terms, latency-bound) distributed "
                                   f"whole, round-robin over the ranks, results exchanged once per pass") if world > 1 else "single GPU"},
        "e2e": {"value": round(e2e_value, 1), "unit": "NFA steps/s", "ms_per_step": round(e2e_ms / K, 4),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "transcript": transcript, "msm": msm,
        "kernel_ms_per_step": kernel_ms, "kernel_share": shares, "wall_ms_per_step": round(wall_ms / K, 4),
    }
    if commit:
        out["commit"] = commit
    if openings:
        out["openings"] = openings
    if also:
        out["also"] = also
    if multi:
        out["multi_gpu"] = multi
    if cpu_line:
        out["cpu_baseline"] = cpu_line
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def isolated_sweep(ctx, gp, w, peak_gbs, reps=5):
    """The same sweeps with nothing else on the GPU: one sum-check over the document table alone, CUDA-event classes of
    libreef_b200 (the in-pass figure moves with what the other streams run beside the sweeps)."""
    import reef_b200
    lib, check = reef_b200.lib, reef_b200._lib.check
    if w["mode"] == "merkle":
        return None
    key, tab, q = ("nlhybrid", gp.hyb_tab, gp.q_hyb[0]) if w["mode"] == "hybrid" else ("nldoc", gp.doc_tab, gp.q_doc[0])
    gp.dh, gp.salt = le32(w["doc_hash"]), le32(w["salt"])
    gp.d_futs = []
    first = le32(w["T"][0]) if w["mode"] == "hybrid" else le32(int(w["udoc"][0]))
    os.environ["REEF_BENCH_SKIP_CALCD"] = "1"
    try:
        for _ in range(2):
            gp._nlookup(key, tab, q[0], q[1], None, first)
        check(lib.reef_profile_enable(ctx._h, 1))
        for _ in range(reps):
            gp._nlookup(key, tab, q[0], q[1], None, first)
        n = 9
        cnt, units, pms = (C.c_uint64 * n)(), (C.c_uint64 * n)(), (C.c_double * n)()
        check(lib.reef_profile_read(ctx._h, n, cnt, units, pms))
        check(lib.reef_profile_enable(ctx._h, 0))
    finally:
        del os.environ["REEF_BENCH_SKIP_CALCD"]
    by = 64.0 * units[0] + 96.0 * units[1]
    ms = pms[0] + pms[1]
    gbs = by / (ms / 1e3) / 1e9 if ms > 0 else 0.0
    return {"achieved": round(gbs, 1), "frac": round(gbs / peak_gbs, 4), "avg_launch_us": round(1e3 * ms / max(1, cnt[0] + cnt[1]), 2),
            "what": "the sweeps of one sum-check over the same table with no other stream running"}


def _ells(w, gp):
    return {"nldoc": (gp.ell_T, gp.ell_doc), "hybrid": (gp.ell_doc + 1,), "merkle": (gp.ell_T,)}[w["mode"]]


def _perms(m, ell):
    """permutations of one nlookup transcript: first absorb in groups of 4 + one per round"""
    return (m + ell + 2 + (m * ell + 253) // 254 + 3) // 4 + ell


def transcript_object(ctx, w, gp, prof, K, clocks):
    """The Fiat-Shamir chain (the time-dominant class of the pass): permutations on the critical path, their measured
    latency, and the dependency floor of one permutation."""
    import reef_b200
    cyc = (C.c_uint64 * 9)()
    inp = pack([1, 2, 3, 4, 5])
    out = C.create_string_buffer(160)
    for _ in range(2):
        reef_b200._lib.check(reef_b200.lib.reef_gputest_poseidon_permute_lp(ctx._h, inp, 16, out, cyc))
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    per = int(cyc[0])
    ells = _ells(w, gp)
    perms_crit = w["steps"] * _perms(w["m"], max(ells))           # the chain of the longest table gates the pass
    perms_all = w["steps"] * sum(_perms(w["m"], e) for e in ells)
    rounds_all = w["steps"] * sum(ells)
    mul_cycles = 345.0                                # dependent lane-parallel multiplication (tools/bench_lp.cu)
    floor_cycles = (56 * 3 + 8 * 4) * mul_cycles      # 56 partial rounds x (w^2, w^4, w^5) + 8 full rounds x (x^5 + MDS row)
    tr_ms = (prof[2][2] + prof[3][2] + prof[4][2]) / K
    return {"bound": "latency: one dependent chain of 255-bit multiplications (Poseidon t=5, 8 full + 56 partial rounds)",
            "permutations_on_critical_path": perms_crit, "cycles_per_permutation": per, "us_per_permutation": round(per / mhz, 2),
            "phases_cycles": {"first_full_rounds": int(cyc[1]), "partial_rounds": int(cyc[2]), "end": int(cyc[3]), "last_full_rounds": int(cyc[4])},
            "dependency_floor_cycles": int(floor_cycles), "frac_of_dependency_floor": round(floor_cycles / per, 3),
            "floor_note": "200 dependent multiplications x 345 cycles (the measured latency of ONE lane-parallel multiplication: 4 "
                          "shared-memory/shuffle hops + 22 IMAD.WIDE deep); round 1: 172 k cycles / 88 us per permutation",
            "critical_path_ms_per_pass": round(perms_crit * per / mhz / 1e3, 3),
            "transcript_kernels_ms_per_pass": round(tr_ms, 3),
            "non_permutation_us_per_round": round(max(0.0, tr_ms * 1e3 - perms_all * per / mhz) / max(1, rounds_all), 2)}


class ShaTranscript:
    """Stand-in Fiat-Shamir transcript of the opening proofs (nova-snark's own lives in the un-vendored crate and stays
    with the Rust caller): SHA-256 chaining, the same absorb / squeeze call pattern on both arms."""

    def __init__(self, label: bytes, modulus: int):
        import hashlib
        self.h = hashlib
        self.state = hashlib.sha256(b"reef-b200-bench-transcript" + label).digest()
        self.p, self.round = modulus, 0

    def absorb(self, label: bytes, data: bytes):
        self.state = self.h.sha256(self.state + len(label).to_bytes(4, "little") + label + len(data).to_bytes(8, "little") + data).digest()

    def absorb_scalars(self, label: bytes, xs):
        self.absorb(label, b"".join(int(x).to_bytes(32, "little") for x in xs))

    def absorb_point(self, label: bytes, P):
        self.absorb(label, bytes(64) if P is None else int(P[0]).to_bytes(32, "little") + int(P[1]).to_bytes(32, "little"))

    def squeeze(self, label: bytes) -> int:
        self.round += 1
        d0 = self.h.sha256(self.state + b"\x00" + label + self.round.to_bytes(4, "little")).digest()
        d1 = self.h.sha256(self.state + b"\x01" + label + self.round.to_bytes(4, "little")).digest()
        self.state = d0
        return int.from_bytes(d0 + d1, "little") % self.p


def opening_phases(ctxs, w, gp, K, no_cpu):
    """The opening proofs behind `prove_consistency` / `CompressedSNARK::prove` (commitment.rs:214-285, 371-393;
    framework.rs:695-698) that follow the folds: Hyrax prove_eval on the resident document table and one inner-product
    argument of the fold-commitment size, through the library's device-resident IPA session (reef_ipa_*: two MSMs and
    the a / b / generator folds per round on the device, transcript on the host).  Wall-clock per proof; proofs compared
    with the oracle's (same transcript).  nova-snark's conventions are parity-unpinned."""
    import reef_b200
    from reef_b200 import snark as G
    out = {}
    ctx = ctxs["pri"]
    rnd = random.Random(4242)

    def pts(raw, n):
        return [(int.from_bytes(raw[i * 64:i * 64 + 32], "little"), int.from_bytes(raw[i * 64 + 32:i * 64 + 64], "little")) for i in range(n)]

    def timed_wall(fn):
        fn()
        ts = []
        for _ in range(K):
            t0 = time.perf_counter()
            res = fn()
            ts.append(time.perf_counter() - t0)
        return res, statistics.median(ts) * 1e3

    # (1) one IPA of the primary fold-commitment size (the openings of W and E inside the compressed SNARK)
    n = w["n_pri"]
    raw = WL.generators("pallas", n + 1)
    gens_raw, gen_c = raw[:64 * n], pts(raw[64 * n:], 1)[0]
    a = [rnd.randrange(FQ) for _ in range(n)]
    b = [rnd.randrange(FQ) for _ in range(n)]
    kb = reef_b200.Bases(ctx, "pallas", gens_raw, 255)          # the commitment key is static: registered once (cached levels)
    a_raw, b_raw = pack(a), pack(b)                              # byte buffers, like every other timed input
    got, ms = timed_wall(lambda: G.ipa_prove(ctx, "pallas", kb, gen_c, a_raw, b_raw, ShaTranscript(b"ipa", FQ)))
    got_f, ms_f = timed_wall(lambda: G.ipa_prove(ctx, "pallas", gens_raw, gen_c, a_raw, b_raw, ShaTranscript(b"ipa", FQ)))
    kb.free()
    if got_f != got:
        raise ParityError("IPA over registered generators differs from the folding session")
    entry = {"n": n, "rounds": n.bit_length() - 1, "ms": round(ms, 3), "ms_folding_generators": round(ms_f, 3),
             "what": "reef_ipa_begin_bases/round/fold/finish over the registered commitment key: per round ONE two-row MSM (L, R) over the "
                     "precomputed window levels, the folds of a, b and of the per-generator weights on the device; `ms_folding_generators` = "
                     "the session that folds the generators every round (reef_ipa_begin), same proof"}
    if not no_cpu:
        from oracle import cport, snark as N
        from oracle.curves import PALLAS
        cm = lambda sc, p_: cport.msm("pallas", p_, sc, threads=cport.max_threads())
        t0 = time.perf_counter()
        exp = N.ipa_prove(PALLAS, pts(gens_raw, n), gen_c, a, b, ShaTranscript(b"ipa", FQ), cm)
        entry["cpu_baseline"] = {"seconds": round(time.perf_counter() - t0, 3), "cores": cport.max_threads(), "kind": "port",
                                 "sample": "oracle/snark.py: MSMs by oracle/c on all threads, scalar folds in Python (not a tuned baseline)"}
        if got != exp:
            raise ParityError("inner-product argument differs from the oracle")
        entry["verified"] = "proof (L, R vectors, a_hat) == oracle under the same transcript"
    out["ipa"] = entry
    # (2) Hyrax prove_eval of the committed document at a random point (prove_consistency, commitment.rs:371-393)
    if w["mode"] != "merkle":
        ell = gp.ell_doc
        rows, cols = WL.hyrax_dims(ell)
        graw = WL.generators("pallas", cols + 1)
        hg, hc = graw[:64 * cols], pts(graw[64 * cols:], 1)[0]
        q = [rnd.randrange(FQ) for _ in range(ell)]
        tab = ctxs["pri"].table_u32(w["udoc"])
        hb = reef_b200.Bases(ctx, "pallas", hg, 255)
        (v, proof), ms = timed_wall(lambda: G.hyrax_prove_eval(ctx, tab, rows, cols, hb, hc, q, ShaTranscript(b"hyrax", FQ)))
        hb.free()
        entry = {"shape": f"{rows} x {cols}", "ms": round(ms, 3),
                 "what": "LZ = L^T M over the resident u32 document table (reef_hyrax_lz) + one IPA of length 2^right"}
        if not no_cpu:
            t0 = time.perf_counter()
            m2 = w["udoc"].reshape(rows, cols).astype(object)
            kl = rows.bit_length() - 1
            Lv = N.eq_table(q[:kl], FQ)
            LZ = [int(x) % FQ for x in (np.asarray(Lv, dtype=object) @ m2)]
            Rv = N.eq_table(q[kl:], FQ)
            ev = N.inner(LZ, Rv, FQ)
            tr = ShaTranscript(b"hyrax", FQ)
            tr.absorb_scalars(b"v", [ev])
            exp = N.ipa_prove(PALLAS, pts(hg, cols), hc, LZ, Rv, tr, cm)
            entry["cpu_baseline"] = {"seconds": round(time.perf_counter() - t0, 3), "cores": cport.max_threads(), "kind": "port",
                                     "sample": "oracle/snark.py (numpy object mat-vec + oracle/c MSMs; not a tuned baseline)"}
            if v != ev or proof != exp:
                raise ParityError("Hyrax prove_eval differs from the oracle")
            entry["verified"] = "evaluation and IPA proof == oracle under the same transcript"
        tab.free()
        out["hyrax_prove_eval"] = entry
    return out


def commit_phase(gp, w, timed_fn, K, Wm, no_cpu):
    """--commit of the same document (run_committer, framework.rs:62-79) from HOST codes: e2e time, verified."""
    gp.commit_setup()
    got = gp.commit()
    ver, cpu = None, None
    if not no_cpu:
        from oracle import cport
        t0 = time.perf_counter()
        if w["mode"] == "merkle":
            cport.lib().oracle_set_fast_poseidon(1)
            try:
                exp = le32(cport.merkle(w["udoc"], threads=cport.max_threads())[-1][0])
            finally:
                cport.lib().oracle_set_fast_poseidon(0)
            cpu_s = time.perf_counter() - t0
            if got != exp:
                raise ParityError("Merkle commitment differs from the oracle")
            ver = "root == oracle/c over all 2^%d leaves" % gp.ell_doc
        else:
            rows, cols = WL.hyrax_dims(gp.ell_doc)
            rnd = random.Random(7)
            sample = sorted({0, rows - 1} | {rnd.randrange(rows) for _ in range(6)})
            for r in sample:
                sc = [int(x) for x in w["udoc"][r * cols:(r + 1) * cols]] + [gp.hy_blind_ints[r]]
                P = cport.msm("pallas", gp.hy_gens, sc, threads=cport.max_threads())
                if got[r * 64:(r + 1) * 64] != (bytes(64) if P is None else le32(P[0]) + le32(P[1])):
                    raise ParityError(f"Hyrax row commitment {r} differs from the oracle")
            cpu_s = (time.perf_counter() - t0) * rows / len(sample)
            # doc_commit_hash: the oracle's PoseidonRO over ALL row commitments (x, y, is_infinity) the GPU produced
            t1 = time.perf_counter()
            elems = b"".join(got[r * 64:(r + 1) * 64] + le32(1 if got[r * 64:(r + 1) * 64] == bytes(64) else 0) for r in range(rows))
            exp_h = cport.poseidon_ro(elems, "fp", FQ, 256)
            cpu_s += time.perf_counter() - t1
            if int.from_bytes(got[rows * 64:rows * 64 + 32], "little") != exp_h:
                raise ParityError("doc_commit_hash (PoseidonRO over the row commitments) differs from the oracle")
            ver = (f"{len(sample)} of {rows} row commitments (blinds included) == oracle/c; doc_commit_hash == oracle/c PoseidonRO over "
                   f"all {rows} rows (parity-unpinned construction, see DESIGN.md)")
        cpu = {"seconds": round(cpu_s, 3), "cores": cport.max_threads(), "kind": "port",
               "sample": "whole tree" if w["mode"] == "merkle" else "sampled rows scaled to all rows + the whole PoseidonRO (textbook rounds)"}
    ms, _ = timed_fn(gp.commit, K, Wm)
    kind = ("merkle_tree.rs:25-78 Poseidon tree" if w["mode"] == "merkle" else
            "commitment.rs:133-212 Hyrax rows %d x %d + doc_commit_hash (PoseidonRO)" % WL.hyrax_dims(gp.ell_doc))
    cmt = cmt_bytes(gp, w, got)
    if w["mode"] != "merkle":
        gp.hy_bases.free()
    return {"what": kind, "e2e_ms": round(ms / K, 3), "chars_per_s": round(w["doc_len"] / (ms / K / 1e3), 1), "verified": ver, "cpu_baseline": cpu,
            "cmt": cmt, "not_included": "CAP key setup (SpartanSNARK::setup, nova-snark) and the OsRng draws: see DESIGN.md"}


def cmt_bytes(gp, w, got):
    """(f3) the .cmt payload of this commitment through the library's bincode writer (host-only, untimed for `e2e_ms`)."""
    import reef_b200
    lib, check = reef_b200.lib, reef_b200._lib.check
    t0 = time.perf_counter()
    if w["mode"] == "merkle":
        ctx = gp.ctxs["doc"]
        root, levels = ctx.merkle_raw(gp.h_doc64, gp.h_tree)
        levels = levels.ctypes.data
        sizes, nl = ctx.last_level_sizes, ctx.last_n_levels
        t0 = time.perf_counter()
        d = np.ascontiguousarray(w["udoc"].astype(np.uint64))
        size = int(lib.reef_cmt_merkle_size(sizes.ctypes.data, nl, len(d)))
        out, n = np.empty(size, dtype=np.uint8), C.c_uint64()
        check(lib.reef_cmt_merkle_write(le32(root), levels, sizes.ctypes.data, nl, d.ctypes.data, len(d), w["doc_len"], out.ctypes.data, size, C.byref(n)))
        return {"kind": "ReefCommitment{merkle}", "bytes": int(n.value), "write_ms": round((time.perf_counter() - t0) * 1e3, 2)}
    from reef_b200._lib import CmtNldoc
    rows, cols = WL.hyrax_dims(gp.ell_doc)
    f = CmtNldoc()
    keep = [C.create_string_buffer(got[:rows * 64], rows * 64), C.create_string_buffer(gp.hy_blinds, rows * 32),
            C.create_string_buffer(got[rows * 64:rows * 64 + 32], 32), C.create_string_buffer(le32(w["salt"]), 32)]
    f.num_vars, f.doc_codes, f.doc_len = gp.ell_doc, w["udoc"].ctypes.data, len(w["udoc"])
    f.row_commitments, f.blinds, f.rows = C.addressof(keep[0]), C.addressof(keep[1]), rows
    f.doc_commit_hash, f.hash_salt = C.addressof(keep[2]), C.addressof(keep[3])
    f.q_len, f.orig_doc_len, f.udoc_len = gp.ell_doc, w["doc_len"], len(w["udoc"])
    size = int(lib.reef_cmt_nldoc_size(C.byref(f)))
    out, n = np.empty(size, dtype=np.uint8), C.c_uint64()
    check(lib.reef_cmt_nldoc_write(C.byref(f), out.ctypes.data, size, C.byref(n)))
    return {"kind": "ReefCommitment{nldoc}", "bytes": int(n.value), "write_ms": round((time.perf_counter() - t0) * 1e3, 2),
            "note": "nova-snark-owned members (generators, CAP keys) are opaque byte strings supplied by the Rust side: empty here"}


def multi_gpu_extras(args, ctxs, gp, w, rank, world, dist, timed_fn, K, Wm, ms_sharded):
    """Measured on the same ranks after the main legs (multi-GPU runs only)."""
    import torch
    import reef_b200
    out = {}
    peaks = load_peaks()
    lib, check = reef_b200.lib, reef_b200._lib.check
    # (a) the SAME world-scaled document on ONE GPU (rank 0 alone): what the sharding of the pass buys
    if rank == 0:
        gp1 = GpuPass(ctxs, w, 0, 1, None)
        gp1.q_nl, gp1.q_doc = gp.q_nl, gp.q_doc
        gp1.sc_dev = gp.sc_dev
        gp1.pair_dev = gp.pair_dev
        gp1.T_tab = gp.T_tab
        gp1.doc_tab = ctxs["doc"].table_u32(w["udoc"])
        streams = [torch.cuda.ExternalStream(c.stream) for c in ctxs.values()]
        torch.cuda.synchronize()
        for _ in range(Wm):
            gp1.run(True)
        e0 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(streams[0])
        for _ in range(K):
            gp1.run(True)
        ends = []
        for s in streams:
            e = torch.cuda.Event(enable_timing=True)
            e.record(s)
            ends.append(e)
        torch.cuda.synchronize()
        ms1 = max(e0.elapsed_time(e) for e in ends) / K
        gp1.doc_tab.free()
        for b in gp1.bases.values():
            b.free()
        out["same_doc_1gpu"] = {"ms_per_step": round(ms1, 4), "sharded_ms_per_step": round(ms_sharded / K, 4),
                                "speedup_vs_1gpu_same_doc": round(ms1 / (ms_sharded / K), 3),
                                "what": f"the same {w['doc_len']}-char document, un-sharded pass on rank 0 alone (other ranks idle)"}
    dist.barrier()
    # (b) throughput-sized MSMs sharded by Pippenger windows, 128-byte all-gather by the library's mailbox kernel
    def sharded_msm(lg, reps):
        n = 1 << lg
        distinct = min(n, 1 << 20)
        pts = WL.generators("pallas", distinct) * (n // distinct)    # beyond 2^20 terms the points repeat (a valid MSM all the same)
        bases = reef_b200.Bases(ctxs["pri"], "pallas", pts)
        del pts
        raw = np.random.default_rng(lg).integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
        raw[:, 3] &= (1 << 61) - 1
        devs = torch.from_numpy(raw.view(np.int64)).cuda()
        del raw
        res, whole = C.create_string_buffer(64), C.create_string_buffer(64)
        check(lib.reef_msm_dev(ctxs["pri"]._h, bases._h, C.c_void_p(devs.data_ptr()), n, whole))          # un-sharded reference + scratch sizing
        dist.barrier()
        ms_m, _ = timed_fn(lambda: check(lib.reef_msm_sharded_dev(ctxs["pri"]._h, bases._h, C.c_void_p(devs.data_ptr()), n, res)), reps, min(Wm, reps))
        if res.raw != whole.raw:
            raise ParityError("window-sharded MSM differs from the un-sharded MSM")
        ms_1, _ = timed_fn(lambda: check(lib.reef_msm_dev(ctxs["pri"]._h, bases._h, C.c_void_p(devs.data_ptr()), n, whole)), reps, min(Wm, reps))
        Wn, c = bases.windows, bases.window_bits
        ops = 10.0 * n * Wn + 14.0 * Wn * (1 << c)
        bases.free()
        del devs
        torch.cuda.empty_cache()
        return {"n": n, "distinct_points": distinct, "windows": Wn, "window_bits": c, "ranks": world, "ms": round(ms_m / reps, 3),
                "mops": round(n / (ms_m / reps) / 1e3, 1), "achieved_modmul_per_s": round(ops / (ms_m / reps / 1e3), 0),
                "frac_of_world_peak": round(ops / (ms_m / reps / 1e3) / (world * peaks["modmul_per_s"]), 4),
                "frac_of_one_gpu_peak": round(ops / (ms_m / reps / 1e3) / peaks["modmul_per_s"], 4),
                "one_gpu_ms": round(ms_1 / reps, 3), "speedup_vs_1gpu": round(ms_1 / ms_m, 3),
                "windows_per_rank_max": -(-Wn // world),
                "verified": "sharded result == un-sharded result on every rank (bit for bit)",
                "exchange": "128-byte XYZZ partial per rank, all-gather by a kernel of the library over NVLink peer mailboxes, combine on every rank"}

    out["msm_sharded"] = sharded_msm(args.msm_large_log2 or 20, K)
    if args.msm_shard_large_log2 and world >= 4:
        # the size at which the replicated bucket reduction is amortised (BASELINE north_star: MSM roofline at 8 GPUs)
        try:
            out["msm_sharded_large"] = sharded_msm(args.msm_shard_large_log2, 2)
        except ParityError:
            raise
        except Exception as e:                                   # memory / time limits of the box: the line survives
            out["msm_sharded_large"] = {"skipped": f"{type(e).__name__}: {e}"[:300]}

    # (c) commit phase split over the ranks: Hyrax rows and Merkle subtrees (one NCCL all-gather each)
    def gather(b):
        mine = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
        outs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(outs, mine)
        return [bytes(o.cpu().numpy().tobytes()) for o in outs]

    base_doc = np.ascontiguousarray(WL.encode(*WL.document("cfg4")))                  # configs[3] document: Hyrax 1024 x 2048
    ell = reef_b200.logmn(len(base_doc))
    rows, cols = WL.hyrax_dims(ell)
    hb = reef_b200.Bases(ctxs["pri"], "pallas", WL.generators("pallas", cols + 1), 255)
    rnd = random.Random(99)
    blinds = [rnd.randrange(FQ) for _ in range(rows)]
    m2 = base_doc.reshape(rows, cols)
    full = hb.msm_rows(m2, rows, cols, entry_bits=8, blinds=blinds)
    got = hb.commit_rows_sharded(m2, rows, cols, 8, blinds, rank, world, gather)
    if got != full:
        raise ParityError("row-split Hyrax commitment differs from the single-GPU one")
    ms_h, _ = timed_fn(lambda: hb.commit_rows_sharded(m2, rows, cols, 8, blinds, rank, world, gather), K, Wm)
    ms_h1, _ = timed_fn(lambda: hb.msm_rows(m2, rows, cols, entry_bits=8, blinds=blinds), K, Wm)
    hb.free()
    per = len(base_doc) // world
    root1 = [ctxs["doc"].merkle_root(base_doc) if rank == 0 else None]
    dist.broadcast_object_list(root1, src=0)
    loc = np.ascontiguousarray(base_doc[rank * per:(rank + 1) * per])
    rootg, _ = reef_b200.MerkleCommitment.build_sharded(ctxs["doc"], loc, len(base_doc), rank, world, gather)
    if rootg != root1[0]:
        raise ParityError("subtree-split Merkle root differs from the single-GPU one")
    ms_t, _ = timed_fn(lambda: reef_b200.MerkleCommitment.build_sharded(ctxs["doc"], loc, len(base_doc), rank, world, gather), K, Wm)
    ms_t1, _ = timed_fn(lambda: ctxs["doc"].merkle_root(base_doc), K, Wm)
    out["commit"] = {"hyrax_rows_split": {"shape": f"{rows} x {cols}", "ms": round(ms_h / K, 3), "one_gpu_ms": round(ms_h1 / K, 3),
                                          "speedup_vs_1gpu": round(ms_h1 / ms_h, 3), "verified": "== single-GPU commitment (all rows)"},
                     "merkle_subtrees_split": {"leaves": len(base_doc), "ms": round(ms_t / K, 3), "one_gpu_ms": round(ms_t1 / K, 3),
                                               "speedup_vs_1gpu": round(ms_t1 / ms_t, 3), "verified": "root == single-GPU root",
                                               "note": "root only (the G roots are gathered, the level slices stay on their ranks)"}}
    return out


def msm_large(ctx, lg, peak):
    """MSM Mop/s vs the modmul roofline at n = 2^lg (BASELINE.json metric, second half)."""
    import torch
    import reef_b200
    n = 1 << lg
    bases = reef_b200.Bases(ctx, "pallas", WL.generators("pallas", n))
    raw = np.random.default_rng(lg).integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= (1 << 61) - 1
    dev = torch.from_numpy(raw.view(np.int64)).cuda()
    out = C.create_string_buffer(64)
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3):
        reef_b200._lib.check(reef_b200.lib.reef_msm_dev(ctx._h, bases._h, C.c_void_p(dev.data_ptr()), n, out))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        reef_b200._lib.check(reef_b200.lib.reef_msm_dev(ctx._h, bases._h, C.c_void_p(dev.data_ptr()), n, out))
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    Wn, c = bases.windows, bases.window_bits
    ops = 10.0 * n * Wn + 14.0 * Wn * (1 << c)
    res = {"n": n, "curve": "pallas", "window_bits": c, "windows": Wn, "ms": round(ms, 3), "mops": round(n / ms / 1e3, 1),
           "achieved_modmul_per_s": round(ops / (ms / 1e3), 0), "frac": round(ops / (ms / 1e3) / peak, 4),
           "scalars": "uniform 255-bit, resident; bases k*G resident with all window levels precomputed"}
    bases.free()
    return res


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's C restatement of the reference's algorithm (never used by the product)
# --------------------------------------------------------------------------------------------
_cpu_tree_cache = {}


def cpu_pass(w, n_steps, threads, collect=None):
    """Runs `n_steps` Nova folds of the workload on the CPU port; returns seconds.  `collect` (a dict)
    receives every output in GpuPass.run_collect()'s layout."""
    from oracle import cport
    from oracle.nlookup import combined_qs, logmn, nlookup_pattern
    cport.lib().oracle_set_fast_poseidon(1)     # neptune hashes with its optimised constants too
    t_total = 0.0
    mode = w["mode"]
    T_arr = np.frombuffer(w["T_bytes"], dtype=np.uint8)
    tabs = []
    if mode == "hybrid":
        half = w["half"]
        hyb = np.zeros((2 * half, 4), dtype=np.uint64)
        limbs = lambda x: [(x >> (64 * k)) & (2 ** 64 - 1) for k in range(4)]
        for i, x in enumerate(w["T"]):
            hyb[i] = limbs(x)
        hyb[len(w["T"]):half] = limbs(w["fill"])
        nd = len(w["udoc"])
        for rep in range(half // nd):
            hyb[half + rep * nd:half + (rep + 1) * nd, 0] = w["udoc"]
        tabs.append(("nlhybrid", np.frombuffer(hyb.tobytes(), dtype=np.uint8), 0, 2 * half, "q_hyb", w["T"][0]))
    else:
        tabs.append(("nl", T_arr, 0, len(w["T"]), "q_nl", w["T"][0]))
        if mode == "nldoc":
            tabs.append(("nldoc", w["udoc"], 1, len(w["udoc"]), "q_doc", int(w["udoc"][0])))
    tree = None
    if mode == "merkle":
        # the tree is the product of --commit: built once, outside the timed pass (as on the GPU arm)
        key = (w["name"], len(w["udoc"]))
        if key not in _cpu_tree_cache:
            from oracle.merkle import MerkleCommitment
            mc = object.__new__(MerkleCommitment)
            mc.doc = [int(x) for x in w["udoc"]]
            mc.tree = cport.merkle(w["udoc"], threads=threads)
            _cpu_tree_cache[key] = mc
        tree = _cpu_tree_cache[key]
    prev = {}
    for s in range(n_steps):
        t0 = time.perf_counter()
        ds = []
        for key, arr, is_u32, n, qk, first in tabs:
            q = w[qk][s]
            vals = [hybrid_value(w, i) for i in q] if key == "nlhybrid" else ([w["T"][i] for i in q] if key == "nl" else [int(w["udoc"][i]) for i in q])
            ell = logmn(n)
            pq, pv = prev[key] if key in prev else ([0] * ell, first)
            cqs = combined_qs(list(q), ell)
            pat = nlookup_pattern(key, len(q), ell, len(cqs))
            query = ([] if key == "nl" else [w["doc_hash"]]) + cqs + vals + list(pq) + [pv]
            claim, rounds, last, nxt = cport.nlookup_raw(arr, is_u32, n, q, pack(query), len(query), cport.ops_words(pat), pack(pq), ell)
            r = [int.from_bytes(rounds[i * 128:i * 128 + 32], "little") for i in range(ell)]
            prev[key] = (r, int.from_bytes(nxt, "little"))
            if key != "nl":
                ds.append(le32(cport.poseidon_hash([pv, w["salt"]], 2)[0]))
                ds.append(le32(cport.poseidon_hash([prev[key][1], w["salt"]], 2)[0]))
            if collect is not None:
                collect.setdefault("sumchecks", []).append((key, claim, rounds, last, nxt))
        wit = None
        if mode == "merkle":
            wit = b"".join(wit_bytes(tree.path_wits(int(q))) for q in w["q_doc"][s])
        sc = w["sc"][s]
        pts = []
        for key, curve, bases in (("Wp", "pallas", w["bases_pri"]), ("Ws", "vesta", w["bases_sec"]),
                                  ("Tp", "pallas", w["bases_pri"]), ("Ts", "vesta", w["bases_sec"])):
            pts.append(cport.msm(curve, bases, sc[key].tobytes(), threads=threads))
        t_total += time.perf_counter() - t0
        if collect is not None:
            collect.setdefault("d", []).extend(ds)
            collect.setdefault("wits", [])
            if wit is not None:
                collect["wits"].append(wit)
            collect.setdefault("msm", []).extend(bytes(64) if P is None else le32(P[0]) + le32(P[1]) for P in pts)
    return t_total


def cpu_baseline(w, threads, collect=None):
    """One WHOLE pass (every fold) of the same workload on the CPU port, timed; its outputs (collect)
    are what the GPU pass is verified against."""
    from oracle import cport
    threads = threads or cport.max_threads()
    sec = cpu_pass(w, w["steps"], threads, collect)
    return {"value": round(w["doc_len"] / sec, 2), "unit": "NFA steps/s", "cores": threads, "kind": "port",
            "seconds_per_pass": round(sec, 3),
            "sample": f"one whole pass = all {w['steps']} Nova folds of the same workload (sum-checks single-threaded as in "
                      f"the reference, MSMs on {threads} threads); no scaling"}


def compare_outputs(got, exp, what):
    """Bit-for-bit comparison of a GPU pass with the CPU restatement; raises ParityError."""
    gs, es = sorted(got["sumchecks"], key=lambda r: r[0]), sorted(exp.get("sumchecks", []), key=lambda r: r[0])
    if len(gs) != len(es):
        raise ParityError(f"{what}: {len(gs)} sum-checks vs {len(es)}")
    for g, e in zip(gs, es):                       # within a tag the folds are in order on both sides (sorted() is stable)
        for name, a, b in zip(("tag", "claim_r", "rounds", "sc_last_claim", "next_running_claim"), g, e):
            if (a if isinstance(a, str) else bytes(a)) != (b if isinstance(b, str) else bytes(b)):
                raise ParityError(f"{what}: {g[0]} sum-check: {name} differs from the oracle")
    for key in ("d", "msm", "wits"):
        if sorted(bytes(x) for x in got[key]) != sorted(bytes(x) for x in exp.get(key, [])):
            raise ParityError(f"{what}: {key} differs from the oracle")
    n_r = sum(len(g[2]) // 128 for g in gs)
    return (f"GPU pass == oracle/c on the same workload, bit for bit: {len(gs)} sum-checks ({', '.join(sorted({g[0] for g in gs}))}: claim_r, "
            f"{n_r} round polynomials and challenges in total, last claims, next running claims), {len(got['d'])} calc_d digests, "
            f"{len(got['msm'])} commitments, {len(got['wits'])} Merkle witness sets (untimed, before the timed legs; resident and host-buffer paths)")


def run_reference(args):
    """The reference's CPU implementation of the path on this box's host cores: the Rust reference cannot be
    built in this image (no cargo, un-vendored crates), so this is the oracle's C restatement of the same
    algorithm in the reference's shape.  Inputs come from workloads.py (no import of reef_b200).  A step is
    one WHOLE pass of the same (world-scaled) workload; only when K whole passes would not end within a few
    minutes (multi-GPU documents) a step becomes one Nova fold scaled to the pass, and the line says so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cport
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = make_workload(args.workload, world=world)
    threads = cport.max_threads()
    t_first = cpu_pass(w, w["steps"], threads)          # warm-up pass (page-in, thread pool), also sizes the run
    whole = t_first * args.steps <= 150.0
    t = 0.0
    for _ in range(args.steps):
        t += cpu_pass(w, w["steps"], threads) if whole else cpu_pass(w, 1, threads) * w["steps"]
    per = t / args.steps
    val = round(w["doc_len"] / per, 2)
    sample = (f"every step = one whole pass (all {w['steps']} Nova folds), no scaling" if whole else
              f"every step = 1 of {w['steps']} Nova folds scaled to the pass (a whole pass takes {t_first:.1f} s on this document)")
    out = {"impl": "reference", "metric": "NFA steps/s proved (= doc_len / hot-path prove time)", "value": val,
           "unit": "NFA steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_passes_run": 1,
           "ms_per_step": round(per * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "u64x4 limbs (255-bit prime fields, exact integer)", "data": "synthetic",
           "config": {"workload": args.workload + ": " + w["desc"]},
           "cpu_baseline": {"value": val, "unit": "NFA steps/s", "cores": threads, "kind": "port",
                            "sample": sample + "; CPU restatement of the reference's algorithm (oracle/c) -- the Rust "
                                               "reference cannot be built in this image"},
           "e2e": {"value": val, "unit": "NFA steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="reef", choices=["reef", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--also", default="cfg2,cfg3,cfg4,cfg5",
                    help="other BASELINE configs timed at N = 1 (value/e2e/commit, each verified) and reported under 'also'; '' = none")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg AND the verification against it")
    ap.add_argument("--msm-shard-large-log2", type=int, default=24,
                    help="at >= 4 GPUs: a second window-sharded MSM of this size (0 = skip)")
    ap.add_argument("--no-openings", action="store_true", help="skip the opening-proof phases (IPA, Hyrax prove_eval)")
    ap.add_argument("--no-commit", action="store_true", help="skip the commit phase of the main workload (profiling runs)")
    ap.add_argument("--msm-large-log2", type=int, default=20, help="size of the stand-alone MSM roofline measurement (0 = skip)")
    ap.add_argument("--debug", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "reef":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_reef(args)


if __name__ == "__main__":
    main()
