#!/usr/bin/env python3
"""bench.py -- Reef prover hot path on B200 (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W [--workload cfg2] [--impl reef|reference]

A "step" is ONE PASS OF THE PROVER HOT PATH over one synthetic document: for each Nova fold of
the `--prove` run the reference would do (framework.rs:405-625 / 642-754):
    nlookup sum-check over the transition table T      ("nl",    r1cs.rs:2088-2100)
    nlookup sum-check over the committed document      ("nldoc", r1cs.rs:2137-2161)
    2 x calc_d                                         (framework.rs:517-553)
    prove_step commitments: commit(W), commit(T) on Pallas and on Vesta (framework.rs:668-675)
metric  = NFA steps/s proved = doc_len / time of one pass        (BASELINE.json)
workload = `target` by default: configs[1] at the 2^20-char document the north_star target names;
          configs[1] at its own 2^16 chars is timed in the same run and reported under "also"
value   = inputs resident in HBM when the timed region starts    (device timed, CUDA events)
e2e     = the same pass through the C ABI with HOST buffers: document/table upload, scalar
          upload and result read-back inside the timed region
--impl reference = the CPU restatement of the reference's algorithm (oracle/c, all host cores).

Multi-GPU (torchrun), weak scaling: ONE document of base_len * G characters.  Its sum-check is
sharded by low index bits; the 96 bytes per rank per round are exchanged by the round kernels
themselves through peer mailboxes over NVLink (P2P stores + system-scope release/acquire; no NCCL
call on that path), every rank runs the same transcript.  The fold commitments (latency-bound at 2^14..2^15 terms) are distributed whole,
round-robin over the ranks; MSMs large enough to be throughput-bound are sharded by Pippenger
windows (one 128-byte all-gather per MSM).
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

FQ = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
FP = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001

WORKLOADS = {
    # name: doc_len, alphabet, nova steps, lookups per step, log2 |T|, primary / secondary MSM sizes
    # "target" = the configuration BASELINE.json's north_star quotes the metric on ("a 2^20-char ascii
    # document / '.*b' at 1 GPU"; BASELINE.md: "cfg-2-style .*b at 2^20 chars"): configs[1] with the
    # document length of the target.  configs[1] itself (2^16 chars) is "cfg2" and is reported beside it.
    "target": dict(doc_len=1 << 20, ab="ascii", steps=2, m=2, t_log=6, n_pri=1 << 15, n_sec=1 << 14,
                   desc="north_star target = configs[1] at 2^20 chars: ascii doc (2^20-1 x 'a' + 'b'), re '.*b', --prove: "
                        "per Nova fold nl(T=2^6) + nldoc(N=2^21,u32) sum-checks, 2 calc_d, commit(W),commit(T) on "
                        "Pallas(2^15) and Vesta(2^14); 2 folds"),
    "cfg2": dict(doc_len=1 << 16, ab="ascii", steps=2, m=2, t_log=6, n_pri=1 << 15, n_sec=1 << 14,
                 desc="ascii 2^16-char doc (65535 x 'a' + 'b'), re '.*b', --prove: per Nova fold "
                      "nl(T=2^6) + nldoc(N=2^17,u32) sum-checks, 2 calc_d, commit(W),commit(T) on "
                      "Pallas(2^15) and Vesta(2^14); 2 folds"),
    "cfg4": dict(doc_len=1 << 20, ab="ascii", doc="cfg4", steps=2, m=4, t_log=8, n_pri=1 << 16, n_sec=1 << 14,
                 desc="ascii 2^20-char doc, per fold nl(T=2^8) + nldoc(N=2^21,u32), 2 calc_d, 4 MSMs; 2 folds"),
    "cfg5": dict(doc_len=1 << 22, ab="ascii", doc="cfg4", steps=1, m=4, t_log=8, n_pri=1 << 16, n_sec=1 << 14,
                 desc="2^22-char doc, nl(T=2^8) + nldoc(N=2^23,u32), 2 calc_d, 4 MSMs; 1 fold"),
}


import workloads as WL                                  # numpy-only input generators shared by both arms and the tests

ASCII_AB = "".join(chr(c) for c in range(128))        # config.rs: the `ascii` alphabet
curve_multiples = WL.curve_multiples


class ParityError(AssertionError):
    """The multi-GPU path produced a result that differs from the single-GPU path."""


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
    rnd = random.Random(1234 + seed_shift)
    # weak scaling: one document of base_len * G characters.  The reference's f32 `logmn` (costs.rs:10-15)
    # mis-rounds 2^22+2 and 2^23+2, so its doc_transform panics on documents of exactly 2^22 / 2^23
    # characters (framework.rs:1007): 64 more characters put the length where logmn is exact; the padded
    # table (2^21 entries per GPU) is unchanged.
    w["doc_len"] = WL.safe_len(w["doc_len"] * world)
    doc_len = w["doc_len"]
    ab, cps = WL.document(w.get("doc", name), doc_len, seed_shift)
    w["udoc"] = np.ascontiguousarray(WL.encode(ab, cps))
    tl = w["t_log"]
    w["T"] = sorted(rnd.randrange(1 << 40) for _ in range(1 << tl))
    w["T_bytes"] = pack(w["T"])
    S, m = w["steps"], w["m"]
    w["q_nl"] = [[rnd.randrange(1 << tl) for _ in range(m)] for _ in range(S)]
    w["q_doc"] = [[rnd.randrange(doc_len + 2) for _ in range(m)] for _ in range(S)]
    w["doc_hash"] = rnd.randrange(FQ)
    w["salt"] = rnd.randrange(FQ)

    def scalars(n, order, witness_like):
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

    # generators k*G: distinct, cheap, and the expected MSM result is checkable
    w["bases_pri"] = WL.generators("pallas", w["n_pri"])     # Pallas: over Fp  (disk-cached k*G, build/gens/)
    w["bases_sec"] = WL.generators("vesta", w["n_sec"])      # Vesta: over Fq
    w["sc"] = [dict(Wp=scalars(w["n_pri"], FQ, True), Tp=scalars(w["n_pri"], FQ, False),
                    Ws=scalars(w["n_sec"], FP, True), Ts=scalars(w["n_sec"], FP, False)) for _ in range(S)]
    return w


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class GpuPass:
    """One prove pass through libreef_b200 with raw buffers (no Python big-int work inside)."""

    def __init__(self, ctxs, w, rank=0, world=1, dist=None):
        """ctxs: dict of libreef_b200 contexts (one CUDA stream each): 'nl', 'doc', 'pri', 'sec'.
        The reference runs the sum-checks on its solver thread and the fold commitments on its
        proving thread (framework.rs:98-110); here each of the independent chains of a
        fold has its own context/stream and host thread."""
        import reef_b200
        import torch
        from concurrent.futures import ThreadPoolExecutor
        self.rb, self.torch, self.ctxs, self.w = reef_b200, torch, ctxs, w
        self.ctx = ctxs["doc"]
        self.lib, self.check = reef_b200.lib, reef_b200._lib.check
        self.rank, self.world, self.dist = rank, world, dist
        self.bases_pri = reef_b200.Bases(ctxs["pri"], "pallas", w["bases_pri"])
        self.bases_sec = reef_b200.Bases(ctxs["sec"], "vesta", w["bases_sec"])
        # commit(T) gets its own context (generators registered there too): it only waits for the
        # cross-term scalars, not for commit(W), so the two commitments of a curve overlap
        self.bases_pri2 = reef_b200.Bases(ctxs["pri2"], "pallas", w["bases_pri"])
        self.bases_sec2 = reef_b200.Bases(ctxs["sec2"], "vesta", w["bases_sec"])
        self.ell_doc = reef_b200.logmn(len(w["udoc"]))
        self.ell_T = w["t_log"]
        self.pool = {k: ThreadPoolExecutor(max_workers=1) for k in ctxs}
        # e2e leg: every host buffer that crosses PCIe inside the timed region is page-locked
        self._pins = []
        self.h_doc = self._pin(self._doc_shard())
        self.h_T = self._pin(np.frombuffer(w["T_bytes"], dtype=np.uint8))
        for s in w["sc"]:
            for k in list(s):
                s[k] = self._pin(s[k])
        # the MSM thread and the sum-check thread issue collectives concurrently: one communicator each
        self.msm_group = dist.new_group() if world > 1 else None
        self._mk_out()

    def _pin(self, arr):
        t = self.torch.from_numpy(np.array(arr, copy=True)).pin_memory()
        self._pins.append(t)
        return t.numpy()

    def _mk_out(self):
        from reef_b200._lib import NlookupOut
        self.o, self.bufs = {}, {}
        for key, ell in (("nl", self.ell_T), ("nldoc", self.ell_doc)):
            b = dict(prev=C.create_string_buffer(32), cq=C.create_string_buffer(64 * 32), claim=C.create_string_buffer(32),
                     rounds=C.create_string_buffer(ell * 128), last=C.create_string_buffer(32), nxt=C.create_string_buffer(32))
            o = NlookupOut()
            o.prev_running_claim = C.addressof(b["prev"]); o.combined_q = C.addressof(b["cq"]); o.combined_q_cap = 64
            o.claim_r = C.addressof(b["claim"]); o.rounds = C.addressof(b["rounds"]); o.rounds_cap = ell
            o.sc_last_claim = C.addressof(b["last"]); o.next_running_claim = C.addressof(b["nxt"])
            self.o[key], self.bufs[key] = o, b

    # -- residency ----------------------------------------------------------------------
    def make_resident(self):
        w, t = self.w, self.torch
        self.doc_tab = self.ctxs["doc"].table_u32(self._doc_shard())
        self.T_tab = self.rb.Table(self.ctxs["nl"], values=w["T"])
        if self.world > 1:
            self.gbuf = t.zeros((self.world + 1) * 96, dtype=t.uint8, device="cuda")
            self._connect_mailboxes()
        self.sc_dev = [{k: t.from_numpy(v.view(np.int64)).cuda() for k, v in s.items()} for s in w["sc"]]
        t.cuda.synchronize()

    def _doc_shard(self):
        return np.ascontiguousarray(self.w["udoc"][self.rank::self.world]) if self.world > 1 else self.w["udoc"]

    def _connect_mailboxes(self):
        """Per-round exchange of the sharded sum-check over NVLink peer memory (reef_b200/csrc/p2p.cu):
        every rank maps every peer's mailbox with CUDA IPC once; NCCL only carries the 64-byte handles."""
        t = self.torch
        ctx = self.ctxs["doc"]
        mine = t.frombuffer(bytearray(ctx.mailbox_create(self.world)), dtype=t.uint8).cuda()
        allh = t.empty(self.world * 64, dtype=t.uint8, device="cuda")
        self.dist.all_gather_into_tensor(allh, mine)
        ctx.mailbox_connect(self.rank, self.world, bytes(allh.cpu().numpy().tobytes()))
        self.dist.barrier()

    def _gather(self, mine_ptr, nbytes, out_ptr):
        """all-gather of `nbytes` per rank: one stream-ordered kernel of the library (P2P stores into the
        peers' mailboxes + system-scope release/acquire), no NCCL call and no host wait"""
        self.ctxs["doc"].p2p_allgather(mine_ptr, nbytes, out_ptr)

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
        self.d_futs.append(("nldoc", self.pool["aux"].submit(self._calc_d, nxt)))
        if self.collect is not None:
            self.collect.append(("nldoc", le32(res.claim_r), b"".join(pack(r) for r in res.rounds), le32(res.sc_last_claim), nxt))
        return pack(res.next_running_q), nxt

    def _calc_d(self, v):
        # calc_d of a new running claim (framework.rs:517-553): an input of the step circuit only, the
        # next fold's sum-check does not wait for it
        d = C.create_string_buffer(32)
        self.check(self.lib.reef_calc_d(self.ctxs["aux"]._h, v, self.salt, d))
        return d.raw

    def _nlookup(self, key, tab, q_arr, v_bytes, prev):
        o, b = self.o[key], self.bufs[key]
        tag = 0 if key == "nl" else 1
        pq = prev[0] if prev else None
        pv = prev[1] if prev else None
        ctx = self.ctxs["nl" if key == "nl" else "doc"]
        self.check(self.lib.reef_nlookup_prove(ctx._h, tag, tab._h, q_arr.ctypes.data, v_bytes, len(q_arr), pq, pv,
                                              self.dh if tag else None, C.byref(o)))
        ell = o.ell
        rounds = b["rounds"].raw
        next_q = b"".join(rounds[i * 128:i * 128 + 32] for i in range(ell))
        nxt = b["nxt"].raw
        self.d_futs.append((key, self.pool["aux"].submit(self._calc_d, nxt)))
        if self.collect is not None:
            self.collect.append((key, b["claim"].raw, rounds[:ell * 128], b["last"].raw, nxt))
        return next_q, nxt

    # An MSM is window-sharded across the ranks only when it is large enough to be throughput-bound
    # (n * windows digit entries); the 2^14..2^15-term fold commitments are latency pipelines, so
    # with G > 1 ranks each WHOLE commitment goes to one rank (round-robin) and the 64-byte results
    # are exchanged once per pass.
    SHARD_MIN_ENTRIES = 1 << 22

    def _msm_local(self, bases, dev_tensor, host_arr, n, resident):
        out = C.create_string_buffer(64)
        ctx = bases.ctx
        if resident:
            self.check(self.lib.reef_msm_dev(ctx._h, bases._h, C.c_void_p(dev_tensor.data_ptr()), n, out))
        else:
            self.check(self.lib.reef_msm(ctx._h, bases._h, host_arr.ctypes.data, n, out))
        return out.raw

    def _msm_sharded(self, bases, dev_tensor, host_arr, n, resident):
        # windows [w0, w1) on this rank, one all-gather of 128-byte partial points
        out = C.create_string_buffer(64)
        ctx = bases.ctx
        t = self.torch
        if not resident:
            dev_tensor = t.from_numpy(host_arr.view(np.int64)).cuda()
        W = bases.windows
        w0, w1 = W * self.rank // self.world, W * (self.rank + 1) // self.world
        part = C.create_string_buffer(128)
        self.check(self.lib.reef_msm_partial_dev(ctx._h, bases._h, C.c_void_p(dev_tensor.data_ptr()), n, w0, w1, part))
        mine = t.frombuffer(bytearray(part.raw), dtype=t.uint8).cuda()
        gathered = [t.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(gathered, mine, group=self.msm_group)
        allp = b"".join(bytes(g.cpu().numpy().tobytes()) for g in gathered)
        self.check(self.lib.reef_msm_combine(ctx._h, bases.curve, allp, self.world, out))
        return out.raw

    def _exchange_points(self, outs, owners):
        """One all-gather per pass: every rank ends up with all commitments (64 B each)."""
        t = self.torch
        k = len(outs)
        mine = bytearray(k * 64)
        for i, o in enumerate(outs):
            if o is not None:
                mine[i * 64:(i + 1) * 64] = o
        buf = t.frombuffer(mine, dtype=t.uint8).cuda()
        allb = t.empty(self.world * k * 64, dtype=t.uint8, device="cuda")
        self.dist.all_gather_into_tensor(allb, buf)
        host = allb.cpu().numpy().tobytes()
        return [outs[i] if owners[i] is None else host[(owners[i] * k + i) * 64:(owners[i] * k + i + 1) * 64] for i in range(k)]

    collect = None

    def run_collect(self, resident: bool = True):
        """One untimed pass that keeps EVERY output (per fold: claim_r, all round polynomials and
        challenges, last claim, next running claim of both sum-checks; the calc_d digests; the four
        commitments) in the same layout as cpu_pass(collect=True) for the bit-for-bit comparison."""
        self.collect = []
        try:
            _, _, outs, ds = self.run(resident)
            sc = list(self.collect)      # each sum-check kind runs on its own thread, in fold order
            got = {"nl": [r[1:] for r in sc if r[0] == "nl"], "nldoc": [r[1:] for r in sc if r[0] == "nldoc"],
                   "d_nl": [d for k, d in ds if k == "nl"], "d_nldoc": [d for k, d in ds if k == "nldoc"], "msm": list(outs)}
        finally:
            self.collect = None
        return got

    def run(self, resident: bool):
        """One pass.  resident=False: every input crosses PCIe inside the call (e2e leg).
        Dependencies kept: sum-check of fold i+1 needs the running claim of fold i; the fold
        commitments of fold i are issued after both sum-checks of fold i (their witness)."""
        w = self.w
        self.dh = le32(w["doc_hash"])
        self.salt = le32(w["salt"])
        if resident:
            doc_tab, T_tab = self.doc_tab, self.T_tab
        else:
            f1 = self.pool["doc"].submit(lambda: self.ctxs["doc"].table_u32(self.h_doc))   # H2D of the document codes
            f2 = self.pool["nl"].submit(self._upload_T)
            doc_tab, T_tab = f1.result(), f2.result()
        prev_nl = prev_doc = None
        msm_futs, owners = [], []
        self.d_futs = []
        for s in range(w["steps"]):
            qn, qd = self.q_nl[s], self.q_doc[s]
            f_nl = self.pool["nl"].submit(self._nlookup, "nl", T_tab, qn[0], qn[1], prev_nl)
            if self.world > 1:
                f_doc = self.pool["doc"].submit(self._nlookup_sharded, doc_tab, w["q_doc"][s], [int(w["udoc"][i]) for i in w["q_doc"][s]], prev_doc)
            else:
                f_doc = self.pool["doc"].submit(self._nlookup, "nldoc", doc_tab, qd[0], qd[1], prev_doc)
            prev_nl, prev_doc = f_nl.result(), f_doc.result()
            sc, scd = w["sc"][s], (self.sc_dev[s] if resident else None)
            for key, pool, bases, n in (("Wp", "pri", self.bases_pri, w["n_pri"]), ("Ws", "sec", self.bases_sec, w["n_sec"]),
                                        ("Tp", "pri2", self.bases_pri2, w["n_pri"]), ("Ts", "sec2", self.bases_sec2, w["n_sec"])):
                args = (bases, scd[key] if resident else None, sc[key], n, resident)
                if self.world == 1:
                    msm_futs.append(self.pool[pool].submit(self._msm_local, *args))
                    owners.append(None)
                elif n * bases.windows >= self.SHARD_MIN_ENTRIES:
                    # one thread issues these collectives, in the same order on every rank
                    msm_futs.append(self.pool["pri"].submit(self._msm_sharded, *args))
                    owners.append(None)
                else:
                    owner = len(owners) % self.world
                    msm_futs.append(self.pool[pool].submit(self._msm_local, *args) if owner == self.rank else None)
                    owners.append(owner)
        outs = [f.result() if f is not None else None for f in msm_futs]
        ds = [(k, f.result()) for k, f in self.d_futs]
        if self.world > 1 and any(o is not None for o in owners):
            outs = self._exchange_points(outs, owners)
        if not resident:
            doc_tab.free()
            T_tab.free()
        return prev_nl, prev_doc, outs, ds

    def _upload_T(self):
        h = C.c_void_p()
        self.check(self.lib.reef_table_upload(self.ctxs["nl"]._h, self.h_T.ctypes.data, len(self.w["T"]), C.byref(h)))
        t = self.rb.Table.__new__(self.rb.Table)
        t.ctx, t._h = self.ctxs["nl"], h
        return t

    def prepare_queries(self):
        w = self.w
        self.q_nl = [(np.asarray(q, dtype=np.uint64), pack(w["T"][i] for i in q)) for q in w["q_nl"]]
        self.q_doc = [(np.asarray(q, dtype=np.uint64), pack(int(w["udoc"][i]) for i in q)) for q in w["q_doc"]]

    def bytes_per_step(self):
        w = self.w
        h2d = w["udoc"].nbytes + len(w["T_bytes"])
        for s in range(w["steps"]):
            h2d += sum(v.nbytes for v in w["sc"][s].values())
            for ell, m in ((self.ell_T, w["m"]), (self.ell_doc, w["m"])):
                h2d += (m + ell + 3) * 32 + ell * 32 + m * 8
        d2h = 0
        for s in range(w["steps"]):
            d2h += 2 * 8192 + 2 * 32 + 4 * 64                        # NlState x2, calc_d x2, 4 points
        return h2d, d2h


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
    ctxs = {k: reef_b200.Context(dev) for k in ("nl", "doc", "pri", "sec", "pri2", "sec2", "aux")}
    # weak scaling: ONE document of base_len * G characters.  Its nldoc sum-check is sharded by
    # low index bits (rank g holds udoc[g::G]); the fold commitments are sharded by Pippenger
    # windows; the tiny T-table sum-check is replicated.
    w = make_workload(args.workload, seed_shift=0, world=world)
    gp = GpuPass(ctxs, w, rank, world, dist)
    gp.prepare_queries()
    gp.make_resident()
    # Untimed, before the timed legs: the whole pass on this workload is compared bit for bit with the
    # CPU restatement of the reference's algorithm (oracle/c) -- every rank against rank 0's CPU run of
    # the SAME (world-scaled) document.  The same CPU run is the cpu_baseline figure of the line.
    verified, cpu_line = None, None
    if not args.no_cpu_baseline:
        exp = [None]
        if rank == 0:
            exp[0] = {}
            cpu_line = cpu_baseline(w, None, exp[0])
        if world > 1:
            dist.broadcast_object_list(exp, src=0)
        got = gp.run_collect(True)
        verified = compare_outputs(got, exp[0], f"rank {rank}")      # ParityError is loud
        got = gp.run_collect(False)
        compare_outputs(got, exp[0], f"rank {rank}, host-buffer (e2e) path")
        if world > 1:
            verified += f"; every one of the {world} ranks checked its own copy of the results"
    streams = {k: torch.cuda.ExternalStream(c.stream) for k, c in ctxs.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(gp, resident, steps, warmup, profile):
        barrier()                               # ranks enter every leg together
        for _ in range(warmup):
            flush.fill_(1)
            gp.run(resident)
        barrier()
        if profile:
            for c in ctxs.values():
                reef_b200._lib.check(reef_b200.lib.reef_profile_enable(c._h, 1))
        launches0 = int(reef_b200.lib.reef_launch_count())
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = {k: torch.cuda.Event(enable_timing=True) for k in streams}
        barrier()
        e0.record(streams["doc"])               # every stream is idle here (barrier above)
        t0 = time.perf_counter()
        w0 = time.time()
        per_step = []
        for _ in range(steps):
            flush.fill_(1)                      # L2 flush between steps; run() returns only after all streams drained
            torch.cuda.current_stream().synchronize()
            ts = time.perf_counter()
            gp.run(resident)
            per_step.append(round((time.perf_counter() - ts) * 1e3, 3))
        if args.debug and rank == 0:
            print(f"[debug] resident={resident} profile={profile} per-step wall ms: {per_step}", file=sys.stderr)
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
        launches = int(reef_b200.lib.reef_launch_count()) - launches0
        clocks = None
        prof = None
        if profile:
            n = 9
            prof = [(0, 0, 0.0)] * n
            for c in ctxs.values():
                cnt, units, pms = (C.c_uint64 * n)(), (C.c_uint64 * n)(), (C.c_double * n)()
                reef_b200._lib.check(reef_b200.lib.reef_profile_read(c._h, n, cnt, units, pms))
                reef_b200._lib.check(reef_b200.lib.reef_profile_enable(c._h, 0))
                prof = [(p[0] + int(cnt[i]), p[1] + int(units[i]), p[2] + float(pms[i])) for i, p in enumerate(prof)]
        return ms, wall * 1e3, launches, clocks, prof

    K, Wm = args.steps, args.warmup
    # value: unprofiled run (event recording around every launch group costs host time);
    # a second, profiled run of the same K steps feeds the per-kernel figures and the clocks.
    clock_windows = []
    clk = sample_clocks_start() if rank == 0 else (None, None)
    ms, wall_ms, launches, _, _ = timed(gp, True, K, Wm, False)
    _, _, _, _, prof = timed(gp, True, K, Wm, True)
    clocks = sample_clocks_stop(*clk, device=dev, windows=clock_windows) if rank == 0 else None
    e2e_ms, _, _, _, _ = timed(gp, False, K, Wm, False)
    also = None
    if world == 1 and args.also and args.also != args.workload:
        # the second workload (configs[1] by default): same pass, value and e2e only
        w2 = make_workload(args.also)
        gp2 = GpuPass(ctxs, w2, rank, world, dist)
        gp2.prepare_queries()
        gp2.make_resident()
        exp2 = {}
        cpu2 = None if args.no_cpu_baseline else cpu_baseline(w2, None, exp2)
        ver2 = None if args.no_cpu_baseline else compare_outputs(gp2.run_collect(True), exp2, args.also)
        ms2, _, launches2, _, _ = timed(gp2, True, K, Wm, False)
        e2e2, _, _, _, _ = timed(gp2, False, K, Wm, False)
        h2d2, d2h2 = gp2.bytes_per_step()
        also = {"workload": args.also + ": " + w2["desc"], "value": round(w2["doc_len"] / (ms2 / K / 1e3), 1),
                "ms_per_step": round(ms2 / K, 4), "gpu_launches": launches2, "verified": ver2, "cpu_baseline": cpu2,
                "e2e": {"value": round(w2["doc_len"] / (e2e2 / K / 1e3), 1), "unit": "NFA steps/s", "ms_per_step": round(e2e2 / K, 4),
                        "h2d_bytes_per_step": h2d2, "d2h_bytes_per_step": d2h2}}
    doc_units = w["doc_len"]            # already base_len * world (one sharded document)
    value = doc_units / (ms / K / 1e3)
    e2e_value = doc_units / (e2e_ms / K / 1e3)
    if rank != 0:
        return
    peaks = load_peaks()
    names = ["sweep_first", "sweep_fold", "round_transcript", "tail", "nl_setup", "msm_sort", "msm_accum", "msm_reduce", "poseidon"]
    total_prof = sum(p[2] for p in prof) or 1.0
    shares = {n: round(p[2] / total_prof, 4) for n, p in zip(names, prof)}
    kernel_ms = {n: round(p[2] / K, 4) for n, p in zip(names, prof)}
    # MLE sweep roofline: algorithmic bytes of the reference-shaped fused schedule (SURVEY 8d:
    # 2 tables x 32 B): round-1 pass reads 2 L elements (64 L bytes); a fold+accumulate pass over
    # an input of length L reads 2 L and writes L elements (96 L bytes).
    sweep_bytes = 64.0 * prof[0][1] + 96.0 * prof[1][1]
    sweep_ms = prof[0][2] + prof[1][2]
    achieved = sweep_bytes / (sweep_ms / 1e3) / 1e9 if sweep_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": "k_sweep (MLE fold+accumulate passes)", "achieved": round(achieved, 1),
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": None,
            "peak_source": peaks["source"], "launches": prof[0][0] + prof[1][0],
            "alg_bytes_per_launch": round(sweep_bytes / max(1, prof[0][0] + prof[1][0])),
            "avg_launch_us": round(1e3 * sweep_ms / max(1, prof[0][0] + prof[1][0]), 2)}
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
    for bases, n in ((gp.bases_pri, w["n_pri"]), (gp.bases_sec, w["n_sec"])):
        Wn, c = bases.windows, bases.window_bits
        ops += 2 * w["steps"] * K * (10.0 * n * Wn + 14.0 * Wn * (1 << c)) / max(1, world)
        n_terms += 2 * w["steps"] * K * n / max(1, world)
    msm = {"bound": "int-alu (255-bit modmul)", "mops": round(n_terms / (msm_ms / 1e3) / 1e6, 2) if msm_ms else None,
           "achieved_modmul_per_s": round(ops / (msm_ms / 1e3), 0) if msm_ms else None, "peak_modmul_per_s": peaks["modmul_per_s"],
           "frac": round(ops / (msm_ms / 1e3) / peaks["modmul_per_s"], 4) if msm_ms else None,
           "peak_source": "tools/bench_fp.cu on this pool (profiles/peak_modmul.json)"}
    # the same MSM kernels on a throughput-sized instance (uniform 255-bit scalars), alone on the GPU
    if world == 1 and args.msm_large_log2:
        msm["large"] = msm_large(ctxs["pri"], args.msm_large_log2, peaks["modmul_per_s"])
    h2d, d2h = gp.bytes_per_step()
    out = {
        "metric": "NFA steps/s proved (= doc_len / hot-path prove time)", "value": round(value, 1), "unit": "NFA steps/s",
        "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": round(ms / K, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 limbs (255-bit prime fields Fq/Fp, exact integer)",
        "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "l2": "256 MiB buffer written between steps (L2 flush)",
                   "timing": "CUDA events: start on an idle stream, end = latest of the library streams; max over ranks",
                   "streams": "7 contexts/streams: nl sum-check | nldoc sum-check | commit(W) Pallas | commit(W) Vesta | commit(T) Pallas | commit(T) Vesta | calc_d "
                              "(fold i+1 sum-checks overlap fold i commitments; commit(T) does not wait for commit(W); calc_d does not gate the next fold)",
                   "verified": verified,
                   "parallelism": (f"1 document of {w['doc_len']} chars: nldoc sum-check sharded by low index bits x{world} "
                                   f"(96 bytes per rank per round; the round kernels themselves store them into the peers' mailboxes over NVLink and "
                                   f"acquire the peers' -- no NCCL call, no extra launch); fold commitments (2^14-2^15 terms, latency-bound) distributed "
                                   f"whole, round-robin over the ranks, results exchanged once per pass; MSMs with >= 2^22 "
                                   f"digit entries are sharded by Pippenger windows (128-byte all-gather)") if world > 1 else "single GPU"},
        "e2e": {"value": round(e2e_value, 1), "unit": "NFA steps/s", "ms_per_step": round(e2e_ms / K, 4),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof, "msm": msm,
        "kernel_ms_per_step": kernel_ms, "kernel_share": shares, "wall_ms_per_step": round(wall_ms / K, 4),
    }
    if also:
        out["also"] = also
    if cpu_line:
        out["cpu_baseline"] = cpu_line
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
def cpu_pass(w, n_steps, threads, collect=None):
    """Runs `n_steps` Nova folds of the workload on the CPU port; returns seconds.  `collect` (a dict)
    receives every output in GpuPass.run_collect()'s layout."""
    from oracle import cport
    from oracle.nlookup import combined_qs, logmn, nlookup_pattern
    cport.lib().oracle_set_fast_poseidon(1)     # neptune hashes with its optimised constants too
    t_total = 0.0
    T_arr = np.frombuffer(w["T_bytes"], dtype=np.uint8)
    prev = {"nl": None, "nldoc": None}
    for s in range(n_steps):
        t0 = time.perf_counter()
        for key, arr, is_u32, n, q, vals in (("nl", T_arr, 0, len(w["T"]), w["q_nl"][s], [w["T"][i] for i in w["q_nl"][s]]),
                                             ("nldoc", w["udoc"], 1, len(w["udoc"]), w["q_doc"][s], [int(w["udoc"][i]) for i in w["q_doc"][s]])):
            ell = logmn(n)
            pq, pv = prev[key] if prev[key] else ([0] * ell, (w["T"][0] if key == "nl" else int(w["udoc"][0])))
            cqs = combined_qs(list(q), ell)
            pat = nlookup_pattern(key, len(q), ell, len(cqs))
            query = ([] if key == "nl" else [w["doc_hash"]]) + cqs + vals + list(pq) + [pv]
            claim, rounds, last, nxt = cport.nlookup_raw(arr, is_u32, n, q, pack(query), len(query), cport.ops_words(pat),
                                                         pack(pq), ell)
            r = [int.from_bytes(rounds[i * 128:i * 128 + 32], "little") for i in range(ell)]
            prev[key] = (r, int.from_bytes(nxt, "little"))
            if collect is not None:
                collect.setdefault(key, []).append((claim, rounds, last, nxt))
        d1 = cport.poseidon_hash([prev["nl"][1], w["salt"]], 2)
        d2 = cport.poseidon_hash([prev["nldoc"][1], w["salt"]], 2)
        sc = w["sc"][s]
        pts = []
        for key, curve, bases in (("Wp", "pallas", w["bases_pri"]), ("Ws", "vesta", w["bases_sec"]),
                                  ("Tp", "pallas", w["bases_pri"]), ("Ts", "vesta", w["bases_sec"])):
            pts.append(cport.msm(curve, bases, sc[key].tobytes(), threads=threads))
        t_total += time.perf_counter() - t0
        if collect is not None:
            collect.setdefault("d_nl", []).append(le32(d1[0]))
            collect.setdefault("d_nldoc", []).append(le32(d2[0]))
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
    for key in ("nl", "nldoc"):
        if len(got[key]) != len(exp[key]):
            raise ParityError(f"{what}: {key}: {len(got[key])} folds vs {len(exp[key])}")
        for f, (g, e) in enumerate(zip(got[key], exp[key])):
            for name, a, b in zip(("claim_r", "rounds", "sc_last_claim", "next_running_claim"), g, e):
                if bytes(a) != bytes(b):
                    raise ParityError(f"{what}: {key} sum-check of fold {f}: {name} differs from the oracle")
    for key in ("d_nl", "d_nldoc", "msm"):
        if [bytes(x) for x in got[key]] != [bytes(x) for x in exp[key]]:
            raise ParityError(f"{what}: {key} differs from the oracle")
    n_r = sum(len(g[1]) // 128 for k in ("nl", "nldoc") for g in got[k])
    return (f"GPU pass == oracle/c on the same workload, bit for bit: {len(got['nl'])} folds x (nl + nldoc sum-checks: claim_r, "
            f"{n_r} round polynomials and challenges in total, last claim, next running claim), {len(got['d_nl']) * 2} calc_d digests, "
            f"{len(got['msm'])} commitments (untimed, before the timed legs)")


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
    for _ in range(max(0, min(args.warmup, 1) - 1)):
        cpu_pass(w, w["steps"], threads)
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
    ap.add_argument("--also", default="cfg2", help="second workload timed (value/e2e only) and reported under 'also'; '' = none")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
