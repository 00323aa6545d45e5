"""Isolated timing of one nlookup sum-check (no concurrent streams): per-kernel-class ms.
Usage (under gpurun): python tools/sweep_probe.py [log2_N ...]"""
import ctypes as C, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import reef_b200

names = ["sweep_first", "sweep_fold", "round", "tail", "nl_setup"]
ctx = reef_b200.Context(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rnd = random.Random(1)
for ell in [int(a) for a in sys.argv[1:]] or [17, 21, 23]:
    n = 1 << ell
    codes = np.random.default_rng(ell).integers(0, 131, size=n, dtype=np.uint32)
    t = ctx.table_u32(codes)
    q = [rnd.randrange(n) for _ in range(4)]
    v = [int(codes[i]) for i in q]
    for _ in range(3):
        ctx.wit_nlookup_gadget(t, q, v, None, None, "nldoc", 5)
    reef_b200._lib.check(reef_b200.lib.reef_profile_enable(ctx._h, 1))
    reps = 5
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        ctx.wit_nlookup_gadget(t, q, v, None, None, "nldoc", 5)
    cnt, units, pms = (C.c_uint64 * 9)(), (C.c_uint64 * 9)(), (C.c_double * 9)()
    reef_b200._lib.check(reef_b200.lib.reef_profile_read(ctx._h, 9, cnt, units, pms))
    reef_b200._lib.check(reef_b200.lib.reef_profile_enable(ctx._h, 0))
    ms = {nm: pms[i] / reps for i, nm in enumerate(names)}
    byts = (64.0 * units[0] + 96.0 * units[1]) / reps
    sw = ms["sweep_first"] + ms["sweep_fold"]
    print(f"ell={ell} " + " ".join(f"{k}={x * 1e3:.1f}us" for k, x in ms.items()) +
          f" | sweeps {sw * 1e3:.1f} us, {byts / (sw / 1e3) / 1e9:.0f} GB/s alg ({int(cnt[0] + cnt[1]) // reps} launches)", flush=True)
    t.free()
