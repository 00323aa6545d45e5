"""MSM throughput probe (run under gpurun): Mop/s and fraction of the measured modmul peak for
uniform 255-bit scalars at several sizes, with the per-class device times.
Usage: python tools/msm_probe.py [log2_n ...]"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import reef_b200
import workloads as WL

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "profiles", "peak_modmul.json")))["modmul_per_s"]
ctx = reef_b200.Context(0)
names = ["sweep_first", "sweep_fold", "round", "tail", "nl_setup", "msm_sort", "msm_accum", "msm_reduce", "poseidon"]
lgs = [int(a) for a in sys.argv[1:]] or [14, 16, 18, 20]
pts_all = WL.generators("pallas", 1 << max(lgs))      # disk-cached k*G (build/gens)
for lg in lgs:
    n = 1 << lg
    b = ctx.bases("pallas", pts_all[:64 * n])
    raw = np.random.default_rng(lg).integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= (1 << 61) - 1
    dev = torch.from_numpy(raw.view(np.int64)).cuda()
    for _ in range(2):
        b.msm_dev(dev.data_ptr(), n)
    reef_b200._lib.check(reef_b200.lib.reef_profile_enable(ctx._h, 1))
    reps = 5
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        b.msm_dev(dev.data_ptr(), n)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps
    cnt, units, pms = (C.c_uint64 * 9)(), (C.c_uint64 * 9)(), (C.c_double * 9)()
    reef_b200._lib.check(reef_b200.lib.reef_profile_read(ctx._h, 9, cnt, units, pms))
    reef_b200._lib.check(reef_b200.lib.reef_profile_enable(ctx._h, 0))
    W, c = b.windows, b.window_bits
    ops = 10.0 * n * W + 14.0 * W * (1 << c)
    dev_ms = sum(pms[i] for i in (5, 6, 7)) / reps
    print(f"msm pallas n=2^{lg} c={c} W={W}: wall {wall * 1e3:.3f} ms, device classes {dev_ms:.3f} ms "
          f"(sort {pms[5] / reps:.3f} accum {pms[6] / reps:.3f} reduce {pms[7] / reps:.3f}) -> {n / wall / 1e6:.1f} Mop/s, "
          f"{ops / wall / 1e9:.2f} G modmul/s = {ops / wall / peak:.3f} of peak", flush=True)
    b.free()
