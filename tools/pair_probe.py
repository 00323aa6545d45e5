"""Latency of the fold commitments (run under gpurun): single MSM vs the W/T pair as one row-batched MSM."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import reef_b200
import workloads as WL
from reef_b200._lib import check, lib

ctx = reef_b200.Context(0)
for curve, lg in (("pallas", 15), ("vesta", 14), ("pallas", 16)):
    n = 1 << lg
    b = reef_b200.Bases(ctx, curve, WL.generators(curve, n))
    rs = np.random.default_rng(lg)
    raw = rs.integers(0, 1 << 63, size=(2 * n, 4), dtype=np.uint64)
    raw[:, 3] &= (1 << 61) - 1
    small = rs.random(n) < 0.85                      # row 0 witness-like
    raw[:n][small, 1:] = 0
    raw[:n][small, 0] &= 0xFFFF
    dev = torch.from_numpy(raw.view(np.int64)).cuda()
    o1, o2 = C.create_string_buffer(64), C.create_string_buffer(128)

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3
    single_w = t(lambda: check(lib.reef_msm_dev(ctx._h, b._h, C.c_void_p(dev.data_ptr()), n, o1)))
    single_t = t(lambda: check(lib.reef_msm_dev(ctx._h, b._h, C.c_void_p(dev.data_ptr() + n * 32), n, o1)))
    pair = t(lambda: check(lib.reef_msm_rows_dev(ctx._h, b._h, C.c_void_p(dev.data_ptr()), 2, n, o2)))
    check(lib.reef_profile_enable(ctx._h, 1))
    for _ in range(5):
        check(lib.reef_msm_rows_dev(ctx._h, b._h, C.c_void_p(dev.data_ptr()), 2, n, o2))
    cnt, units, pms = (C.c_uint64 * 9)(), (C.c_uint64 * 9)(), (C.c_double * 9)()
    check(lib.reef_profile_read(ctx._h, 9, cnt, units, pms))
    check(lib.reef_profile_enable(ctx._h, 0))
    print(f"{curve} n=2^{lg}: single(W-like) {single_w:.3f} ms, single(uniform) {single_t:.3f} ms, pair {pair:.3f} ms "
          f"(pair classes: sort {pms[5]/5:.3f} accum {pms[6]/5:.3f} reduce {pms[7]/5:.3f})", flush=True)
    b.free()
