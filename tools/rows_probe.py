"""Where the Hyrax row commitment spends its time (run under gpurun): per-class device times of reef_msm_rows_u32."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import reef_b200
import workloads as WL
from reef_b200._lib import check, lib

ctx = reef_b200.Context(0)
for rows, cols, bits in ((1024, 2048, 8), (2048, 4096, 21)):
    gens = WL.generators("pallas", cols + 1)
    b = reef_b200.Bases(ctx, "pallas", gens, 255)
    codes = torch.from_numpy(np.random.default_rng(1).integers(0, 1 << bits, size=rows * cols, dtype=np.uint32).astype(np.int32)).pin_memory().numpy().view(np.uint32)
    blinds = np.random.default_rng(2).integers(0, 1 << 62, size=(rows, 4), dtype=np.uint64)
    blinds[:, 3] &= (1 << 61) - 1
    out = torch.empty(rows * 64, dtype=torch.uint8).pin_memory().numpy()
    f = lambda: check(lib.reef_msm_rows_u32(ctx._h, b._h, codes.ctypes.data, rows, cols, bits, blinds.ctypes.data, out.ctypes.data))
    for _ in range(2):
        f()
    t0 = time.perf_counter()
    for _ in range(5):
        f()
    wall = (time.perf_counter() - t0) / 5 * 1e3
    check(lib.reef_profile_enable(ctx._h, 1))
    for _ in range(3):
        f()
    cnt, units, pms = (C.c_uint64 * 9)(), (C.c_uint64 * 9)(), (C.c_double * 9)()
    check(lib.reef_profile_read(ctx._h, 9, cnt, units, pms))
    check(lib.reef_profile_enable(ctx._h, 0))
    print(f"rows {rows} x {cols} ({bits}-bit): wall {wall:.3f} ms; sort {pms[5]/3:.3f} accum {pms[6]/3:.3f} reduce {pms[7]/3:.3f} ms; c={b.window_bits} W={b.windows}", flush=True)
    b.free()
