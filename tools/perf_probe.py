"""Quick timing probe (run under gpurun): prints per-call latencies of the hot-path entries."""
import sys, time, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import reef_b200
from oracle.fields import FQ

ctx = reef_b200.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
rnd = random.Random(0)

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3, sum(ts) / len(ts) * 1e3

for ell in (10, 14, 17, 21, 23):
    n = 1 << ell
    codes = np.random.default_rng(ell).integers(0, 131, size=n, dtype=np.uint32)
    t = ctx.table_u32(codes)
    q = [rnd.randrange(n) for _ in range(4)]
    v = [int(codes[i]) for i in q]
    best, avg = timeit(lambda: ctx.wit_nlookup_gadget(t, q, v, None, None, "nldoc", 5))
    print(f"nlookup u32  ell={ell:2d}  best {best:8.3f} ms  avg {avg:8.3f} ms   ({256*n/best/1e6:8.1f} GB/s alg)")
    t.free()
for ell in (17, 21):
    n = 1 << ell
    raw = np.random.default_rng(ell).integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= (1 << 61) - 1
    dev = torch.from_numpy(raw.view(np.int64)).cuda()
    t = reef_b200.Table(ctx, dev_ptr=dev.data_ptr(), n=n, is_u32=False)
    q = [rnd.randrange(n) for _ in range(4)]
    vals = [int.from_bytes(raw[i].tobytes(), "little") for i in q]
    best, avg = timeit(lambda: ctx.wit_nlookup_gadget(t, q, vals, None, None, "nl"))
    print(f"nlookup Fq   ell={ell:2d}  best {best:8.3f} ms  avg {avg:8.3f} ms   ({256*n/best/1e6:8.1f} GB/s alg)")
for lg in (12, 16, 20, 21):
    n = 1 << lg
    doc = np.random.default_rng(lg).integers(0, 6, size=n, dtype=np.uint64)
    ddoc = torch.from_numpy(doc.view(np.int64)).cuda()
    total = int(reef_b200.lib.reef_merkle_tree_elems(n))
    lev = torch.empty(total * 4, dtype=torch.int64, device="cuda")
    import ctypes as C
    root = C.create_string_buffer(32); sizes = np.zeros(64, dtype=np.uint64); nl = C.c_uint32()
    def f():
        reef_b200._lib.check(reef_b200.lib.reef_merkle_build_dev(ctx._h, C.c_void_p(ddoc.data_ptr()), n, C.c_void_p(lev.data_ptr()), sizes.ctypes.data, C.byref(nl), root))
    best, avg = timeit(f, n=3, warm=1)
    print(f"merkle dev   n=2^{lg:2d}  best {best:8.3f} ms  avg {avg:8.3f} ms   ({(n-1)/best/1e3:8.2f} Mperm/s)")
rows = [rnd.randrange(FQ) for _ in range(2)]
best, avg = timeit(lambda: ctx.calc_d(rows[0], rows[1]))
print(f"calc_d              best {best:8.3f} ms  avg {avg:8.3f} ms")
best, avg = timeit(lambda: ctx.poseidon_sponge([("A", 4), ("S", 1)], [1, 2, 3, 4]))
print(f"sponge A4S1 (warp5) best {best:8.3f} ms  avg {avg:8.3f} ms")
# ---- MSM
from oracle.curves import PALLAS
for lg in (12, 15):
    n = 1 << lg
    pts = PALLAS.multiples(n)
    b = ctx.bases("pallas", pts)
    raw = np.random.default_rng(lg).integers(0, 1 << 63, size=(n, 4), dtype=np.uint64); raw[:, 3] &= (1 << 61) - 1
    dev = torch.from_numpy(raw.view(np.int64)).cuda()
    best, avg = timeit(lambda: b.msm_dev(dev.data_ptr(), n), n=5, warm=2)
    print(f"msm pallas n=2^{lg} c={b.window_bits} W={b.windows}  best {best:8.3f} ms  avg {avg:8.3f} ms  ({n/best/1e3:8.2f} Mop/s)")
# ---- commitment path of --commit: Hyrax row commitments and the prove_eval mat-vec
import random as _r
for (lr, lc, ab_bits, name) in ((8, 9, 8, "cfg2 256x512"), (10, 11, 8, "cfg3/4 1024x2048")):
    rows, cols = 1 << lr, 1 << lc
    pts = PALLAS.multiples(cols + 1)
    b = ctx.bases("pallas", pts, 255)
    M = np.random.default_rng(lr).integers(0, 131, size=rows * cols, dtype=np.uint32)
    blinds = [_r.Random(lr).randrange(1, PALLAS.order) for _ in range(rows)]
    best, avg = timeit(lambda: b.msm_rows(M, rows, cols, ab_bits, blinds), n=3, warm=1)
    print(f"hyrax commit {name} (u32 codes + blinds, host in/out) best {best:8.3f} ms  avg {avg:8.3f} ms  ({rows * cols / best / 1e3:8.1f} M terms/s)")
    t = ctx.table_u32(M)
    L = [_r.Random(7).randrange(FQ) for _ in range(rows)]
    best, avg = timeit(lambda: ctx.hyrax_lz(t, rows, cols, L), n=3, warm=1)
    print(f"hyrax LZ mat-vec {name} best {best:8.3f} ms  avg {avg:8.3f} ms")
    t.free(); b.free()
