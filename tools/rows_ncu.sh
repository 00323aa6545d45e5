#!/bin/bash
# ncu --set full of the combine / scan / row kernels inside one Hyrax row commitment 1024 x 2048 (run under gpurun)
mkdir -p gpurun_out
cat > tools/_rows_tmp.py <<'PY'
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import random, numpy as np, reef_b200, workloads as WL
from oracle.fields import FQ
ctx = reef_b200.Context(0)
rnd = random.Random(1)
rows, cols = 1024, 2048
b = reef_b200.Bases(ctx, "pallas", WL.generators("pallas", cols + 1), 255)
blinds = [rnd.randrange(FQ) for _ in range(rows)]
codes = np.full(rows * cols, ord("a"), dtype=np.uint32); codes[-1] = ord("b")
import ctypes as C
from reef_b200._lib import check, lib
out = np.zeros(rows * 64, dtype=np.uint8)
bl = np.zeros((rows, 4), dtype=np.uint64)
for _ in range(2):
    check(lib.reef_msm_rows_u32(ctx._h, b._h, codes.ctypes.data, rows, cols, 8, bl.ctypes.data, out.ctypes.data))
PY
ncu --set full --clock-control none --import-source on -k regex:'k_accum_next|k_scan|k_rows_affine|k_rows_final|k_accum_first' -s 9 -c 9 -o /tmp/r2z_rows python tools/_rows_tmp.py > gpurun_out/r2z_ncu_rows.log 2>&1
ncu -i /tmp/r2z_rows.ncu-rep --page raw --csv > gpurun_out/r2z_prof_rows_raw.csv 2>/dev/null
python tools/summarize_ncu_raw.py gpurun_out/r2z_prof_rows_raw.csv gpurun_out/r2z_ncu_rows_summary.csv
cat gpurun_out/r2z_ncu_rows_summary.csv | cut -d, -f1-5,8-11,15-23
rm -f tools/_rows_tmp.py
tail -2 gpurun_out/r2z_ncu_rows.log
