#!/bin/bash
# ncu --set full of the kernels added in round 2 (run under gpurun)
mkdir -p gpurun_out
cat > /tmp/new_kernels.py <<'PY'
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import random, numpy as np, reef_b200, workloads as WL
from oracle.fields import FP, FQ
ctx = reef_b200.Context(0)
rnd = random.Random(1)
e = [rnd.randrange(FP) for _ in range(240)]
for _ in range(2):
    ctx.poseidon_ro(e, "fp")
rows, cols = 1024, 2048
gens = WL.generators("pallas", cols + 1)
b = reef_b200.Bases(ctx, "pallas", gens, 255)
codes = np.random.default_rng(3).integers(0, 256, size=rows * cols, dtype=np.uint32)
blinds = [rnd.randrange(FQ) for _ in range(rows)]
for _ in range(2):
    b.doc_commit(codes, rows, cols, 8, blinds)
n = 1 << 13
kb = reef_b200.Bases(ctx, "pallas", WL.generators("pallas", n), 255)
from reef_b200 import snark as G
s = G.Ipa(ctx, "pallas", kb, (FP - 1, 2), [rnd.randrange(FQ) for _ in range(n)], [rnd.randrange(FQ) for _ in range(n)])
L, R = s.round(); s.fold(5, pow(5, -1, FQ)); s.round()
PY
cp /tmp/new_kernels.py tools/_new_kernels_tmp.py
ncu --set full --clock-control none --import-source on -k regex:'k_poseidon_ro_fast|k_bitsum_partial_warp|k_rows_final|k_rows_affine|k_ipa_scalars|k_ipa_weights|k_digits_rows|k_cross_term' -c 16 -o /tmp/r02_new python tools/_new_kernels_tmp.py > gpurun_out/r02_ncu_new.log 2>&1
ncu -i /tmp/r02_new.ncu-rep --page raw --csv > gpurun_out/r02_prof_new_raw.csv 2>/dev/null
rm -f tools/_new_kernels_tmp.py
tail -3 gpurun_out/r02_ncu_new.log
wc -l gpurun_out/r02_prof_new_raw.csv
