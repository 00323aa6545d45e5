#!/bin/bash
# per-launch times (ncu, cold caches, serialised) of one Hyrax row commitment 1024 x 2048: random bytes and the
# all-equal 'aaaa...b' document of the target workload (run under gpurun)
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
which = sys.argv[1]
if which == "random":
    codes = np.random.default_rng(3).integers(0, 256, size=rows * cols, dtype=np.uint32)
else:
    codes = np.full(rows * cols, ord("a"), dtype=np.uint32); codes[-1] = ord("b")
for _ in range(3):
    b.doc_commit(codes, rows, cols, 8, blinds)
PY
for w in random aaaab; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_rows_launches_$w.csv python tools/_rows_tmp.py $w > /dev/null 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/r2z_rows_launches_$w.csv") if l.startswith('"'))]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
L = [(r[ki].split("(")[0], float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0)) for r in rows[1:]]
n = len(L) // 3
print("$w: launches per call", n, "total us of the last call %.1f" % sum(t for _, t in L[-n:]))
for k, t in L[-n:]: print("   %-40s %9.1f us" % (k[:40], t))
PY
done
rm -f tools/_rows_tmp.py
