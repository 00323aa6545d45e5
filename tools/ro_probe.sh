#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_poseidon_ro.py tests/test_gpu_nifs.py -m gpu -x -q --timeout 600 2>&1 | tail -3
REEF_RO_TEXTBOOK=1 python -m pytest tests/test_poseidon_ro.py -m gpu -x -q --timeout 600 2>&1 | tail -2
for tb in 0 1; do
REEF_RO_TEXTBOOK=$tb python - <<'PY'
import os, time, random, reef_b200
from oracle.fields import FP
ctx = reef_b200.Context(0)
rnd = random.Random(1)
for n in (24, 3072):
    e = [rnd.randrange(FP) for _ in range(n)]
    ctx.poseidon_ro(e, "fp")
    t0 = time.perf_counter(); ctx.poseidon_ro(e, "fp"); dt = time.perf_counter() - t0
    print(f"REEF_RO_TEXTBOOK={os.environ['REEF_RO_TEXTBOOK']} poseidon_ro n={n}: {dt*1e3:.3f} ms ({dt*1e6/((n+23)//24):.1f} us per permutation)")
PY
done
