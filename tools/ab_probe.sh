#!/bin/bash
# A/B of the programmatic-dependent-launch chain and a quick parity subset (run under gpurun)
mkdir -p gpurun_out
TAG=${1:-ab}
python -m pytest tests/test_poseidon_ro.py tests/test_gpu_mle.py tests/test_gpu_msm.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1
tail -5 gpurun_out/${TAG}_pytest.log
REEF_PDL=0 python tools/perf_probe.py > gpurun_out/${TAG}_perf_pdl0.txt 2>&1
REEF_PDL=1 python tools/perf_probe.py > gpurun_out/${TAG}_perf_pdl1.txt 2>&1
grep -h "nlookup\|msm\|calc_d" gpurun_out/${TAG}_perf_pdl0.txt | head -12
echo ---
grep -h "nlookup\|msm\|calc_d" gpurun_out/${TAG}_perf_pdl1.txt | head -12
python - <<'PY' > gpurun_out/${TAG}_ro.txt 2>&1
import time, random, numpy as np, reef_b200, workloads as WL
from oracle.fields import FP
ctx = reef_b200.Context(0)
rnd = random.Random(1)
for n in (24, 768, 3072, 6144):
    e = [rnd.randrange(FP) for _ in range(n)]
    ctx.poseidon_ro(e, "fp")
    t0 = time.perf_counter(); ctx.poseidon_ro(e, "fp"); dt = time.perf_counter() - t0
    print(f"poseidon_ro n={n}: {dt*1e3:.3f} ms ({dt*1e6/((n+23)//24):.1f} us per permutation)")
PY
cat gpurun_out/${TAG}_ro.txt
