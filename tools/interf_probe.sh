#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-ip}
python -m pytest tests/test_gpu_msm.py -m gpu -x -q -s --timeout 900 2>&1 | grep -E "partition|passed|failed|Error" | tail -5
L="--steps 5 --no-cpu-baseline --also= --no-commit --no-openings --msm-large-log2 0"
run() { name=$1; shift; env "$@" python bench.py $L 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$name', 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], json.dumps(d['kernel_ms_per_step']))
"; }
run reserve12          REEF_RESERVE_SMS=12 > gpurun_out/${TAG}.txt
run reserve0           REEF_RESERVE_SMS=0 >> gpurun_out/${TAG}.txt
run reserve20          REEF_RESERVE_SMS=20 >> gpurun_out/${TAG}.txt
run reserve12_b        REEF_RESERVE_SMS=12 >> gpurun_out/${TAG}.txt
run reserve12_skiplast REEF_RESERVE_SMS=12 REEF_BENCH_SKIP_MSM=last >> gpurun_out/${TAG}.txt
cat gpurun_out/${TAG}.txt
python bench.py --steps 5 --also= --no-commit --no-openings --msm-large-log2 0 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('verified', d['ms_per_step'], d['config']['verified'][:80])
    elif 'rror' in line: print(line)
"
