#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-ip}
L="--steps 5 --no-cpu-baseline --also= --no-commit --no-openings --msm-large-log2 0"
run() { name=$1; shift; env "$@" python bench.py $L 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$name', 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], json.dumps(d['kernel_ms_per_step']))
"; }
: > gpurun_out/${TAG}.txt
for c in 8 32 1; do
  run full_conn$c CUDA_DEVICE_MAX_CONNECTIONS=$c >> gpurun_out/${TAG}.txt
  run skiplast_conn$c CUDA_DEVICE_MAX_CONNECTIONS=$c REEF_BENCH_SKIP_MSM=last >> gpurun_out/${TAG}.txt
  run skipall_conn$c CUDA_DEVICE_MAX_CONNECTIONS=$c REEF_BENCH_SKIP_MSM=1 >> gpurun_out/${TAG}.txt
done
cat gpurun_out/${TAG}.txt
