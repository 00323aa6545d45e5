#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-ip}
python bench.py --steps 5 --also=cfg3,cfg4 --no-commit --no-openings --msm-large-log2 0 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('verified', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['config']['verified'][:60]); print([(a['workload'][:5], a['ms_per_step'], a['verified'][:30]) for a in d['also']])
    elif 'rror' in line or 'Traceback' in line: print(line)
"
L="--steps 5 --no-cpu-baseline --also= --no-commit --no-openings --msm-large-log2 0"
run() { name=$1; shift; env "$@" python bench.py $L 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$name', 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
"; }
run full X=1
run skiplast REEF_BENCH_SKIP_MSM=last
run skipall REEF_BENCH_SKIP_MSM=1
