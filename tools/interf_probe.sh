#!/bin/bash
L="--steps 5 --no-cpu-baseline --also= --no-commit --no-openings --msm-large-log2 0"
run() { name=$1; shift; env "$@" python bench.py $L 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$name', 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
"; }
for i in 1 2 3; do
run base X=1
run both REEF_MSM_POLITE=1 REEF_RESERVE_SMS=12
done
run base_skiplast REEF_BENCH_SKIP_MSM=last
run both_skiplast REEF_MSM_POLITE=1 REEF_RESERVE_SMS=12 REEF_BENCH_SKIP_MSM=last
