#!/bin/bash
L="--steps 5 --no-cpu-baseline --also= --no-commit --no-openings --msm-large-log2 0"
run() { name=$1; shift; env "$@" python bench.py $L 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$name', 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
"; }
for i in 1 2; do
run polite0 REEF_MSM_POLITE=0
run polite1 REEF_MSM_POLITE=1
run polite0_skiplast REEF_MSM_POLITE=0 REEF_BENCH_SKIP_MSM=last
run polite1_skiplast REEF_MSM_POLITE=1 REEF_BENCH_SKIP_MSM=last
done
python bench.py --steps 3 --also= --no-commit --no-openings --msm-large-log2 0 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('verified', d['ms_per_step'], d['config']['verified'][:50])
    elif 'rror' in line: print(line)
"
