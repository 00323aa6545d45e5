#!/bin/bash
L="--steps 5 --no-cpu-baseline --also= --no-commit --no-openings --msm-large-log2 0"
run() { name=$1; shift; env "$@" python bench.py $L 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$name', 'ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], json.dumps({k:v for k,v in d['kernel_ms_per_step'].items() if k.startswith('msm')}))
"; }
for dl in 0 1 2 -1 3; do
run cdelta$dl REEF_MSM_C_DELTA=$dl
done
REEF_MSM_C_DELTA=1 python tools/pair_probe.py 2>&1 | grep "n=2^15\|n=2^14"
REEF_MSM_C_DELTA=2 python tools/pair_probe.py 2>&1 | grep "n=2^15\|n=2^14"
