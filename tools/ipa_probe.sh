#!/bin/bash
run() { name=$1; shift; env "$@" bash -c 'python bench.py --also=$BENCH_ALSO --steps 5 --no-cpu-baseline --no-commit --no-openings --msm-large-log2 0' 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$name', 'target', d['ms_per_step'], d['e2e']['ms_per_step'], [(a['workload'][:5], a['ms_per_step'], a['e2e']['ms_per_step']) for a in d['also']])
    elif 'rror' in line: print(line)
"; }
run async_all REEF_BENCH_ASYNC_UPLOAD=1 BENCH_ALSO=cfg2,cfg3,cfg4,cfg5
run sync_all REEF_BENCH_ASYNC_UPLOAD=0 BENCH_ALSO=cfg2,cfg3,cfg4,cfg5
python -m pytest tests/test_gpu_witness.py tests/test_gpu_mle.py tests/test_gpu_lifetime.py -m gpu -x -q --timeout 900 2>&1 | tail -2
