#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_witness.py tests/test_gpu_mle.py tests/test_gpu_lifetime.py tests/test_gpu_fullsize.py -m gpu -x -q --timeout 900 2>&1 | tail -3
python bench.py --steps 5 --also=cfg5 --no-commit --no-openings --msm-large-log2 0 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['config']['verified'][:40]); print([(a['workload'][:5], a['ms_per_step'], a['e2e']['ms_per_step'], a['verified'][:20]) for a in d['also']])
    elif 'rror' in line: print(line)
"
