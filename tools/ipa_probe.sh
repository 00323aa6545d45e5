#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_snark.py tests/test_gpu_nifs.py tests/test_gpu_msm.py -m gpu -x -q --timeout 900 2>&1 | tail -5
python bench.py --steps 3 --also= --no-commit --msm-large-log2 0 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('ms_per_step', d['ms_per_step'], json.dumps(d['openings'])[:1500])
    elif 'rror' in line: print(line)
"
