#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_snark.py tests/test_gpu_witness.py tests/test_gpu_lifetime.py -m gpu -x -q --timeout 900 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 5 --also= --no-commit --msm-large-log2 0 --no-cpu-baseline 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); o=d['openings']; print('ms_per_step', d['ms_per_step'], 'ipa', o['ipa']['ms'], o['ipa']['ms_folding_generators'], 'hyrax', o['hyrax_prove_eval']['ms'])
    elif 'rror' in line: print(line)
"
done
