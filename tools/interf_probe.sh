#!/bin/bash
python bench.py --steps 5 --no-cpu-baseline --also=cfg4 --no-commit --no-openings --msm-large-log2 0 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline'].get('isolated'))
    elif 'rror' in line or 'Trace' in line: print(line)
"
