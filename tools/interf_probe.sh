#!/bin/bash
python bench.py --steps 5 --no-cpu-baseline --also= --no-commit --no-openings 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'large', d['msm']['large'])
"
