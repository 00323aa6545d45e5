#!/bin/bash
python bench.py --workload cfg3 --also= --steps 3 --msm-large-log2 0 --no-openings 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('cfg3', d['ms_per_step'], d['e2e']['ms_per_step'], json.dumps(d['commit'])[:600])
    elif 'rror' in line: print(line)
"
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2
