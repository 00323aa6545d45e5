#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -c 300 gpurun_out/r2g_bench.err
python - <<'PY'
import json
for line in open("gpurun_out/r2g_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "commit", d["commit"]["e2e_ms"], "ipa", d["openings"]["ipa"]["ms"], d["openings"]["hyrax_prove_eval"]["ms"])
        for a in d["also"]:
            print(a["workload"][:5], a["value"], a["ms_per_step"], "e2e", a["e2e"]["value"], a["e2e"]["ms_per_step"], "commit", a["commit"]["e2e_ms"])
PY
