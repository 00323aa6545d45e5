#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -c 300 gpurun_out/r2g_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_reference.json 2>/dev/null
python - <<'PY'
import json
for line in open("gpurun_out/r2g_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "commit", d["commit"]["e2e_ms"], "ipa", d["openings"]["ipa"]["ms"], d["openings"]["hyrax_prove_eval"]["ms"], "large", d["msm"]["large"]["frac"], "roofline", d["roofline"]["frac"])
        for a in d["also"]:
            print(a["workload"][:5], a["value"], a["ms_per_step"], "e2e", a["e2e"]["value"], a["e2e"]["ms_per_step"], "commit", a["commit"]["e2e_ms"])
for line in open("gpurun_out/r2g_bench_reference.json"):
    if line.startswith("{"):
        d = json.loads(line); print("REF", d["value"], d["ms_per_step"])
PY
