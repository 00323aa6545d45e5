#!/bin/bash
# rows probe + parity tests + smoke + one bench line (run under gpurun)
TAG=${1:-r2x}
mkdir -p gpurun_out
python tools/rows_probe.py > gpurun_out/${TAG}_rows_probe.txt 2>&1; cat gpurun_out/${TAG}_rows_probe.txt
python -m pytest tests -m gpu -q --tb=short --timeout 900 -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -8 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for line in open("gpurun_out/${TAG}_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "commit", d["commit"]["e2e_ms"], "ipa", d["openings"]["ipa"]["ms"], d["openings"]["hyrax_prove_eval"]["ms"], "large", d["msm"]["large"]["frac"], "roofline", d["roofline"]["frac"], "verified", d["config"].get("verified"))
        for a in d["also"]:
            print(a["workload"][:5], a["value"], a["ms_per_step"], "e2e", a["e2e"]["value"], a["e2e"]["ms_per_step"], "commit", a["commit"]["e2e_ms"])
PY
