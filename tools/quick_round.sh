#!/bin/bash
# parity tests + one bench line (run under gpurun)
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short --timeout 900 -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for line in open("gpurun_out/${TAG}_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], json.dumps(d.get("openings"))[:1500])
        print(json.dumps(d.get("commit"))[:900])
        print(json.dumps(d["kernel_ms_per_step"]))
PY
