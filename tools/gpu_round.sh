#!/bin/bash
# One GPU round trip: parity tests, bench line, ncu launch list, ncu full captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [full]
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short --timeout 900 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
if [ "$2" == "full" ]; then
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  # launch list of one timed step (3 warm-up steps precede it); per-launch times are cold-cache
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also '' --msm-large-log2 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
  # full capture of every sweep launch of one pass (2 folds x 11 launches), after the warm-up passes
  ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 66 -c 22 -o /tmp/${TAG}_prof_sweep \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also '' --msm-large-log2 0 > gpurun_out/${TAG}_ncu_sweep.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_accum_first -s 8 -c 2 -o /tmp/${TAG}_prof_msm \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also '' --msm-large-log2 0 > gpurun_out/${TAG}_ncu_msm.log 2>&1
  # reports stay on the box (gpurun_out is capped at 64 MiB): bring back the raw-metric CSV pages
  ncu -i /tmp/${TAG}_prof_sweep.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_sweep_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_prof_msm.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_msm_raw.csv 2>/dev/null
  python tools/sweep_probe.py 17 21 23 > gpurun_out/${TAG}_sweep_probe.txt 2>&1
  python tools/perf_probe.py > gpurun_out/${TAG}_perf_probe.txt 2>&1
fi
ls -la gpurun_out | tail -14
