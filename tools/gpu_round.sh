#!/bin/bash
# One GPU round trip: parity tests, bench line, ncu launch list, ncu full captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [full]
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short --timeout 900 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
tail -c 600 gpurun_out/${TAG}_bench_cfg2.json
if [ "$2" == "full" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_sweep -c 4 -o gpurun_out/${TAG}_prof_sweep \
      python bench.py --workload cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_sweep.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_accum_first -s 8 -c 2 -o gpurun_out/${TAG}_prof_msm \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_msm.log 2>&1
  python bench.py --workload cfg4 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
  tail -c 300 gpurun_out/${TAG}_bench_cfg4.json
fi
ls -la gpurun_out | tail -12
