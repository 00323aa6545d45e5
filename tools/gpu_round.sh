#!/bin/bash
# One GPU round trip: parity tests, smoke, bench line, ncu launch list, ncu full captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [full]
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short --timeout 900 > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
tail -2 gpurun_out/${TAG}_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 600 gpurun_out/${TAG}_bench.json
if [ "$2" == "full" ]; then
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  LIGHT="--no-cpu-baseline --also= --msm-large-log2 0 --no-commit --no-openings"
  # launch list of one timed step (3 warm-up steps precede it); per-launch times are cold-cache
  REEF_RESERVE_SMS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 3 $LIGHT > gpurun_out/${TAG}_ncu_launch.log 2>&1
  python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launch_summary.csv > /dev/null 2>&1
  # full captures, taken after the untimed verification + warm-up launches
  ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 66 -c 22 -o /tmp/${TAG}_prof_sweep \
      python bench.py --steps 1 --warmup 3 $LIGHT > gpurun_out/${TAG}_ncu_sweep.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_accum_first|k_accum_next|k_bitsum_partial|k_bitsum_final|k_scatter|k_hist' \
      -s 216 -c 27 -o /tmp/${TAG}_prof_msm \
      python bench.py --steps 1 --warmup 3 $LIGHT > gpurun_out/${TAG}_ncu_msm.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_round|k_tail|k_nl_begin' -s 90 -c 30 -o /tmp/${TAG}_prof_transcript \
      python bench.py --steps 1 --warmup 3 $LIGHT > gpurun_out/${TAG}_ncu_transcript.log 2>&1
  # reports stay on the box (gpurun_out is capped at 64 MiB): bring back the raw-metric CSV pages
  for k in sweep msm transcript; do
    ncu -i /tmp/${TAG}_prof_$k.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_${k}_raw.csv 2>/dev/null
    if [ "$(stat -c %s /tmp/${TAG}_prof_$k.ncu-rep 2>/dev/null || echo 99999999)" -lt 12000000 ]; then cp /tmp/${TAG}_prof_$k.ncu-rep gpurun_out/; fi
  done
  python tools/sweep_probe.py 17 21 23 > gpurun_out/${TAG}_sweep_probe.txt 2>&1
  python tools/perf_probe.py > gpurun_out/${TAG}_perf_probe.txt 2>&1
fi
ls -la gpurun_out | tail -14
