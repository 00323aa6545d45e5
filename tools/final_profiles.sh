#!/bin/bash
# final-code evidence: per-launch list of the row commitment and of one timed bench step (run under gpurun)
mkdir -p gpurun_out
bash tools/rows_launches.sh > gpurun_out/r2z_rows_launches_after.txt 2>&1
cat gpurun_out/r2z_rows_launches_after.txt
LIGHT="--no-cpu-baseline --also= --msm-large-log2 0 --no-commit --no-openings"
REEF_RESERVE_SMS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2z_launches.csv \
    python bench.py --steps 1 --warmup 3 $LIGHT > gpurun_out/r2z_ncu_launch.log 2>&1
python tools/summarize_launches.py gpurun_out/r2z_launches.csv gpurun_out/r2z_launch_summary.csv
head -24 gpurun_out/r2z_launch_summary.csv
rm -f gpurun_out/r2z_launches.csv
