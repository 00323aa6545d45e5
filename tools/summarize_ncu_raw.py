#!/usr/bin/env python3
"""Condense `ncu -i X.ncu-rep --page raw --csv` (one row per captured launch) into the handful of metrics the
roofline discussion uses.  usage: summarize_ncu_raw.py <raw.csv> <out.csv>"""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe_fma_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
]

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = []
for name, short in COLS:
    if name in hdr:
        i = hdr.index(name)
        idx.append((i, short + ("_" + units[i].replace("/", "_per_") if units[i] and short in ("time", "dram_rd", "dram_wr") else "")))
out = [",".join(s for _, s in idx)]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    vals = []
    for i, s in idx:
        v = r[i]
        if s == "kernel":
            v = '"' + v.split("(")[0].replace("void ", "").replace("reef::", "") + '"'
        vals.append(v.replace(",", ""))
    out.append(",".join(vals))
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
