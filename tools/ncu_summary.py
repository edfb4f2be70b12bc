#!/usr/bin/env python3
"""Developer helper: selected metrics of an ncu --set full report as a CSV for profiles/ (one column per launch).
usage: ncu_summary.py REPORT.ncu-rep OUT.csv"""
import csv
import subprocess
import sys

SEL = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
       "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
       "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
SEL += ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s for s in
        ("long_scoreboard", "no_instruction", "wait", "short_scoreboard", "branch_resolving", "math_pipe_throttle", "not_selected",
         "dispatch_stall", "barrier", "lg_throttle", "membar", "drain")]

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "") for r in rows[2:]])
    for s in SEL:
        if s in hdr:
            i = hdr.index(s)
            w.writerow([s, units[i]] + [r[i] for r in rows[2:]])
print("wrote", out, len(rows) - 2, "launches")
