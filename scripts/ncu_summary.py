#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): one column per kernel launch, the metrics the
roofline discussion in DESIGN.md uses.  usage: ncu_summary.py <rep> [title]"""
import csv, io, subprocess, sys
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
names = [r[ki].split("(")[0].split("::")[-1] for r in data]
print(f"# {sys.argv[2] if len(sys.argv) > 2 else rep}\n")
print("| metric | unit | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(names))
for m in METRICS:
    if m not in hdr:
        continue
    i = hdr.index(m)
    print(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
