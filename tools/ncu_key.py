#!/usr/bin/env python
"""Print the handful of ncu metrics the profiles/ summaries quote: tools/ncu_key.py report.ncu-rep [launch index]."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "sm__throughput.avg.pct", "smsp__inst_executed.sum ",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__average_warps_issue_stalled", "launch__registers_per_thread ",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg ", "sm__cycles_elapsed.max ",
        "lts__t_bytes.sum ", "lts__t_sector_hit_rate", "launch__occupancy_limit", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed", "l1tex__t_sector_hit_rate", "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_fmalite"]
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, vals = rows[0], rows[2 + idx]
for h, v in zip(hdr, vals):
    hh = h + " "
    if any(w in hh for w in WANT) or h in ("Kernel Name",):
        if "issue_stalled" in h and float(v or 0) < 0.05:
            continue
        print(f"{h:90s} {v}")
