#!/usr/bin/env python
"""Export the handful of ncu metrics the round summaries quote from a .ncu-rep into a small CSV
(profiles/ keeps these; the .ncu-rep files stay in gpurun_out/).  usage: ncu_summary.py report.ncu-rep out.csv"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
with open(sys.argv[2], "w") as f:
    w = csv.writer(f)
    for r in rows[2:]:
        w.writerow(["Kernel Name", "", r[h.index("Kernel Name")]])
        for i, n in enumerate(h):
            if n in WANT or ("issue_stalled" in n and "per_issue_active" in n and "not_issued" not in n and float(r[i] or 0) > 0.05):
                w.writerow([n, units[i], r[i]])
