#!/usr/bin/env python3
"""Trim an `ncu -i X.ncu-rep --page raw --csv` dump to the columns the roofline notes cite.

    ncu -i gpurun_out/r1_prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_raw_summary.py raw.csv > profiles/rNN_ncu_full_summary.csv
"""
import csv
import sys

WANT = [
    "ID", "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(w) for w in WANT if w in hdr]
out = csv.writer(sys.stdout)
out.writerow([hdr[i] for i in idx])
out.writerow([units[i] for i in idx])
for r in rows[2:]:
    out.writerow([r[i][:64] for i in idx])
