#!/usr/bin/env python
"""Selected metrics of every kernel launch in an .ncu-rep as one column per launch (read here, no GPU).
usage: python tools/ncu_table.py report.ncu-rep [extra metric ...]"""
import csv, io, subprocess, sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size"]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
ki = h.index("Kernel Name")
print("%-78s %s" % ("Kernel Name", [r[ki].split("(")[0].replace("aur::<", "")[-44:] for r in data]))
for m in METRICS + sys.argv[2:]:
    if m in h:
        i = h.index(m)
        print("%-78s %s" % (m + (" [" + units[i] + "]" if units[i] else ""), [r[i] for r in data]))
