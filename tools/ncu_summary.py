#!/usr/bin/env python3
"""Summary CSV of an ncu --set full report for profiles/ (one row per captured launch, the columns bench.py and the docs cite):
   ncu_summary.py <report.ncu-rep> "<header comment>" > profiles/<name>.csv"""
import csv
import io
import subprocess
import sys

METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
           "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
rep, comment = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keep = [hdr.index(c) for c in ["Kernel Name", "Block Size", "Grid Size"]] + [hdr.index(m) for m in METRICS if m in hdr]
w = csv.writer(sys.stdout)
print("# " + comment)
for r in [hdr, units] + rows[2:]:
    row = [r[i] for i in keep]
    row[0] = row[0][:80]
    w.writerow(row)
