"""Prints a few raw metrics of the last kernel in an .ncu-rep (run where ncu is available)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
                        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__cluster_size"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
for k, un, x in zip(h, u, v):
    if k in want:
        print(f"{k} [{un}] = {x}")
