#!/usr/bin/env python
"""One-line-per-launch summary CSV of an .ncu-rep (`ncu -i X --page raw --csv`): the metrics DESIGN.md / bench.py quote."""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct"]
out = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    if first:
        out.writerow(["report"] + [f"{k} [{units[idx[k]]}]" if k in idx and units[idx[k]] else k for k in KEYS])
        first = False
    for r in rows[2:]:
        out.writerow([rep.split("/")[-1]] + [r[idx[k]] if k in idx else "" for k in KEYS])
