#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics that decide what bounds a kernel.   usage: ncu_read.py raw.csv [regex]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
pat = sys.argv[2] if len(sys.argv) > 2 else (
    r"^gpu__time_duration.sum$|smsp__issue_active.avg.pct|^smsp__inst_executed.sum$|sm__warps_active.avg.pct|"
    r"smsp__average_warps_issue_stalled.*_per_issue_active|launch__occupancy_limit|launch__registers|launch__grid_size|"
    r"sm__inst_executed_pipe_[a-z0-9_]+.avg.pct_of_peak_sustained_active$|^dram__bytes_(read|write).sum$|"
    r"l1tex__data_pipe_lsu_wavefronts.avg.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|dram__throughput.avg.pct|lts__t_sectors.avg.pct|sm__throughput.avg.pct")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:90])
    for k in hdr:
        if re.search(pat, k):
            v = d[k]
            try:
                if float(v.replace(",", "")) == 0:
                    continue
            except ValueError:
                pass
            print(f"  {k} = {v}")
