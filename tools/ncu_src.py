#!/usr/bin/env python
"""Per-instruction view of an `ncu --page source --csv --print-source sass` dump: top instructions by a column.
usage: ncu_src.py src.csv [column] [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = sys.argv[2] if len(sys.argv) > 2 else "L1 Wavefronts Shared"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
ix = {h: i for i, h in enumerate(hdr)}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(num(r[ix[col]]) for r in body)
tot_inst = sum(num(r[ix["Instructions Executed"]]) for r in body)
print(f"total {col} = {tot:.0f}; total warp instructions = {tot_inst:.0f}")
for r in sorted(body, key=lambda r: -num(r[ix[col]]))[:n]:
    print(f"{r[ix['Address']][-5:]} {num(r[ix[col]]):12.0f} ({100 * num(r[ix[col]]) / max(tot, 1):4.1f}%) inst {num(r[ix['Instructions Executed']]):10.0f}  "
          f"ideal {r[ix['L1 Wavefronts Shared Ideal']]:>10}  {r[ix['Source']][:90]}")
