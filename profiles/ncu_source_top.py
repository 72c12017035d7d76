#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top source lines / SASS by stall samples and by
instructions executed.  Usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K | python ncu_source_top.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]
si, sa, ie = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[hi + 1:]:
    try:
        data.append((float(r[sa]), float(r[ie]), r[si][:100], r))
    except (ValueError, IndexError):
        continue
ts, ti = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"total samples {ts:.0f}  warp-instructions {ti:.0f}")
agg = {}
for i, c in stall_cols:
    agg[c] = sum(float(d[3][i] or 0) for d in data)
print("stall mix:", ", ".join(f"{c[6:]} {100 * v / max(1, sum(agg.values())):.0f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
for d in sorted(data, key=lambda d: -d[0])[:n]:
    top = max(stall_cols, key=lambda ic: float(d[3][ic[0]] or 0))[1][6:]
    print(f"{100 * d[0] / ts:5.1f}% smp  {100 * d[1] / ti:5.1f}% inst  [{top:>10}]  {d[2]}")
