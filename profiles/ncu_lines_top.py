#!/usr/bin/env python
"""Top CUDA source lines by stall samples:
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K | python profiles/ncu_lines_top.py [N]"""
import collections
import csv
import sys

n = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter(), ""])
path, hdr = "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        path = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        s, e = float(r[si] or 0), float(r[ie] or 0)
    except ValueError:
        continue
    key = (path, r[0])
    a = agg[key]
    a[0] += s
    a[1] += e
    for i, c in stall:
        try:
            a[2][c[6:]] += float(r[i] or 0)
        except ValueError:
            pass
    if r[1].strip():
        a[3] = r[1].strip()[:110]
ts = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {ts:.0f}, warp-instructions {ti:.0f}")
for (p, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    top = ", ".join(f"{k} {100 * v / max(1, a[0]):.0f}%" for k, v in a[2].most_common(3))
    print(f"{100 * a[0] / ts:5.1f}% smp {100 * a[1] / ti:5.1f}% inst  {p}:{ln:>4}  [{top}]  {a[3]}")
