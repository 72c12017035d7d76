#!/usr/bin/env python
"""Per CUDA source line: share of the executed warp instructions and of the stall samples of one kernel.
Usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --print-source cuda,sass | python profiles/ncu_lines_by_source.py [N]"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
cur = hdr = None
inst, smp, src = collections.Counter(), collections.Counter(), {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
    elif r[0] == "Line No":
        hdr = r
    elif hdr is not None and cur and r[0] != "Function Name":
        try:
            ln, ie, sa = int(r[0]), float(r[hdr.index("Instructions Executed")]), float(r[hdr.index("# Samples")])
        except (ValueError, IndexError):
            continue
        inst[(cur, ln)] += ie
        smp[(cur, ln)] += sa
        src[(cur, ln)] = r[1].strip()[:110]
ti, ts = sum(inst.values()), sum(smp.values())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
print(f"warp instructions {ti:.0f}, stall samples {ts:.0f}\n")
print("| file:line | % instructions | % samples | source |")
print("|---|---|---|---|")
for k, _ in sorted(smp.items(), key=lambda kv: -kv[1])[:n]:
    print(f"| {k[0]}:{k[1]} | {100 * inst[k] / ti:.1f} | {100 * smp[k] / ts:.1f} | `{src[k].replace('|', '/')}` |")
