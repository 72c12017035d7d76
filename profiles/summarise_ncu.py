#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` into a short per-kernel table (markdown).
Usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python profiles/summarise_ncu.py <hands>"""
import csv
import sys

hands = float(sys.argv[1]) if len(sys.argv) > 1 else None
rows = list(csv.reader(sys.stdin))
h = rows[0]
units = rows[1]
want = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "warp-inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]
idx = [(h.index(k), lab, units[h.index(k)]) for k, lab in want if k in h]
ki = h.index("Kernel Name")
print("| kernel | " + " | ".join(f"{lab} ({u})" if u else lab for _, lab, u in idx) + (" | warp-inst/hand |" if hands else " |"))
print("|---|" + "---|" * (len(idx) + (1 if hands else 0)))
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("mb::<unnamed>::", "").replace("void ", "")
    cells = []
    for i, lab, _ in idx:
        try:
            v = float(r[i])
            cells.append(f"{v:.4g}")
        except ValueError:
            cells.append(r[i])
    extra = ""
    if hands:
        try:
            extra = f" {float(r[h.index('smsp__inst_executed.sum')]) / hands:.0f} |"
        except (ValueError, KeyError):
            extra = " |"
    print(f"| {name} | " + " | ".join(cells) + " |" + extra)
