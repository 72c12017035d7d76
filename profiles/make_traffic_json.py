#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv | python profiles/make_traffic_json.py <hands> "<source note>"` ->
per-kernel DRAM bytes / L2 sectors / warp instructions of one launch each (the file bench.py's `roofline.traffic` reads)."""
import csv
import json
import sys

hands = int(sys.argv[1])
rows = list(csv.reader(sys.stdin))
h = rows[0]
col = {k: h.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                               "lts__t_sectors.sum", "smsp__inst_executed.sum")}
units = rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6, "sector": 1.0, "inst": 1.0, "": 1.0}
out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mb::<unnamed>::", "").replace("unnamed>::", "").strip()
    name = name.split("<")[0]                                 # template arguments off: one entry per kernel family
    f = lambda k: float(r[col[k]]) * scale.get(units[col[k]], 1.0)
    out[name] = {"time_ms": f("gpu__time_duration.sum"), "dram_read_bytes": f("dram__bytes_read.sum"),
                 "dram_write_bytes": f("dram__bytes_write.sum"), "hands": hands, "lts_sectors": f("lts__t_sectors.sum"),
                 "warp_inst": f("smsp__inst_executed.sum")}
    # utilisation of the units that can bound a kernel (percent of peak), when the capture has them
    for key, metric in (("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        ("l1_data_pipe_tensor_operand_pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                        ("l1_data_pipe_lsu_pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"),
                        ("l2_pct", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed")):
        if metric in h:
            try:
                out[name][key] = float(r[h.index(metric)])
            except ValueError:
                pass
json.dump({"source": sys.argv[2] if len(sys.argv) > 2 else "", "units": {"dram_*": "bytes per launch", "time_ms": "cold-cache serialised ncu time"},
           "kernels": out}, sys.stdout, indent=1)
print()
