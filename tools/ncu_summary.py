#!/usr/bin/env python
"""Turn an .ncu-rep (raw page) into a small markdown table of the metrics the roofline argument uses."""
import csv
import subprocess
import sys

rep, out_md, title = sys.argv[1], sys.argv[2], sys.argv[3]
KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_dim_x",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
lines = [f"# {title}\n"]
for r in rows[2:]:
    lines.append(f"## `{r[idx['Kernel Name']][:140]}`\n")
    lines.append("| metric | value |\n|---|---|")
    for k in KEEP:
        if k in idx:
            lines.append(f"| `{k}` | {r[idx[k]]} {units[idx[k]]} |")
    if "dram__bytes_read.sum" in idx:
        def gb(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        t = gb(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
            gb(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        lines.append(f"| **DRAM traffic (read+write) per launch** | {t / 1e9:.3f} GB |")
    lines.append("")
open(out_md, "w").write("\n".join(lines))
print("\n".join(lines))
