#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (markdown)."""
import csv
import sys
from collections import OrderedDict

src, out_md, title, command = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0,
                                          "msecond": 1e3}.get(r[ui], 1.0)
    name = r[ki].split("(")[0][:110]
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + v)
total = sum(t for _, t in agg.values())
lines = [f"# {title}\n", f"Command: `{command}`",
         "(per-launch times are cold-cache and serialised — compare SHARES, not absolutes). Raw CSV beside this file.\n",
         "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{name}` | {n} | {t:.1f} | {100 * t / total:.1f}% |")
open(out_md, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
