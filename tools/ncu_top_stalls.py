#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: per kernel, the instructions with the most stall samples."""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels, cur, hdr = [], None, None
i = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
        hdr = rows[i + 1]
        i += 2
        continue
    if cur is not None and r:
        cur["rows"].append(r)
    i += 1
for k in kernels:
    h = {n: j for j, n in enumerate(hdr)}
    tot = sum(int(r[h["# Samples"]] or 0) for r in k["rows"])
    print("=" * 100)
    print(k["name"][:120], "total samples", tot)
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(int(r[h[n]] or 0) for r in k["rows"]) for n in stall_cols}
    print("  stall reasons:", ", ".join(f"{n[6:]}={v}" for n, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
    top = sorted(k["rows"], key=lambda r: -int(r[h["# Samples"]] or 0))[:topn]
    for r in top:
        reasons = sorted(((n[6:], int(r[h[n]] or 0)) for n in stall_cols), key=lambda x: -x[1])[:3]
        print(f"  {int(r[h['# Samples']]):7d} {100.0 * int(r[h['# Samples']]) / max(tot, 1):5.1f}%  {r[h['Source']].strip()[:70]:70s} "
              + " ".join(f"{a}={b}" for a, b in reasons if b))
