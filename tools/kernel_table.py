#!/usr/bin/env python
"""Side-by-side per-launch times of bench.py --dump-kernels files:  tools/kernel_table.py a.json b.json ..."""
import json, sys
ds = [json.load(open(f)) for f in sys.argv[1:]]
maps = [{r["name"]: r for r in d["launches"]} for d in ds]
names = [r["name"] for r in ds[0]["launches"]]
tot = [0.0] * len(ds)
for n in names:
    row = f"{n:18s}"
    for i, m in enumerate(maps):
        r = m.get(n)
        if r is None:
            row += "      -   "
            continue
        tot[i] += r["ms"]
        row += f" {r['ms']:7.3f}"
    r = maps[0][n]
    row += f"   {r['bytes'] / r['ms'] / 1e6:7.0f} GB/s {r['flops'] / r['ms'] / 1e9:6.0f} TF/s"
    print(row)
print("total".ljust(18) + "".join(f" {t:7.3f}" for t in tot))
