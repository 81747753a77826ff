#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...).
usage: python tools/launch_summary.py profiles/r02_launches_bench.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).split("::")[-1]
    grid = int(r[8].strip("()").split(",")[0])
    key = (name, grid if name.startswith("split") else None)
    agg[key][0] += 1
    agg[key][1] += float(r[14]) / 1e3
total = sum(v[1] for v in agg.values())
print("all launches")
for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    g = f"{grid:5d}" if grid is not None else "    *"
    print(f"  {name:30s} grid {g}  n={n:4d}  total {us / 1e3:9.3f} ms  avg {us / n:8.1f} us  share {100 * us / total:5.1f} %")
single = {k: v for k, v in agg.items() if k[0] in ("hist_insert_kernel", "map_gather_kernel", "map_unique_fast_param_kernel") or k == ("split2_kernel", 148)}
call = sum(v[1] / v[0] for v in single.values())
print("\nsingle-call path only (split grid 148) -- shares comparable to stage_ms of the bench line")
for (name, grid), (n, us) in sorted(single.items(), key=lambda kv: -kv[1][1] / kv[1][0]):
    g = f"{grid:5d}" if grid is not None else "    *"
    print(f"  {name:30s} grid {g}  avg {us / n:8.1f} us  share of one call {100 * (us / n) / call:5.1f} %")
