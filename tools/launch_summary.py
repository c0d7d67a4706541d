#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV).
Usage: launch_summary.py launches.csv [steps]            totals over the whole list divided by `steps`
       launch_summary.py launches.csv --step MARKER [k]  only the k-th (default: last) window between two consecutive
                                                         launches of the kernel whose name contains MARKER (one step)"""
import collections, csv, re, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    rows.append((name, v))
steps = 1.0
if len(sys.argv) > 2 and sys.argv[2] == "--step":
    marks = [i for i, (n, _) in enumerate(rows) if sys.argv[3] in n]
    k = int(sys.argv[4]) if len(sys.argv) > 4 else len(marks) - 2
    rows = rows[marks[k] + 1:marks[k + 1] + 1]
    print(f"one step: launches {marks[k] + 1} .. {marks[k + 1]} of the list (window between two '{sys.argv[3]}' launches)")
elif len(sys.argv) > 2:
    steps = float(sys.argv[2])
agg, tot = collections.OrderedDict(), 0.0
for name, v in rows:
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"total {tot/steps:.1f} us of kernel time per step, {len(rows)/steps:g} launches per step "
      f"(ncu: cold-cache, serialised launches -- compare shares)")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/steps:10.1f} us/step {100*v/tot:5.1f}%  x{n/steps:<4g} avg {v/n:8.1f} us  {k[:70]}")
