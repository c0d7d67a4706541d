#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV).  Usage: launch_summary.py launches.csv [steps]"""
import collections, csv, re, sys
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"total {tot/steps:.1f} us per step over {steps:g} steps (ncu: cold-cache, serialised launches -- compare shares)")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v/steps:10.1f} us/step {100*v/tot:5.1f}%  x{n/steps:<4g} avg {v/n:8.1f} us  {k[:70]}")
