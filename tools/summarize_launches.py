#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
  lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
  name = re.sub(r"\(.*", "", row["Kernel Name"])
  v = float(row["Metric Value"].replace(",", ""))
  v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
  agg[name][0] += 1
  agg[name][1] += v
total = sum(v[1] for v in agg.values())
print("# %s: %d launches, %.3f ms of kernel time (ncu-serialised, cold caches: compare shares)" %
      (path, sum(v[0] for v in agg.values()), total))
print("%-58s %7s %12s %7s" % ("kernel", "count", "total ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print("%-58s %7d %12.3f %6.1f%%" % (k[:58], v[0], v[1], 100 * v[1] / total))
