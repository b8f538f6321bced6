#!/usr/bin/env python
"""Per-kernel summary of an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
h = rows[0]
k, v = h.index("Kernel Name"), h.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[k].split("(")[0].replace("void ", ""), []).append(float(r[v]) / 1e3)
print(len(rows) - 1, "launches")
for n, l in d.items():
    print(f"{n:44s} n={len(l):3d} first={l[0]:9.1f} us  median={sorted(l)[len(l)//2]:9.1f}  max={max(l):9.1f}")
