#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list: python tools/launch_sums.py list.csv [divisor]
(divisor = number of steps x ranks the list covers, to get time per rank and step)."""
import collections, csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = rows[0]
k, v = h.index("Kernel Name"), h.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[k].split("(")[0].replace("void ", ""), []).append(float(r[v]) / 1e3)
tot = 0
for n, l in d.items():
    print(f"{n:40s} n={len(l):5d} sum/div={sum(l)/div:9.1f} us  mean={sum(l)/len(l):8.1f}")
    if not n.startswith(("k_generate", "k_prepare", "k_key")):
        tot += sum(l)
print("total (without store build) / div = %.1f us" % (tot / div))
