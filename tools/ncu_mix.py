#!/usr/bin/env python
"""Dynamic instruction mix of one kernel from `ncu --page source --csv` (executed warp-level
instructions by opcode): python tools/ncu_mix.py src.csv"""
import collections
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iexec = hdr.index("Source"), hdr.index("Instructions Executed")
c = collections.Counter()
for r in rows[2:]:
    if len(r) > iexec and r[iexec].isdigit():
        t = r[isrc].split()
        op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
        c[op.split(".")[0]] += int(r[iexec])
tot = sum(c.values())
print("total", tot)
for k, v in c.most_common(28):
    print(f"{k:10s} {v:12d} {100*v/tot:5.1f}%")
