#!/usr/bin/env python
"""Hottest SASS instructions of one kernel of an ncu report (by stall samples), with the executed
instruction count: ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv;
python tools/ncu_hot.py src.csv [top]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
tot = sum(int(r[isamp]) for r in body)
texec = sum(int(r[iexec]) for r in body)
print(f"instructions (static) {len(body)}  executed (warp-level) {texec}  samples {tot}")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for n, r in enumerate(body):
    r.append(n)
for r in sorted(body, key=lambda r: -int(r[isamp]))[:top]:
    why = sorted(((int(r[i]), hdr[i][6:]) for i in stall if r[i].isdigit() and int(r[i])), reverse=True)[:2]
    print(f"#{r[-1]:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}%  exec {int(r[iexec]):9d}  {r[isrc].strip()[:60]:60s} {why}")
