#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: per-opcode executed instructions of the hot loop and the top stall sites.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME | python tools/ncu_hot.py [hot_fraction]"""
import csv, sys
from collections import Counter
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
rows = list(csv.reader(sys.stdin))
# split per kernel instance
inst, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; inst.append((r[1], cur)); hdr = None; continue
    if cur is None: continue
    if hdr is None: hdr = r; cur.append(hdr); continue
    cur.append(r)
name, tab = inst[0]
hdr = tab[0]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = [r for r in tab[1:] if len(r) > iE and r[iE].isdigit()]
tot = sum(int(r[iE]) for r in data)
n = max(int(r[iE]) for r in data)
hot = [r for r in data if int(r[iE]) > frac * n]
print(name[:80]); print("total executed", tot, "| static instrs", len(data), "| hot instrs", len(hot), "| max exec/instr", n)
c, cs = Counter(), Counter()
for r in hot:
    t = r[iS].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[iE]); cs[op] += int(r[iSm])
for k, v in c.most_common(): print("  %-10s %7.1f per iter   stall samples %d" % (k, v / n, cs[k]))
print("  sum per iteration %.1f  (hot share of all executed %.2f)" % (sum(c.values()) / n, sum(c.values()) / tot))
print("top stall sites:")
for r in sorted(data, key=lambda r: -int(r[iSm]))[:18]: print("  %6s %9s  %s" % (r[iSm], r[iE], r[iS][:100]))
