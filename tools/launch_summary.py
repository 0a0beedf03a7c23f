#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share."""
import csv, sys, re
from collections import OrderedDict
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "")
    v = float(r[iV].replace(",", ""))
    v = v / 1e3 if r[iU] in ("nsecond", "ns") else (v * 1e3 if r[iU] in ("msecond", "ms") else v)   # -> us
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-34s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %8d %12.1f %10.2f %6.1f%%" % (k, n, t, t / n, 100 * t / tot))
print("%-34s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
