#!/usr/bin/env python
"""Attribute the executed-instruction counts of an `ncu --page source --csv` dump to CUDA source lines, using the line
table of `nvdisasm -g -c <cubin>` (the two list the SASS of a kernel in the same order).
usage: sass_lines.py <nvdisasm.txt> <ncu_source.csv> <kernel-substring> <source.cu> [units]"""
import csv, re, sys
from collections import Counter
sass, ncu, kern, srcf = sys.argv[1:5]
units = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.strip().startswith('.section') and '.text.' in l and kern in l][0]
cur, seq = None, []
for l in lines[start + 1:]:
    if l.strip().startswith('.section'):
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(srcf.split('/')[-1]) else -1
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        seq.append(cur)
rows = list(csv.reader(open(ncu)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r][0]
hdr = rows[hi]
iE, iSm = hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
data = []
for r in rows[hi + 1:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) > iE and r[iE].isdigit():
        data.append(r)
print("sass instrs", len(seq), "ncu instrs", len(data))
byline, st = Counter(), Counter()
for ln, r in zip(seq, data):
    byline[ln] += int(r[iE]); st[ln] += int(r[iSm])
tot = sum(byline.values())
src = open(srcf).read().split('\n')
for ln, v in byline.most_common(45):
    text = src[ln - 1].strip()[:100] if ln and ln > 0 else '(other file / inlined header)'
    print("%4s %5.1f%% %7.1f/unit  stalls %5d  %s" % (ln, 100 * v / tot, v / units, st[ln], text))
