#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K` per CUDA source line:
share of executed warp instructions, of stall samples and of shared-memory wavefronts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = ''
items = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        def g(name):
            try:
                return float(r[hdr.index(name)].replace(',', ''))
            except (ValueError, IndexError):
                return 0.0
        items.append((cur, int(r[0]), r[1].strip()[:90], g('Instructions Executed'), g('# Samples'), g('L1 Wavefronts Shared'), g('L1 Wavefronts Shared Ideal')))
tot = sum(x[3] for x in items) or 1
tw = sum(x[4] for x in items) or 1
tl = sum(x[5] for x in items) or 1
print("total warp instructions %.0f, stall samples %.0f, shared wavefronts %.0f (ideal %.0f)" % (tot, tw, tl, sum(x[6] for x in items)))
for x in sorted(items, key=lambda x: -(x[3] / tot + x[4] / tw))[:top]:
    print("%-18s:%-4d %5.1f%% inst %5.1f%% stall %5.1f%% smem-wf (x%.1f) | %s" % (x[0], x[1], 100 * x[3] / tot, 100 * x[4] / tw, 100 * x[5] / tl, x[5] / x[6] if x[6] else 0, x[2]))
