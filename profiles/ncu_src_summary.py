#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` output per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = ''
items = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 8 and r[0].isdigit():
        try:
            items.append((cur, int(r[0]), r[1].strip()[:95], float(r[7] or 0), float(r[4] or 0)))
        except ValueError:
            pass
tot = sum(x[3] for x in items) or 1
tw = sum(x[4] for x in items) or 1
print("total warp instructions %.0f, stall samples %.0f" % (tot, tw))
for x in sorted(items, key=lambda x: -(x[3] / tot + x[4] / tw))[:top]:
    print("%-16s:%-4d %5.1f%% inst %5.1f%% stall | %s" % (x[0], x[1], 100 * x[3] / tot, 100 * x[4] / tw, x[2]))
