#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the last N launches."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
last_n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
recs = [(r[ki], float(r[vi].replace(',', '')), r[ui]) for r in rows[hdr + 1:] if len(r) > vi]
if last_n: recs = recs[-last_n:]
agg = collections.OrderedDict()
for k, v, u in recs:
    k = k.split('(')[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
scale = 1e3 if recs[0][2] == 'ns' else 1.0
for k, (c, v) in agg.items():
    print("%-28s %3d %10.1f us %5.1f%%" % (k, c, v / scale, 100 * v / tot))
print("total %.1f us over %d launches" % (tot / scale, len(recs)))
