#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, block, grid) count, total, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = (r[4].split('(')[0].replace('void ', ''), r[7], r[8])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
print("launches %d  total %.1f us" % (len(rows), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %-13s %-14s n=%4d %10.1f us %5.1f%% avg %8.1f us" % (k[0][:34], k[1], k[2], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
