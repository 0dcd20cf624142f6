#!/usr/bin/env python
"""Top stall hot spots from `ncu -i X.ncu-rep --page source --csv [--launch-skip k --launch-count 1]` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
data = [r for r in rows[2:] if len(r) > max(iS, iSrc, iEx) and r[iS].isdigit()]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:n]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols if c < len(r)), reverse=True)[:2]
    print("%5d %6s %8s  %-72s %s" % (i, r[iS], r[iEx], r[iSrc].strip()[:72], st))
