#!/usr/bin/env python
"""Hottest SASS lines of a kernel from `ncu -i X.ncu-rep --page source --csv` (needs --import-source on / -lineinfo).

  ncu -i rep.ncu-rep --page source --csv > src.csv
  python profiles/hotlines.py src.csv 12,14 25      # blocks 12 and 14 (every launch appears twice in the csv), top 25 lines

Per block: total warp-stall samples, the stall-reason totals, and per line its samples, the two dominant stall reasons and
the theoretical / ideal L2 sectors (uncoalesced accesses show a ratio > 1).  Samples cover ALL warps of the CTA, so spinning
waiters (idle role warps, the final barrier) show up next to the real hot spots."""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
blocks=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur={'hdr':None,'data':[]}; blocks.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r; continue
    if len(r)==len(cur['hdr']): cur['data'].append(r)
print(len(blocks),'kernels')
which=[int(x) for x in sys.argv[2].split(',')]
ntop=int(sys.argv[3]) if len(sys.argv)>3 else 25
for b in which:
    blk=blocks[b]; hdr=blk['hdr']; ix={h:i for i,h in enumerate(hdr)}; data=blk['data']
    tot=sum(int(r[ix['# Samples']]) for r in data)
    print('== kernel', b, 'total samples', tot)
    stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg={h:sum(int(r[ix[h]]) for r in data) for h in stalls}
    print('   stall totals:', ' '.join('%s=%.1f%%'%(h[6:],100*v/max(tot,1)) for h,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8]))
    for r in sorted(data, key=lambda r:-int(r[ix['# Samples']]))[:ntop]:
        s=int(r[ix['# Samples']])
        st=sorted(((int(r[ix[h]]),h) for h in stalls), reverse=True)[:2]
        print('%6d %5.1f%%  %-64s  %s  L2sect %s/%s' % (s, 100*s/tot, r[ix['Source']].strip()[:64], ' '.join('%s=%d'%(h[6:],v) for v,h in st), r[ix['L2 Theoretical Sectors Global']], r[ix['L2 Theoretical Sectors Global Ideal']]))
