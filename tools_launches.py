import csv, sys, collections
lines=open(sys.argv[1]).read().splitlines()
i=[k for k,l in enumerate(lines) if l.startswith('"ID"')][0]
rows=list(csv.DictReader(lines[i:]))
skip=int(sys.argv[2]) if len(sys.argv)>2 else 0
agg=collections.OrderedDict()
for r in rows[skip:]:
    name=r['Kernel Name'].split('(')[0].replace('void <unnamed>::','').replace('<unnamed>::','')
    t=float(r['Metric Value'].replace(',',''))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(a[1] for a in agg.values())
for k,(n,t) in agg.items(): print(f"{k[:70]:70s} n={n:3d} avg={t/n/1e3:10.1f} us share={100*t/tot:5.1f}%")
