"""Aggregate an `ncu --csv` launch list per kernel (time share + instruction counts)."""
import collections
import csv
import sys

lines = open(sys.argv[1]).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[i:]))
agg = collections.OrderedDict()
for r in rows:
    name = r['Kernel Name'].split('(')[0].replace('void <unnamed>::', '').replace('<unnamed>::', '')
    if 'scheid' in name:
        continue
    v = float(r['Metric Value'].replace(',', ''))
    agg.setdefault(name, {}).setdefault(r['Metric Name'], []).append(v)
tot = sum(sum(a['gpu__time_duration.sum']) / len(a['gpu__time_duration.sum']) for a in agg.values())
for k, a in agg.items():
    t = a['gpu__time_duration.sum']
    n = a.get('smsp__inst_executed.sum', [0])
    print(f"{k[:44]:44s} n={len(t):2d} avg={sum(t)/len(t)/1e3:9.1f} us share={100*sum(t)/len(t)/tot:5.1f}% inst={sum(n)/len(n)/1e6:8.1f} M")
print(f"sum of per-kernel averages: {tot/1e3:.1f} us")
