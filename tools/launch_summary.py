#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. usage: launch_summary.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r: hdr = r; start = i + 1; break
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start:]:
    if len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d['Metric Name'] != 'gpu__time_duration.sum': continue
    v = float(d['Metric Value'].replace(',', '')); u = d['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else (v * 1e6 if u == 's' else v))
    nm = d['Kernel Name'][:60]
    agg[nm][0] += 1; agg[nm][1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]/1e3:10.3f} ms {100*a[1]/tot:5.1f}%  n={a[0]:5d}  avg {a[1]/a[0]:9.1f} us  {k}")
print(f"{tot/1e3:10.3f} ms total")
