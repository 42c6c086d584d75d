#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row['Kernel Name'][:64]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    if sum(v) / tot > 0.003:
        print(f"{k:66s} n={len(v):3d} mean={sum(v)/len(v):9.1f} us  share={sum(v)/tot*100:5.1f}%")
