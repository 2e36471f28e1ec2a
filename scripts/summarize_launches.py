#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = r['Kernel Name']
    name = re.sub(r'\(.*$', '', name)
    val = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}.get(unit, 1e-3)
    rows.append((name, val * scale))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
tot = sum(v for _, v in rows)
agg = collections.OrderedDict()
for n, v in rows:
    c, s = agg.get(n, (0, 0.0))
    agg[n] = (c + 1, s + v)
print('launches %d  total %.1f us' % (len(rows), tot))
for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%6.2f%%  %9.1f us  x%-4d avg %8.1f us  %s' % (100 * s / tot, s, c, s / c, n[:110]))
