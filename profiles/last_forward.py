#!/usr/bin/env python
"""Print the kernels of the last complete forward in an ncu launch list (eager bench run):
from the last lang_embed_kernel to the following select_kernel.  usage: last_forward.py launches.csv"""
import csv
import re
import sys

f = sys.argv[1]
lines = [l for l in open(f) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r.get('Metric Name') == 'gpu__time_duration.sum']
names = [re.sub(r'\(.*', '', re.sub(r'<.*', '', r['Kernel Name'])).replace('void ', '').replace('vog::', '') for r in rows]
starts = [i for i, n in enumerate(names) if n.startswith('lang_embed')]
ends = [i for i, n in enumerate(names) if n.startswith('select_kernel')]
pairs = [(s, min(e for e in ends if e > s)) for s in starts if any(e > s for e in ends)]
s, e = pairs[-1]
tot = 0.0
out = []
for i in range(s, e + 1):
    v = float(rows[i]['Metric Value'].replace(',', ''))
    tot += v
    out.append(f"{names[i][:28]} {rows[i]['Grid Size']} {v / 1e3:.1f}")
print(f'{f}: {e - s + 1} launches, {tot / 1e3:.1f} us (serialised, cold cache)')
print(' | '.join(out))
