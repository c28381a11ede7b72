#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total,
share, average (ns -> us).  usage: summarize_launches.py launches.csv [first_id last_id]"""
import collections
import csv
import re
import sys


def main():
    f = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else None
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else None
    with open(f) as fh:
        lines = [l for l in fh if not l.startswith('==')]
    rows = []
    for r in csv.DictReader(lines):
        try:
            rows.append((int(r['ID']), r['Kernel Name'], float(r['Metric Value'].replace(',', '')), r['Grid Size']))
        except Exception:
            pass
    if lo is not None:
        rows = [r for r in rows if lo <= r[0] <= (hi if hi is not None else 10 ** 9)]
    agg = collections.OrderedDict()
    for i, n, v, g in rows:
        n = re.sub(r'\(.*', '', re.sub(r'<.*', '', n))[:48]
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f'{f}: {len(rows)} launches, {tot / 1e3:.1f} us total (serialised, cold cache)')
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'  {n:48s} n={c:4d} total={t / 1e3:9.1f} us share={t / tot * 100:5.1f}% avg={t / c / 1e3:8.1f} us')


if __name__ == '__main__':
    main()
