#!/usr/bin/env python
"""Top stall sites of one kernel from an .ncu-rep (source page, needs -lineinfo + --import-source on).
usage: ncu_stalls.py report.ncu-rep [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
h = rows[hdr]
data = [r for r in rows[hdr + 1:] if len(r) == len(h)]
ia, isamp = h.index('Source'), h.index('# Samples')
stall = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
tot = sum(int(r[isamp]) for r in data)
print(f'{rep}: {len(data)} SASS instructions, {tot} samples')
agg = {}
for r in data:
    for i in stall:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print('  by reason: ' + ', '.join(f'{k[6:]} {v * 100 // max(tot, 1)}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for idx, r in sorted(enumerate(data), key=lambda x: -int(x[1][isamp]))[:top]:
    st = {h[i][6:]: int(r[i]) for i in stall if int(r[i]) > 0}
    print(f'  #{idx:5d} {int(r[isamp]) * 100 / max(tot, 1):5.1f}%  {r[ia].strip()[:60]:60s} {st}')
