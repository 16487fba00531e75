"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, time and share.

    python profiles/summarise_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md
"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors='replace')) if len(r) > 14 and r[12] == 'gpu__time_duration.sum']
agg = defaultdict(lambda: [0, 0.0, ''])
for r in rows:
    name = r[4].split('(')[0].replace('void ', '')[:70]
    a = agg[name]
    a[0] += 1
    a[1] += float(r[14].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[13], 1e-3)
    a[2] = f'grid {r[8]} block {r[7]}'
total = sum(v[1] for v in agg.values())
print(f'| kernel | launches | total us | avg us | share | last geometry |\n|---|---|---|---|---|---|')
for name, (n, t, geo) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{name}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / total:.1f} % | {geo} |')
print(f'\n{len(rows)} launches, {total:.1f} us of device time (ncu per-launch times are cold-cache and serialised: compare shares).')
