"""Per-source-line executed warp-instruction counts: joins `nvdisasm -g` line info of the built cubin with the
per-SASS-instruction counts of an ncu report (`ncu -i rep --page source --csv`).

    python profiles/attribute_lines.py <nvdisasm -g -c dump of the kernel> <ncu source-page csv> <num_envs>
"""
import collections
import csv
import re
import sys

lines, cur = [], None
for line in open(sys.argv[1]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', line):
        lines.append(cur)
rows = list(csv.reader(open(sys.argv[2])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h, body = rows[hi], rows[hi + 1:]
ci = h.index('Instructions Executed')
E = float(sys.argv[3])
assert len(lines) == len(body), (len(lines), len(body))
agg = collections.Counter()
for loc, r in zip(lines, body):
    agg[loc] += (int(r[ci]) if r[ci].isdigit() else 0) / E
print(f'total warp-instructions per env: {sum(agg.values()):.1f}')
src = {}
for (f, l), c in sorted(agg.items()):
    if c >= float(sys.argv[4]) if len(sys.argv) > 4 else c >= 2.0:
        if f not in src:
            try:
                src[f] = open(f'/root/repo/gym_d2d_b200/csrc/{f}').read().split('\n')
            except OSError:
                src[f] = []
        t = src[f][l - 1].strip()[:90] if l - 1 < len(src[f]) else ''
        print(f'{f[:20]:20s} {l:4d} {c:7.1f}  {t}')
