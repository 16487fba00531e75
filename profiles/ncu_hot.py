"""Top stalled SASS instructions of one kernel: `python profiles/ncu_hot.py report.ncu-rep [N]` (needs -lineinfo / --import-source)."""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix['Address']], 16)
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('samples', tot, 'SASS instructions', len(data))
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:top]:
    s = int(r[ix['# Samples']])
    main = sorted(((k, int(r[ix[k]] or 0)) for k in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{int(r[ix['Address']], 16) - base:05x} {100 * s / tot:5.1f}% {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:58]:58s} {main}")
