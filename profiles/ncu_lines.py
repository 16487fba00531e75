"""Per-source-line view of an ncu report captured with --import-source on:
    python profiles/ncu_lines.py report.ncu-rep <num_envs> [min_instr_per_env]
prints, for every CUDA source line, warp-instructions per env, the share of stall samples and the dominant stall reasons."""
import csv
import subprocess
import sys

rep, E = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, tot_s, tot_i, agg = '', None, 0, 0.0, {}
for r in rows:
    if len(r) >= 2 and r[0] in ('File Name', 'File Path'):
        fname = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        ins = float(d.get('Instructions Executed') or 0)
        smp = int(d.get('# Samples') or 0)
        stalls = {k[6:]: int(v or 0) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k}
        tot_s += smp
        tot_i += ins
        key = (fname, int(r[0]))
        if key not in agg:
            agg[key] = [0.0, 0, {}, r[1].strip()[:70]]
        a = agg[key]
        a[0] += ins / E
        a[1] += smp
        for k, v in stalls.items():
            a[2][k] = a[2].get(k, 0) + v
print(f'total {tot_i / E:.0f} warp-instructions per env, {tot_s} samples')
for (f, ln), (ins, smp, st, text) in sorted(agg.items()):
    if ins >= thr or smp >= 0.01 * tot_s:
        top = ' '.join(f'{k}={v}' for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f'{f[:18]:18s} {ln:4d} {ins:7.0f} {100.0 * smp / max(tot_s, 1):5.1f}%  {top:45s} | {text}')
