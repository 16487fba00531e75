"""SASS evidence for the hot kernels: `python profiles/sass_excerpt.py > profiles/sass_r02.md` (needs cuobjdump; no GPU).

For the instantiation each bench leg launches it lists the static instruction mix that matters for this path - MUFU (lg2 / ex2 / rcp
on the SFU pipe), shared-memory atomics and 128-bit shared loads / stores (the binned per-RB tables), global loads / stores, fp64
(the rare recomputation pass), local-memory spills - the programmatic-dependent-launch instructions (PREEXIT =
griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait) with the instructions around them, and confirms what is NOT there:
no tensor-core instructions (HMMA / UTCMMA) - north_star says so, DESIGN.md explains why; the bulk-copy engine (UBLKCP, tracked by
SYNCS mbarrier instructions) and cp.async (LDGSTS) appear in the dense kernel only, which stages each env's input rows with them.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / 'gym_d2d_b200' / 'libd2d_b200.so'
WANT = [
    ('configs[1] / large_batch step: d2d_step_warp_kernel<PLE2=1, EXACT=0, WPB=4, FULL=1, SPEC=1, MODE=0>', r'd2d_step_warp_kernelILb1ELb0ELi4ELb1ELb1ELi0E'),
    ('large_batch (E >= 65536): d2d_step_warp_kernel<1, 0, 8, 1, 1, 0>', r'd2d_step_warp_kernelILb1ELb0ELi8ELb1ELb1ELi0E'),
    ('episode_loop: d2d_step_warp_kernel<1, 0, 8, 1, 1, 3> (d2d_episode, drawn actions)', r'd2d_step_warp_kernelILb1ELb0ELi8ELb1ELb1ELi3E'),
    ('fused_rollout: d2d_step_warp_kernel<1, 0, 4, 1, 1, 4> (d2d_rollout)', r'd2d_step_warp_kernelILb1ELb0ELi4ELb1ELb1ELi4E'),
    ('dense_cell: d2d_step_dense_kernel<PLE2=1, LPT=2, BT=320, FULL=1, EXACT=0>', r'd2d_step_dense_kernelILb1ELi2ELi320ELb1ELb0E'),
    ('d2d_reset_kernel', r'd2d_reset_kernel'),
]
GROUPS = collections.OrderedDict([
    ('MUFU (SFU)', r'^MUFU'), ('ATOMS (shared atomics)', r'^ATOMS'), ('LDS.128', r'^LDS\.128'), ('LDS (other)', r'^LDS(?!\.128)'),
    ('STS.128', r'^STS\.128'), ('STS (other)', r'^STS(?!\.128)'), ('LDG', r'^LDG'), ('STG', r'^STG'), ('RED/ATOMG (statistics)', r'^(REDG|RED|ATOMG)'),
    ('SHFL / VOTE / REDUX / MATCH', r'^(SHFL|VOTE|REDUX|MATCH)'), ('fp64 (DFMA / DADD / DMUL)', r'^(DFMA|DADD|DMUL)'),
    ('IMAD (incl. Philox)', r'^IMAD'), ('BAR / barrier', r'^(BAR|WARPSYNC)'), ('STL / LDL (spills)', r'^(STL|LDL)'),
    ('PREEXIT (griddepcontrol.launch_dependents)', r'^PREEXIT'), ('ACQBULK (griddepcontrol.wait)', r'^ACQBULK'),
    ('tensor core (HMMA, UTCMMA)', r'^(HMMA|UTC)'),
    ('bulk copy / TMA (UBLKCP, UTMA*)', r'^(UTMA|UBLKCP)'),
    ('cp.async (LDGSTS)', r'^LDGSTS'),
    ('mbarrier (SYNCS.*)', r'^SYNCS'),
])


def main():
    out = subprocess.run(['cuobjdump', '-sass', str(LIB)], capture_output=True, text=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m and name:
            ins = m.group(2).strip()
            ins = re.sub(r'^@!?U?P\d+\s+', '', ins)
            funcs[name].append(ins)
    print('# SASS excerpt of the hot kernels (round 2)\n')
    print(f'`cuobjdump -sass {LIB.relative_to(ROOT)}` (sm_100a cubin), summarised by `profiles/sass_excerpt.py`.\n')
    for title, pat in WANT:
        hits = [f for f in funcs if re.search(pat, f)]
        if not hits:
            print(f'## {title}\n\n(not found)\n')
            continue
        f = hits[0]
        code = funcs[f]
        print(f'## {title}\n\n`{f}` - {len(code)} SASS instructions\n')
        print('| group | count |\n|---|---|')
        for g, rx in GROUPS.items():
            print(f'| {g} | {sum(1 for i in code if re.search(rx, i))} |')
        for key, what in (('PREEXIT', 'launch_dependents'), ('ACQBULK', 'wait')):
            idx = [k for k, i in enumerate(code) if i.startswith(key)]
            for k in idx[:2]:
                lo, hi = max(0, k - 4), min(len(code), k + 5)
                print(f'\n`{key}` ({what}) at instruction {k}:\n\n```')
                for q in range(lo, hi):
                    print(('>> ' if q == k else '   ') + code[q])
                print('```')
        print()


if __name__ == '__main__':
    main()
