"""Summarise one kernel of an .ncu-rep: `python profiles/ncu_summary.py report.ncu-rep [envs]` -> the counters DESIGN.md cites."""
import csv
import subprocess
import sys

rep = sys.argv[1]
envs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[-1]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for k in want:
    if k in m:
        print(f'{k:82s} {m[k][0]:>16s} {m[k][1]}')
if envs and 'smsp__inst_executed.sum' in m:
    print(f'{"warp-instructions per env":82s} {float(m["smsp__inst_executed.sum"][0].replace(",", "")) / envs:16.1f}')
for k in sorted(m):
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
        v = float(m[k][0].replace(',', ''))
        if v >= 0.05:
            print(f'{k:82s} {v:16.2f}')
