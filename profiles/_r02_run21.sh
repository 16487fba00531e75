# round 2, GPU calls 21+: dense kernel experiments, one change at a time
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "dense or config3 or block or overflow or crowded or golden or fixture" 2>&1 | tail -12
{
echo "== $D2D_LABEL"; timeout 300 python profiles/time_step.py 65536 5 dense
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab21.log
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:d2d_step_dense -s 2 -c 1 python profiles/prof_step.py 65536 4 dense 2>&1 | grep -E "inst_executed|time_duration|issue_active|barrier|scoreboard|wavefronts|conflicts" | tee -a gpurun_out/r02_ab21.log
if [ -n "$D2D_AB_INLINE" ]; then
echo "== inline passes (D2D_B200_DEFER=0)" | tee -a gpurun_out/r02_ab21.log
D2D_B200_DEFER=0 timeout 300 python profiles/time_step.py 65536 5 dense 2>&1 | grep -v "^$" | tee -a gpurun_out/r02_ab21.log
fi
