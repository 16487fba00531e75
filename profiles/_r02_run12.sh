# round 2, GPU call 12: dense kernel with the fp64 passes queued and drained at the end of the kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "dense or config3 or block_min" 2>&1 | tail -4
{
echo "== dense deferred (queue + tail drain)"; timeout 300 python profiles/time_step.py 65536 5 dense
echo "== dense inline";                       D2D_B200_DEFER=0 timeout 300 python profiles/time_step.py 65536 5 dense
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab12.log
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:d2d_step_dense -s 2 -c 1 python profiles/prof_step.py 65536 4 dense 2>&1 | grep -E "inst_executed|time_duration|issue_active|barrier|long_score" | tee -a gpurun_out/r02_ab12.log
