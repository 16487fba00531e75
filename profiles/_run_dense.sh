cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in ${VARIANTS:-256}; do echo "== D2D_B200_DENSE=$v"; D2D_B200_DENSE=$v timeout 300 python profiles/time_step.py 65536 5 dense; done
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio --clock-control none -k regex:d2d_step_dense -s 2 -c 1 python profiles/prof_step.py 16384 4 dense 2>&1 | grep -E "inst_executed|time_duration|issue_active|barrier" | tail -5
