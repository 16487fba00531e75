cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in ${VARIANTS:-320}; do echo "== D2D_B200_DENSE=$v"; D2D_B200_DENSE=$v timeout 300 python profiles/time_step.py 65536 5 dense; done
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio --clock-control none -k regex:d2d_ -s 8 -c 3 python profiles/prof_step.py 65536 4 dense 2>&1 | grep -E "d2d_|inst_executed|time_duration|issue_active|barrier" | tail -16
