# A/B harness: `bash profiles/ab.sh [lib ...]` times every listed build of libd2d_b200 (paths under gym_d2d_b200/_variants, or
# "main") at E = 4096 and 131072 and reads the instruction / issue counters of one launch with ncu.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in "$@"; do
  if [ $lib = main ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/$lib.so; fi
  for E in 4096 131072; do
    echo "== $lib E=$E"; timeout 120 python profiles/time_step.py $E 20
  done
  timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:d2d_step_warp -c 2 python profiles/time_step.py 131072 1 2>&1 | grep -E "inst_executed|time_duration|issue_active|warps_active|wavefronts" | tail -5
done 2>&1 | grep -v "^+" | tee -a gpurun_out/ab.log
