# round 2, GPU call 48: compute-sanitizer over late-wait chains on reduced grids (2- and 4-warp blocks)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for W in 2 4; do for LG in 3 0; do
  D2D_B200_WPB=$W D2D_B200_LATE_GRID=$LG timeout 400 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_late.py > gpurun_out/san_r02_late_mem_wpb${W}_g$LG.log 2>&1; echo "memcheck wpb=$W late_grid=$LG rc=$? $(grep -c '^ok' gpurun_out/san_r02_late_mem_wpb${W}_g$LG.log) workloads; $(grep 'ERROR SUMMARY' gpurun_out/san_r02_late_mem_wpb${W}_g$LG.log)"
done; done
for W in 2 4; do
  D2D_B200_WPB=$W D2D_B200_LATE_GRID=3 timeout 500 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_late.py > gpurun_out/san_r02_late_race_wpb$W.log 2>&1; echo "racecheck wpb=$W rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/san_r02_late_race_wpb$W.log)"
done
