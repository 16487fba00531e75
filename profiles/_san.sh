cd $GRAFT_REPO_ROOT
for W in 2 4 8; do
  D2D_B200_WPB=$W timeout 300 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_run.py > gpurun_out/san_mem_wpb$W.log 2>&1; echo "memcheck wpb=$W rc=$? $(grep -c '^ok' gpurun_out/san_mem_wpb$W.log) workloads; $(grep 'ERROR SUMMARY' gpurun_out/san_mem_wpb$W.log)"
done
for W in 2 4; do
  D2D_B200_WPB=$W timeout 400 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_run.py > gpurun_out/san_race_wpb$W.log 2>&1; echo "racecheck wpb=$W rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/san_race_wpb$W.log)"
done
ncu --set full --clock-control none -k regex:d2d_step_dense -s 2 -c 1 -f -o gpurun_out/prof_dense_final python profiles/prof_step.py 65536 4 dense > gpurun_out/prof_dense_final.log 2>&1
python profiles/ncu_summary.py gpurun_out/prof_dense_final.ncu-rep 65536 > gpurun_out/ncu_final_dense.txt 2>&1; rm -f gpurun_out/prof_dense_final.ncu-rep; head -3 gpurun_out/ncu_final_dense.txt
