# round 2, GPU call 22: full gpu tier + compute-sanitizer on the dense kernel with TMA-staged inputs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_tests_22.log 2>&1
tail -4 gpurun_out/r02_tests_22.log
D2D_B200_GRID=2 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_dense.py > gpurun_out/san_r02_dense_mem.log 2>&1; echo "memcheck rc=$? $(grep -c '^ok' gpurun_out/san_r02_dense_mem.log) workloads; $(grep 'ERROR SUMMARY' gpurun_out/san_r02_dense_mem.log)"
D2D_B200_GRID=2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_dense.py > gpurun_out/san_r02_dense_race.log 2>&1; echo "racecheck rc=$? $(grep -c '^ok' gpurun_out/san_r02_dense_race.log) workloads; $(grep 'RACECHECK SUMMARY' gpurun_out/san_r02_dense_race.log)"
grep -m5 -A6 "hazard\|Invalid" gpurun_out/san_r02_dense_race.log gpurun_out/san_r02_dense_mem.log | head -40
