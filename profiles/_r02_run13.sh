# round 2, GPU call 13: state check (gpu tier) + per-source-line attribution of the dense kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_tests_13.log 2>&1
tail -3 gpurun_out/r02_tests_13.log
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:d2d_step_dense -s 2 -c 1 -o gpurun_out/r02_dense13 python profiles/prof_step.py 65536 4 dense > /dev/null 2>&1
python profiles/ncu_lines.py gpurun_out/r02_dense13.ncu-rep 65536 10 > gpurun_out/r02_dense13_lines.txt 2>&1
python profiles/ncu_summary.py gpurun_out/r02_dense13.ncu-rep 65536 > gpurun_out/r02_dense13_summary.txt 2>&1
tail -5 gpurun_out/r02_dense13_lines.txt
