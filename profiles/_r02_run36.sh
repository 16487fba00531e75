# round 2, GPU call 36: per-source-line attribution of the SPEC dense kernel with the shared-memory constants of the fp64 pass
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:d2d_step_dense -s 2 -c 1 -o gpurun_out/r02_dense36 python profiles/prof_step.py 65536 4 dense > /dev/null 2>&1
python profiles/ncu_lines.py gpurun_out/r02_dense36.ncu-rep 65536 10 > gpurun_out/r02_dense36_lines.txt 2>&1
python profiles/ncu_summary.py gpurun_out/r02_dense36.ncu-rep 65536 > gpurun_out/ncu_r02d_step_dense.txt 2>&1
tail -3 gpurun_out/r02_dense36_lines.txt
