# round 2, GPU call 54: ncu counters of the pipelined dense kernel
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
D2D_B200_DENSE_PIPE=1 timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:d2d_step_dense -s 2 -c 1 -o gpurun_out/r02_densepipe python profiles/prof_step.py 65536 4 dense > /dev/null 2>&1
python profiles/ncu_summary.py gpurun_out/r02_densepipe.ncu-rep 65536 > gpurun_out/ncu_r02_densepipe.txt 2>&1
cat gpurun_out/ncu_r02_densepipe.txt
python profiles/ncu_lines.py gpurun_out/r02_densepipe.ncu-rep 65536 > gpurun_out/r02_densepipe_lines.txt 2>&1
