# round 2, GPU call 2: whole gpu tier (no -x), then --set full captures with source of the three kernels to work on
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_tests2.log 2>&1
tail -25 gpurun_out/r02_tests2.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:d2d_step_dense -s 2 -c 1 -o gpurun_out/r02_dense python profiles/prof_step.py 65536 4 dense > gpurun_out/r02_ncu_dense.log 2>&1
timeout 400 $NCU -k regex:d2d_step_warp -s 3 -c 1 -o gpurun_out/r02_warp131072 python profiles/prof_step.py 131072 6 > gpurun_out/r02_ncu_warp.log 2>&1
timeout 400 $NCU -k regex:d2d_step_warp -s 2 -c 1 -o gpurun_out/r02_episode python profiles/prof_step.py 131072 2 episode > gpurun_out/r02_ncu_episode.log 2>&1
ls -la gpurun_out/*.ncu-rep
