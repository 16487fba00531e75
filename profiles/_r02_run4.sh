# round 2, GPU call 4: gpu tier, A/B: dense deferred vs inline fp64 pass; episode / rollout after Philox-7 + batched scalars; ncu of dense
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_tests4.log 2>&1
tail -12 gpurun_out/r02_tests4.log
{
echo "== dense deferred";        timeout 300 python profiles/time_step.py 65536 5 dense
echo "== dense inline";          D2D_B200_DEFER=0 timeout 300 python profiles/time_step.py 65536 5 dense
echo "== warp E=131072";         timeout 300 python profiles/time_step.py 131072 20
echo "== warp E=4096";           timeout 300 python profiles/time_step.py 4096 40
echo "== episode E=131072";      timeout 300 python profiles/time_many.py 131072 10 16 episode
echo "== rollout E=131072";      timeout 300 python profiles/time_many.py 131072 10 16 rollout
echo "== rollout / many / episode E=4096"; timeout 300 python profiles/time_many.py 4096 10 200 rollout; timeout 300 python profiles/time_many.py 4096 10 200; timeout 300 python profiles/time_many.py 4096 10 200 episode
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab4.log
timeout 400 ncu --set full --clock-control none --import-source on -f -k regex:d2d_step_dense -s 2 -c 1 -o gpurun_out/r02_dense_defer python profiles/prof_step.py 65536 4 dense > gpurun_out/r02_ncu_dense2.log 2>&1
ls -la gpurun_out/*.ncu-rep
