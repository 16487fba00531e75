# round 2, GPU call 56: pipelined dense kernel: running counters; skipping the empty slot
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$PWD/gym_d2d_b200/_variants
D2D_B200_DENSE_PIPE=1 timeout 900 python -m pytest tests -m gpu -q -x -k "dense or config3 or spec or properties" 2>&1 | tail -3
{
echo "== barrier kernel"; timeout 200 python profiles/time_step.py 65536 8 dense
echo "== pipe kernel (skip)"; D2D_B200_DENSE_PIPE=1 timeout 200 python profiles/time_step.py 65536 8 dense
echo "== pipe kernel (no skip)"; D2D_B200_LIB=$V/pipe_noskip.so D2D_B200_DENSE_PIPE=1 timeout 200 python profiles/time_step.py 65536 8 dense
} 2>&1 | grep -v "^$" | cut -c1-200 | tee gpurun_out/r02_ab56.log
