# round 2, GPU call 53: pipelined dense kernel (full / empty mbarriers over three bin buffers): parity tests and timing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
D2D_B200_DENSE_PIPE=1 timeout 900 python -m pytest tests -m gpu -q -x -k "dense or config3 or spec or oracle or properties" 2>&1 | tail -5
{
echo "== barrier kernel"; timeout 200 python profiles/time_step.py 65536 8 dense
echo "== pipe kernel"; D2D_B200_DENSE_PIPE=1 timeout 200 python profiles/time_step.py 65536 8 dense
echo "== pipe kernel, generic"; D2D_B200_SPEC=0 D2D_B200_DENSE_PIPE=1 timeout 200 python profiles/time_step.py 65536 8 dense
} 2>&1 | grep -v "^$" | cut -c1-200 | tee gpurun_out/r02_ab53.log
