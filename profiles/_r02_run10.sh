# round 2, GPU call 10: reset kernel with one instruction stream; dense fused walk A/B
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_tests10.log 2>&1
tail -5 gpurun_out/r02_tests10.log
{
echo "== reset"; timeout 300 python profiles/time_reset.py 131072; timeout 300 python profiles/time_reset.py 65536 dense
echo "== dense main / fused walk"; timeout 300 python profiles/time_step.py 65536 5 dense; D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/fusedwalk.so timeout 300 python profiles/time_step.py 65536 5 dense
D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/fusedwalk.so timeout 600 python -m pytest tests -m gpu -q -k "dense or config3" 2>&1 | tail -3
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab10.log
