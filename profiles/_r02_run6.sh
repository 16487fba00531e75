# round 2, GPU call 6: per-warp tickets - the gpu tier (under a watchdog: a wrong ticket must not hang the box), then timings with / without
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "ticket or pdl" ) > gpurun_out/r02_tests6a.log 2>&1
tail -8 gpurun_out/r02_tests6a.log
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02_tests6.log 2>&1
tail -8 gpurun_out/r02_tests6.log
{
for E in 148 1024 2048 4096 16384 131072; do
echo "== E=$E tickets on / off"; timeout 120 python profiles/time_step.py $E 40; D2D_B200_TICKET=0 timeout 120 python profiles/time_step.py $E 40
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab6.log
