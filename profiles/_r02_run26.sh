# round 2, GPU call 26: late wait on / off across the one-wave batch sizes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 296 512 768 1280 1536 1792 2048 2304 2560 3072 3584 4608; do
  echo "== E=$E late wait on / off"
  timeout 120 python profiles/time_step.py $E 20
  D2D_B200_LATE_WAIT=0 timeout 120 python profiles/time_step.py $E 20
done
} 2>&1 | grep -v "^$" | cut -c1-120 | tee gpurun_out/r02_ab26.log
