# round 2, GPU call 49: default-ordering steps (wait before the first input load) on reduced grids
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 1024 2048 4096 8192 16384; do
  for G in 0 888 592 444 296 148; do
    echo "== fresh E=$E GRID=$G"; D2D_B200_GRID=$G timeout 120 python profiles/time_step.py $E 40 fresh
  done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab49.log
