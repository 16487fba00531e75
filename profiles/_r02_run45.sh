# round 2, GPU call 45: 2-warp blocks beyond 2048 envs under the late wait, on fewer blocks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 2560 3072 4096 6144 8192 16384; do
  for G in 148 296 444 592 888; do
    echo "== E=$E WPB=2 LATE_GRID=$G"; D2D_B200_WPB=2 D2D_B200_LATE_GRID=$G timeout 120 python profiles/time_step.py $E 40
  done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab45.log
