# round 2, GPU call 50: fused launches (rollout / step_many / episode) of one-wave batches on reduced grids
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 4096 8192; do
for M in rollout many episode; do
  for G in 0 888 592 444 296; do
    echo "== $M E=$E GRID=$G"; D2D_B200_GRID=$G timeout 120 python profiles/time_many.py $E 10 40 $M
  done
done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab50.log
