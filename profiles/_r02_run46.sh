# round 2, GPU call 46: a 20-step chain (the driver's bench arguments): per-launch latency against steady-state rate
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 4096; do
  for K in 20 10; do
  for G in 0 148 222 296 370 444 518 592; do
    echo "== E=$E WPB=4 LATE_GRID=$G"; D2D_B200_LATE_GRID=$G timeout 120 python profiles/time_chain.py $E $K
  done
  for G in 296 444 592 740; do
    echo "== E=$E WPB=2 LATE_GRID=$G"; D2D_B200_WPB=2 D2D_B200_LATE_GRID=$G timeout 120 python profiles/time_chain.py $E $K
  done
  done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab46.log
