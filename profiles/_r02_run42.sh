# round 2, GPU call 42: where a late-wait step at 148 blocks spends its period (timeline), and what the step-counter load costs (A/B)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$PWD/gym_d2d_b200/_variants
{
for G in 148 296; do
for E in 4096 8192; do
  echo "== main E=$E LATE_GRID=$G"; D2D_B200_LATE_GRID=$G timeout 120 python profiles/time_step.py $E 40
  echo "== nocount E=$E LATE_GRID=$G"; D2D_B200_LIB=$V/nocount.so D2D_B200_LATE_GRID=$G timeout 120 python profiles/time_step.py $E 40
done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab42.log
D2D_B200_LIB=$V/timeline.so D2D_B200_LATE_GRID=148 timeout 120 python profiles/timeline.py 4096 > gpurun_out/timeline_late148_E4096.txt 2>&1
tail -14 gpurun_out/timeline_late148_E4096.txt
