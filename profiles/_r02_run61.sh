# round 2, GPU call 61: late-wait steps with ONE warp per block at griddepcontrol.wait (the others wait for its shared-memory flag)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$PWD/gym_d2d_b200/_variants
D2D_B200_LIB=$V/onewaiter.so timeout 300 python -m pytest tests -m gpu -q -x -k "late_wait" 2>&1 | tail -2
{
for E in 4096 8192 16384; do
  echo "== main E=$E"; timeout 120 python profiles/time_step.py $E 40 | cut -c1-70
  echo "== one waiter E=$E"; D2D_B200_LIB=$V/onewaiter.so timeout 120 python profiles/time_step.py $E 40 | cut -c1-70
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab61.log
