# round 2, GPU call 28: per-warp tickets against the late wait at large batches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 65536 131072 262144; do
  echo "== E=$E tickets (default) / late wait (D2D_B200_TICKET=0)"
  timeout 120 python profiles/time_step.py $E 10
  D2D_B200_TICKET=0 timeout 120 python profiles/time_step.py $E 10
done
} 2>&1 | grep -v "^$" | cut -c1-120 | tee gpurun_out/r02_ab28.log
