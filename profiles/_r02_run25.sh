# round 2, GPU call 25: late wait (per-link outputs stored ahead of griddepcontrol.wait when the predecessor's buffers are disjoint)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "late_wait or ticket or ordering or pdl or graph" 2>&1 | tail -4
{
for E in 148 1024 2048 4096 5120; do
  echo "== E=$E late wait on / off"
  timeout 120 python profiles/time_step.py $E 20
  D2D_B200_LATE_WAIT=0 timeout 120 python profiles/time_step.py $E 20
done
echo "== E=131072 (tickets: unaffected)"; timeout 120 python profiles/time_step.py 131072 10
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab25.log
