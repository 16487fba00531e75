# round 2, GPU call 60: default-ordering steps below two envs per warp on half the blocks: gpu tier, chains and policy loops
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
{
for E in 2560 4096 6144 8192; do
  for G in 0 ""; do
    if [ -z "$G" ]; then unset D2D_B200_FRESH_GRID; L=policy; else export D2D_B200_FRESH_GRID=$G; L=full; fi
    echo "== fresh chain E=$E grid=$L"; timeout 120 python profiles/time_step.py $E 40 fresh | cut -c1-60
    echo "== policy loop E=$E grid=$L"; timeout 120 python profiles/time_policy_loop.py $E | cut -c1-60
  done
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab60.log
