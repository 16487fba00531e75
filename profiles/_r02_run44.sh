# round 2, GPU call 44: the latency shape (2-warp blocks) on fewer blocks under the late wait
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "late_wait or ticket or shapes" 2>&1 | tail -3
{
for E in 148 512 1024 1792 2048; do
  for G in 0 148 296 444 592; do
    echo "== E=$E LATE_GRID=$G"; D2D_B200_LATE_GRID=$G timeout 120 python profiles/time_step.py $E 40
  done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab44.log
