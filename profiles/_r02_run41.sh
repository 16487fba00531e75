# round 2, GPU call 41: balanced env split + a smaller grid for late-wait steps (D2D_B200_LATE_GRID sweep), parity tests of the shapes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "late_wait or ticket or shapes or episode or many or rollout or oracle" 2>&1 | tail -3
{
for E in 2560 3072 4096 5632 8192 16384 32768 65536 131072; do
  for G in 0 148 296 444 518 592; do
    echo "== E=$E LATE_GRID=$G"; D2D_B200_LATE_GRID=$G timeout 120 python profiles/time_step.py $E 40
  done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab41.log
