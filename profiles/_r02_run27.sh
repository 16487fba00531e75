# round 2, GPU call 27: per-warp tickets against the late wait beyond one wave
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 2048 5632 6144 8192 12288 16384 32768; do
  echo "== E=$E tickets (default) / late wait (D2D_B200_TICKET=0)"
  timeout 120 python profiles/time_step.py $E 20
  D2D_B200_TICKET=0 timeout 120 python profiles/time_step.py $E 20
done
} 2>&1 | grep -v "^$" | cut -c1-120 | tee gpurun_out/r02_ab27.log
