# round 2, GPU call 30: fp64 pass ahead of the per-env scalars (under the late wait it overlaps the wait)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "late_wait or ticket or shapes or episode or many or rollout" 2>&1 | tail -3
{
for E in 1024 4096 4608 16384 131072; do
  echo "== E=$E"; timeout 120 python profiles/time_step.py $E 20
done
echo "== episode / rollout E=131072"; timeout 200 python profiles/time_many.py 131072 10 episode; timeout 200 python profiles/time_many.py 131072 10 rollout
echo "== rollout E=4096"; timeout 200 python profiles/time_many.py 4096 10 rollout
} 2>&1 | grep -v "^$" | cut -c1-130 | tee gpurun_out/r02_ab30.log
