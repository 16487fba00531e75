# round 2, GPU call 47: dense kernel with rotating warp roles (scheduler balance), dict API through the pinned slot buffers
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$PWD/gym_d2d_b200/_variants
timeout 1200 python -m pytest tests -m gpu -q -x -k "dense or dict or D2DEnv or gym or adapter or subset or overrides or downlink or partial or config or reference" 2>&1 | tail -3
{
for lib in main norot rotskip main norot; do
  if [ $lib = main ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=$V/$lib.so; fi
  echo "== $lib dense"; timeout 200 python profiles/time_step.py 65536 8 dense
done
unset D2D_B200_LIB
echo "== main dense generic"; D2D_B200_SPEC=0 timeout 200 python profiles/time_step.py 65536 8 dense
} 2>&1 | grep -v "^$" | cut -c1-200 | tee gpurun_out/r02_ab47.log
timeout 300 ncu --metrics smsp__inst_executed.sum,smsp__inst_executed.max,smsp__inst_executed.avg,smsp__inst_executed.min,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__issue_active.max.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio --clock-control none -k regex:d2d_step_dense -s 6 -c 1 python profiles/time_step.py 65536 1 dense 2>&1 | grep -E "smsp__|gpu__time" | tee -a gpurun_out/r02_ab47.log
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('dict_api', d.get('dict_api')); print('dense', d['dense_cell']['ms_per_step'], d['dense_cell']['roofline']['frac'])" | tee -a gpurun_out/r02_ab47.log
