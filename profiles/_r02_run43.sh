# round 2, GPU call 43: late-grid policy in place: gpu tier, step timings across batch sizes, the latency shape on fewer blocks
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_tests_43.log 2>&1
tail -4 gpurun_out/r02_tests_43.log
{
for E in 1024 2048 2560 4096 8192 16384 32768 131072; do
  echo "== E=$E"; timeout 120 python profiles/time_step.py $E 40
done
for E in 1024 1792 2048; do
  for W in 2 4; do for G in 74 148 296; do
    echo "== E=$E WPB=$W LATE_GRID=$G"; D2D_B200_WPB=$W D2D_B200_LATE_GRID=$G timeout 120 python profiles/time_step.py $E 40
  done; done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab43.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02e_driver_args.json 2> gpurun_out/bench_r02e_driver_args.err
tail -c 1500 gpurun_out/bench_r02e_driver_args.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02e_driver_args.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','frac')}) for k,v in d.items() if k in ('value','ms_per_step','roofline','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api','clocks')})
PY
