# round 2, GPU call 29: full gpu tier + bench line with the late wait
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r02_tests_29.log 2>&1
tail -4 gpurun_out/r02_tests_29.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02d_driver_args.json 2> gpurun_out/bench_r02d_driver_args.err
tail -3 gpurun_out/bench_r02d_driver_args.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02d_driver_args.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','frac','sm_mhz','reasons')}) for k,v in d.items() if k in ('value','ms_per_step','roofline','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api','clocks','gpu_launches')})
PY
