# round 2, GPU call 24: bench line with the driver's arguments (new dense kernel, device-side sleep ahead of the timed region)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b_driver_args.json 2> gpurun_out/bench_r02b_driver_args.err ) 2>&1 | tail -3
tail -3 gpurun_out/bench_r02b_driver_args.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02b_driver_args.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','frac','sm_mhz','reasons')}) for k,v in d.items() if k in ('value','ms_per_step','roofline','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api','clocks','cpu_baseline','gpu_launches')})
PY
