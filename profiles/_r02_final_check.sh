# round 2, final check of the committed build: gpu tier, smoke, the driver's bench command
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02d_tests_final.log 2>&1
tail -4 gpurun_out/r02d_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02g_driver_args.json 2> gpurun_out/bench_r02g_driver_args.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02g_driver_args.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','frac','sm_mhz','reasons')}) for k,v in d.items() if k in ('value','ms_per_step','roofline','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api','clocks','gpu_launches')})
PY
