# round 2, last check of the committed build: the whole gpu tier, smoke, the driver's bench command
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02g_tests_final.log 2>&1
grep -E "passed|failed" gpurun_out/r02g_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02i_driver_args.json 2> gpurun_out/bench_r02i_driver_args.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02i_driver_args.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','vs_steps_only','frac')}) for k,v in d.items() if k in ('value','ms_per_step','roofline','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api','gpu_launches')})
PY
