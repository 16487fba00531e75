# round 2, final check of the committed build (int16 upload in the e2e leg): gpu tier, smoke, the driver's bench command, the defaults
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02e_tests_final.log 2>&1
grep -E "passed|failed" gpurun_out/r02e_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02h_driver_args.json 2> gpurun_out/bench_r02h_driver_args.err
timeout 900 python bench.py > gpurun_out/bench_r02h_default.json 2> gpurun_out/bench_r02h_default.err
python - <<'PY'
import json
for f in ('bench_r02h_driver_args', 'bench_r02h_default'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','frac','h2d_bytes_per_step')}) for k,v in d.items() if k in ('value','ms_per_step','roofline','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api')})
PY
