# round 2, third evidence pass (late-wait grid, dict API), final build of the round (1 GPU): gpu tier, launch list, --set full summaries, bench lines
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02c_tests_final.log 2>&1
tail -5 gpurun_out/r02c_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# launch list of the bench command (per-launch times under ncu are cold-cache and serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:d2d_ -c 700 --csv --log-file gpurun_out/launches_r02c.csv \
    python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-e2e --dict-steps 0 > gpurun_out/bench_under_ncu_r02c.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:d2d_step_warp -s 3 -c 1 -o gpurun_out/r02h_warp4096 python profiles/prof_step.py 4096 6 > /dev/null 2>&1
timeout 300 $NCU -k regex:d2d_step_warp -s 3 -c 1 -o gpurun_out/r02h_warp131072 python profiles/prof_step.py 131072 6 > /dev/null 2>&1
timeout 300 $NCU -k regex:d2d_step_warp -s 2 -c 1 -o gpurun_out/r02h_episode python profiles/prof_step.py 131072 2 episode > /dev/null 2>&1
timeout 300 $NCU -k regex:d2d_step_dense -s 2 -c 1 -o gpurun_out/r02h_dense python profiles/prof_step.py 65536 4 dense > /dev/null 2>&1
timeout 300 $NCU -k regex:d2d_reset -s 1 -c 1 -o gpurun_out/r02h_reset python profiles/time_reset.py 131072 > /dev/null 2>&1
for f in warp4096:4096 warp131072:131072 episode:1441792 dense:65536 reset:131072; do
  n=${f%%:*}; e=${f##*:}
  python profiles/ncu_summary.py gpurun_out/r02h_$n.ncu-rep $e > gpurun_out/ncu_r02d_$n.txt 2>&1
done
rm -f gpurun_out/r02h_warp4096.ncu-rep gpurun_out/r02h_reset.ncu-rep gpurun_out/r02h_episode.ncu-rep gpurun_out/r02h_warp131072.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02f_driver_args.json 2> gpurun_out/bench_r02f_driver_args.err
timeout 900 python bench.py > gpurun_out/bench_r02f_default.json 2> gpurun_out/bench_r02f_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02f_ref.json 2> gpurun_out/bench_r02f_ref.err
ls -la gpurun_out | tail -20
python - <<'PY'
import json
for f in ('bench_r02f_driver_args', 'bench_r02f_default'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','frac')}) for k,v in d.items() if k in ('value','ms_per_step','roofline','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api','clocks')})
PY
