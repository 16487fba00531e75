# round 2, GPU call 5: gpu tier; dense walk unroll A/B; reset kernel; bench at the driver's arguments
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_tests5.log 2>&1
tail -12 gpurun_out/r02_tests5.log
{
echo "== dense unroll 2 (main)"; timeout 300 python profiles/time_step.py 65536 5 dense
echo "== dense unroll 1";        D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/unroll1.so timeout 300 python profiles/time_step.py 65536 5 dense
echo "== dense unroll 4";        D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/unroll4.so timeout 300 python profiles/time_step.py 65536 5 dense
echo "== reset";                 timeout 300 python profiles/time_reset.py 131072; timeout 300 python profiles/time_reset.py 65536 dense
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab5.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench5.json 2> gpurun_out/r02_bench5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench5.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','window_ms')}) for k,v in d.items() if k in ('value','ms_per_step','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api')})
PY
tail -3 gpurun_out/r02_bench5.err
