# round 2, GPU call 9: gpu tier with the final ticket policy; 64-register variant of the 4-warp shape; bench at the driver's arguments
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_tests9.log 2>&1
tail -6 gpurun_out/r02_tests9.log
{
echo "== E=131072 main / minb4=8"; timeout 120 python profiles/time_step.py 131072 30; D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/minb4_8.so timeout 120 python profiles/time_step.py 131072 30
echo "== E=16384 main / minb4=8"; timeout 120 python profiles/time_step.py 16384 30; D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/minb4_8.so timeout 120 python profiles/time_step.py 16384 30
echo "== episode / rollout E=131072"; timeout 300 python profiles/time_many.py 131072 10 16 episode; timeout 300 python profiles/time_many.py 131072 10 16 rollout
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab9.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench9.json 2> gpurun_out/r02_bench9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench9.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','window_ms','launch')}) for k,v in d.items() if k in ('value','ms_per_step','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api')})
PY
tail -3 gpurun_out/r02_bench9.err
