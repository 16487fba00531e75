# round 2, GPU call 3: gpu tier, then A/B timings: dense SPEC / generic / 640-thread shape; warp kernel at 64 registers; episode / rollout
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_tests3.log 2>&1
tail -12 gpurun_out/r02_tests3.log
{
echo "== dense SPEC";            timeout 300 python profiles/time_step.py 65536 5 dense
echo "== dense generic";         D2D_B200_SPEC=0 timeout 300 python profiles/time_step.py 65536 5 dense
echo "== dense 640 x 1 SPEC";    D2D_B200_DENSE=640 timeout 300 python profiles/time_step.py 65536 5 dense
echo "== warp E=131072 main";    timeout 300 python profiles/time_step.py 131072 20
echo "== warp E=131072 minb8=4"; D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/minb8_4.so timeout 300 python profiles/time_step.py 131072 20
echo "== warp E=4096 stable / fresh"; timeout 300 python profiles/time_step.py 4096 40; timeout 300 python profiles/time_step.py 4096 40 fresh
echo "== warp E=1024 stable / fresh"; timeout 300 python profiles/time_step.py 1024 40; timeout 300 python profiles/time_step.py 1024 40 fresh
echo "== episode E=131072";      timeout 300 python profiles/time_many.py 131072 10 16 episode
echo "== rollout E=131072";      timeout 300 python profiles/time_many.py 131072 10 16 rollout
echo "== many E=131072";         timeout 300 python profiles/time_many.py 131072 10 16
echo "== rollout / many / episode E=4096"; timeout 300 python profiles/time_many.py 4096 10 200 rollout; timeout 300 python profiles/time_many.py 4096 10 200; timeout 300 python profiles/time_many.py 4096 10 200 episode
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab3.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench3.json'))
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','window_ms')}) for k,v in d.items() if k in ('value','ms_per_step','e2e','fused_rollout','large_batch','episode_loop','dense_cell','dict_api')})
PY
tail -3 gpurun_out/r02_bench3.err
