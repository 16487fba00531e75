# round 2, 8-GPU call after the late-wait grid: the bench at N = 8 exactly as the driver launches it
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02h_8gpu.json 2> gpurun_out/bench_r02h_8gpu.err
tail -c 3000 gpurun_out/bench_r02h_8gpu.err | tail -5
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02h_8gpu.json') if l.startswith('{')][-1])
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','stats_allreduces','allreduce_world','frac')}) for k,v in d.items() if k in ('value','n_gpus','ms_per_step','e2e','fused_rollout','large_batch','episode_loop','dense_cell','roofline')})
PY
