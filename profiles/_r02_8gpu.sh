# round 2, 8-GPU call: the device-guard test (needs two GPUs), then the bench at N = 8 exactly as the driver launches it, then N = 2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "another_device" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_r02_8gpu.err
tail -c 3000 gpurun_out/bench_r02_8gpu.err | tail -5
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_8gpu.json') if l.startswith('{')][-1])
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','us_per_step','ms_per_episode','vs_steps_only','d2h_gbs_per_gpu','stats_allreduces','allreduce_world','host_link_peak','numa')}) for k,v in d.items() if k in ('value','n_gpus','ms_per_step','e2e','fused_rollout','large_batch','episode_loop','dense_cell')})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_2gpu.json') if l.startswith('{')][-1])
print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ('value','ms_per_step','ms_per_episode','d2h_gbs_per_gpu','stats_allreduces')}) for k,v in d.items() if k in ('value','n_gpus','ms_per_step','e2e','large_batch','episode_loop','dense_cell')})
PY
