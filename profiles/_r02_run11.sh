# round 2, GPU call 11: e2e with the kernel storing straight into the mapped pinned slot buffers vs the staged copy-out; bench timing hygiene
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --large-envs-per-gpu 0 --dense-envs-per-gpu 0 --fused-steps 0 --dict-steps 0 --skip-cpu-baseline --episodes 0"
for i in 1 2; do
echo "== staged";  timeout 300 $B | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:d['e2e'][k] for k in ('value','d2h_gbs_per_gpu','window_ms')}, d['e2e']['host_link_peak']['d2h_gbs_per_gpu'])"
echo "== direct";  D2D_B200_HOST_DIRECT=1 timeout 300 $B | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:d['e2e'][k] for k in ('value','d2h_gbs_per_gpu','window_ms')}, d['e2e']['host_link_peak']['d2h_gbs_per_gpu'])"
done 2>&1 | tee gpurun_out/r02_ab11.log
D2D_B200_HOST_DIRECT=1 timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "obs_dyn" 2>&1 | tail -3
