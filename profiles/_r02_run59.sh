# round 2, GPU call 59: int16 action upload: parity test and the e2e number at one GPU
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "host or int16 or dict" 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --dict-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']; print('e2e', e['value'], e['window_ms'], e['h2d_bytes_per_step'], e['d2h_gbs_per_gpu'], e['host_link_peak']['d2h_gbs_per_gpu'])"
done 2>&1 | tee gpurun_out/r02_ab59.log
