# round 2, GPU call 51: e2e with cached host io descriptors; e2e at half the batch (is the host loop or the link the limit?)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "host" 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --dict-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']; print('e2e', e['value'], e['window_ms'], e['d2h_gbs_per_gpu'], e['host_link_peak']['d2h_gbs_per_gpu'])"
done 2>&1 | tee gpurun_out/r02_ab51.log
