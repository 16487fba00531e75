# round 2, GPU call 52: e2e with the copy-out alternating between two streams
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for S in 1 2 1 2; do
D2D_B200_OUT_STREAMS=$S timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --dict-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']; print('out streams $S: e2e', e['value'], e['window_ms'], e['d2h_gbs_per_gpu'], e['host_link_peak']['d2h_gbs_per_gpu'])"
done 2>&1 | tee gpurun_out/r02_ab52.log
