# round 2, GPU call 1: the whole gpu tier, then the bench at the driver's arguments
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_tests1.log 2>&1
tail -15 gpurun_out/r02_tests1.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -c 6000 gpurun_out/r02_bench1.json; tail -5 gpurun_out/r02_bench1.err
