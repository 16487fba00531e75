# round 2, 8-GPU call: does the upload's size matter to the download rate when all eight GPUs copy at once?
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for H in 819200 409600 0; do
  T=$(python -c "import time; print(time.time() + 25)")
  for g in 0 1 2 3 4 5 6 7; do python profiles/link_probe.py $g $H 2478080 $T & done
  wait
done 2>&1 | sort | tee gpurun_out/r02_link8.log
nvidia-smi topo -m 2>&1 | head -20 >> gpurun_out/r02_link8.log
