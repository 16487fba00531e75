# round 2, GPU call 23: dense kernel bounds: without the fp64 pass; reciprocal with two Newton steps
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -m gpu -q -x -k "unaligned_rows" 2>&1 | tail -2
for lib in main noresc rcp2; do
  if [ $lib = main ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/$lib.so; fi
  echo "== dense $lib"; timeout 300 python profiles/time_step.py 65536 5 dense
done
export D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/rcp2.so
timeout 900 python -m pytest tests -m gpu -q -x -k "dense or band_edge or threshold or config3" 2>&1 | tail -2
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab23.log
