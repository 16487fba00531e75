cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
for lib in main wpb2 wpb1; do
  if [ $lib = main ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/$lib.so; fi
  for E in 1024 4096 16384; do echo "== $lib E=$E"; timeout 120 python profiles/time_step.py $E 20 | cut -c1-110; done
done
done
