cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for lib in main earlywait; do
  if [ $lib = main ]; then unset D2D_B200_LIB; else export D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/$lib.so; fi
  for E in 4096 131072; do echo "== $lib E=$E"; timeout 120 python profiles/time_step.py $E 20 | cut -c1-100; done
done
done
