# round 2, GPU call 32: dense kernel with BASELINE config #3 counts as immediates + compile-time cue (variant builds)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
echo "== dense main"; timeout 300 python profiles/time_step.py 65536 5 dense
echo "== dense cue0"; D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/cue0.so timeout 300 python profiles/time_step.py 65536 5 dense
echo "== dense specx + cue0"; D2D_B200_LIB=$PWD/gym_d2d_b200/_variants/specx.so timeout 300 python profiles/time_step.py 65536 5 dense
} 2>&1 | grep -v "^$" | cut -c1-120 | tee gpurun_out/r02_ab32.log
