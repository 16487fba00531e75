# round 2, GPU call 58: pipelined dense kernel at three blocks per SM: sleeping poll instead of a try_wait spin
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=$PWD/gym_d2d_b200/_variants
D2D_B200_DENSE_PIPE=1 timeout 900 python -m pytest tests -m gpu -q -x -k "dense or config3 or spec or properties" 2>&1 | tail -2
{
echo "== barrier kernel"; timeout 200 python profiles/time_step.py 65536 8 dense
echo "== pipe kernel sleep 100"; D2D_B200_DENSE_PIPE=1 timeout 200 python profiles/time_step.py 65536 8 dense
echo "== pipe kernel sleep 30"; D2D_B200_LIB=$V/sleep30.so D2D_B200_DENSE_PIPE=1 timeout 200 python profiles/time_step.py 65536 8 dense
echo "== pipe kernel sleep 300"; D2D_B200_LIB=$V/sleep300.so D2D_B200_DENSE_PIPE=1 timeout 200 python profiles/time_step.py 65536 8 dense
} 2>&1 | grep -v "^$" | cut -c1-120 | tee gpurun_out/r02_ab58.log
D2D_B200_DENSE_PIPE=1 timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:d2d_step_dense -s 2 -c 1 -o gpurun_out/r02_densepipe python profiles/prof_step.py 65536 4 dense > /dev/null 2>&1
python profiles/ncu_summary.py gpurun_out/r02_densepipe.ncu-rep 65536 > gpurun_out/ncu_r02_densepipe3.txt 2>&1
grep -E "duration|inst_executed.sum|issue_active|eligible|stalled|per env" gpurun_out/ncu_r02_densepipe3.txt
python profiles/ncu_lines.py gpurun_out/r02_densepipe.ncu-rep 65536 > gpurun_out/r02_densepipe3_lines.txt 2>&1
sort -k4 -n -r gpurun_out/r02_densepipe3_lines.txt | head -6 | cut -c1-170
