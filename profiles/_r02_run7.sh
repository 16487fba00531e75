# round 2, GPU call 7: ticket threshold and launch shape at large batches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 6144 8192 12288 32768 65536; do
echo "== E=$E tickets from 1 env per warp / off"; D2D_B200_TICKET_MIN=1 timeout 120 python profiles/time_step.py $E 40; D2D_B200_TICKET=0 timeout 120 python profiles/time_step.py $E 40
done
for E in 32768 65536 131072 262144; do
echo "== E=$E WPB=4 tickets / WPB=8 tickets"; D2D_B200_WPB=4 timeout 120 python profiles/time_step.py $E 30;  D2D_B200_WPB=8 timeout 120 python profiles/time_step.py $E 30
done
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab7.log
