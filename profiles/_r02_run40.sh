# round 2, GPU call 40: E = 4096 with fewer, fuller warps UNDER THE LATE WAIT (no tickets): two launches co-resident
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for E in 4096; do
  for W in 4 8 2; do
    for G in 0 1036 888 740 592 518 512 444 342 296 256 148; do
      echo "== E=$E WPB=$W GRID=$G"; D2D_B200_WPB=$W D2D_B200_GRID=$G timeout 120 python profiles/time_step.py $E 40
    done
  done
done
for E in 2048 3072 6144 8192; do
  for G in 0 518 444 296; do
     echo "== E=$E WPB=4 GRID=$G"; D2D_B200_WPB=4 D2D_B200_GRID=$G timeout 120 python profiles/time_step.py $E 40
  done
done
} 2>&1 | grep -v "^$" | cut -c1-150 | tee gpurun_out/r02_ab40.log
