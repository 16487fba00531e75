# round 2, GPU call 8: E = 4096 with fewer, fuller warps + tickets; sizes around the ticket threshold; the shape switch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for G in 0 1036 888 740 683 592 512 444 296; do
echo "== E=4096 grid cap $G (0 = default)"; D2D_B200_GRID=$G timeout 120 python profiles/time_step.py 4096 60
done
for E in 4608 5120 5632; do echo "== E=$E on / off"; timeout 120 python profiles/time_step.py $E 40; D2D_B200_TICKET=0 timeout 120 python profiles/time_step.py $E 40; done
echo "== E=131072 default"; timeout 120 python profiles/time_step.py 131072 30
echo "== E=2048 / 3072 grid caps"; for G in 0 512 256; do D2D_B200_GRID=$G timeout 120 python profiles/time_step.py 2048 60; done; for G in 0 512 384; do D2D_B200_GRID=$G timeout 120 python profiles/time_step.py 3072 60; done
} 2>&1 | grep -v "^$" | tee gpurun_out/r02_ab8.log
