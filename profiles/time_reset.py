"""Device time of the standalone reset kernel: `python profiles/time_reset.py E [dense]` -> us per d2d_reset (CUDA events)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gym_d2d_b200 as G  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
cfg = dict(num_rbs=100, num_cues=100, num_due_pairs=500, path_loss_model=G.FreeSpacePathLoss) if 'dense' in sys.argv else {}
env = G.VecD2DEnv(E, cfg, device='cuda', seed=0)
mask = torch.ones(E, dtype=torch.uint8, device='cuda')
for _ in range(3):
    env.reset(mask=mask)
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 20
s.record()
for _ in range(iters):
    env.reset(mask=mask)
t.record()
torch.cuda.synchronize()
us = s.elapsed_time(t) * 1e3 / iters
nbytes = E * env.num_devices * 8
print(f'd2d_reset E={E} V={env.num_devices}: {us:.1f} us per reset, {nbytes / us / 1e3:.0f} GB/s of positions written ({nbytes / 1e6:.0f} MB)')
