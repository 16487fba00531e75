"""Tiny driver for ncu captures: `python profiles/prof_step.py E [steps] [dense] [episode]` runs a few fused steps / episodes."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gym_d2d_b200 as G  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = dict(num_rbs=100, num_cues=100, num_due_pairs=500, path_loss_model=G.FreeSpacePathLoss) if 'dense' in sys.argv else {}
env = G.VecD2DEnv(E, cfg, device='cuda', seed=0)
env.reset()
acts = [env.sample_actions() for _ in range(4)]
outs = [env.alloc_outputs() for _ in range(4)]
torch.cuda.synchronize()
if 'episode' in sys.argv:          # d2d_episode: reset + uncounted step + 10 counted steps with on-device sampled actions per launch
    o = env.alloc_many_outputs(11)
    for i in range(steps):
        env.episode(10, out=o)
else:
    for i in range(steps):
        env.step(acts[i % 4], out=outs[i % 4], inputs_stable=True)
torch.cuda.synchronize()
print('done', E, steps, env.step_geometry(), env.stats())
