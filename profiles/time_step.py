"""Quick device-time probe: `python profiles/time_step.py E [iters] [dense] [fresh]` -> us per fused step (CUDA events, graph of 32)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gym_d2d_b200 as G  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 20
dense = 'dense' in sys.argv
cfg = dict(num_rbs=100, num_cues=100, num_due_pairs=500, path_loss_model=G.FreeSpacePathLoss) if dense else {}
env = G.VecD2DEnv(E, cfg, device='cuda', seed=0)
env.reset()
ring = 32 if E <= 16384 and not dense else (8 if not dense else 4)
B = 32 * env.num_links + 8 * env.num_devices + 5
acts = [env.sample_actions() for _ in range(ring)]
outs = [env.alloc_outputs() for _ in range(ring)]
for a, o in zip(acts, outs):
    env.step(a, out=o)
g = env.capture_steps(acts, outs, inputs_stable='fresh' not in sys.argv)      # 'fresh': every step orders itself the default way
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(iters):
    g.replay()
t.record()
torch.cuda.synchronize()
us = s.elapsed_time(t) * 1e3 / (iters * ring)
print(f'E={E} {us:.2f} us/step  {E / us * 1e6:.3e} env-steps/s  frac={B * E / us / 1e3 / 6546.2:.3f}  geom={env.step_geometry()} rescues/env-step={env.stats()["rescues"] / max(1.0, env.stats()["env_steps"]):.4f}')
