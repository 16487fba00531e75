"""Device time of d2d_step_many: `python profiles/time_many.py E T [iters]` -> us per env-step-batch (CUDA events)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gym_d2d_b200 as G  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
env = G.VecD2DEnv(E, {}, device='cuda', seed=0)
env.reset()
ring = max(2, min(16, (256 << 20) // (E * T * 1405) + 1))      # > 126 MB of outputs in flight: per-step I/O never sits in L2
acts = [torch.stack([env.sample_actions() for _ in range(T)]).contiguous() for _ in range(ring)]
outs = [env.alloc_many_outputs(T) for _ in range(ring)]
mode = 'episode' if 'episode' in sys.argv else 'rollout' if 'rollout' in sys.argv else 'many'
if mode == 'episode':          # T counted steps + reset + the uncounted reset step per launch: T + 1 slices
    outs = [env.alloc_many_outputs(T + 1) for _ in range(ring)]


def launch(i):
    if mode == 'episode':
        env.episode(T, out=outs[i % ring])
    elif mode == 'rollout':
        env.rollout(T, action_seed=3, first_step_index=1 + i * T, out=outs[i % ring], inputs_stable=True)
    else:
        env.step_many(acts[i % ring], outs[i % ring], inputs_stable=True)


for i in range(ring):
    launch(i)
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(iters):
    launch(i)
t.record()
torch.cuda.synchronize()
us = s.elapsed_time(t) * 1e3 / (iters * T)
B = 32 * env.num_links + 5 + 8 * env.num_devices / T
print(f'{mode} E={E} T={T} ring={ring} {us:.2f} us/step  {E / us * 1e6:.3e} env-steps/s  frac={B * E / us / 1e3 / 6546.2:.3f} (bytes/env-step {B:.0f})')
