"""Host cost of an eager step: `python profiles/time_eager.py [E]` -> us per VecD2DEnv.step call in a Python loop (no CUDA graph)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gym_d2d_b200 as G  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = G.VecD2DEnv(E, {}, device='cuda', seed=0)
env.reset()
acts = [env.sample_actions() for _ in range(8)]
outs = [env.alloc_outputs() for _ in range(8)]
for stable in (False, True):
    for i in range(200):
        env.step(acts[i % 8], out=outs[i % 8], inputs_stable=stable)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(5000):
        env.step(acts[i % 8], out=outs[i % 8], inputs_stable=stable)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f'E={E} inputs_stable={stable}: host {1e6 * (t1 - t0) / 5000:.2f} us per call, {1e6 * (t2 - t0) / 5000:.2f} us per step incl. the drain')
