"""A policy-in-the-loop chain: `python profiles/time_policy_loop.py [E]` -> us per {torch kernel that rewrites the actions; default-ordering
step} pair, replayed from a CUDA graph of 32 pairs (the step cannot overlap its predecessor: what an isolated launch costs)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gym_d2d_b200 as G  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = G.VecD2DEnv(E, {}, device='cuda', seed=0)
env.reset()
acts = env.sample_actions()
outs = [env.alloc_outputs() for _ in range(4)]
nvec = env._nvec_dev.to(torch.int32)


def pair(i):
    acts.add_(1).remainder_(nvec)            # the "policy": two small torch kernels that rewrite the actions
    env.step(acts, out=outs[i % 4])          # default ordering


for i in range(8):
    pair(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
        for i in range(32):
            pair(i)
torch.cuda.current_stream().wait_stream(side)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(40):
    g.replay()
t.record()
torch.cuda.synchronize()
print(f'E={E}: {s.elapsed_time(t) * 1e3 / (40 * 32):.2f} us per (policy kernels + step)  geom={env.step_geometry()}')
