"""One K-step chain, timed like bench.py's headline: `python profiles/time_chain.py E [K]` -> us per step of ONE replay of a CUDA graph of
K step kernels behind a device-side sleep (median of 15) - the pipeline's fill and drain count, unlike profiles/time_step.py's steady state."""
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import gym_d2d_b200 as G  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = G.VecD2DEnv(E, {}, device='cuda', seed=0)
env.reset()
acts = [env.sample_actions() for _ in range(32)]
outs = [env.alloc_outputs() for _ in range(32)]
for a, o in zip(acts, outs):
    env.step(a, out=o, inputs_stable=True)
g = env.capture_steps(acts[:K], outs[:K], inputs_stable=True)
g.replay()
torch.cuda.synchronize()
ts = []
for _ in range(15):
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(600_000)
    s.record()
    g.replay()
    t.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(t) * 1e3 / K)
print(f'E={E} K={K}: {statistics.median(ts):.2f} us/step (min {min(ts):.2f} max {max(ts):.2f})')
