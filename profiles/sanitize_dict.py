"""compute-sanitizer workload for the dict API (D2DEnv through the pinned slot buffers) and the pipelined host steps."""
import sys
import numpy as np
sys.path.insert(0, '.')
import gym_d2d_b200 as G
env = G.D2DEnv({}, seed=4)
obs = env.reset()
keys = list(obs)
rng = np.random.default_rng(0)
for it in range(6):
    sub = keys if it % 2 == 0 else [keys[i] for i in rng.permutation(len(keys))[:20]]
    o, r, d, info = env.step({k: int(rng.integers(0, 500)) for k in sub})
env.step({'mbs:cue00': 3, keys[0]: 1})          # downlink: the env is rebuilt on the general-topology kernel
env.close()
vec = G.VecD2DEnv(500, {}, device='cuda', seed=1)
vec.reset()
slots = [vec.host_slot_buffers(k) for k in range(4)]
for i in range(12):
    if i >= 4:
        vec.step_host_wait(i % 4)
    slots[i % 4]['actions'][:] = rng.integers(0, 500, size=(500, vec.num_links), dtype=np.int32)
    vec.step_host_async(slots[i % 4]['actions'], slots[i % 4], i % 4)
for k in range(4):
    vec.step_host_wait(k)
vec.close()
print('ok')
