import numpy as np, torch, sys
sys.path.insert(0, '.')
import gym_d2d_b200 as G
for kw, E in [({}, 300), (dict(num_rbs=1, num_cues=20, num_due_pairs=30), 64), (dict(num_rbs=3, num_cues=4, num_due_pairs=5), 200),
              (dict(num_rbs=16, num_cues=25, num_due_pairs=40), 40), (dict(num_rbs=100, num_cues=100, num_due_pairs=500), 6),
              (dict(reward_fn=G.CueSinrShannonRewardFunction), 100)]:
    env = G.VecD2DEnv(E, dict(kw), device='cuda', seed=1, info=True)
    env.reset()
    for _ in range(3):
        a = env.sample_actions()
        a[::3, ::4] = -1
        env.step(a)
    acts = torch.stack([env.sample_actions() for _ in range(4)]).contiguous()
    env.step_many(acts)
    torch.cuda.synchronize()
    print('ok', kw, env.stats()['env_steps'])
    env.close()
