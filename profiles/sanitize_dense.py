"""Workload for compute-sanitizer (memcheck / racecheck) on the dense kernel: few blocks stepping many envs each (the
multi-buffered tables rotate), crowded RBs (overflow list + its extra barrier), absent agents, FULL and info instantiations,
the shadowing (general-topology) kernel.  D2D_B200_GRID=2 python profiles/sanitize_dense.py"""
import functools
import sys

import torch

sys.path.insert(0, '.')
import gym_d2d_b200 as G  # noqa: E402

for kw, E, info in [(dict(num_rbs=4, num_cues=40, num_due_pairs=60), 24, True), (dict(num_rbs=4, num_cues=40, num_due_pairs=60), 24, False),
                    (dict(num_rbs=100, num_cues=100, num_due_pairs=500), 8, False),
                    (dict(num_rbs=3, num_cues=4, num_due_pairs=5, path_loss_model=functools.partial(G.ShadowingPathLoss, d0_m=5.0)), 16, True)]:
    env = G.VecD2DEnv(E, dict(kw), device='cuda', seed=1, info=info)
    env.reset()
    for s in range(3):
        a = env.sample_actions()
        a[::3, ::4] = -1
        if s == 1:
            a[::2] = a[::2] % 21                 # everybody on RB 0
        env.step(a)
    torch.cuda.synchronize()
    print('ok', E, info, env.step_geometry(), env.stats()['env_steps'])
    env.close()
