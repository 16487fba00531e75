"""compute-sanitizer workload for the late-wait grid: chains of flagged steps over rotating output sets, on few blocks (so that every
warp steps many envs and the wait sits behind its whole range), broken by resets; plus a fused episode and a rollout."""
import sys
import torch
sys.path.insert(0, '.')
import gym_d2d_b200 as G
for kw, E in [({}, 300), ({}, 37), (dict(num_rbs=3, num_cues=4, num_due_pairs=5), 1200), (dict(num_rbs=16, num_cues=25, num_due_pairs=32), 150)]:
    env = G.VecD2DEnv(E, dict(kw), device='cuda', seed=2)
    env.reset()
    acts = [env.sample_actions() for _ in range(3)]
    outs = [env.alloc_outputs() for _ in range(3)]
    for it in range(3):
        for k in range(5):
            env.step(acts[k % 3], out=outs[k % 3], inputs_stable=True)
        env.reset()
    g = env.capture_steps(acts, outs, inputs_stable=True)
    g.replay(); g.replay()
    o = env.alloc_many_outputs(4)
    env.episode(3, out=o)
    torch.cuda.synchronize()
    print('ok', kw, E, env.step_geometry(), env.stats()['env_steps'])
    env.close()
