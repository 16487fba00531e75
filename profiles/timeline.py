"""Phase timeline of one warp (debug build with -DD2D_TIMELINE): SM-clock deltas between phase marks."""
import ctypes, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ['D2D_B200_LIB'] = str(Path(__file__).resolve().parent.parent / 'gym_d2d_b200/_variants/lib_timeline.so')
import torch
import gym_d2d_b200 as G
from gym_d2d_b200 import _lib
E = int(sys.argv[1]) if len(sys.argv) > 1 else 128
env = G.VecD2DEnv(E, {}, device='cuda')
env.reset()
acts = [env.sample_actions() for _ in range(8)]
names = ['start', 'prologue done', 'pdl wait done', 'inputs arrived', 'decode+rank issued', 'sorted by RB', 'walks done', 'epilogue done',
         'reward done', 'stores issued', 'loop exit']
lib = ctypes.CDLL(os.environ['D2D_B200_LIB'])
for rep in range(4):
    for a in acts:
        env.step(a)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    lib.d2d_debug_timeline(buf)
    t = list(buf)
    print(f'rep {rep} (E={E}): ' + ', '.join(f'{names[i]} +{t[i] - t[i - 1]}' for i in range(1, 11)) + f'  total {t[10] - t[0]} cycles')
