"""Generate tests/golden/reset_distribution.npz by RUNNING THE UNMODIFIED REFERENCE's position sampler.

    python tests/golden/gen_reset_distribution.py        (build container: /root/reference must exist)

The device-side reset (d2d_reset / d2d_episode) draws from a counter-based Philox stream while the reference draws
from Python's global Mersenne Twister (position.py:18-45), so value parity is impossible: what is pinned here is the
DISTRIBUTION of the reference's own sampler - quantile tables of

  cue_r2       (r / R)^2 of get_random_position(R)                          (uniform on [0, 1] if uniform in the disc)
  cue_theta    atan2(y, x) / (2 pi) mod 1 of the same draws                 (uniform on [0, 1])
  off_r2       (|rx - tx| / d)^2 of get_random_position_nearby(R, tx, d) for transmitters with |tx| <= R - d
  off_theta    direction of rx - tx for the same pairs
  edge_dr      |rx| - |tx| for transmitters in the edge band |tx| > R - d, where the in-cell rejection loop
               (position.py:38-44) skews the offset inwards
  edge_off_r2  (|rx - tx| / d)^2 for the same edge pairs

for the default cell (R = 500 m, d = 20 m).  Both the oracle's restatement of the product's draw scheme and the CUDA
kernels are tested against these tables (two-sample Kolmogorov-Smirnov bound).
"""
from __future__ import annotations

import random
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import ref_runner as R  # noqa: E402

LEVELS = (np.arange(2000) + 0.5) / 2000.0


def main() -> None:
    assert R.import_reference() is not None, 'reference not available'
    from gym_d2d.position import get_random_position, get_random_position_nearby
    random.seed(20261017)
    Rc, d, n = 500.0, 20.0, 200_000
    cue = np.array([get_random_position(Rc).as_tuple() for _ in range(n)])
    inner, edge = [], []
    while len(inner) < n or len(edge) < n:
        tx = get_random_position(Rc)
        band = (tx.x ** 2 + tx.y ** 2) ** 0.5 > Rc - d
        if band and len(edge) < n:
            edge.append((tx.as_tuple(), get_random_position_nearby(Rc, tx, d).as_tuple()))
        elif not band and len(inner) < n:
            inner.append((tx.as_tuple(), get_random_position_nearby(Rc, tx, d).as_tuple()))
    inner, edge = np.array(inner), np.array(edge)

    def q(x):
        return np.quantile(np.asarray(x, np.float64), LEVELS)

    off_i = inner[:, 1] - inner[:, 0]
    off_e = edge[:, 1] - edge[:, 0]
    out = dict(levels=LEVELS, cell_radius_m=Rc, d2d_radius_m=d, samples=n,
               cue_r2=q((cue ** 2).sum(-1) / Rc ** 2),
               cue_theta=q(np.mod(np.arctan2(cue[:, 1], cue[:, 0]) / (2 * np.pi), 1.0)),
               off_r2=q((off_i ** 2).sum(-1) / d ** 2),
               off_theta=q(np.mod(np.arctan2(off_i[:, 1], off_i[:, 0]) / (2 * np.pi), 1.0)),
               edge_dr=q(np.sqrt((edge[:, 1] ** 2).sum(-1)) - np.sqrt((edge[:, 0] ** 2).sum(-1))),
               edge_off_r2=q((off_e ** 2).sum(-1) / d ** 2),
               edge_fraction=float(1.0 - ((Rc - d) / Rc) ** 2))
    np.savez_compressed(HERE / 'reset_distribution.npz', **out)
    print({k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items()})


if __name__ == '__main__':
    main()
