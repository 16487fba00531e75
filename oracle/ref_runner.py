"""Drive the UNMODIFIED reference (davidcotton/gym-d2d) on given positions and actions.

TEST INFRASTRUCTURE ONLY.  The reference is imported from, in order, ``$D2D_REFERENCE_SRC``,
``/root/reference/src`` (build container only) or ``baseline/_ref`` (the git-ignored
``pip install --target`` copy that travels to the GPU box), behind the ~60-line ``gym`` stand-in
under ``oracle/gym_stub`` (the image has no gym).  Used to (a) pin the C oracle, (b) generate the
golden fixtures in tests/golden, (c) time the reference's own ``env.step`` loop as the CPU baseline.
"""
from __future__ import annotations

import importlib
import os
import sys
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

_HERE = Path(__file__).resolve().parent
_ROOT = _HERE.parent
_ref = None


def reference_src() -> Optional[Path]:
    cands = [os.environ.get('D2D_REFERENCE_SRC'), '/root/reference/src', str(_ROOT / 'baseline' / '_ref')]
    for c in cands:
        if c and (Path(c) / 'gym_d2d' / 'simulator.py').exists():
            return Path(c)
    return None


def import_reference():
    """Returns the reference's ``gym_d2d`` package, or None when no copy is reachable."""
    global _ref
    if _ref is not None:
        return _ref
    src = reference_src()
    if src is None:
        return None
    try:
        importlib.import_module('gym')
    except ImportError:
        sys.path.insert(0, str(_HERE / 'gym_stub'))
    sys.path.insert(0, str(src))
    _ref = importlib.import_module('gym_d2d')
    return _ref


def make_env(env_config: Optional[dict] = None):
    """gym.make('D2DEnv-v0', env_config=...) on the reference (gym_d2d/__init__.py:8-11)."""
    assert import_reference() is not None, 'reference not available'
    import gym
    return gym.make('D2DEnv-v0', env_config=dict(env_config or {}))


def set_positions(env, positions: np.ndarray) -> None:
    """positions (V,2) in devices.py:20-25 order, applied with device.py:82-83 set_position."""
    from gym_d2d.position import Position
    for dev, (x, y) in zip(env.simulator.devices.values(), np.asarray(positions, dtype=np.float64)):
        dev.set_position(Position(float(x), float(y)))


def link_keys(env) -> List[str]:
    """Canonical link order: CUEs then DUE pairs (envs/d2d_env.py:55-60)."""
    devs = env.simulator.devices
    return [f'{c}:mbs' for c in devs.cues.keys()] + [f'{t}:{r}' for (t, r) in devs.dues.keys()]


def step(env, actions, keys: Optional[List[str]] = None) -> Dict[str, np.ndarray]:
    """One reference env.step.  ``actions`` are ints for ``keys`` (default: all links, canonical order).
    Returns arrays in the order of ``keys`` plus the raw per-agent obs matrix (n, 6n)."""
    keys = keys if keys is not None else link_keys(env)
    raw = {k: int(a) for k, a in zip(keys, actions)}
    obs, rewards, done, info = env.step(raw)
    out = {
        'rb': np.array([info[k]['rb'] for k in keys], np.int64),
        'tx_pwr_dbm': np.array([info[k]['tx_pwr_dbm'] for k in keys], np.int64),
        'sinr_db': np.array([info[k]['sinr_db'] for k in keys]),
        'snr_db': np.array([info[k]['snr_db'] for k in keys]),
        'rate_bps': np.array([info[k]['rate_bps'] for k in keys]),
        'capacity_mbps': np.array([info[k]['capacity_mbps'] for k in keys]),
        'reward': np.array([rewards[k] for k in keys]),
        'per_agent_obs': np.stack([obs[k] for k in keys]) if keys else np.zeros((0, 0)),
        'done': bool(done['__all__']),
    }
    return out
