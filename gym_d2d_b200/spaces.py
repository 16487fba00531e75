"""Observation / action space descriptors.

When `gym` (or `gymnasium`) is installed its own space classes are used so that `env.action_space` is a real
gym space; the image this was built in has neither (and no network), so light stand-ins with the same
attributes and `sample()` are provided.  Only what envs/d2d_env.py:36-40,56,58 and envs/obs_fn.py:41 touch.
"""
from __future__ import annotations

import random as _random

try:  # pragma: no cover - depends on the environment
    from gym.spaces import Box, Dict, Discrete  # type: ignore  # noqa: F401
    HAVE_GYM = True
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gymnasium.spaces import Box, Dict, Discrete  # type: ignore  # noqa: F401
        HAVE_GYM = True
    except Exception:  # noqa: BLE001
        HAVE_GYM = False

        class Discrete:  # type: ignore[no-redef]
            def __init__(self, n: int) -> None:
                self.n = int(n)

            def sample(self) -> int:
                return _random.randrange(self.n)

            def contains(self, x) -> bool:
                return 0 <= int(x) < self.n

            def __repr__(self) -> str:
                return f'Discrete({self.n})'

        class Box:  # type: ignore[no-redef]
            def __init__(self, low, high, shape=None, dtype='float32') -> None:
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape or ()), dtype

            def __repr__(self) -> str:
                return f'Box({self.low}, {self.high}, {self.shape}, {self.dtype})'

        class Dict:  # type: ignore[no-redef]
            def __init__(self, spaces=None) -> None:
                self.spaces = dict(spaces or {})

            def __getitem__(self, key):
                return self.spaces[key]

            def sample(self):
                return {k: s.sample() for k, s in self.spaces.items()}

            def __repr__(self) -> str:
                return f'Dict({self.spaces})'
