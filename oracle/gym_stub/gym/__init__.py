"""Minimal stand-in for the `gym` package.  TEST INFRASTRUCTURE ONLY.

The image has neither `gym` nor `gymnasium` and no network, so the unmodified reference
(`import gym` at gym_d2d/__init__.py:1, envs/d2d_env.py:5-6, envs/obs_fn.py:4) cannot be
imported.  This stub supplies exactly the surface those lines touch so that the reference
can be *run* as the parity oracle and CPU baseline.  It carries no physics and is never
imported by the product package.
"""
from . import spaces  # noqa: F401
from .spaces import Space  # noqa: F401
from .envs import registration as _registration


class Env:
    metadata = {}
    observation_space = None
    action_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError


def make(id, **kwargs):  # noqa: A002 - gym's own signature
    return _registration.make(id, **kwargs)
