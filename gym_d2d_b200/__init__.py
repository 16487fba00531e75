"""gym_d2d_b200 - B200-native (sm_100a) batched implementation of GymD2D's per-step radio physics.

Drop-in surface of davidcotton/gym-d2d for the step path:

    import gym_d2d_b200 as gym_d2d
    env = gym_d2d.make('D2DEnv-v0', env_config={...})     # == gym.make(...) when gym is installed
    obs = env.reset(); obs, rewards, done, info = env.step({'cue00:mbs': 17, ...})

plus the batched tensor API `VecD2DEnv(num_envs, env_config)`.  The arithmetic runs only in the CUDA
library (gym_d2d_b200/libd2d_b200.so, C ABI in include/d2d_b200.h); there is no CPU fallback.
"""
from .config import EPISODE_LENGTH, EnvConfig  # noqa: F401
from .plugins import (AreaType, CostHataPathLoss, CueSinrShannonRewardFunction, DownlinkTrafficModel,  # noqa: F401
                      FreeSpacePathLoss, LinearObsFunction, LogDistancePathLoss, ShadowingPathLoss,
                      ShannonRewardFunction, SystemCapacityRewardFunction, UnsupportedPluginError,
                      UplinkTrafficModel)

ENV_ID = 'D2DEnv-v0'      # gym_d2d/__init__.py:8-11

__all__ = ['make', 'D2DEnv', 'VecD2DEnv', 'EnvConfig', 'ENV_ID', 'LogDistancePathLoss', 'FreeSpacePathLoss',
           'LinearObsFunction', 'SystemCapacityRewardFunction', 'UnsupportedPluginError']


def __getattr__(name):   # torch is imported only when an env class is actually requested
    if name == 'D2DEnv':
        from .d2d_env import D2DEnv
        return D2DEnv
    if name == 'VecD2DEnv':
        from .vec_env import VecD2DEnv
        return VecD2DEnv
    raise AttributeError(name)


def make(id: str = ENV_ID, **kwargs):  # noqa: A002 - gym's signature
    """gym.make('D2DEnv-v0', env_config=...) without requiring gym."""
    if id != ENV_ID:
        raise ValueError(f'unknown environment id {id!r}; this package registers {ENV_ID!r}')
    from .d2d_env import D2DEnv
    return D2DEnv(**kwargs)


def _register_with_gym() -> None:
    for mod in ('gym', 'gymnasium'):
        try:
            registration = __import__(f'{mod}.envs.registration', fromlist=['register'])
            registration.register(id=ENV_ID, entry_point='gym_d2d_b200.d2d_env:D2DEnv')
        except Exception:  # noqa: BLE001 - gym absent (this image) or id already registered
            pass


_register_with_gym()
