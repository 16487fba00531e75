"""Plugin *names* of the reference, resolved to kernel enums.

In the reference the `path_loss_model`, `obs_fn`, `reward_fn` and `traffic_model` entries of `env_config`
are Python CLASSES that the env instantiates and calls per link / per step (envs/env_config.py:21-22,
envs/d2d_env.py:27-28, simulator.py:58-59).  Here the arithmetic lives in sm_100a kernels, so the same
keys take marker classes with the same names; they carry no arithmetic and are only mapped to enums.
The reference's own classes are accepted by identity of module + name when they are importable.  Any
other class - a user subclass such as examples/custom_path_loss.py, or a built-in the kernels do not
implement yet - is rejected at construction: no Python callback ever runs inside a step.
"""
from __future__ import annotations

import functools
from typing import Any, Tuple

from . import _lib


class UnsupportedPluginError(ValueError):
    """Raised at construction for plugin classes the CUDA path cannot honour."""


# ---- path loss (path_loss.py) ---------------------------------------------------------------------
class PathLoss:
    """Marker base, mirrors path_loss.py:12-25 (ctor takes the carrier frequency, simulator.py:59)."""
    kernel_enum: int = -1

    def __init__(self, carrier_freq_GHz: float) -> None:
        self.carrier_freq_GHz = float(carrier_freq_GHz)


class LogDistancePathLoss(PathLoss):
    """path_loss.py:42-66: PL = 10 ple log10(d) + 10 ple log10(f) + 10 ple log10(4 pi / c)."""
    kernel_enum = _lib.PL_LOG_DISTANCE

    def __init__(self, carrier_freq_GHz: float, ple: float = 2.0) -> None:
        super().__init__(carrier_freq_GHz)
        self.ple = float(ple)


class FreeSpacePathLoss(PathLoss):
    """Free-space path loss = LogDistancePathLoss with ple = 2 (path_loss.py:43,45,51).  Not a separate class in
    the reference snapshot; its oracle is reference LogDistancePathLoss(f, ple=2.0)."""
    kernel_enum = _lib.PL_FREE_SPACE
    ple = 2.0


class ShadowingPathLoss(PathLoss):      # path_loss.py:69-81 (stochastic per evaluation) - not implemented
    pass


class CostHataPathLoss(PathLoss):       # path_loss.py:90-123 - not implemented
    pass


# ---- observation / reward (envs/obs_fn.py, envs/reward_fn.py) --------------------------------------
class ObsFunction:
    kernel_enum: int = -1


class LinearObsFunction(ObsFunction):
    """envs/obs_fn.py:35-61: per link (tx_x, tx_y, rx_x, rx_y, sinr_dB, snr_dB)."""
    kernel_enum = _lib.OBS_LINEAR


class RewardFunction:
    kernel_enum: int = -1


class SystemCapacityRewardFunction(RewardFunction):
    """envs/reward_fn.py:22-44."""
    kernel_enum = _lib.REWARD_SYSTEM_CAPACITY

    def __init__(self, min_capacity_mbps: float = 0.0) -> None:
        self.min_capacity_mbps = float(min_capacity_mbps)


class ShannonRewardFunction(RewardFunction):          # envs/reward_fn.py:47-57 - not implemented
    pass


class CueSinrShannonRewardFunction(RewardFunction):   # envs/reward_fn.py:60-78 - not implemented
    pass


# ---- traffic models (traffic_model.py): instantiated by the reference but never called (simulator.py:58,78)
class TrafficModel:
    def __init__(self, num_rbs: int) -> None:
        self.num_rbs = num_rbs


class UplinkTrafficModel(TrafficModel):
    pass


class DownlinkTrafficModel(TrafficModel):
    pass


_REFERENCE_MODULES = {
    'path_loss': ('gym_d2d.path_loss',),
    'obs_fn': ('gym_d2d.envs.obs_fn',),
    'reward_fn': ('gym_d2d.envs.reward_fn',),
}
_OURS = {
    'path_loss': {'LogDistancePathLoss': LogDistancePathLoss, 'FreeSpacePathLoss': FreeSpacePathLoss},
    'obs_fn': {'LinearObsFunction': LinearObsFunction},
    'reward_fn': {'SystemCapacityRewardFunction': SystemCapacityRewardFunction},
}
_KNOWN_UNSUPPORTED = {'ShadowingPathLoss', 'CostHataPathLoss', 'ShannonRewardFunction',
                      'CueSinrShannonRewardFunction'}


def _unwrap_partial(obj: Any) -> Tuple[Any, dict]:
    kwargs = {}
    while isinstance(obj, functools.partial):
        if obj.args:
            raise UnsupportedPluginError('functools.partial plugins may only bind keyword arguments')
        kwargs = {**obj.keywords, **kwargs}
        obj = obj.func
    return obj, kwargs


def _resolve(kind: str, obj: Any):
    """Map a plugin class (ours, or the reference's by module + name) to our marker class + kwargs."""
    cls, kwargs = _unwrap_partial(obj)
    if not isinstance(cls, type):
        raise TypeError(f'{kind} must be a class (as in the reference), got {type(obj).__name__}')
    ours = _OURS[kind]
    if cls in ours.values():
        return cls, kwargs
    if cls.__module__ in _REFERENCE_MODULES[kind] and cls.__name__ in ours:
        return ours[cls.__name__], kwargs
    if cls.__name__ in _KNOWN_UNSUPPORTED:
        raise UnsupportedPluginError(
            f'{kind} {cls.__name__} is a reference built-in that the CUDA path does not implement yet; '
            f'supported: {sorted(ours)}')
    raise UnsupportedPluginError(
        f'{kind} {cls.__module__}.{cls.__qualname__} is a custom Python plugin: the batched CUDA path cannot call '
        f'Python per link/step and rejects it at construction; supported: {sorted(ours)}')


def resolve_path_loss(obj: Any) -> Tuple[int, float]:
    """-> (d2d_path_loss_model enum, path-loss exponent)."""
    cls, kwargs = _resolve('path_loss', obj)
    extra = set(kwargs) - {'ple'}
    if extra:
        raise UnsupportedPluginError(f'unsupported path-loss arguments {sorted(extra)}')
    if cls is FreeSpacePathLoss:
        if 'ple' in kwargs and float(kwargs['ple']) != 2.0:
            raise UnsupportedPluginError('FreeSpacePathLoss has a fixed exponent of 2')
        return cls.kernel_enum, 2.0
    return cls.kernel_enum, float(kwargs.get('ple', 2.0))   # path_loss.py:43 default


def resolve_obs_fn(obj: Any) -> int:
    cls, kwargs = _resolve('obs_fn', obj)
    if kwargs:
        raise UnsupportedPluginError(f'unsupported obs_fn arguments {sorted(kwargs)}')
    return cls.kernel_enum


def resolve_reward_fn(obj: Any) -> Tuple[int, float]:
    """-> (d2d_reward_fn enum, min_capacity_mbps)."""
    cls, kwargs = _resolve('reward_fn', obj)
    extra = set(kwargs) - {'min_capacity_mbps'}
    if extra:
        raise UnsupportedPluginError(f'unsupported reward_fn arguments {sorted(extra)}')
    return cls.kernel_enum, float(kwargs.get('min_capacity_mbps', 0.0))   # envs/reward_fn.py:23 default
