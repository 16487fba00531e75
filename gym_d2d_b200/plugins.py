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

import enum
import functools
import math
from typing import Any, Dict, Tuple

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


class ShadowingPathLoss(PathLoss):
    """path_loss.py:69-81: log-distance plus gauss(0, chi_dB) beyond d0_m, drawn at every evaluation.  The CUDA path draws from
    a counter-based Philox stream (seeded by VecD2DEnv's seed), so it matches the reference in distribution only."""
    kernel_enum = _lib.PL_SHADOWING

    def __init__(self, carrier_freq_GHz: float, ple: float = 2.0, d0_m: float = 100.0, chi_dB: float = 2.7) -> None:
        super().__init__(carrier_freq_GHz)
        self.ple, self.d0_m, self.chi_dB = float(ple), float(d0_m), float(chi_dB)


class AreaType(enum.Enum):
    """path_loss.py:84-87."""
    RURAL = 0
    SUBURBAN = 1
    URBAN = 2


class CostHataPathLoss(PathLoss):
    """path_loss.py:90-123: Lb = 46.3 + 33.9 log10(f) - 13.82 log10(h_tx) - a(h_rx, f) + (44.9 - 6.55 log10(h_tx)) log10(d_km) + C.

    With one transmitter antenna height (every transmitter of the uplink + sidelink step path is a UserEquipment) this is
    a log-distance law: exponent ple = (44.9 - 6.55 log10(h_tx)) / 10 and a constant that depends on the RECEIVER's height.
    cost_hata_terms() folds both on the host; the kernels run their general-exponent path."""
    kernel_enum = _lib.PL_COST_HATA

    def __init__(self, carrier_freq_GHz: float, area_type: AreaType = AreaType.SUBURBAN) -> None:
        super().__init__(carrier_freq_GHz)
        self.area_type = area_type


def cost_hata_terms(carrier_freq_GHz: float, area_type: int, h_tx_m: float, h_rx_m: float) -> Tuple[float, float]:
    """-> (ple, K_dB) with PL = 10 ple log10(d_m) + K_dB, restating path_loss.py:100-123 for one (h_tx, h_rx) pair."""
    f = carrier_freq_GHz * 1000.0                                             # :100 MHz
    if area_type == AreaType.URBAN.value:                                     # :116-120
        a_hc = 8.29 * math.log10(1.54 * h_rx_m) ** 2 - 1.1 if f >= 200 else 3.2 * math.log10(11.75 * h_rx_m) ** 2 - 4.97
    else:
        a_hc = (1.1 * math.log10(f) - 0.7) * h_rx_m - (1.56 * math.log10(f) - 0.8)   # :122
    c = 3.0 if area_type == AreaType.URBAN.value else 0.0                     # :107
    b = 44.9 - 6.55 * math.log10(h_tx_m)
    a = 46.3 + 33.9 * math.log10(f) - 13.82 * math.log10(h_tx_m) - a_hc + c   # :108 without the distance term
    return b / 10.0, a - 3.0 * b                                              # log10(d_km) = log10(d_m) - 3


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


class ShannonRewardFunction(RewardFunction):
    """envs/reward_fn.py:47-57: per agent log2(1 + 10^(sinr/10)) if sinr >= min_sinr else -1."""
    kernel_enum = _lib.REWARD_SHANNON

    def __init__(self, min_sinr: float = -70.0) -> None:
        self.min_sinr = float(min_sinr)


class CueSinrShannonRewardFunction(RewardFunction):
    """envs/reward_fn.py:60-78: per agent -1 if another action on its RB is a cellular link with sinr < threshold."""
    kernel_enum = _lib.REWARD_CUE_SINR_SHANNON

    def __init__(self, sinr_threshold_dB: float = 0.0) -> None:
        self.sinr_threshold_dB = float(sinr_threshold_dB)


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
    'path_loss': {'LogDistancePathLoss': LogDistancePathLoss, 'FreeSpacePathLoss': FreeSpacePathLoss,
                  'CostHataPathLoss': CostHataPathLoss, 'ShadowingPathLoss': ShadowingPathLoss},
    'obs_fn': {'LinearObsFunction': LinearObsFunction},
    'reward_fn': {'SystemCapacityRewardFunction': SystemCapacityRewardFunction, 'ShannonRewardFunction': ShannonRewardFunction,
                  'CueSinrShannonRewardFunction': CueSinrShannonRewardFunction},
}
_KNOWN_UNSUPPORTED: set = set()


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
    """-> (d2d_path_loss_model enum, path-loss exponent; for CostHata the AreaType value instead of the exponent)."""
    cls, kwargs = _resolve('path_loss', obj)
    if cls is CostHataPathLoss:
        extra = set(kwargs) - {'area_type'}
        if extra:
            raise UnsupportedPluginError(f'unsupported CostHataPathLoss arguments {sorted(extra)}')
        area = kwargs.get('area_type', AreaType.SUBURBAN)                     # path_loss.py:91 default
        return cls.kernel_enum, float(getattr(area, 'value', area))          # ours or the reference's AreaType member
    extra = set(kwargs) - ({'ple', 'd0_m', 'chi_dB'} if cls is ShadowingPathLoss else {'ple'})
    if extra:
        raise UnsupportedPluginError(f'unsupported path-loss arguments {sorted(extra)}')
    if cls is FreeSpacePathLoss:
        if 'ple' in kwargs and float(kwargs['ple']) != 2.0:
            raise UnsupportedPluginError('FreeSpacePathLoss has a fixed exponent of 2')
        return cls.kernel_enum, 2.0
    return cls.kernel_enum, float(kwargs.get('ple', 2.0))   # path_loss.py:43 default


def shadowing_params(obj: Any) -> Tuple[float, float]:
    """-> (d0_m, chi_dB) of a ShadowingPathLoss plugin (path_loss.py:70 defaults), or (0, 0) for the other models."""
    cls, kwargs = _resolve('path_loss', obj)
    if cls is not ShadowingPathLoss:
        return 0.0, 0.0
    return float(kwargs.get('d0_m', 100.0)), float(kwargs.get('chi_dB', 2.7))


def resolve_obs_fn(obj: Any) -> int:
    cls, kwargs = _resolve('obs_fn', obj)
    if kwargs:
        raise UnsupportedPluginError(f'unsupported obs_fn arguments {sorted(kwargs)}')
    return cls.kernel_enum


_REWARD_PARAM: Dict[type, Tuple[str, float]] = {
    SystemCapacityRewardFunction: ('min_capacity_mbps', 0.0),       # envs/reward_fn.py:23
    ShannonRewardFunction: ('min_sinr', -70.0),                     # envs/reward_fn.py:48
    CueSinrShannonRewardFunction: ('sinr_threshold_dB', 0.0),       # envs/reward_fn.py:61
}


def resolve_reward_fn(obj: Any) -> Tuple[int, float]:
    """-> (d2d_reward_fn enum, the class's one parameter: min_capacity_mbps / min_sinr / sinr_threshold_dB)."""
    cls, kwargs = _resolve('reward_fn', obj)
    name, default = _REWARD_PARAM[cls]
    extra = set(kwargs) - {name}
    if extra:
        raise UnsupportedPluginError(f'unsupported reward_fn arguments {sorted(extra)}')
    return cls.kernel_enum, float(kwargs.get(name, default))
