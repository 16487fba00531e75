"""ctypes front-end of the CPU oracle (oracle/d2d_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package (gym_d2d_b200) never does.  See d2d_oracle.h for parity status.

The device defaults below restate /root/reference/src/gym_d2d/device.py:12-41 independently of the
product's own tables (gym_d2d_b200/config.py) so that the two can be checked against each other.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Optional

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / '_build' / 'libd2d_oracle.so'
_lib = None


class _Device(C.Structure):
    _fields_ = [('is_bs', C.c_int32), ('_pad', C.c_int32),
                ('tx_antenna_gain_dBi', C.c_double), ('rx_antenna_gain_dBi', C.c_double),
                ('thermal_noise_dBm', C.c_double), ('noise_figure_dB', C.c_double),
                ('sinr_dB', C.c_double), ('ix_margin_dB', C.c_double),
                ('body_loss_dB', C.c_double), ('cable_loss_dB', C.c_double),
                ('masthead_amplifier_gain_dB', C.c_double),
                ('num_subcarriers', C.c_double), ('subcarrier_spacing_kHz', C.c_double),
                ('antenna_height_m', C.c_double)]


class _Cfg(C.Structure):
    _fields_ = [('num_rbs', C.c_int32), ('num_cues', C.c_int32), ('num_due_pairs', C.c_int32),
                ('n_pwr_cue', C.c_int32), ('n_pwr_due', C.c_int32), ('path_loss_model', C.c_int32),
                ('carrier_freq_GHz', C.c_double), ('ple', C.c_double), ('min_capacity_mbps', C.c_double),
                ('area_type', C.c_int32), ('downlinks', C.c_int32), ('n_pwr_mbs', C.c_int32), ('_pad', C.c_int32),
                ('shadow_d0_m', C.c_double), ('shadow_chi_dB', C.c_double), ('rng_seed', C.c_uint64),
                ('first_global_env', C.c_uint64), ('rng_step', C.c_uint64)]


# device.py:12-41 (merged with DEFAULT_DEVICE_CONFIG :12-16)
UE_DEFAULTS = dict(is_bs=0, tx_antenna_gain_dBi=0.0, rx_antenna_gain_dBi=0.0, thermal_noise_dBm=-104.5,
                   noise_figure_dB=7.0, sinr_dB=-10.0, ix_margin_dB=3.0, body_loss_dB=3.0,
                   cable_loss_dB=0.0, masthead_amplifier_gain_dB=0.0,
                   num_subcarriers=12, subcarrier_spacing_kHz=15.0, antenna_height_m=1.5)
BS_DEFAULTS = dict(is_bs=1, tx_antenna_gain_dBi=17.5, rx_antenna_gain_dBi=17.5, thermal_noise_dBm=-118.4,
                   noise_figure_dB=2.0, sinr_dB=-7.0, ix_margin_dB=2.0, body_loss_dB=0.0,
                   cable_loss_dB=2.0, masthead_amplifier_gain_dB=2.0,
                   num_subcarriers=12, subcarrier_spacing_kHz=15.0, antenna_height_m=23.0)


@dataclass
class OracleConfig:
    """The hot-path subset of envs/env_config.py:12-27 plus the decode moduli of envs/d2d_env.py:31-35."""
    num_rbs: int = 25
    num_cues: int = 25
    num_due_pairs: int = 25
    cell_radius_m: float = 500.0
    d2d_radius_m: float = 20.0
    due_min_tx_power_dBm: int = 0
    due_max_tx_power_dBm: int = 20
    cue_max_tx_power_dBm: int = 23
    carrier_freq_GHz: float = 2.1
    num_subcarriers: int = 12
    subcarrier_spacing_kHz: int = 15
    ple: float = 2.0
    min_capacity_mbps: float = 0.0
    mbs_max_tx_power_dBm: int = 46
    downlinks: bool = False                    # append the DOWNLINK links 'mbs:cueXX' (envs/d2d_env.py:87-89): N = 2C + D
    path_loss_model: str = 'log_distance'      # or 'cost_hata' (path_loss.py:90-123), 'shadowing' (path_loss.py:69-81)
    shadow_d0_m: float = 100.0                 # path_loss.py:70
    shadow_chi_dB: float = 2.7                 # path_loss.py:70
    rng_seed: int = 0                          # shadowing draws: the product's counter-based scheme (d2d_oracle.h)
    first_global_env: int = 0
    rng_step: int = 0
    area_type: int = 1                         # path_loss.py:84-87 AreaType value (CostHata only; default SUBURBAN, :91)
    # per-device overrides {device_id: {field: value}} as a device_config_file's 'config' dicts would give
    device_overrides: Dict[str, dict] = field(default_factory=dict)

    @property
    def num_links(self) -> int:
        return self.num_cues + self.num_due_pairs + (self.num_cues if self.downlinks else 0)

    @property
    def num_devices(self) -> int:
        return 1 + self.num_cues + 2 * self.num_due_pairs

    def device_ids(self):
        """simulator.py:34-48 ids in devices.py:20-25 order."""
        ids = ['mbs'] + [f'cue{i:02d}' for i in range(self.num_cues)]
        for i in range(0, 2 * self.num_due_pairs, 2):
            ids += [f'due{i:02d}', f'due{i + 1:02d}']
        return ids

    def link_keys(self):
        """'tx:rx' keys in the canonical link order of envs/d2d_env.py:55-60."""
        keys = [f'cue{i:02d}:mbs' for i in range(self.num_cues)]
        keys += [f'due{i:02d}:due{i + 1:02d}' for i in range(0, 2 * self.num_due_pairs, 2)]
        if self.downlinks:
            keys += [f'mbs:cue{i:02d}' for i in range(self.num_cues)]       # envs/d2d_env.py:87-89 DOWNLINK actions
        return keys


def build(force: bool = False) -> Path:
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < (_HERE / 'd2d_oracle.c').stat().st_mtime:
        subprocess.run(['make', '-C', str(_HERE), '-B'], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        d = C.c_double
        for name, args, res in [
            ('d2d_oracle_dB_to_linear', [d], d), ('d2d_oracle_linear_to_dB', [d], d),
            ('d2d_oracle_pl_constant_dB', [d, d], d), ('d2d_oracle_log_distance_pl', [d, d, d], d),
            ('d2d_oracle_distance', [d, d, d, d], d),
            ('d2d_oracle_eirp_dBm', [C.POINTER(_Device), d], d),
            ('d2d_oracle_rx_signal_level_dBm', [C.POINTER(_Device), d, d], d),
            ('d2d_oracle_rx_sensitivity_dBm', [C.POINTER(_Device)], d),
            ('d2d_oracle_rb_bandwidth_kHz', [C.POINTER(_Device)], d),
        ]:
            getattr(L, name).argtypes = args
            getattr(L, name).restype = res
        L.d2d_oracle_decode_action.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.d2d_oracle_decode_action.restype = None
        vp = C.c_void_p
        L.d2d_oracle_shadow_normal.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.d2d_oracle_shadow_normal.restype = d
        L.d2d_oracle_cost_hata_pl.argtypes = [d, d, C.c_int, d, d]
        L.d2d_oracle_cost_hata_pl.restype = d
        L.d2d_oracle_step_batch.argtypes = [C.POINTER(_Cfg), C.POINTER(_Device), C.c_int64] + [vp] * 11 + [C.c_int]
        L.d2d_oracle_step_batch.restype = C.c_int
        L.d2d_oracle_per_agent_obs.argtypes = [vp, vp, C.c_int32, vp]
        L.d2d_oracle_per_agent_obs.restype = None
        L.d2d_oracle_philox4x32_10.argtypes = [vp, vp, vp]
        L.d2d_oracle_philox4x32_10.restype = None
        L.d2d_oracle_reset_positions.argtypes = [C.POINTER(_Cfg), d, d, C.c_uint64, C.c_uint64, C.c_int64, vp]
        L.d2d_oracle_reset_positions.restype = None
        L.d2d_oracle_sample_actions.argtypes = [C.POINTER(_Cfg), C.c_int32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, vp]
        L.d2d_oracle_sample_actions.restype = None
        _lib = L
    return _lib


def make_device(kind: str, **overrides) -> _Device:
    base = dict(BS_DEFAULTS if kind == 'bs' else UE_DEFAULTS)
    base.update(overrides)
    return _Device(**{k: (int(v) if k == 'is_bs' else float(v)) for k, v in base.items()})


def _c_cfg(cfg: OracleConfig) -> _Cfg:
    return _Cfg(num_rbs=cfg.num_rbs, num_cues=cfg.num_cues, num_due_pairs=cfg.num_due_pairs,
                n_pwr_cue=cfg.cue_max_tx_power_dBm + 1,                               # envs/d2d_env.py:33
                n_pwr_due=cfg.due_max_tx_power_dBm - cfg.due_min_tx_power_dBm + 1,    # envs/d2d_env.py:32
                carrier_freq_GHz=cfg.carrier_freq_GHz, ple=cfg.ple, min_capacity_mbps=cfg.min_capacity_mbps,
                path_loss_model={'cost_hata': 2, 'shadowing': 3}.get(cfg.path_loss_model, 0), area_type=int(cfg.area_type),
                shadow_d0_m=cfg.shadow_d0_m, shadow_chi_dB=cfg.shadow_chi_dB, rng_seed=cfg.rng_seed,
                first_global_env=cfg.first_global_env, rng_step=cfg.rng_step,
                downlinks=int(bool(cfg.downlinks)), n_pwr_mbs=cfg.mbs_max_tx_power_dBm + 1)            # envs/d2d_env.py:34


def device_table(cfg: OracleConfig):
    """simulator.py:18-50: per-device config = defaults + env-level subcarrier fields + file overrides."""
    ids = cfg.device_ids()
    arr = (_Device * len(ids))()
    base = dict(num_subcarriers=cfg.num_subcarriers, subcarrier_spacing_kHz=cfg.subcarrier_spacing_kHz)
    for i, id_ in enumerate(ids):
        kw = dict(base)
        if id_ in cfg.device_overrides:   # simulator.py:31: a file 'config' REPLACES the env-level dict
            kw = {k: v for k, v in cfg.device_overrides[id_].items() if k in UE_DEFAULTS}
        arr[i] = make_device('bs' if i == 0 else 'ue', **kw)
    return arr


def step_batch(cfg: OracleConfig, positions, actions, active=None, nthreads: int = 1) -> Dict[str, np.ndarray]:
    """Float64 oracle step over E envs.  positions (E,V,2) float64, actions (E,N) int32."""
    positions = np.ascontiguousarray(positions, dtype=np.float64)
    actions = np.ascontiguousarray(actions, dtype=np.int32)
    E = positions.shape[0]
    N, V = cfg.num_links, cfg.num_devices
    assert positions.shape == (E, V, 2), positions.shape
    assert actions.shape == (E, N), actions.shape
    if active is not None:
        active = np.ascontiguousarray(active, dtype=np.uint8)
        assert active.shape == (E, N)
    out = dict(rb=np.empty((E, N), np.int32), tx_pwr_dbm=np.empty((E, N), np.int32),
               sinr_db=np.empty((E, N)), snr_db=np.empty((E, N)), rate_bps=np.empty((E, N)),
               capacity_mbps=np.empty((E, N)), obs=np.empty((E, N, 6)), reward=np.empty((E,)))
    ccfg, devs = _c_cfg(cfg), device_table(cfg)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    st = lib().d2d_oracle_step_batch(C.byref(ccfg), devs, E, p(positions), p(actions),
                                     p(active) if active is not None else None,
                                     p(out['rb']), p(out['tx_pwr_dbm']), p(out['sinr_db']), p(out['snr_db']),
                                     p(out['rate_bps']), p(out['capacity_mbps']), p(out['obs']), p(out['reward']),
                                     int(nthreads))
    out['status'] = st
    return out


def agent_rewards(cfg: OracleConfig, res: Dict[str, np.ndarray], kind: str, param: float, active=None) -> np.ndarray:
    """The reference's per-agent reward functions, restated on the oracle's float64 step results (numpy; integer /
    comparison logic plus one log2).  Returns (E, N); absent agents (no entry in the reference's dict) get 0.

    kind 'shannon' - ShannonRewardFunction (envs/reward_fn.py:47-57): log2(1 + 10^(sinr/10)) if sinr >= min_sinr else -1.
    kind 'cue_sinr_shannon' - CueSinrShannonRewardFunction (envs/reward_fn.py:60-78): -1 if some OTHER action on the
        agent's RB is a non-SIDELINK (CUE) link with sinr < sinr_threshold_dB, else log2(1 + 10^(sinr/10))."""
    sinr, rb = res['sinr_db'], res['rb']
    E, N = sinr.shape
    act = np.ones((E, N), bool) if active is None else np.asarray(active, bool)
    shannon = np.log2(1.0 + np.power(10.0, sinr / 10.0))                 # conversion.py:4-13 + reward_fn.py:56,76
    if kind == 'shannon':
        out = np.where(sinr >= param, shannon, -1.0)                      # reward_fn.py:56
    elif kind == 'cue_sinr_shannon':
        is_cue = np.arange(N)[None, :] < cfg.num_cues                    # link_type != SIDELINK (reward_fn.py:71)
        weak = act & is_cue & (sinr < param)                             # reward_fn.py:72-73
        nweak = np.zeros((E, int(rb.max()) + 2), np.int64)
        ee = np.broadcast_to(np.arange(E)[:, None], (E, N))
        np.add.at(nweak, (ee[weak], rb[weak]), 1)                        # weak CUE links per RB (actions.py:27-31 grouping)
        others = nweak[ee, np.where(act, rb, 0)] - weak                  # .difference({action}) (reward_fn.py:69)
        out = np.where(others > 0, -1.0, shannon)
    else:
        raise ValueError(kind)
    return np.where(act, out, 0.0)


def per_agent_obs(table: np.ndarray, present) -> np.ndarray:
    table = np.ascontiguousarray(table, dtype=np.float64)
    present = np.ascontiguousarray(present, dtype=np.int32)
    n = present.shape[0]
    out = np.empty((n, 6 * n))
    lib().d2d_oracle_per_agent_obs(table.ctypes.data_as(C.c_void_p), present.ctypes.data_as(C.c_void_p), n,
                                   out.ctypes.data_as(C.c_void_p))
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, np.uint32)
    lib().d2d_oracle_philox4x32_10(ctr.ctypes.data_as(C.c_void_p), key.ctypes.data_as(C.c_void_p),
                                   out.ctypes.data_as(C.c_void_p))
    return out


def reset_positions(cfg: OracleConfig, seed: int, first_global_env: int, num_envs: int) -> np.ndarray:
    out = np.empty((num_envs, cfg.num_devices, 2))
    ccfg = _c_cfg(cfg)
    lib().d2d_oracle_reset_positions(C.byref(ccfg), cfg.cell_radius_m, cfg.d2d_radius_m, seed, first_global_env,
                                     num_envs, out.ctypes.data_as(C.c_void_p))
    return out


def sample_actions(cfg: OracleConfig, seed: int, first_global_env: int, step_index: int, num_envs: int) -> np.ndarray:
    """The product's counter-based Discrete(n).sample() (d2d_sample_actions / d2d_episode), restated: int32 [E][N]."""
    out = np.empty((num_envs, cfg.num_links), np.int32)
    ccfg = _c_cfg(cfg)
    lib().d2d_oracle_sample_actions(C.byref(ccfg), cfg.num_links, seed & 0xFFFFFFFFFFFFFFFF, first_global_env, step_index, num_envs,
                                    out.ctypes.data_as(C.c_void_p))
    return out


def random_positions(cfg: OracleConfig, num_envs: int, rng: np.random.Generator, fp32_exact: bool = True):
    """Synthetic scenario generator for tests/bench: CUE and DUE-tx uniform in the cell disc, DUE-rx
    uniform in the d2d disc around its tx and re-drawn until inside the cell (position.py:18-45),
    MBS at the origin (simulator.py:63-64).  With fp32_exact the coordinates are rounded to float32 so
    that the oracle and the fp32 device state see bit-identical inputs."""
    E, Cn, D, V = num_envs, cfg.num_cues, cfg.num_due_pairs, cfg.num_devices
    pos = np.zeros((E, V, 2))

    def disc(shape, radius):
        th = 2 * np.pi * rng.random(shape)
        r = radius * np.sqrt(rng.random(shape))
        return np.stack([r * np.cos(th), r * np.sin(th)], axis=-1)

    pos[:, 1:1 + Cn] = disc((E, Cn), cfg.cell_radius_m)
    tx = disc((E, D), cfg.cell_radius_m)
    if fp32_exact:
        tx = tx.astype(np.float32).astype(np.float64)
    rx = tx + disc((E, D), cfg.d2d_radius_m)
    for _ in range(200):
        bad = (rx ** 2).sum(-1) > cfg.cell_radius_m ** 2
        if not bad.any():
            break
        rx[bad] = tx[bad] + disc((int(bad.sum()),), cfg.d2d_radius_m)
    pos[:, 1 + Cn::2] = tx
    pos[:, 2 + Cn::2] = rx
    if fp32_exact:
        pos = pos.astype(np.float32).astype(np.float64)
    return pos


def random_actions(cfg: OracleConfig, num_envs: int, rng: np.random.Generator) -> np.ndarray:
    """Uniform actions over the reference's Discrete spaces (envs/d2d_env.py:36-40)."""
    n_cue = cfg.num_rbs * (cfg.cue_max_tx_power_dBm + 1)
    n_due = cfg.num_rbs * (cfg.due_max_tx_power_dBm - cfg.due_min_tx_power_dBm + 1)
    a = np.empty((num_envs, cfg.num_links), np.int32)
    C, D = cfg.num_cues, cfg.num_due_pairs
    a[:, :C] = rng.integers(0, n_cue, (num_envs, C))
    a[:, C:C + D] = rng.integers(0, n_due, (num_envs, D))
    if cfg.downlinks:
        a[:, C + D:] = rng.integers(0, cfg.num_rbs * (cfg.mbs_max_tx_power_dBm + 1), (num_envs, C))
    return a
