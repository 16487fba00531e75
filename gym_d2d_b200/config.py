"""Host-side configuration: the reference's EnvConfig surface and the folding of its per-device
link-budget dicts into the per-link constant table the kernels consume.

Mirrors envs/env_config.py:12-37 (same 16 keys, same defaults, unknown key -> TypeError, JSON
device_config_file), simulator.py:18-50 (device creation and the per-device 'config' override rule) and
device.py:12-41,51-80,93-95,134-140,158-162 (the link-budget arithmetic, evaluated once per device here
instead of per link per step).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

from . import _lib
from .plugins import (LogDistancePathLoss, UnsupportedPluginError, UplinkTrafficModel, cost_hata_terms, resolve_path_loss,
                      shadowing_params)

EPISODE_LENGTH = 10                 # envs/d2d_env.py:16
BASE_STATION_ID = 'mbs'             # simulator.py:15

# device.py:12-16
DEFAULT_DEVICE_CONFIG = {'num_PRB': 1, 'num_subcarriers': 12, 'subcarrier_spacing_kHz': 15.0}
# device.py:17-29
DEFAULT_BASE_STATION_CONFIG = {**DEFAULT_DEVICE_CONFIG, 'max_tx_power_dBm': 46.0, 'antenna_height_m': 23.0,
                               'tx_antenna_gain_dBi': 17.5, 'rx_antenna_gain_dBi': 17.5,
                               'thermal_noise_dBm': -118.4, 'noise_figure_dB': 2.0, 'sinr_dB': -7.0,
                               'ix_margin_dB': 2.0, 'cable_loss_dB': 2.0, 'masthead_amplifier_gain_dB': 2.0}
# device.py:30-41
DEFAULT_UE_CONFIG = {**DEFAULT_DEVICE_CONFIG, 'max_tx_power_dBm': 23.0, 'antenna_height_m': 1.5,
                     'tx_antenna_gain_dBi': 0.0, 'rx_antenna_gain_dBi': 0.0, 'thermal_noise_dBm': -104.5,
                     'noise_figure_dB': 7.0, 'sinr_dB': -10.0, 'ix_margin_dB': 3.0,
                     'control_channel_overhead_dB': 1.0, 'body_loss_dB': 3.0}


@dataclass
class EnvConfig:
    """Same fields and defaults as the reference dataclass (envs/env_config.py:12-27)."""
    num_rbs: int = 25
    num_cues: int = 25
    num_due_pairs: int = 25
    cell_radius_m: float = 500.0
    d2d_radius_m: float = 20.0
    due_min_tx_power_dBm: int = 0
    due_max_tx_power_dBm: int = 20
    cue_max_tx_power_dBm: int = 23
    mbs_max_tx_power_dBm: int = 46
    path_loss_model: Any = LogDistancePathLoss
    traffic_model: Any = UplinkTrafficModel
    carrier_freq_GHz: float = 2.1
    num_subcarriers: int = 12
    subcarrier_spacing_kHz: int = 15
    channel_bandwidth_MHz: float = 20.0
    device_config_file: Optional[Path] = None
    devices: Dict[str, dict] = field(init=False, default_factory=dict)
    # not a reference field: set by VecD2DEnv(downlink=True) to append the DOWNLINK links 'mbs:cueXX' that D2DEnv.step
    # accepts (envs/d2d_env.py:87-89) after the canonical uplink + sidelink ones
    downlinks: bool = field(init=False, default=False)

    def __post_init__(self) -> None:
        self.devices = self.load_device_config()

    def load_device_config(self) -> dict:
        """envs/env_config.py:32-37: only a pathlib.Path is honoured."""
        if isinstance(self.device_config_file, Path):
            with self.device_config_file.open(mode='r') as fid:
                return json.load(fid)
        return {}

    # ---- derived sizes -----------------------------------------------------------------------
    @property
    def num_links(self) -> int:
        return self.num_cues + self.num_due_pairs + (self.num_cues if self.downlinks else 0)

    @property
    def num_devices(self) -> int:
        return 1 + self.num_cues + 2 * self.num_due_pairs

    @property
    def num_pwr_actions(self) -> Dict[str, int]:
        """envs/d2d_env.py:31-35."""
        return {'due': self.due_max_tx_power_dBm - self.due_min_tx_power_dBm + 1,
                'cue': self.cue_max_tx_power_dBm + 1,
                'mbs': self.mbs_max_tx_power_dBm + 1}

    def device_ids(self) -> List[str]:
        """simulator.py:34-48 ids in devices.py:20-25 insertion order."""
        ids = [BASE_STATION_ID] + [f'cue{i:02d}' for i in range(self.num_cues)]
        for i in range(0, 2 * self.num_due_pairs, 2):
            ids += [f'due{i:02d}', f'due{i + 1:02d}']
        return ids

    def link_ids(self) -> List[Tuple[str, str]]:
        """Canonical link order: CUE j -> MBS, then DUE pairs (envs/d2d_env.py:55-60)."""
        ids = self.device_ids()
        C = self.num_cues
        links = [(ids[1 + j], BASE_STATION_ID) for j in range(C)]
        links += [(ids[1 + C + 2 * d], ids[2 + C + 2 * d]) for d in range(self.num_due_pairs)]
        if self.downlinks:
            links += [(BASE_STATION_ID, ids[1 + j]) for j in range(C)]      # envs/d2d_env.py:87-89
        return links

    def link_keys(self) -> List[str]:
        return [f'{t}:{r}' for t, r in self.link_ids()]

    # ---- per-device config dicts (simulator.py:25-48) -----------------------------------------------
    def device_configs(self) -> Dict[str, dict]:
        base_cfg = {'num_subcarriers': self.num_subcarriers, 'subcarrier_spacing_kHz': self.subcarrier_spacing_kHz}
        cue_cfg = {**base_cfg, 'max_tx_power_dBm': self.cue_max_tx_power_dBm}
        due_cfg = {**base_cfg, 'max_tx_power_dBm': self.due_max_tx_power_dBm}
        out = {}
        for idx, id_ in enumerate(self.device_ids()):
            if idx == 0:
                defaults, env_level = DEFAULT_BASE_STATION_CONFIG, base_cfg
            elif idx <= self.num_cues:
                defaults, env_level = DEFAULT_UE_CONFIG, cue_cfg
            else:
                defaults, env_level = DEFAULT_UE_CONFIG, due_cfg
            # simulator.py:31: a file entry's 'config' REPLACES the env-level dict before the merge
            override = self.devices.get(id_, {}).get('config', env_level)
            out[id_] = {**defaults, **override}
        return out


def _eirp_offset(cfg: dict, is_bs: bool) -> float:
    """eirp_dBm(p) - p: device.py:60 then :135 (BS) or :159 (UE)."""
    off = cfg['tx_antenna_gain_dBi'] - cfg['ix_margin_dB']
    return off - cfg['cable_loss_dB'] + cfg['masthead_amplifier_gain_dB'] if is_bs else off - cfg['body_loss_dB']


def _rx_offset(cfg: dict, is_bs: bool) -> float:
    """rx_signal_level_dBm(e, pl) - (e - pl): device.py:72 then :137-140 (BS) or :161-162 (UE)."""
    off = cfg['rx_antenna_gain_dBi']
    return off - cfg['cable_loss_dB'] + cfg['masthead_amplifier_gain_dB'] if is_bs else off - cfg['body_loss_dB']


def cost_hata_fold(config: EnvConfig) -> Optional[Tuple[float, Dict[str, float]]]:
    """CostHataPathLoss (path_loss.py:90-123) as (ple, {receiver device id: K_dB}), or None for the other models.  The
    exponent depends on the transmitter's antenna height, so every transmitter must share one height (they are all
    UserEquipments on the uplink + sidelink path; a device_config_file may still move receivers' heights freely)."""
    pl_enum, area = resolve_path_loss(config.path_loss_model)
    if pl_enum != _lib.PL_COST_HATA:
        return None
    dev_cfg = config.device_configs()
    tx_heights = {float(dev_cfg[tx]['antenna_height_m']) for tx, _ in config.link_ids()}
    if len(tx_heights) != 1:
        raise UnsupportedPluginError('CostHataPathLoss: the CUDA path needs one antenna height for all transmitters, got '
                                     f'{sorted(tx_heights)}')
    h_tx = tx_heights.pop()
    ple, consts = 0.0, {}
    for _, rx in config.link_ids():
        ple, consts[rx] = cost_hata_terms(float(config.carrier_freq_GHz), int(area), h_tx, float(dev_cfg[rx]['antenna_height_m']))
    return ple, consts


def link_table(config: EnvConfig) -> List[dict]:
    """One dict per link (canonical order) with the fields of d2d_link_t."""
    dev_cfg = config.device_configs()
    hata = cost_hata_fold(config)
    rows = []
    for j, (tx_id, rx_id) in enumerate(config.link_ids()):
        tx, rx = dev_cfg[tx_id], dev_cfg[rx_id]
        rx_is_bs = rx_id == BASE_STATION_ID
        rows.append(dict(
            tx_eirp_offset_dB=float(_eirp_offset(tx, tx_id == BASE_STATION_ID)),
            rx_offset_dB=float(_rx_offset(rx, rx_is_bs)),
            rx_noise_dBm=float(rx['thermal_noise_dBm']),                                               # device.py:117-119
            rx_sensitivity_dBm=float(rx['noise_figure_dB'] + rx['thermal_noise_dBm'] + rx['sinr_dB']),   # device.py:74-80
            tx_rb_bandwidth_kHz=float(int(tx['num_subcarriers']) * int(tx['subcarrier_spacing_kHz'])),   # device.py:85-95
            link_type=(_lib.LINK_UPLINK if j < config.num_cues else
                       _lib.LINK_SIDELINK if j < config.num_cues + config.num_due_pairs else _lib.LINK_DOWNLINK),
            path_loss_const_dB=float(hata[1][rx_id]) if hata else 0.0))
    return rows


def to_c_config(config: EnvConfig, num_envs: int, cuda_device: int, obs_enum: int, reward_enum: int,
                reward_param: float, rng_seed: int = 0, first_global_env: int = 0) -> '_lib.D2DConfig':
    pl_enum, ple = resolve_path_loss(config.path_loss_model)
    hata = cost_hata_fold(config)
    if hata:
        ple = hata[0]
    npw = config.num_pwr_actions
    for name in ('num_rbs', 'num_cues', 'num_due_pairs'):
        v = getattr(config, name)
        if not isinstance(v, int) or isinstance(v, bool) or v < 0:
            raise ValueError(f'{name} must be a non-negative int, got {v!r}')
    return _lib.D2DConfig(abi_version=_lib.ABI_VERSION, cuda_device=cuda_device, num_envs=num_envs,
                          num_rbs=config.num_rbs, num_cues=config.num_cues, num_due_pairs=config.num_due_pairs,
                          n_pwr_cue=npw['cue'], n_pwr_due=npw['due'], episode_length=EPISODE_LENGTH,
                          path_loss_model=pl_enum, obs_fn=obs_enum, reward_fn=reward_enum,
                          num_downlinks=config.num_cues if config.downlinks else 0, n_pwr_mbs=npw['mbs'],
                          carrier_freq_GHz=float(config.carrier_freq_GHz), ple=ple,
                          cell_radius_m=float(config.cell_radius_m), d2d_radius_m=float(config.d2d_radius_m),
                          min_capacity_mbps=float(reward_param) if reward_enum == _lib.REWARD_SYSTEM_CAPACITY else 0.0,
                          reward_param=float(reward_param), shadow_d0_m=shadowing_params(config.path_loss_model)[0],
                          shadow_chi_dB=shadowing_params(config.path_loss_model)[1],
                          rng_seed=int(rng_seed) & 0xFFFFFFFFFFFFFFFF, first_global_env=int(first_global_env))
