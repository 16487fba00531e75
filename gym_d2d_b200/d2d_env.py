"""D2DEnv: the reference's single-environment dict API (envs/d2d_env.py) as an E = 1 view over the CUDA path.

Same surface as the reference class: `D2DEnv(env_config)`, `reset() -> {key: ndarray(6N,)}`,
`step({'tx:rx': int}) -> (obs, rewards, {'__all__': done}, info)`, `render()`, `save_device_config(Path)`,
attributes `observation_space`, `action_space`, `num_pwr_actions`, `actions`, `state`, `num_steps`.
All arithmetic runs in the sm_100a step kernel through `d2d_step_host`; this file is key parsing and
dict building only.

Documented divergences from the reference:
  * custom / unimplemented plugin classes are rejected at construction (north star);
  * negative or out-of-range integer actions raise ValueError (the reference decodes them silently, B.9);
  * the first 'mbs:cueXX' downlink key (B.8) moves the env to the general-topology kernel (2C + D links);
  * device positions come from a Philox stream keyed by `seed`, not Python's global `random` (section 3.2).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .config import BASE_STATION_ID, EPISODE_LENGTH
from .spaces import Box, Dict as DictSpace, Discrete
from .vec_env import VecD2DEnv


class D2DEnv:
    metadata = {'render.modes': ['human']}

    def __init__(self, env_config: Optional[dict] = None, device: Any = 'cuda', seed: int = 0) -> None:
        env_config = env_config or {}
        self._ctor = (dict(env_config), device, seed)      # to rebuild the env with DOWNLINK links on the first 'mbs:' key
        # VecD2DEnv pops 'obs_fn' / 'reward_fn' from the caller's dict exactly like envs/d2d_env.py:27-28
        self.vec = VecD2DEnv(1, env_config, device=device, seed=seed, info=True, exact_positions=True)
        self._bind_vec()
        self.actions: Optional[Dict[str, Tuple[int, int]]] = None
        self.state: Optional[dict] = None
        self.num_steps = 0
        self._present: List[int] = []
        import random
        self._host_rng = random.Random(seed)             # receivers re-drawn around file-placed transmitters (reset)

    def _bind_vec(self) -> None:
        cfg = self.vec.config
        self.config = cfg
        r = cfg.cell_radius_m
        # envs/obs_fn.py:36-41
        self.observation_space = Box(low=-r, high=r, shape=(6 * (cfg.num_cues + cfg.num_due_pairs),))
        self.num_pwr_actions = cfg.num_pwr_actions                               # envs/d2d_env.py:31-35
        self.action_space = DictSpace({k: Discrete(cfg.num_rbs * n) for k, n in
                                       (('due', self.num_pwr_actions['due']), ('cue', self.num_pwr_actions['cue']),
                                        ('mbs', self.num_pwr_actions['mbs']))})   # envs/d2d_env.py:36-40
        self.device_ids = cfg.device_ids()
        self.link_keys = cfg.link_keys()
        self._link_index = {k: i for i, k in enumerate(self.link_keys)}
        self._device_set = set(self.device_ids)
        self._cues = set(self.device_ids[1:1 + cfg.num_cues])
        self._due_tx = {t for (t, _r) in cfg.link_ids()[cfg.num_cues:]}
        # the library's pinned slot buffers (d2d_host_slot_buffers): the step's actions go up and all of its results come back
        # with ONE copy each way instead of one pageable copy per output array
        outputs = ('obs', 'capacity_mbps', 'reward', 'done', 'rate_bps', 'rb', 'tx_pwr_dbm') + (('agent_reward',) if self.vec.per_agent_reward else ())
        self._host = self.vec.host_slot_buffers(0, outputs=outputs)
        self._nvec = [int(n) for n in self.vec.action_nvec]
        self._views: Dict[tuple, tuple] = {}                  # per set of present agents: (rows, per-agent row order, id pairs)

    def _enable_downlink(self) -> None:
        """envs/d2d_env.py:87-89: a key whose transmitter is neither a DUE nor a CUE is a DOWNLINK action of the MBS.  The
        canonical uplink / sidelink link indices are unchanged; C downlink links are appended."""
        cfg, device, seed = self._ctor
        old = self.vec
        new = VecD2DEnv(1, dict(cfg), device=device, seed=seed, info=True, exact_positions=True, downlink=True)
        new.set_positions(old.positions_f64)
        new.step_count.copy_(old.step_count)
        new._episode = old._episode
        old.close()
        self.vec = new
        self._bind_vec()

    # ---- helpers ------------------------------------------------------------------------------
    def _decode_action(self, key: str, action: Any) -> int:
        """Type rule of envs/d2d_env.py:93-101; the integer itself is decoded on the GPU."""
        if not isinstance(action, (int, np.integer)) or isinstance(action, bool):
            raise ValueError(f'Unable to decode action type "{type(action)}"')
        j = self._link_index.get(key)
        if j is None:
            tx_id, _, rx_id = key.partition(':')
            for id_ in (tx_id, rx_id):
                if id_ not in self._device_set:
                    raise KeyError(id_)                               # devices.py:28
            raise KeyError(key)
        n = self._nvec[j]
        a = int(action)
        if not 0 <= a < n:
            raise ValueError(f'action {a} for "{key}" is outside Discrete({n})')
        return a

    def _step_arrays(self, raw_actions: Dict[str, Any]) -> List[str]:
        if not self.vec.config.downlinks and any(isinstance(k, str) and k.startswith(BASE_STATION_ID + ':') and
                                                 k.partition(':')[2] in self._cues for k in raw_actions):
            self._enable_downlink()                                    # envs/d2d_env.py:87-89: DOWNLINK actions of the MBS
        acts = [-1] * self.config.num_links                            # -1: agent absent (Appendix B.8)
        keys = []
        index = self._link_index
        for key, action in raw_actions.items():                       # caller's insertion order (envs/d2d_env.py:75)
            a = self._decode_action(key, action)                      # (raises for keys that are no link of this env)
            acts[index[key]] = a
            keys.append(key)
        self._host['actions'][0] = acts
        self.vec.step_host_async(self._host['actions'], self._host, 0)
        self.vec.step_host_wait(0)
        return keys

    def _view(self, keys: List[str]) -> tuple:
        """Index tables of one set of present agents, built once per distinct key order: the agents' link rows, the row order of
        every agent's observation (envs/obs_fn.py:43-53: own 6-tuple first, then the others in the action dict's order) and the
        (tx, rx) id pairs the reference keys its state by."""
        tk = tuple(keys)
        v = self._views.get(tk)
        if v is None:
            rows = np.asarray([self._link_index[k] for k in keys], dtype=np.intp)
            n = len(keys)
            order = np.empty((n, n), dtype=np.intp)
            for i in range(n):
                order[i, 0] = rows[i]
                order[i, 1:i + 1] = rows[:i]
                order[i, i + 1:] = rows[i + 1:]
            ids = [tuple(k.split(':')) for k in keys]
            if len(self._views) >= 64:
                self._views.clear()
            v = self._views[tk] = (rows, order, ids)
        return v

    def _obs_dict(self, keys: List[str]) -> Dict[str, np.ndarray]:
        """envs/obs_fn.py:43-53: own 6-tuple then every other present link's, in the action dict's order."""
        table = self._host['obs'][0].astype(np.float64)
        _rows, order, _ids = self._view(keys)
        flat = table[order].reshape(len(keys), -1)                     # one gather for all agents; each agent's obs is its row
        return dict(zip(keys, flat))

    def _state(self, keys: List[str]) -> dict:
        h = self._host
        rows, _order, ids = self._view(keys)
        dyn = h['obs'][0, rows, 4:6].astype(np.float64)
        return {
            'sinrs_db': dict(zip(ids, dyn[:, 0].tolist())),
            'snrs_db': dict(zip(ids, dyn[:, 1].tolist())),
            'rate_bps': dict(zip(ids, h['rate_bps'][0, rows].astype(np.float64).tolist())),
            'capacity_mbps': dict(zip(ids, h['capacity_mbps'][0, rows].astype(np.float64).tolist())),
        }

    def _position_nearby(self, anchor) -> Tuple[float, float]:
        """get_random_position_nearby (position.py:31-45) for one receiver, on the host (E = 1 adapter only)."""
        import math
        cfg = self.config
        rng = self._host_rng
        while True:
            theta = 2 * math.pi * rng.random()
            r = cfg.d2d_radius_m * math.sqrt(rng.random())
            x, y = float(anchor[0]) + r * math.cos(theta), float(anchor[1]) + r * math.sin(theta)
            if x * x + y * y <= cfg.cell_radius_m ** 2:
                return x, y

    # ---- gym surface ----------------------------------------------------------------------------
    def reset(self) -> Dict[str, np.ndarray]:
        """envs/d2d_env.py:45-52: new positions, then one uncounted step with random actions."""
        self.num_steps = 0
        self.vec.reset(mask=np_mask_all(self.vec))                    # positions + counters only
        file_devices = self.config.devices
        if file_devices:                                               # simulator.py:65-66
            pos = self.vec.positions_f64[0].cpu().numpy()
            placed = set()
            for idx, id_ in enumerate(self.device_ids):
                if idx and id_ in file_devices and 'position' in file_devices[id_]:
                    pos[idx] = file_devices[id_]['position']
                    placed.add(id_)
            # simulator.py:70-73: a DUE receiver that is NOT in the file is drawn around its transmitter's FINAL position - which
            # may have come from the file - so re-draw those receivers (position.py:31-45) instead of leaving them next to the
            # transmitter position the device-side reset had drawn
            index = {id_: i for i, id_ in enumerate(self.device_ids)}
            for tx_id, rx_id in self.config.link_ids()[self.config.num_cues:self.config.num_cues + self.config.num_due_pairs]:
                if tx_id in placed and rx_id not in placed:
                    pos[index[rx_id]] = self._position_nearby(pos[index[tx_id]])
            self.vec.set_positions(pos[None])
        raw = {k: self.action_space['cue' if k.startswith('cue') else 'due'].sample() for k in self.link_keys
               if not k.startswith(BASE_STATION_ID + ':')}             # envs/d2d_env.py:54-60: uplinks and sidelinks only
        self.vec._bind(False)                                          # the reset step is not counted
        try:
            keys = self._step_arrays(raw)
        finally:
            self.vec._bind(True)
        self.actions = self._actions_view(keys)
        self.state = self._state(keys)
        return self._obs_dict(keys)

    def step(self, raw_actions: Dict[str, Any]):
        """envs/d2d_env.py:62-71."""
        keys = self._step_arrays(raw_actions)
        self.num_steps += 1
        self.actions = self._actions_view(keys)
        self.state = self._state(keys)
        obs = self._obs_dict(keys)
        if self.vec.per_agent_reward:                                  # envs/reward_fn.py:47-78: one reward per agent
            rewards = dict(zip(keys, self._host['agent_reward'][0, self._view(keys)[0]].astype(np.float64).tolist()))
        else:
            reward = float(self._host['reward'][0])
            rewards = {k: reward for k in keys}                        # envs/reward_fn.py:44
        game_over = {'__all__': self.num_steps >= EPISODE_LENGTH}      # envs/d2d_env.py:68
        # envs/d2d_env.py:106-116
        st = self.state
        info = {k: {'rb': a[0], 'tx_pwr_dbm': a[1], 'snr_db': snr, 'sinr_db': sinr, 'rate_bps': rate, 'capacity_mbps': cap}
                for k, a, snr, sinr, rate, cap in zip(keys, self.actions.values(), st['snrs_db'].values(), st['sinrs_db'].values(),
                                                      st['rate_bps'].values(), st['capacity_mbps'].values())}
        return obs, rewards, game_over, info

    def _actions_view(self, keys: List[str]) -> Dict[str, Tuple[int, int]]:
        h = self._host
        rows = self._view(keys)[0]
        return dict(zip(keys, zip(h['rb'][0, rows].tolist(), h['tx_pwr_dbm'][0, rows].tolist())))

    def render(self, mode: str = 'human') -> None:
        assert self.state is not None and self.actions is not None, \
            'Initialise environment with `reset()` before calling `render()`'
        print(self._obs_dict(list(self.actions.keys())))

    # ---- device config I/O (envs/d2d_env.py:124-134, envs/env_config.py:32-37) -----------------------
    def device_positions(self) -> Dict[str, Tuple[float, float]]:
        pos = self.vec.positions_f64[0].cpu().numpy()
        return {id_: (float(pos[i, 0]), float(pos[i, 1])) for i, id_ in enumerate(self.device_ids)}

    def set_device_positions(self, positions: Dict[str, Tuple[float, float]]) -> None:
        """Batch form of Device.set_position (device.py:82-83) for this env."""
        pos = self.vec.positions_f64[0].cpu().numpy()
        index = {id_: i for i, id_ in enumerate(self.device_ids)}
        for id_, xy in positions.items():
            pos[index[id_]] = xy                                       # KeyError on unknown id, like devices.py:28
        self.vec.set_positions(pos[None])

    def save_device_config(self, config_file: Path) -> None:
        positions = self.device_positions()
        configs = self.config.device_configs()
        doc = {id_: {'position': positions[id_], 'config': configs[id_]} for id_ in self.device_ids}
        with config_file.open(mode='w') as fid:
            json.dump(doc, fid)

    def close(self) -> None:
        self.vec.close()


def np_mask_all(vec: VecD2DEnv):
    import torch
    return torch.ones((vec.num_envs,), dtype=torch.uint8, device=vec.device)
