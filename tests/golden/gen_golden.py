"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container (where /root/reference exists):

    python tests/golden/gen_golden.py

The reference (davidcotton/gym-d2d) is imported from /root/reference/src behind the gym stand-in in
oracle/gym_stub; nothing from this repository's product code or oracle arithmetic is involved in
producing the numbers - oracle.d2d_oracle is used only for its input generators (positions/actions).
The fixtures pin SINR / SNR / rate / capacity / observation / reward / action decode, none of which
the reference's own tests pin.
"""
from __future__ import annotations

import json
import random
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import d2d_oracle as O  # noqa: E402  (input generators only)
from oracle import ref_runner as R  # noqa: E402


def run_cases(name, env_kwargs, cfg, num_envs, steps, seed, fp32_exact=True):
    """num_envs independent scenarios x steps random actions through the reference."""
    rng = np.random.default_rng(seed)
    env = R.make_env(dict(env_kwargs))
    env.reset()
    keys = R.link_keys(env)
    N = len(keys)
    pos = O.random_positions(cfg, num_envs, rng, fp32_exact=fp32_exact)
    out = dict(positions=pos, actions=np.zeros((steps, num_envs, N), np.int32))
    for k in ['rb', 'tx_pwr_dbm']:
        out[k] = np.zeros((steps, num_envs, N), np.int64)
    for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
        out[k] = np.zeros((steps, num_envs, N))
    out['reward'] = np.zeros((steps, num_envs))
    out['agent0_obs'] = np.zeros((steps, num_envs, 6 * N))     # obs of the first agent (reference layout)
    out['agentlast_obs'] = np.zeros((steps, num_envs, 6 * N))  # obs of the last agent
    for s in range(steps):
        act = O.random_actions(cfg, num_envs, rng)
        out['actions'][s] = act
        for e in range(num_envs):
            R.set_positions(env, pos[e])
            ref = R.step(env, act[e], keys)
            for k in ['rb', 'tx_pwr_dbm', 'sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
                out[k][s, e] = ref[k]
            assert np.all(ref['reward'] == ref['reward'][0])    # same scalar for every agent
            out['reward'][s, e] = ref['reward'][0]
            out['agent0_obs'][s, e] = ref['per_agent_obs'][0]
            out['agentlast_obs'][s, e] = ref['per_agent_obs'][-1]
    out['keys'] = np.array(keys)
    np.savez_compressed(HERE / f'{name}.npz', **out)
    print(name, 'ok', {k: v.shape for k, v in out.items() if hasattr(v, 'shape')})


def appendix_c():
    """The hand-checkable 2/2/2 scenario of SURVEY.md Appendix C."""
    env = R.make_env(dict(num_rbs=2, num_cues=2, num_due_pairs=2))
    env.reset()
    pos = np.array([[0, 0], [100, 0], [0, -200], [50, 50], [60, 50], [-300, 40], [-300, 25]], float)
    R.set_positions(env, pos)
    keys = R.link_keys(env)
    act = [23, 34, 20, 5]
    ref = R.step(env, act, keys)
    doc = dict(env_config=dict(num_rbs=2, num_cues=2, num_due_pairs=2), positions=pos.tolist(), keys=keys,
               actions=act, rb=ref['rb'].tolist(), tx_pwr_dbm=ref['tx_pwr_dbm'].tolist(),
               sinr_db=ref['sinr_db'].tolist(), snr_db=ref['snr_db'].tolist(), rate_bps=ref['rate_bps'].tolist(),
               capacity_mbps=ref['capacity_mbps'].tolist(), reward=float(ref['reward'][0]),
               per_agent_obs={k: ref['per_agent_obs'][i].tolist() for i, k in enumerate(keys)})
    (HERE / 'appendix_c.json').write_text(json.dumps(doc, indent=1))
    print('appendix_c ok')


def subset_and_order():
    """Appendix B.8: a subset of agents, in a caller-chosen order; absent agents do not transmit."""
    cfg = O.OracleConfig(num_rbs=3, num_cues=4, num_due_pairs=5)
    rng = np.random.default_rng(77)
    env = R.make_env(dict(num_rbs=3, num_cues=4, num_due_pairs=5))
    env.reset()
    pos = O.random_positions(cfg, 1, rng)[0]
    R.set_positions(env, pos)
    keys_all = R.link_keys(env)
    order = [7, 0, 5, 2, 8]                       # canonical link indices, caller order
    keys = [keys_all[i] for i in order]
    act_all = O.random_actions(cfg, 1, rng)[0]
    ref = R.step(env, [act_all[i] for i in order], keys)
    doc = dict(env_config=dict(num_rbs=3, num_cues=4, num_due_pairs=5), positions=pos.tolist(),
               keys_all=keys_all, order=order, actions_all=act_all.tolist(),
               **{k: ref[k].tolist() for k in ['rb', 'tx_pwr_dbm', 'sinr_db', 'snr_db', 'rate_bps',
                                               'capacity_mbps', 'reward']},
               per_agent_obs=ref['per_agent_obs'].tolist())
    (HERE / 'subset_order.json').write_text(json.dumps(doc, indent=1))
    print('subset_order ok')


def overrides_penalty():
    """Appendix B.3: a device_config_file that raises the MBS 'sinr_dB' so CUE capacities gate to 0 and
    SystemCapacityRewardFunction returns -1 whenever a DUE shares an RB with a CUE; plus a per-device
    antenna-gain override on one DUE.  Exercises simulator.py:31 and envs/env_config.py:32-37."""
    random.seed(4242)
    base = dict(num_rbs=3, num_cues=3, num_due_pairs=4)
    env = R.make_env(dict(base))
    env.reset()
    with tempfile.TemporaryDirectory() as td:
        f = Path(td) / 'dev.json'
        env.save_device_config(f)
        doc = json.loads(f.read_text())
    doc['mbs']['config']['sinr_dB'] = 500.0
    doc['due02']['config']['tx_antenna_gain_dBi'] = 4.5
    doc['due03']['config']['rx_antenna_gain_dBi'] = 2.25
    doc['cue01']['config']['thermal_noise_dBm'] = -101.0
    (HERE / 'overrides_device_config.json').write_text(json.dumps(doc, indent=1))
    env = R.make_env(dict(base, device_config_file=HERE / 'overrides_device_config.json'))
    env.reset()
    keys = R.link_keys(env)
    cfg = O.OracleConfig(**base)
    rng = np.random.default_rng(5)
    steps = 12
    acts = np.stack([O.random_actions(cfg, 1, rng)[0] for _ in range(steps)])
    for s in range(0, steps, 3):   # every third step: CUEs on RB 0, DUEs on RBs 1-2 -> no penalty
        acts[s, :3] = 0 * 24 + rng.integers(0, 24, 3)
        acts[s, 3:] = rng.integers(1, 3, 4) * 21 + rng.integers(0, 21, 4)
    res = [R.step(env, a, keys) for a in acts]
    out = dict(env_config=base, keys=keys, actions=acts.tolist(),
               **{k: [r[k].tolist() for r in res] for k in ['rb', 'tx_pwr_dbm', 'sinr_db', 'snr_db', 'rate_bps',
                                                            'capacity_mbps']},
               reward=[float(r['reward'][0]) for r in res])
    (HERE / 'overrides_penalty.json').write_text(json.dumps(out, indent=1))
    print('overrides_penalty ok; rewards', out['reward'])


def fixed_scenario_10k():
    """BASELINE config #4: one fixed scenario from a device_config_file produced by the reference's own
    reset() + save_device_config(), then 10 000 steps of NumPy default_rng(0) actions.  Every 50th
    step is stored in full; reward and total capacity are stored for all 10 000 steps."""
    random.seed(20261017)
    env = R.make_env({})
    env.reset()
    env.save_device_config(HERE / 'fixed_device_config.json')
    env = R.make_env(dict(device_config_file=HERE / 'fixed_device_config.json'))
    env.reset()
    keys = R.link_keys(env)
    cfg = O.OracleConfig()
    rng = np.random.default_rng(0)
    T, every = 10000, 50
    acts = np.stack([O.random_actions(cfg, 1, rng)[0] for _ in range(T)])
    reward = np.zeros(T)
    capsum = np.zeros(T)
    sinrsum = np.zeros(T)
    full = {k: [] for k in ['rb', 'tx_pwr_dbm', 'sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']}
    for t in range(T):
        r = R.step(env, acts[t], keys)
        reward[t] = r['reward'][0]
        capsum[t] = r['capacity_mbps'].sum()
        sinrsum[t] = r['sinr_db'].sum()
        if t % every == 0:
            for k in full:
                full[k].append(r[k])
    np.savez_compressed(HERE / 'fixed_scenario_10k.npz', actions=acts.astype(np.int16), reward=reward,
                        capsum=capsum, sinrsum=sinrsum, every=every, keys=np.array(keys),
                        **{k: np.stack(v) for k, v in full.items()})
    print('fixed_scenario_10k ok', reward[:3], capsum[:3])


def cost_hata():
    """CostHataPathLoss (path_loss.py:90-123) through the unmodified reference, SUBURBAN (the class
    default, what `path_loss_model=CostHataPathLoss` gives through simulator.py:59) and URBAN (functools.partial)."""
    import functools
    from gym_d2d.path_loss import AreaType, CostHataPathLoss
    base = dict(num_rbs=3, num_cues=5, num_due_pairs=7)
    cfg = O.OracleConfig(**base)
    rng = np.random.default_rng(41)
    num_envs, steps = 4, 3
    pos = O.random_positions(cfg, num_envs, rng)
    acts = np.stack([O.random_actions(cfg, num_envs, rng) for _ in range(steps)])
    out = dict(positions=pos, actions=acts.astype(np.int32))
    for name, model in [('suburban', CostHataPathLoss), ('urban', functools.partial(CostHataPathLoss, area_type=AreaType.URBAN))]:
        env = R.make_env(dict(base, path_loss_model=model))
        env.reset()
        keys = R.link_keys(env)
        res = {k: np.zeros((steps, num_envs, len(keys))) for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']}
        rew = np.zeros((steps, num_envs))
        for s in range(steps):
            for e in range(num_envs):
                R.set_positions(env, pos[e])
                ref = R.step(env, acts[s, e], keys)
                for k in res:
                    res[k][s, e] = ref[k]
                rew[s, e] = ref['reward'][0]
        for k in res:
            out[f'{name}_{k}'] = res[k]
        out[f'{name}_reward'] = rew
    np.savez_compressed(HERE / 'cost_hata.npz', **out)
    print('cost_hata ok', out['suburban_sinr_db'][0, 0, :4], out['urban_sinr_db'][0, 0, :4])


def downlink():
    """Appendix B.8 / envs/d2d_env.py:87-89: 'mbs:cueXX' keys create DOWNLINK actions (MBS transmitter, 'mbs' action space with
    47 power levels).  Uplinks use RBs 0-1, downlinks RBs 2-3 (an uplink and a downlink on one RB put the MBS transmitter at
    distance 0 from the MBS receiver: math domain error in the reference), sidelinks anywhere; a few agents are absent."""
    base = dict(num_rbs=4, num_cues=5, num_due_pairs=6)
    cfg = O.OracleConfig(**base, downlinks=True)
    rng = np.random.default_rng(61)
    num_envs, steps = 4, 3
    pos = O.random_positions(cfg, num_envs, rng)
    C, D, N = 5, 6, 16
    env = R.make_env(dict(base))
    env.reset()
    keys = cfg.link_keys()
    out = dict(positions=pos, actions=np.zeros((steps, num_envs, N), np.int32), active=np.zeros((steps, num_envs, N), np.uint8))
    res = {k: np.zeros((steps, num_envs, N)) for k in ['rb', 'tx_pwr_dbm', 'sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']}
    rew = np.zeros((steps, num_envs))
    for s in range(steps):
        for e in range(num_envs):
            act = np.zeros(N, np.int64)
            act[:C] = rng.integers(0, 2, C) * 24 + rng.integers(0, 24, C)                  # uplink: RB 0-1
            act[C:C + D] = rng.integers(0, 4, D) * 21 + rng.integers(0, 21, D)             # sidelink: any RB
            act[C + D:] = rng.integers(2, 4, C) * 47 + rng.integers(0, 47, C)              # downlink: RB 2-3, up to 46 dBm
            active = rng.random(N) < 0.8
            active[0] = active[C] = active[C + D] = True
            present = [i for i in range(N) if active[i]]
            R.set_positions(env, pos[e])
            ref = R.step(env, [act[i] for i in present], [keys[i] for i in present])
            out['actions'][s, e] = act
            out['active'][s, e] = active
            for k in res:
                res[k][s, e, present] = ref[k]
            assert np.all(ref['reward'] == ref['reward'][0])
            rew[s, e] = ref['reward'][0]
    out.update(res)
    out['reward'] = rew
    out['keys'] = np.array(keys)
    np.savez_compressed(HERE / 'downlink.npz', **out)
    print('downlink ok', res['sinr_db'][0, 0], rew[0])


def shadowing_stats():
    """ShadowingPathLoss (path_loss.py:69-81) draws gauss(0, chi) from Python's global RNG at every evaluation, so its
    results can only be pinned as DISTRIBUTIONS: per-link mean / std of SINR_dB, SNR_dB and capacity over K steps of one
    fixed scenario with fixed actions, from the unmodified reference."""
    from gym_d2d.path_loss import ShadowingPathLoss
    base = dict(num_rbs=3, num_cues=4, num_due_pairs=6, cell_radius_m=500.0, d2d_radius_m=150.0)   # some D2D links beyond d0 = 100 m
    cfg = O.OracleConfig(**{k: v for k, v in base.items()})
    rng = np.random.default_rng(81)
    random.seed(81)
    pos = O.random_positions(cfg, 1, rng)[0]
    act = O.random_actions(cfg, 1, rng)[0]
    env = R.make_env(dict(base, path_loss_model=ShadowingPathLoss))
    env.reset()
    keys = R.link_keys(env)
    R.set_positions(env, pos)
    K = 6000
    sinr = np.zeros((K, len(keys))); snr = np.zeros((K, len(keys))); cap = np.zeros((K, len(keys)))
    for t in range(K):
        ref = R.step(env, act, keys)
        sinr[t], snr[t], cap[t] = ref['sinr_db'], ref['snr_db'], ref['capacity_mbps']
    np.savez_compressed(HERE / 'shadowing_stats.npz', positions=pos, actions=act.astype(np.int32), K=K,
                        sinr_mean=sinr.mean(0), sinr_std=sinr.std(0), snr_mean=snr.mean(0), snr_std=snr.std(0),
                        cap_mean=cap.mean(0), cap_std=cap.std(0), corr_sinr_snr=np.array([np.corrcoef(sinr[:, j], snr[:, j])[0, 1]
                                                                                     if snr[:, j].std() > 0 else 0.0 for j in range(len(keys))]))
    print('shadowing_stats ok', sinr.std(0).round(2), snr.std(0).round(2))


def reward_plugins():
    """the reference's two per-agent reward functions (envs/reward_fn.py:47-78), run through the unmodified
    reference with `reward_fn=<class>` in env_config (envs/d2d_env.py:28).  Stores every agent's reward."""
    from gym_d2d.envs.reward_fn import CueSinrShannonRewardFunction, ShannonRewardFunction
    base = dict(num_rbs=4, num_cues=6, num_due_pairs=9)
    cfg = O.OracleConfig(**base)
    rng = np.random.default_rng(31)
    num_envs, steps = 5, 4
    pos = O.random_positions(cfg, num_envs, rng)
    acts = np.stack([O.random_actions(cfg, num_envs, rng) for _ in range(steps)])
    out = dict(positions=pos, actions=acts.astype(np.int32))
    for name, cls in [('shannon', ShannonRewardFunction), ('cue_sinr_shannon', CueSinrShannonRewardFunction)]:
        env = R.make_env(dict(base, reward_fn=cls))
        env.reset()
        keys = R.link_keys(env)
        rew = np.zeros((steps, num_envs, len(keys)))
        sinr = np.zeros((steps, num_envs, len(keys)))
        for s in range(steps):
            for e in range(num_envs):
                R.set_positions(env, pos[e])
                ref = R.step(env, acts[s, e], keys)
                rew[s, e] = ref['reward']
                sinr[s, e] = ref['sinr_db']
        out[f'{name}_reward'] = rew
        out['sinr_db'] = sinr
        out['keys'] = np.array(keys)
    # a subset of agents (absent agents get no reward entry) for the CUE-SINR rule
    env = R.make_env(dict(base, reward_fn=CueSinrShannonRewardFunction))
    env.reset()
    keys = R.link_keys(env)
    present = [0, 2, 3, 6, 7, 8, 11, 14]
    R.set_positions(env, pos[0])
    ref = R.step(env, [acts[0, 0, i] for i in present], [keys[i] for i in present])
    out['subset_present'] = np.array(present)
    out['subset_cue_sinr_shannon_reward'] = ref['reward']
    out['subset_sinr_db'] = ref['sinr_db']
    np.savez_compressed(HERE / 'reward_plugins.npz', **out)
    print('reward_plugins ok', out['shannon_reward'][0, 0, :4], out['cue_sinr_shannon_reward'][0, 0, :4],
          'penalised', float((out['cue_sinr_shannon_reward'] == -1).mean()))


if __name__ == '__main__':
    assert R.import_reference() is not None, 'the reference must be importable to (re)generate fixtures'
    if len(sys.argv) > 1:                      # regenerate single fixtures: python gen_golden.py reward_plugins ...
        for fn in sys.argv[1:]:
            globals()[fn]()
        sys.exit(0)
    appendix_c()
    subset_and_order()
    overrides_penalty()
    run_cases('default_25_25_25', {}, O.OracleConfig(), num_envs=6, steps=4, seed=11)
    run_cases('default_fp64_positions', {}, O.OracleConfig(), num_envs=4, steps=2, seed=12, fp32_exact=False)
    kw = dict(num_rbs=8, num_cues=6, num_due_pairs=30)
    run_cases('dense_small_8_6_30', kw, O.OracleConfig(**kw), num_envs=3, steps=3, seed=13)
    kw = dict(num_rbs=100, num_cues=100, num_due_pairs=500)
    run_cases('dense_100_100_500', kw, O.OracleConfig(**kw), num_envs=1, steps=1, seed=14)
    kw = dict(num_rbs=1, num_cues=1, num_due_pairs=1)
    run_cases('tiny_1_1_1', kw, O.OracleConfig(**kw), num_envs=4, steps=3, seed=15)
    fixed_scenario_10k()
    reward_plugins()
    cost_hata()
    downlink()
    shadowing_stats()
