"""CPU tier: pin the oracle (oracle/d2d_oracle.c) against
  (1) every known-answer vector the reference's own tests hold for this path,
  (2) the committed fixtures produced by running the unmodified reference (tests/golden/gen_golden.py),
  (3) the live reference when a copy is importable (build container / baseline/_ref).
"""
import ctypes as C
import json

import numpy as np
import pytest

from oracle import d2d_oracle as O
from oracle import ref_runner as R
from tests._util import assert_rel

TIGHT = 1e-9   # the oracle is float64 like the reference: it must agree far below the 1e-4 product tolerance


# ---- (1) the reference's own golden vectors ------------------------------------------------------------
def test_conversion_vectors():
    """test/gym_d2d/test_conversion.py:6-23"""
    L = O.lib()
    for dB, lin in [(1, 1.258925), (2, 1.584893), (30, 1000.0)]:
        assert L.d2d_oracle_dB_to_linear(dB) == pytest.approx(lin, rel=1e-6)
    for lin, dB in [(2, 3.0103), (3, 4.771213), (5, 6.9897), (30, 14.771213)]:
        assert L.d2d_oracle_linear_to_dB(lin) == pytest.approx(dB, rel=1e-6)


def test_path_loss_vectors():
    """test/gym_d2d/test_path_loss.py:8-27"""
    L = O.lib()
    assert L.d2d_oracle_pl_constant_dB(2.0, 2.0) == pytest.approx(38.46838313516298)
    assert L.d2d_oracle_pl_constant_dB(2.1, 2.0) == pytest.approx(38.892169116561746)
    assert L.d2d_oracle_pl_constant_dB(2.2, 2.0) == pytest.approx(39.2962368383275)
    assert L.d2d_oracle_log_distance_pl(L.d2d_oracle_distance(250, 0, 0, 0), 2.1, 2.0) == pytest.approx(86.85097)
    assert L.d2d_oracle_log_distance_pl(L.d2d_oracle_distance(0, 500, 0, 0), 2.1, 2.0) == pytest.approx(92.87156)


def test_device_vectors():
    """test/gym_d2d/test_device.py:71-99 (EIRP, sensitivity, noise floor from the default dicts)"""
    L = O.lib()
    ue, bs = O.make_device('ue'), O.make_device('bs')
    assert L.d2d_oracle_eirp_dBm(C.byref(ue), 12.0) == pytest.approx(12 + 0.0 - 3.0 - 3.0)
    assert L.d2d_oracle_eirp_dBm(C.byref(bs), 46.0) == pytest.approx(46 + 17.5 - 2.0 - 2.0 + 2.0)
    assert L.d2d_oracle_rx_sensitivity_dBm(C.byref(ue)) == pytest.approx((7.0 - 104.5) - 10.0)
    assert L.d2d_oracle_rx_sensitivity_dBm(C.byref(bs)) == pytest.approx((2.0 - 118.4) - 7.0)
    assert L.d2d_oracle_rb_bandwidth_kHz(C.byref(ue)) == 180.0


def test_decode_is_python_floor_divmod():
    """envs/d2d_env.py:95-96 with Python // and % semantics (Appendix B.9: a = -1 -> rb -1, p n-1)"""
    L = O.lib()
    rb, pw = C.c_int64(), C.c_int64()
    for a in [-50, -22, -21, -1, 0, 1, 20, 21, 524, 599, 1174, 10 ** 6]:
        for n in (21, 24, 47):
            L.d2d_oracle_decode_action(a, n, C.byref(rb), C.byref(pw))
            assert (rb.value, pw.value) == (a // n, a % n)


def test_rb_grouping_semantics():
    """test/gym_d2d/test_actions.py:24-48: links interfere iff they share the RB.  With every link on its own
    RB there is no interference and SINR == SNR (up to rounding)."""
    cfg = O.OracleConfig(num_rbs=8, num_cues=3, num_due_pairs=4)
    rng = np.random.default_rng(3)
    pos = O.random_positions(cfg, 5, rng)
    act = np.zeros((5, 7), np.int32)
    act[:, :3] = np.arange(3) * 24 + 5
    act[:, 3:] = (3 + np.arange(4)) * 21 + 7
    out = O.step_batch(cfg, pos, act)
    np.testing.assert_allclose(out['sinr_db'], out['snr_db'], rtol=1e-12)
    act[:, 3] = 0 * 21 + 7          # DUE 0 joins CUE 0's RB: only those two links change
    out2 = O.step_batch(cfg, pos, act)
    changed = np.abs(out2['sinr_db'] - out['sinr_db']) > 1e-9
    assert changed[:, [0, 3]].all() and not changed[:, [1, 2, 4, 5, 6]].any()


# ---- (2) committed reference outputs ----------------------------------------------------------------------
def _cfg_from(kw):
    return O.OracleConfig(**{k: v for k, v in kw.items() if k in ('num_rbs', 'num_cues', 'num_due_pairs')})


def test_appendix_c(golden_dir):
    doc = json.loads((golden_dir / 'appendix_c.json').read_text())
    cfg = _cfg_from(doc['env_config'])
    out = O.step_batch(cfg, np.array([doc['positions']]), np.array([doc['actions']], np.int32))
    assert out['rb'][0].tolist() == doc['rb'] and out['tx_pwr_dbm'][0].tolist() == doc['tx_pwr_dbm']
    for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
        assert_rel(out[k][0], doc[k], TIGHT, k)
    assert_rel(out['reward'][0], doc['reward'], TIGHT, 'reward')
    # SURVEY Appendix C hand check: cue00 SNR = 23 - 6 - (20 log10(100) + 38.892169) + 17.5 + 118.4
    assert out['snr_db'][0, 0] == pytest.approx(74.00783088343826, rel=1e-12)
    pa = O.per_agent_obs(out['obs'][0], np.arange(cfg.num_links))
    for i, key in enumerate(doc['keys']):
        assert_rel(pa[i], doc['per_agent_obs'][key], TIGHT, f'obs[{key}]')


@pytest.mark.parametrize('name,kw', [
    ('default_25_25_25', {}), ('default_fp64_positions', {}),
    ('dense_small_8_6_30', dict(num_rbs=8, num_cues=6, num_due_pairs=30)),
    ('dense_100_100_500', dict(num_rbs=100, num_cues=100, num_due_pairs=500)),
    ('tiny_1_1_1', dict(num_rbs=1, num_cues=1, num_due_pairs=1)),
])
def test_batched_fixtures(golden_dir, name, kw):
    g = np.load(golden_dir / f'{name}.npz')
    cfg = O.OracleConfig(**kw)
    N = cfg.num_links
    for s in range(g['actions'].shape[0]):
        out = O.step_batch(cfg, g['positions'], g['actions'][s])
        assert out['status'] == 0
        np.testing.assert_array_equal(out['rb'], g['rb'][s])
        np.testing.assert_array_equal(out['tx_pwr_dbm'], g['tx_pwr_dbm'][s])
        for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps', 'reward']:
            assert_rel(out[k], g[k][s], TIGHT, f'{name}.{k}')
        for e in range(g['positions'].shape[0]):
            pa = O.per_agent_obs(out['obs'][e], np.arange(N))
            assert_rel(pa[0], g['agent0_obs'][s, e], TIGHT, 'agent0 obs')
            assert_rel(pa[-1], g['agentlast_obs'][s, e], TIGHT, 'last agent obs')


def test_subset_and_caller_order(golden_dir):
    doc = json.loads((golden_dir / 'subset_order.json').read_text())
    cfg = _cfg_from(doc['env_config'])
    order = doc['order']
    active = np.zeros((1, cfg.num_links), np.uint8)
    active[0, order] = 1
    out = O.step_batch(cfg, np.array([doc['positions']]), np.array([doc['actions_all']], np.int32), active=active)
    for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
        assert_rel(out[k][0, order], doc[k], TIGHT, k)
    assert_rel(out['reward'][0], doc['reward'][0], TIGHT, 'reward')
    assert (out['capacity_mbps'][0, [i for i in range(cfg.num_links) if i not in order]] == 0).all()
    assert_rel(O.per_agent_obs(out['obs'][0], order), doc['per_agent_obs'], TIGHT, 'per-agent obs in caller order')


def test_device_overrides_and_penalty(golden_dir):
    doc = json.loads((golden_dir / 'overrides_penalty.json').read_text())
    dev = json.loads((golden_dir / 'overrides_device_config.json').read_text())
    cfg = O.OracleConfig(**doc['env_config'], device_overrides={k: v['config'] for k, v in dev.items()})
    pos = np.array([[dev[i]['position'] for i in cfg.device_ids()]])
    pos[0, 0] = 0.0
    rewards = []
    for s, act in enumerate(doc['actions']):
        out = O.step_batch(cfg, pos, np.array([act], np.int32))
        for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
            assert_rel(out[k][0], doc[k][s], TIGHT, k)
        assert_rel(out['reward'][0], doc['reward'][s], TIGHT, 'reward')
        rewards.append(out['reward'][0])
    assert any(r == -1.0 for r in rewards) and any(r > 0 for r in rewards)   # both reward branches covered


def test_fixed_scenario_10k(golden_dir):
    """BASELINE config #4 against the reference's stored outputs: all 10 000 rewards / capacity sums, and the
    full per-link outputs of every 50th step."""
    g = np.load(golden_dir / 'fixed_scenario_10k.npz')
    dev = json.loads((golden_dir / 'fixed_device_config.json').read_text())
    cfg = O.OracleConfig()
    pos1 = np.array([dev[i]['position'] for i in cfg.device_ids()])
    pos1[0] = 0.0
    acts = g['actions'].astype(np.int32)
    T = acts.shape[0]
    out = O.step_batch(cfg, np.broadcast_to(pos1, (T,) + pos1.shape), acts, nthreads=4)
    assert_rel(out['reward'], g['reward'], TIGHT, 'reward[10k]')
    assert_rel(out['capacity_mbps'].sum(1), g['capsum'], TIGHT, 'capsum[10k]')
    assert_rel(out['sinr_db'].sum(1), g['sinrsum'], 1e-7, 'sinrsum[10k]')
    every = int(g['every'])
    for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
        assert_rel(out[k][::every], g[k], TIGHT, k)
    np.testing.assert_array_equal(out['rb'][::every], g['rb'])
    np.testing.assert_array_equal(out['tx_pwr_dbm'][::every], g['tx_pwr_dbm'])


# ---- (3) live reference -----------------------------------------------------------------------------------
@pytest.mark.skipif(R.reference_src() is None, reason='no copy of the reference is reachable')
@pytest.mark.parametrize('rbs,cues,dues', [(25, 25, 25), (3, 2, 5), (40, 30, 60)])
def test_live_reference(rbs, cues, dues):
    cfg = O.OracleConfig(num_rbs=rbs, num_cues=cues, num_due_pairs=dues)
    env = R.make_env(dict(num_rbs=rbs, num_cues=cues, num_due_pairs=dues))
    env.reset()
    rng = np.random.default_rng(rbs)
    for s in range(4):
        pos = O.random_positions(cfg, 1, rng, fp32_exact=bool(s % 2))
        act = O.random_actions(cfg, 1, rng)
        R.set_positions(env, pos[0])
        ref = R.step(env, act[0])
        out = O.step_batch(cfg, pos, act)
        np.testing.assert_array_equal(out['rb'][0], ref['rb'])
        np.testing.assert_array_equal(out['tx_pwr_dbm'][0], ref['tx_pwr_dbm'])
        for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
            assert_rel(out[k][0], ref[k], TIGHT, k)
        assert_rel(out['reward'][0], ref['reward'][0], TIGHT, 'reward')
        assert_rel(O.per_agent_obs(out['obs'][0], np.arange(cfg.num_links)), ref['per_agent_obs'], TIGHT, 'obs')


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    assert O.philox4x32_10([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_reset_positions_geometry():
    cfg = O.OracleConfig()
    pos = O.reset_positions(cfg, seed=7, first_global_env=0, num_envs=64)
    assert (pos[:, 0] == 0).all()
    assert ((pos ** 2).sum(-1) <= cfg.cell_radius_m ** 2 * (1 + 1e-12)).all()
    tx, rx = pos[:, 1 + cfg.num_cues::2], pos[:, 2 + cfg.num_cues::2]
    assert (np.sqrt(((tx - rx) ** 2).sum(-1)) <= cfg.d2d_radius_m * (1 + 1e-12)).all()
    # shard independence: env g is the same whichever slice it is drawn in
    np.testing.assert_array_equal(O.reset_positions(cfg, 7, 40, 8), pos[40:48])


def test_reward_plugins_match_reference_fixture(golden_dir):
    """ShannonRewardFunction / CueSinrShannonRewardFunction (envs/reward_fn.py:47-78): the oracle's restatement against
    per-agent rewards produced by the unmodified reference (tests/golden/reward_plugins.npz)."""
    g = np.load(golden_dir / "reward_plugins.npz")
    cfg = O.OracleConfig(num_rbs=4, num_cues=6, num_due_pairs=9)
    steps, E, N = g['actions'].shape
    for s in range(steps):
        res = O.step_batch(cfg, g['positions'], g['actions'][s])
        np.testing.assert_allclose(res['sinr_db'], g['sinr_db'][s], rtol=1e-9)
        np.testing.assert_allclose(O.agent_rewards(cfg, res, 'shannon', -70.0), g['shannon_reward'][s], rtol=1e-9)
        np.testing.assert_allclose(O.agent_rewards(cfg, res, 'cue_sinr_shannon', 0.0), g['cue_sinr_shannon_reward'][s], rtol=1e-9)
    assert (g['cue_sinr_shannon_reward'] == -1).any() and (g['cue_sinr_shannon_reward'] > 0).any()
    # a subset of agents: absent agents neither transmit nor count as weak CUEs
    present = g['subset_present']
    active = np.zeros((1, N), np.uint8)
    active[0, present] = 1
    res = O.step_batch(cfg, g['positions'][:1], g['actions'][0, :1], active=active)
    rew = O.agent_rewards(cfg, res, 'cue_sinr_shannon', 0.0, active=active)
    np.testing.assert_allclose(rew[0, present], g['subset_cue_sinr_shannon_reward'], rtol=1e-9)
    assert (rew[0, np.setdiff1d(np.arange(N), present)] == 0).all()


def test_cost_hata_known_answers_and_fixture(golden_dir):
    """CostHataPathLoss: the reference's own known-answer vectors (test/gym_d2d/test_path_loss.py:42-52: URBAN, f = 2.1 GHz,
    default BaseStation (23 m) and UserEquipment (1.5 m) heights, both directions) and step results of the unmodified reference (cost_hata.npz)."""
    L = O.lib()
    for d, h_tx, h_rx, want in [(250.0, 23.0, 1.5, 121.44557455875727), (250.0, 1.5, 23.0, 114.35415557446962),
                                (500.0, 23.0, 1.5, 132.2768393081241), (500.0, 1.5, 23.0, 127.5231950610599)]:
        assert L.d2d_oracle_cost_hata_pl(d, 2.1, 2, h_tx, h_rx) == pytest.approx(want, rel=1e-12)   # test_path_loss.py:42-52
    g = np.load(golden_dir / 'cost_hata.npz')
    steps = g['actions'].shape[0]
    for name, area in [('suburban', 1), ('urban', 2)]:
        cfg = O.OracleConfig(num_rbs=3, num_cues=5, num_due_pairs=7, path_loss_model='cost_hata', area_type=area)
        for s in range(steps):
            res = O.step_batch(cfg, g['positions'], g['actions'][s])
            for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
                np.testing.assert_allclose(res[k], g[f'{name}_{k}'][s], rtol=1e-9, atol=1e-12, err_msg=f'{name} {k}')
            np.testing.assert_allclose(res['reward'], g[f'{name}_reward'][s], rtol=1e-9)


def test_downlink_actions_match_reference_fixture(golden_dir):
    """'mbs:cueXX' DOWNLINK actions (envs/d2d_env.py:87-89; Appendix B.8): MBS transmitter, 47 power levels, decoded and
    stepped by the unmodified reference together with uplinks, sidelinks and absent agents (tests/golden/downlink.npz)."""
    g = np.load(golden_dir / 'downlink.npz')
    cfg = O.OracleConfig(num_rbs=4, num_cues=5, num_due_pairs=6, downlinks=True)
    assert cfg.link_keys() == [str(k) for k in g['keys']]
    for s in range(g['actions'].shape[0]):
        res = O.step_batch(cfg, g['positions'], g['actions'][s], active=g['active'][s])
        on = g['active'][s] > 0
        np.testing.assert_array_equal(res['rb'][on], g['rb'][s][on])
        np.testing.assert_array_equal(res['tx_pwr_dbm'][on], g['tx_pwr_dbm'][s][on])
        for k in ['sinr_db', 'snr_db', 'rate_bps', 'capacity_mbps']:
            np.testing.assert_allclose(res[k][on], g[k][s][on], rtol=1e-9, atol=1e-12, err_msg=k)
        np.testing.assert_allclose(res['reward'], g['reward'][s], rtol=1e-9)
    assert g['tx_pwr_dbm'][..., 11:].max() > 23            # downlink powers beyond any UE's range were drawn


def test_shadowing_distribution_matches_reference(golden_dir):
    """ShadowingPathLoss (path_loss.py:69-81): the oracle uses the product's counter-based draws, the reference Python's global
    RNG, so the comparison is distributional: per-link mean and standard deviation of SINR_dB / SNR_dB / capacity over K
    steps of one fixed scenario (tests/golden/shadowing_stats.npz), and the independence of the SINR and SNR draws."""
    g = np.load(golden_dir / 'shadowing_stats.npz')
    cfg = O.OracleConfig(num_rbs=3, num_cues=4, num_due_pairs=6, d2d_radius_m=150.0, path_loss_model='shadowing', rng_seed=5)
    K = 4000
    sinr = np.zeros((K, cfg.num_links)); snr = np.zeros_like(sinr); cap = np.zeros_like(sinr)
    for t in range(K):
        cfg.rng_step = t
        r = O.step_batch(cfg, g['positions'][None], g['actions'][None])
        sinr[t], snr[t], cap[t] = r['sinr_db'][0], r['snr_db'][0], r['capacity_mbps'][0]
    se = lambda std: 6.0 * np.maximum(std, 1e-9) * np.sqrt(1.0 / K + 1.0 / float(g['K']))     # six standard errors
    assert (np.abs(sinr.mean(0) - g['sinr_mean']) <= se(g['sinr_std']) + 1e-9).all()
    assert (np.abs(snr.mean(0) - g['snr_mean']) <= se(g['snr_std']) + 1e-9).all()
    assert (np.abs(cap.mean(0) - g['cap_mean']) <= se(g['cap_std']) + 1e-9).all()
    np.testing.assert_allclose(sinr.std(0), g['sinr_std'], rtol=0.08, atol=1e-9)
    np.testing.assert_allclose(snr.std(0), g['snr_std'], rtol=0.08, atol=1e-9)
    far = g['snr_std'] > 1e-6                              # links longer than d0 = 100 m: SNR std = chi = 2.7 dB
    assert far.any() and (~far).any()
    np.testing.assert_allclose(snr.std(0)[far], 2.7, rtol=0.08)
    corr = np.array([np.corrcoef(sinr[:, j], snr[:, j])[0, 1] for j in np.nonzero(far)[0]])
    assert (np.abs(corr - g['corr_sinr_snr'][far]) < 0.1).all()          # separate evaluations: (nearly) uncorrelated


# ---- the product's counter-based draw schemes, restated by the oracle, against the reference's own sampler ---------------------
def test_reset_scheme_matches_reference_sampler_distribution(golden_dir):
    """tests/golden/reset_distribution.npz holds quantile tables of the UNMODIFIED reference's get_random_position /
    get_random_position_nearby (position.py:18-45; generated by tests/golden/gen_reset_distribution.py): radial and angular
    laws of a CUE, of a DUE receiver's offset, and the inward skew the in-cell rejection loop gives receivers of transmitters
    near the cell edge.  The oracle's restatement of the device-side reset must follow the same laws (two-sample KS)."""
    from tests._util import check_reset_distribution
    fx = np.load(golden_dir / 'reset_distribution.npz')
    cfg = O.OracleConfig()
    pos = O.reset_positions(cfg, seed=2026, first_global_env=123, num_envs=8192)
    res = check_reset_distribution(pos, cfg.num_cues, fx)
    assert set(res) == {'cue_r2', 'cue_theta', 'off_r2', 'off_theta', 'edge_dr', 'edge_off_r2'}
    assert (pos[:, 0] == 0).all()                                                  # simulator.py:63-64
    # a wrong law is caught: receivers drawn WITHOUT the rejection loop's inward skew
    bad = pos.copy()
    tx = bad[:, 1 + cfg.num_cues::2]
    rng = np.random.default_rng(0)
    th, r = 2 * np.pi * rng.random(tx.shape[:2]), 20.0 * np.sqrt(rng.random(tx.shape[:2]))
    bad[:, 2 + cfg.num_cues::2] = tx + np.stack([r * np.cos(th), r * np.sin(th)], -1)
    with pytest.raises(AssertionError):
        check_reset_distribution(bad, cfg.num_cues, fx)


def test_live_reference_sampler_matches_fixture(golden_dir):
    """The fixture itself against a fresh run of the reference's sampler, when a copy of the reference is importable."""
    from oracle import ref_runner as R
    if R.import_reference() is None:
        pytest.skip('reference not available')
    import random
    from gym_d2d.position import get_random_position
    from tests._util import ks_against_quantiles
    fx = np.load(golden_dir / 'reset_distribution.npz')
    random.seed(1)
    p = np.array([get_random_position(500.0).as_tuple() for _ in range(50000)])
    assert ks_against_quantiles((p ** 2).sum(-1) / 500.0 ** 2, fx['levels'], fx['cue_r2']) < 0.01


def test_action_scheme_is_uniform_like_discrete_sample():
    """envs/d2d_env.py:54-60 samples gym.spaces.Discrete(n): uniform over 0 .. n - 1.  The product's counter-based draw
    (restated by d2d_oracle_sample_actions) must be uniform, differ between steps / envs / links, and not depend on sharding."""
    cfg = O.OracleConfig()
    E = 20000
    a = np.stack([O.sample_actions(cfg, 42, 0, t, E) for t in range(4)])
    n_cue, n_due = 25 * 24, 25 * 21
    assert a[..., :25].min() == 0 and a[..., :25].max() == n_cue - 1
    assert a[..., 25:].min() == 0 and a[..., 25:].max() == n_due - 1
    for block, n in ((a[..., :25], n_cue), (a[..., 25:], n_due)):
        counts = np.bincount(block.ravel(), minlength=n)
        expect = block.size / n
        assert np.abs(counts - expect).max() < 6 * np.sqrt(expect)               # every value equally likely
    assert not np.array_equal(a[0], a[1]) and not np.array_equal(a[2], a[3])     # the two halves of one Philox block differ too
    # successive steps of one link are uncorrelated
    x = a[:, :, 3].astype(np.float64)
    assert abs(np.corrcoef(x[0], x[1])[0, 1]) < 0.03 and abs(np.corrcoef(x[1], x[2])[0, 1]) < 0.03
    # sharding invariance: global env g draws the same action whatever the shard offset
    b = O.sample_actions(cfg, 42, 5000, 1, 100)
    np.testing.assert_array_equal(b, a[1, 5000:5100])
