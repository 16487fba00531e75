"""Shared helpers for the parity tests (oracle = checker; never the thing under test)."""
from __future__ import annotations

import numpy as np

from oracle import d2d_oracle as O

RTOL = 1e-4   # BASELINE.json north_star: floats within 1e-4 RELATIVE of the reference's float64 results


def rel_err(got, ref):
    """Pure relative error |got - ref| / |ref| (0 where both are exactly 0)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    diff = np.abs(got - ref)
    den = np.abs(ref)
    out = np.zeros_like(diff)
    nz = den > 0
    out[nz] = diff[nz] / den[nz]
    out[~nz & (diff > 0)] = np.inf
    return out


def assert_rel(got, ref, rtol=RTOL, what=''):
    err = rel_err(got, ref)
    worst = float(err.max()) if err.size else 0.0
    assert worst <= rtol, f'{what}: max relative error {worst:.3e} > {rtol:g} at {np.unravel_index(err.argmax(), err.shape)}'
    return worst


def oracle_cfg(**kw) -> 'O.OracleConfig':
    return O.OracleConfig(**kw)


def check_against_oracle(out, ref, rtol=RTOL, pos_rtol=RTOL):
    """out: dict of numpy arrays from the CUDA path; ref: oracle step_batch dict (float64)."""
    worst = {}
    np.testing.assert_array_equal(out['rb'], ref['rb'], err_msg='rb must be bit-exact')
    np.testing.assert_array_equal(out['tx_pwr_dbm'], ref['tx_pwr_dbm'], err_msg='tx_pwr must be bit-exact')
    worst['sinr_db'] = assert_rel(out['obs'][..., 4], ref['sinr_db'], rtol, 'sinr_db')
    worst['snr_db'] = assert_rel(out['obs'][..., 5], ref['snr_db'], rtol, 'snr_db')
    worst['obs_pos'] = assert_rel(out['obs'][..., :4], ref['obs'][..., :4], pos_rtol, 'obs positions')
    worst['capacity'] = assert_rel(out['capacity_mbps'], ref['capacity_mbps'], rtol, 'capacity_mbps')
    worst['rate'] = assert_rel(out['rate_bps'], ref['rate_bps'], rtol, 'rate_bps')
    worst['reward'] = assert_rel(out['reward'], ref['reward'], rtol, 'reward')
    return worst


def ks_against_quantiles(sample, levels, quantiles):
    """Two-sample Kolmogorov-Smirnov distance between `sample` and the sample a quantile table (levels, quantiles) was
    made from: max |F_sample(q_i) - level_i| over the table's points."""
    x = np.sort(np.asarray(sample, np.float64))
    f = np.searchsorted(x, quantiles, side='right') / x.size
    return float(np.max(np.abs(f - levels)))


def check_reset_distribution(positions, num_cues, fixture, alpha_c=1.95):
    """positions [E][V][2] drawn by the product's reset scheme against tests/golden/reset_distribution.npz (quantile tables of
    the UNMODIFIED reference's get_random_position / get_random_position_nearby, position.py:18-45).  alpha_c = 1.95 is the
    two-sample KS coefficient for alpha = 0.001: D < c sqrt((n + m) / (n m))."""
    pos = np.asarray(positions, np.float64)
    Rc, d, m = float(fixture['cell_radius_m']), float(fixture['d2d_radius_m']), int(fixture['samples'])
    lv = fixture['levels']
    cue = pos[:, 1:1 + num_cues].reshape(-1, 2)
    tx = pos[:, 1 + num_cues::2].reshape(-1, 2)
    rx = pos[:, 2 + num_cues::2].reshape(-1, 2)
    off = rx - tx
    band = np.sqrt((tx ** 2).sum(-1)) > Rc - d
    stats = {
        'cue_r2': (cue ** 2).sum(-1) / Rc ** 2,
        'cue_theta': np.mod(np.arctan2(cue[:, 1], cue[:, 0]) / (2 * np.pi), 1.0),
        'off_r2': (off[~band] ** 2).sum(-1) / d ** 2,
        'off_theta': np.mod(np.arctan2(off[~band, 1], off[~band, 0]) / (2 * np.pi), 1.0),
        'edge_dr': np.sqrt((rx[band] ** 2).sum(-1)) - np.sqrt((tx[band] ** 2).sum(-1)),
        'edge_off_r2': (off[band] ** 2).sum(-1) / d ** 2,
    }
    out = {}
    for name, sample in stats.items():
        n = sample.size
        assert n > 1000, f'{name}: only {n} samples'
        dist = ks_against_quantiles(sample, lv, fixture[name])
        bound = alpha_c * np.sqrt((n + m) / (n * m)) + 0.5 / lv.size      # + the table's own resolution
        assert dist < bound, f'{name}: KS distance {dist:.4f} >= {bound:.4f} (n = {n})'
        out[name] = (dist, bound)
    # the share of transmitters in the edge band is itself a property of the uniform-in-disc draw
    assert abs(band.mean() - float(fixture['edge_fraction'])) < 4 * np.sqrt(0.0784 * 0.9216 / band.size)
    return out
