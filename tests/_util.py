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
