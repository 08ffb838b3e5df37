"""Parity AT THE SIZE of BASELINE.json configs C3, C3b and C4 (the catalogue sizes, walker counts,
bounds and fixed parameters the metric is quoted on), against the NumPy oracle on a few walkers --
the oracle makes ~20 NumPy passes over the catalogue per walker, so three walkers take about a second.

Reference semantics matched: ``analysis/runner.py:264-286`` (Gaussian sum / fixed-background mixture),
``analysis/model.py:391-456`` (fitted Gaussian background), ``analysis/model.py:93-223`` with a free
centre; omega Cen bounds and ``v_sys = 232.5`` fixed from ``bin/run_test_5139_center.py:157-165``.
"""
import numpy as np
import pytest

from common import RTOL
from mcmc_dynamics_b200 import configs, synthetic
from oracle import harness
from oracle import reference_np as ref

pytestmark = pytest.mark.gpu

N_ORACLE_WALKERS = 3


def check_against_oracle(model, truth, n_walkers, lnlike_background=None):
    theta = synthetic.initial_ball(truth, model.fitted_parameters, n_walkers, seed=5, scale=0.05)
    half = n_walkers // 2
    got = model.lnprob(theta[:half])                     # the call emcee makes: one half-ensemble
    assert got.shape == (half,) and np.all(np.isfinite(got))
    oracle = harness.oracle_for(model, lnlike_background=lnlike_background)
    want = oracle.lnprob_many(theta[:N_ORACLE_WALKERS])
    assert np.all(np.isfinite(want))
    err = harness.relative_error(got[:N_ORACLE_WALKERS], want)
    assert err < RTOL, err
    # the arithmetic variants agree over the whole half-ensemble (size-independent property)
    model.math_mode = 'plain'
    plain = model.lnprob(theta[:half])
    model.math_mode = 'fast'
    assert harness.relative_error(plain, got) < 1e-11
    return theta, got


def test_c3_fixed_background_mixture_at_size():
    """10^5 stars + SingleStars(M = 2000) background + pmember, 256 walkers."""
    name, model, truth, n_walkers = configs.config_c3()
    assert model.n_data == 100_000 and n_walkers == 256
    # the background column itself (M x N log-mean-exp, background/single_stars.py:72-77) against the
    # oracle on a subsample of the stars (the oracle materialises M x N like the reference)
    v_bg = np.asarray(model.background.v.value)
    idx = np.random.default_rng(0).choice(model.n_data, 3000, replace=False)
    v = np.asarray(model.v.value)[idx]
    verr = np.asarray(model.verr.value)[idx]
    want_bg = ref.single_stars_background(v_bg, v, verr)
    got_bg = np.asarray(model.lnlike_background)[idx]
    assert np.max(np.abs(got_bg - want_bg) / np.maximum(1.0, np.abs(want_bg))) < RTOL
    theta, got = check_against_oracle(model, truth, n_walkers)
    # prior rejection and evaluation in one call
    theta[1, model.fitted_parameters.index('sigma_max')] = -1.0
    out = model.lnprob(theta[:8])
    # (8 walkers per call use another launch geometry than 128: same values up to summation order)
    assert out[1] == -np.inf and np.allclose(out[[0, 2, 3]], got[[0, 2, 3]], rtol=1e-12, atol=0)


def test_c3b_fitted_gaussian_background_at_size():
    """10^5 stars, ModelFitGB (v_back, sigma_back, f_back sampled), 256 walkers."""
    name, model, truth, n_walkers = configs.config_c3b()
    assert model.n_data == 100_000 and model.n_fitted_parameters == 9
    check_against_oracle(model, truth, n_walkers)


def test_c4_free_centre_omega_cen_at_size():
    """3 x 10^5 stars, free centre + rotation axis, v_sys fixed 232.5, omega Cen bounds, 128 walkers."""
    name, model, truth, n_walkers = configs.config_c4()
    assert model.n_data == 300_000 and 'v_sys' not in model.fitted_parameters
    assert model.parameters['v_sys'].value == 232.5 and model.parameters['r_peak'].max == 500
    theta, got = check_against_oracle(model, truth, n_walkers)
    # the bounds of bin/run_test_5139_center.py:157-165 reject exactly at the edge
    names = model.fitted_parameters
    th = theta[:4].copy()
    th[0, names.index('sigma_max')] = 100.0          # inclusive
    th[1, names.index('sigma_max')] = 100.0000001
    th[2, names.index('v_maxx')] = -100.0000001
    out = model.lnprob(th)
    assert np.isfinite(out[0]) and out[1] == -np.inf and out[2] == -np.inf and np.isclose(out[3], got[3], rtol=1e-12, atol=0)
