"""Batched per-radial-bin fits: one segmented launch against one model object per bin
(bin/run_tests.py:81-97, bin/run.py:179-190)."""
import numpy as np
import pytest

from mcmc_dynamics_b200 import synthetic
from mcmc_dynamics_b200.analysis import ConstantFit, ModelFit, RadialBinsFit
from oracle import harness

pytestmark = pytest.mark.gpu


def _binned(n_stars=2500, seed=4, cls=ConstantFit):
    data, truth = synthetic.mock_cluster(n_stars, seed=seed)
    data.make_radial_bins(truth['ra_center'], truth['dec_center'], nstars=50, dlogr=0.1)
    fit = RadialBinsFit(data, model_class=cls)
    fit.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    fit.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    return fit, truth


@pytest.mark.parametrize('cls', [ConstantFit, ModelFit])
def test_segmented_lnprob_equals_per_bin_models_and_oracle(cls):
    fit, truth = _binned(cls=cls)
    assert fit.n_bins >= 5 and fit.bin_sizes.sum() == 2500
    n_walkers = 24
    theta = np.stack([synthetic.initial_ball(truth, fit.fitted_parameters, n_walkers, seed=10 + b, scale=0.3)
                      for b in range(fit.n_bins)])
    theta[2, 5, fit.fitted_parameters.index('sigma_max')] = -1.0       # one rejected walker in one bin
    got = fit.lnprob(theta)
    assert got.shape == (fit.n_bins, n_walkers)
    assert got[2, 5] == -np.inf
    for b in range(fit.n_bins):
        single = fit.bin_model(b)
        assert single.n_data == fit.bin_sizes[b]
        alone = single.lnprob(theta[b])
        assert harness.relative_error(got[b], alone) < 1e-13          # same kernel, same arithmetic per star
        want = harness.oracle_for(single).lnprob_many(theta[b])
        assert harness.relative_error(got[b], want) < 1e-9
    ll = fit.lnlike(theta)
    assert np.isfinite(ll[2, 5]) and np.allclose(np.delete(ll.ravel(), 2 * n_walkers + 5),
                                                 np.delete(got.ravel(), 2 * n_walkers + 5))


@pytest.mark.parametrize('path', ['resident', 'graph'])
def test_binned_device_sampler_matches_independent_runs_statistically(path, monkeypatch):
    monkeypatch.setenv('MCD_NO_RESIDENT_CHAIN', '1' if path == 'graph' else '0')
    monkeypatch.setenv('MCD_FORCE_RESIDENT_CHAIN', '1' if path == 'resident' else '0')
    fit, truth = _binned(n_stars=1500, seed=9)
    fit.parameters['sigma_max'].set(initials='rng.lognormal(mean=2.3, sigma=0.3, size=n)')
    fit.parameters['v_maxx'].set(initials='rng.normal(loc=0, scale=3, size=n)')
    fit.parameters['v_maxy'].set(initials='rng.normal(loc=0, scale=3, size=n)')
    n_walkers, n_steps, n_burn = 32, 400, 150
    engine = fit(n_walkers=n_walkers, n_steps=n_steps, seed=5)
    chain = engine.chain
    assert chain.shape == (fit.n_bins, n_walkers, n_steps, fit.n_fitted_parameters)
    lnp = engine.lnprobability
    assert lnp.shape == (fit.n_bins, n_walkers, n_steps) and np.all(np.isfinite(lnp))
    # stored lnprob belongs to the stored positions, bin by bin
    again = fit.lnprob(np.ascontiguousarray(chain[:, :, -1, :]))
    assert np.allclose(again, lnp[:, :, -1], rtol=1e-12, atol=0)
    # each bin's posterior agrees with a stand-alone device run of that bin
    j = fit.fitted_parameters.index('sigma_max')
    for b in (0, fit.n_bins - 1):
        single = fit.bin_model(b)
        pos = chain[b, :, 0, :]
        alone = single(n_walkers=n_walkers, n_steps=n_steps, pos=np.ascontiguousarray(pos), sampler='device', seed=77,
                       prefix=None)
        a = chain[b, :, n_burn:, j].ravel()
        c = alone.chain[:, n_burn:, j].ravel()
        assert abs(np.median(a) - np.median(c)) < 0.5 * 0.5 * (a.std() + c.std())
        assert 0.6 < a.std() / c.std() < 1.6
    frac = engine.naccepted / float(n_steps)
    assert frac.shape == (fit.n_bins, n_walkers) and 0.15 < frac.mean() < 0.9


def test_bins_need_labels_and_plain_models():
    data, truth = synthetic.mock_cluster(100, seed=1)
    with pytest.raises(IOError):
        RadialBinsFit(data)


@pytest.mark.parametrize('path', ['resident', 'graph'])
def test_binned_fits_with_a_background_object(path, monkeypatch):
    """``ConstantFit(data_i, parameters=parameters, background=background)`` per radial bin
    (bin/run.py:186): the fixed-background mixture (analysis/runner.py:272-286) in one segmented launch
    equals one model object per bin and the oracle; the batched device sampler runs on it."""
    from mcmc_dynamics_b200.background import SingleStars
    monkeypatch.setenv('MCD_NO_RESIDENT_CHAIN', '1' if path == 'graph' else '0')
    monkeypatch.setenv('MCD_FORCE_RESIDENT_CHAIN', '1' if path == 'resident' else '0')
    columns, truth = synthetic.mock_cluster(1800, seed=12, as_reader=False)
    columns, sample_field = synthetic.add_background(columns, truth, seed=112)
    data = synthetic.reader_from_columns(columns)
    data.make_radial_bins(truth['ra_center'], truth['dec_center'], nstars=60, dlogr=0.1)
    background = SingleStars(sample_field(300, seed=5))
    fit = RadialBinsFit(data, model_class=ConstantFit, background=background)
    fit.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    fit.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    n_walkers = 20
    theta = np.stack([synthetic.initial_ball(truth, fit.fitted_parameters, n_walkers, seed=30 + b, scale=0.3)
                      for b in range(fit.n_bins)])
    got = fit.lnprob(theta)
    assert got.shape == (fit.n_bins, n_walkers) and np.all(np.isfinite(got))
    plain = RadialBinsFit(data, model_class=ConstantFit)
    plain.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    plain.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    assert not np.allclose(got, plain.lnprob(theta))                  # the mixture is really evaluated
    for b in (0, fit.n_bins // 2, fit.n_bins - 1):
        single = fit.bin_model(b)
        assert harness.relative_error(got[b], single.lnprob(theta[b])) < 1e-13
        want = harness.oracle_for(single).lnprob_many(theta[b])
        assert harness.relative_error(got[b], want) < 1e-9
    engine = fit(n_walkers=n_walkers, n_steps=40, pos=theta, seed=3)
    assert engine.engine[0] == path
    chain, lnp = engine.chain, engine.lnprobability
    assert chain.shape == (fit.n_bins, n_walkers, 40, fit.n_fitted_parameters) and np.all(np.isfinite(lnp))
    again = fit.lnprob(np.ascontiguousarray(chain[:, :, -1, :]))
    assert np.allclose(again, lnp[:, :, -1], rtol=1e-12, atol=0)
    assert 0.1 < (engine.naccepted / 40.0).mean() < 0.95


def test_segmented_lnprob_with_a_free_centre():
    """Bins that also fit the centre (the free-centre kernels in a segmented launch)."""
    data, truth = synthetic.mock_cluster(1500, seed=14)
    data.make_radial_bins(truth['ra_center'], truth['dec_center'], nstars=80, dlogr=0.1)
    fit = RadialBinsFit(data, model_class=ConstantFit)
    fit.parameters['ra_center'].set(value=truth['ra_center'], fixed=False)
    fit.parameters['dec_center'].set(value=truth['dec_center'], fixed=False)
    assert 'ra_center' in fit.fitted_parameters
    n_walkers = 12
    theta = np.stack([synthetic.initial_ball(truth, fit.fitted_parameters, n_walkers, seed=50 + b, scale=0.2)
                      for b in range(fit.n_bins)])
    got = fit.lnprob(theta)
    for b in (0, fit.n_bins - 1):
        single = fit.bin_model(b)
        assert harness.relative_error(got[b], single.lnprob(theta[b])) < 1e-12
        want = harness.oracle_for(single).lnprob_many(theta[b])
        assert harness.relative_error(got[b], want) < 1e-9
