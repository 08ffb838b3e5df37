"""GPU parity: every kernel variant, through the C ABI, against the NumPy oracle on the same seeded
inputs (tolerance: 1e-9 relative, BASELINE.json north_star)."""
import numpy as np
import pytest

from common import RTOL, VARIANTS, build
from oracle import harness

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('math_mode', ['fast', 'plain'])
@pytest.mark.parametrize('free_centre', [False, True])
@pytest.mark.parametrize('variant', VARIANTS)
def test_lnprob_matches_oracle(variant, free_centre, math_mode):
    model, oracle, theta, _ = build(variant, n_stars=3001, free_centre=free_centre, math_mode=math_mode)
    th = theta(48)
    got = model.lnprob(th)
    want = oracle.lnprob_many(th)
    assert np.all(np.isfinite(want))
    err = harness.relative_error(got, want)
    assert err < RTOL, (variant, free_centre, math_mode, err)
    # lnlike (no prior) agrees too, and the scalar call returns a float
    got1 = model.lnlike(th[3])
    assert isinstance(got1, float)
    assert abs(got1 - oracle.lnlike(th[3])) <= RTOL * max(1.0, abs(got1))


@pytest.mark.parametrize('n_walkers', [1, 2, 7, 16, 50, 100, 257, 600])
def test_walker_counts(n_walkers):
    model, oracle, theta, _ = build('ModelFit', n_stars=1500)
    th = theta(n_walkers)
    err = harness.relative_error(model.lnprob(th), oracle.lnprob_many(th))
    assert err < RTOL


@pytest.mark.parametrize('n_stars', [1, 2, 15, 16, 17, 255, 256, 257, 1000, 4097])
def test_ragged_star_counts(n_stars):
    model, oracle, theta, _ = build('ModelFitGB', n_stars=n_stars, free_centre=True)
    th = theta(20)
    err = harness.relative_error(model.lnprob(th), oracle.lnprob_many(th))
    assert err < RTOL


def test_prior_rejection_is_exact_minus_inf():
    model, oracle, theta, _ = build('ModelFit', n_stars=500)
    th = theta(16)
    names = model.fitted_parameters
    th[2, names.index('sigma_max')] = -1.0          # below min = 0
    th[5, names.index('a')] = -0.5
    got = model.lnprob(th)
    want = oracle.lnprob_many(th)
    assert got[2] == -np.inf and got[5] == -np.inf
    assert harness.relative_error(got, want) < RTOL
    # bounds are inclusive (parameter.py:691-692)
    model.parameters['v_sys'].set(min=-3.0, max=3.0)
    th2 = theta(4)
    th2[:, names.index('v_sys')] = [3.0, -3.0, 3.0000001, 0.0]
    got2 = model.lnprob(th2)
    assert np.isfinite(got2[0]) and np.isfinite(got2[1]) and got2[2] == -np.inf and np.isfinite(got2[3])
    # a fixed parameter outside its bounds rejects everything (runner.py:207-216)
    model.parameters['v_sys'].set(value=0.0, fixed=True)
    model.parameters['v_sys'].min = 1.0
    model.parameters['v_sys'].max = 2.0
    th3 = theta(4)                      # v_sys is fixed now: one column fewer
    assert th3.shape[1] == len(names) - 1
    assert np.all(model.lnprob(th3) == -np.inf)


def test_repack_after_parameter_edit():
    model, oracle, theta, truth = build('ModelFit', n_stars=800)
    th = theta(8)
    a = model.lnprob(th)
    # fix v_sys: theta loses a column and the routing must be recompiled
    names = model.fitted_parameters
    model.parameters['v_sys'].set(value=0.5, fixed=True)
    oracle2 = harness.oracle_for(model)
    th2 = np.delete(th, names.index('v_sys'), axis=1)
    b = model.lnprob(th2)
    assert harness.relative_error(b, oracle2.lnprob_many(th2)) < RTOL
    assert not np.allclose(a, b)
    # move the fixed centre: packed geometry must follow
    model.parameters['ra_center'].set(value=truth['ra_center'] + 0.01)
    oracle3 = harness.oracle_for(model)
    assert harness.relative_error(model.lnprob(th2), oracle3.lnprob_many(th2)) < RTOL


def test_units_of_parameters_are_honoured():
    """a and r_peak may legally be given in another angular unit."""
    from mcmc_dynamics_b200 import units as u
    from mcmc_dynamics_b200.parameter import Parameters
    model, oracle, theta, _ = build('ModelFit', n_stars=600)
    th = theta(6)
    ref_val = model.lnprob(th)
    pars = Parameters().load(model.parameters_file)
    pars2 = Parameters()
    for name, par in pars.items():
        unit = 'arcmin' if name in ('a', 'r_peak') else par.unit
        pars2.add(name, unit=unit, fixed=par.fixed, min=par.min, max=par.max, initials=par.initials)
    pars2['ra_center'].set(value=model.parameters['ra_center'].value, fixed=True)
    pars2['dec_center'].set(value=model.parameters['dec_center'].value, fixed=True)
    model2 = type(model)(model.data, parameters=pars2)
    names = model2.fitted_parameters
    th_arcmin = th.copy()
    for name in ('a', 'r_peak'):
        th_arcmin[:, names.index(name)] /= 60.0
    assert harness.relative_error(model2.lnprob(th_arcmin), ref_val) < RTOL
    assert harness.relative_error(model2.lnprob(th_arcmin), harness.oracle_for(model2).lnprob_many(th_arcmin)) < RTOL


def test_no_sum_per_star():
    model, oracle, theta, _ = build('ModelFitConstantBackground', n_stars=700)
    th = theta(3)
    got = model.lnlike(th[1], no_sum=True)
    want = oracle.lnlike(th[1], no_sum=True)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < RTOL
    assert abs(got.sum() - model.lnlike(th[1])) < 1e-9 * abs(got.sum())


def test_mixture_corner_cases():
    """p = 1, p = 0 and hopeless members: -inf / finite exactly where the reference has them
    (runner.py:280-286, SURVEY.md 8c KAT 4)."""
    import numpy as np
    from mcmc_dynamics_b200 import synthetic
    from mcmc_dynamics_b200.analysis import ModelFit
    from mcmc_dynamics_b200.background import Gaussian
    columns, truth = synthetic.mock_cluster(64, seed=3, as_reader=False)
    columns['pmember'] = np.full(64, 0.5)
    columns['pmember'][0] = 1.0
    columns['pmember'][1] = 0.0
    columns['v'][2] = 4000.0          # member likelihood underflows by far; background carries it
    columns['pmember'][3] = 1.0
    columns['v'][3] = 3000.0          # p = 1 and member term ~ -1e5 below the background: -inf
    for mode in ('fast', 'plain'):
        model = ModelFit(synthetic.reader_from_columns(columns), background=Gaussian(0.0, 500.0), math_mode=mode)
        model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
        model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
        th = synthetic.initial_ball(truth, model.fitted_parameters, 4, seed=1)
        oracle = harness.oracle_for(model)
        want = oracle.lnprob_many(th)
        got = model.lnprob(th)
        assert np.all(want == -np.inf) and np.all(got == -np.inf), (mode, got, want)
        columns2 = dict(columns)
        columns2['pmember'] = columns['pmember'].copy()
        columns2['pmember'][3] = 0.999
        model = ModelFit(synthetic.reader_from_columns(columns2), background=Gaussian(0.0, 500.0), math_mode=mode)
        model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
        model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
        oracle = harness.oracle_for(model)
        want = oracle.lnprob_many(th)
        got = model.lnprob(th)
        assert np.all(np.isfinite(want))
        assert harness.relative_error(got, want) < RTOL, mode


def test_torch_ops_match_host_entry_points():
    import torch
    model, oracle, theta, _ = build('ModelFit', n_stars=2000, free_centre=True)
    th = theta(64)
    host = model.lnprob(th)
    dev = model.lnprob_tensor(torch.as_tensor(th, device='cuda:0'))
    assert dev.is_cuda and dev.dtype == torch.float64
    assert np.array_equal(dev.cpu().numpy(), host)        # same kernel, same reduction order
    part = model.pack().lnprob_partial_tensor(torch.as_tensor(th, device='cuda:0'))
    assert np.array_equal(part.cpu().numpy(), host)


def test_large_catalogue_properties():
    """At a size the oracle cannot finish quickly: additivity over star shards and agreement of the
    two arithmetic variants (size-independent properties)."""
    from mcmc_dynamics_b200 import synthetic
    from mcmc_dynamics_b200.analysis import ModelFit
    n = 1_000_000
    columns, truth = synthetic.mock_cluster(n, seed=9, as_reader=False)

    def make(cols, mode='fast'):
        m = ModelFit(synthetic.reader_from_columns(cols), math_mode=mode)
        m.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
        m.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
        return m
    whole = make(columns)
    th = synthetic.initial_ball(truth, whole.fitted_parameters, 64, seed=2)
    total = whole.lnprob(th)
    cut = 377_123
    parts = sum(make({k: v[s] for k, v in columns.items()}).lnprob(th) for s in (slice(0, cut), slice(cut, n)))
    assert harness.relative_error(parts, total) < 1e-12
    plain = make(columns, 'plain').lnprob(th)
    assert harness.relative_error(plain, total) < 1e-11
    # oracle on a subsample of walkers (one literal NumPy pass per walker)
    oracle = harness.oracle_for(whole)
    assert harness.relative_error(total[:3], oracle.lnprob_many(th[:3])) < RTOL
