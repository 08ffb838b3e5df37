"""Host-side boundary, no GPU: parameter/config compatibility, model-class protocol and error
behaviour, the host ensemble sampler, and that the C-ABI library loads and exports every symbol
``include/mcd_b200.h`` declares (no compute calls)."""
import json
import os
import re

import numpy as np
import pytest

from mcmc_dynamics_b200 import DataReader, Parameters, _native, pack, sampler, synthetic
from mcmc_dynamics_b200 import units as u
from mcmc_dynamics_b200.analysis import (ConstantFit, ConstantFitGB, ModelFit, ModelFitGB,
                                         ModelFitConstantBackground)
from oracle import reference_np as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------
# C ABI
# ---------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, 'include', 'mcd_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(mcd_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(native_lib, name), 'libmcd_b200.so does not export ' + name
    assert declared == set(_native.SYMBOLS), declared.symmetric_difference(_native.SYMBOLS)
    assert native_lib.mcd_abi_version() == _native.ABI_VERSION


def test_struct_layout_matches_header():
    """ctypes mirrors of mcd_pack_desc / mcd_info: field order and sizes follow the header."""
    header = open(os.path.join(ROOT, 'include', 'mcd_b200.h')).read()
    body = header[header.index('typedef struct mcd_pack_desc {'):header.index('} mcd_pack_desc;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    fields = re.findall(r'\b(?:int32_t|int64_t|double|const double \*|const int64_t \*)\s*\*?(\w+)(?:\[\w+\])?;', body)
    assert fields == [name for name, _ in _native.PackDesc._fields_]
    import ctypes
    assert ctypes.sizeof(_native.PackDesc) == 16 + 8 + 7 * 8 + 11 * 4 + 4 + 2 * 11 * 8 + 2 * 16 * 8 + 8 + 8 + 8 + 8


@pytest.mark.skipif(__import__('torch').cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_gpu_fails_loudly():
    data, truth = synthetic.mock_cluster(50, seed=1)
    model = ModelFit(data)
    with pytest.raises(_native.NativeError, match='no CUDA device|CPU fallback'):
        model.lnprob(model.get_initials(4))


# ---------------------------------------------------------------------------------------------
# parameters / config
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('cls,table', [(ConstantFit, 'constant'), (ConstantFitGB, 'constant_with_background'),
                                       (ModelFit, 'model'), (ModelFitGB, 'model_with_background'),
                                       (ModelFitConstantBackground, 'model_with_background')])
def test_default_parameter_files_match_reference_tables(cls, table):
    pars = cls.default_parameters()
    want = ref.DEFAULT_TABLES[table]
    assert list(pars.keys()) == [row[0] for row in want]
    for name, unit, lo, hi in want:
        par = pars[name]
        assert par.min == lo and par.max == hi and not par.fixed
        assert (par.unit is None and unit is None) or str(par.unit).replace(' ', '') == unit
    # value = (min + max) / 2 if both bounds finite else 0 (parameter.py:794-798)
    assert pars['ra_center'].value == 180.0 and pars['dec_center'].value == 0.0


def test_parameters_json_round_trip(tmp_path):
    pars = ModelFit.default_parameters()
    pars['ra_center'].set(value=201.697, fixed=True)
    pars['a'].set(value=u.Quantity(0.5, u.arcmin))          # converted to the parameter's unit
    assert pars['a'].value == pytest.approx(30.0)
    text = pars.dumps()
    doc = json.loads(text.replace('Infinity', '1e999'))
    assert [row[0] for row in doc['params']] == list(pars.keys())
    again = Parameters().loads(text)
    for name in pars:
        assert again[name].value == pars[name].value and again[name].fixed == pars[name].fixed
        assert again[name].min == pars[name].min and again[name].max == pars[name].max
    path = tmp_path / 'p.json'
    with open(path, 'w') as f:
        pars.dump(f)
    assert Parameters().load(str(path))['ra_center'].fixed


def test_initials_and_lnprior_expressions():
    pars = Parameters(rng_seed=7)
    pars.add('sigma_max', unit='km/s', min=0, initials='rng.lognormal(mean=2.3, sigma=0.5, size=n)')
    pars.add('v_sys', unit='km/s', value=3.0, min=-10, max=10, lnprior='norm.logpdf(val, loc=0, scale=2)')
    start = pars['sigma_max'].evaluate_initials(100)
    assert start.shape == (100,) and np.all(start > 0)
    assert pars['v_sys'].evaluate_lnprior(11.0) == -np.inf
    from scipy import stats
    assert pars['v_sys'].evaluate_lnprior(1.0) == pytest.approx(stats.norm.logpdf(1.0, 0, 2))
    draws = pars['v_sys'].evaluate_initials(2000)             # truncnorm(loc=value, scale=1) in [-10, 10]
    assert abs(draws.mean() - 3.0) < 0.1 and draws.min() >= -10 and draws.max() <= 10


# ---------------------------------------------------------------------------------------------
# model-class protocol (no GPU work)
# ---------------------------------------------------------------------------------------------
def test_constructor_errors_follow_the_reference():
    data, _ = synthetic.mock_cluster(20, seed=1)
    with pytest.raises(AssertionError):
        ModelFit({'v': [1.0]})                                 # not a DataReader (runner.py:66)
    with pytest.raises(AssertionError):
        ModelFit(data, bogus=1)                                # unknown kwargs (runner.py:56)
    with pytest.raises(IOError):
        ModelFit(DataReader({'v': [1.0], 'verr': [1.0]}))      # no coordinates (runner.py:72)
    pars = ConstantFit.default_parameters()
    with pytest.raises(IOError):
        ModelFit(data, parameters=pars)                        # missing a, r_peak (runner.py:89)
    with pytest.raises(AssertionError):
        ModelFitGB(data)                                       # density column missing (runner.py:76)


def test_fitted_parameters_and_batched_prior():
    data, truth = synthetic.mock_cluster(20, seed=1)
    model = ModelFit(data)
    assert model.fitted_parameters == ['v_sys', 'sigma_max', 'a', 'v_maxx', 'ra_center', 'dec_center', 'v_maxy',
                                       'r_peak']
    model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    assert model.n_fitted_parameters == 6 and model.n_data == 20
    theta = synthetic.initial_ball(truth, model.fitted_parameters, 5)
    theta[1, 1] = -1.0
    lp = model.lnprior(theta)
    assert lp.shape == (5,) and lp[1] == -np.inf and np.all(lp[[0, 2, 3, 4]] == 0)
    assert model.lnprior(theta[0]) == 0 and model.lnprior(theta[1]) == -np.inf
    values = model.fetch_parameter_values(theta[0])
    assert list(values) == list(model.parameters) and values['ra_center'].value == truth['ra_center']
    with pytest.raises(AssertionError):
        model.fetch_parameter_values(np.append(theta[0], 1.0))     # runner.py:178
    # oracle prior agrees row by row
    from oracle import harness
    assert np.array_equal(harness.oracle_for(model).lnprior_many(theta), lp)


def test_batched_box_prior_matches_the_oracle_on_random_parameter_tables():
    """Property test of the host prior (runner.py:206-217, parameter.py:691-692): random bounds, random
    fixed / free splits -- fixed parameters are bounds-checked too -- and walkers placed inside, outside and
    exactly on the bounds (inclusive)."""
    from hypothesis import given, settings, strategies as st
    from oracle import harness
    data, truth = synthetic.mock_cluster(12, seed=1)
    names = ['v_sys', 'sigma_max', 'a', 'v_maxx', 'ra_center', 'dec_center', 'v_maxy', 'r_peak']

    bound = st.one_of(st.just(None), st.floats(-50.0, 50.0, allow_nan=False).map(lambda x: round(x, 3)))
    row = st.tuples(st.booleans(), bound, bound, st.sampled_from(['inside', 'below', 'above', 'at_min', 'at_max']))

    @settings(max_examples=60, deadline=None)
    @given(st.lists(row, min_size=len(names), max_size=len(names)), st.integers(0, 2 ** 31 - 1))
    def check(rows, seed):
        rng = np.random.default_rng(seed)
        model = ModelFit(data)
        placements = {}
        for name, (fixed, lo, hi, where) in zip(names, rows):
            lo, hi = (-np.inf if lo is None else lo), (np.inf if hi is None else hi)
            if lo > hi:
                lo, hi = hi, lo
            if lo == hi:                                    # min == max raises, as in the reference (parameter.py:800-801)
                hi = lo + 1.0
            centre = 0.5 * (lo + hi) if np.isfinite(lo) and np.isfinite(hi) else (lo + 1.0 if np.isfinite(lo) else (
                hi - 1.0 if np.isfinite(hi) else 0.0))
            model.parameters[name].set(value=centre, min=lo, max=hi, fixed=fixed)
            placements[name] = (lo, hi, centre, where)
        if model.n_fitted_parameters == 0:
            return
        theta = np.empty((6, model.n_fitted_parameters))
        for j, name in enumerate(model.fitted_parameters):
            lo, hi, centre, where = placements[name]
            theta[:, j] = centre
            k = int(rng.integers(0, 6))                     # one walker per parameter leaves the middle
            if where == 'below' and np.isfinite(lo):
                theta[k, j] = lo - abs(lo) * 1e-12 - 1e-9
            elif where == 'above' and np.isfinite(hi):
                theta[k, j] = hi + abs(hi) * 1e-12 + 1e-9
            elif where == 'at_min' and np.isfinite(lo):
                theta[k, j] = lo
            elif where == 'at_max' and np.isfinite(hi):
                theta[k, j] = hi
        got = model.lnprior(theta)
        want = harness.oracle_for(model).lnprior_many(theta)
        assert np.array_equal(got, want), (rows, theta, got, want)
        assert model.lnprior(theta[0]) == want[0]
    check()


def test_descriptor_routing_and_units():
    data, truth = synthetic.mock_cluster(20, seed=1)
    model = ModelFitGB(synthetic.reader_from_columns(dict(
        synthetic.mock_cluster(20, seed=1, as_reader=False)[0], density=np.ones(20))))
    model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    model.parameters['v_sys'].set(value=1.0, fixed=True)
    desc, keep = model._descriptor()
    free = model.fitted_parameters
    assert desc.n_theta == len(free) == 8 and desc.n_stars == 20
    slots = dict(zip(_native.PARAM_SLOTS, list(desc.slot)))
    assert slots['v_sys'] == -1 and slots['ra_center'] == -1
    assert slots['sigma_max'] == free.index('sigma_max') and slots['f_back'] == free.index('f_back')
    scale = dict(zip(_native.PARAM_SLOTS, list(desc.unit_scale)))
    assert scale['a'] == pytest.approx(1 / 60.0) and scale['r_peak'] == pytest.approx(1 / 60.0) and scale['v_sys'] == 1.0
    assert desc.lower[free.index('sigma_max')] == 0.0 and desc.upper[free.index('f_back')] == 1.0
    assert desc.fixed_prior_ok == 1
    model.parameters['v_sys'].min = 2.0                       # fixed value 1.0 now violates its bounds
    assert model._descriptor()[0].fixed_prior_ok == 0
    model.parameters['v_sys'].min = -np.inf
    # per-walker constraint (reference: asteval per call, analysis/runner.py:163-176): the constrained
    # parameter leaves theta and comes back as an extra kernel column evaluated on the host per walker
    model.parameters['a'].set(expr='2 * r_peak')
    free = model.fitted_parameters
    assert 'a' not in free and pack.derived_parameters(model.parameters) == ['a']
    desc, keep = model._descriptor()
    assert desc.n_theta == len(free) + 1
    slots = dict(zip(_native.PARAM_SLOTS, list(desc.slot)))
    assert slots['a'] == len(free) and slots['r_peak'] == free.index('r_peak')
    assert desc.lower[len(free)] == 0.0 and desc.upper[len(free)] == np.inf      # bounds of `a` itself
    theta = np.abs(np.random.default_rng(0).normal(size=(5, len(free)))) + 0.1
    theta[:, free.index('f_back')] = 0.5
    full = model._device_theta(theta)
    assert full.shape == (5, len(free) + 1)
    assert np.array_equal(full[:, -1], 2 * theta[:, free.index('r_peak')])
    # expressions the vectorised evaluation cannot handle fall back to one evaluation per walker
    model.parameters['a'].set(expr='max(2 * r_peak, 1.0)')
    model._derived = pack.derived_parameters(model.parameters)
    assert np.array_equal(model._device_theta(theta)[:, -1], np.maximum(2 * theta[:, free.index('r_peak')], 1.0))
    # the prior bounds-checks the constrained value per walker (parameter.py:691-692 via runner.py:207-216)
    model.parameters['a'].set(expr='2 * r_peak', max=1.0)
    lnp = model.lnprior(theta)
    assert np.array_equal(np.isfinite(lnp), 2 * theta[:, free.index('r_peak')] <= 1.0)
    # a constraint on fixed parameters only stays a plain fixed value
    model.parameters['r_peak'].set(value=3.0, fixed=True)
    assert pack.derived_parameters(model.parameters) == []
    assert model._descriptor()[0].n_theta == len(model.fitted_parameters)


def test_default_parameter_files_do_not_freeze_the_generator():
    """The reference's config/*.json carry ``rng_seed: null`` and no ``random_state``
    (config/model.json:1-5), so initials drawn from the defaults differ from object to object
    (parameter.py:73-74); ``Parameters(rng_seed=...)`` stays reproducible (parameter.py:207-209)."""
    import json
    from mcmc_dynamics_b200 import config
    from mcmc_dynamics_b200.parameter import Parameters
    for name in config.SETS:
        with open(config.default_file(name)) as f:
            state = json.load(f)
        assert 'random_state' not in state and state['unique_symbols'] == {'rng_seed': None}
    a = Parameters().load(open(config.default_file('model')))
    b = Parameters().load(open(config.default_file('model')))
    assert not np.array_equal(a['v_sys'].evaluate_initials(8), b['v_sys'].evaluate_initials(8))
    c = Parameters(rng_seed=5).load(open(config.default_file('model')))
    d = Parameters(rng_seed=5).load(open(config.default_file('model')))
    assert np.array_equal(c['v_sys'].evaluate_initials(8), d['v_sys'].evaluate_initials(8))
    # a user's own snapshot keeps the generator state, as in the reference (parameter.py:458-466)
    snap = json.loads(c.dumps())
    assert snap['random_state'] is not None
    e = Parameters().loads(c.dumps())
    assert np.array_equal(c['sigma_max'].evaluate_initials(4), e['sigma_max'].evaluate_initials(4))


def test_superfluous_free_parameters_are_prior_checked_only():
    """ModelFitConstantBackground loads model_with_background.json whose v_back, sigma_back are not
    model parameters (model.py:526,529): they stay free dimensions of theta."""
    cols = dict(synthetic.mock_cluster(20, seed=1, as_reader=False)[0], density=np.ones(20))

    class Flat(object):                                       # background callable evaluated on the host
        def __call__(self, v, verr):
            return np.full(len(np.asarray(getattr(v, 'value', v))), -5.0)
    model = ModelFitConstantBackground(synthetic.reader_from_columns(cols), background=Flat())
    assert 'v_back' in model.fitted_parameters and 'v_back' not in model.MODEL_PARAMETERS
    desc, _ = model._descriptor()
    assert desc.n_theta == 11 and dict(zip(_native.PARAM_SLOTS, list(desc.slot)))['v_back'] == -1
    assert desc.lower[model.fitted_parameters.index('sigma_back')] == 0.0


# ---------------------------------------------------------------------------------------------
# host ensemble sampler (emcee stand-in), vectorised calls
# ---------------------------------------------------------------------------------------------
def test_host_sampler_recovers_a_gaussian():
    mean = np.array([1.0, -2.0, 0.5])
    sigma = np.array([0.5, 2.0, 1.0])
    calls = []

    def lnprob(theta):
        calls.append(theta.shape)
        return -0.5 * np.sum(((theta - mean) / sigma) ** 2, axis=1)
    s = sampler.HostEnsembleSampler(32, 3, lnprob, seed=3)
    rng = np.random.default_rng(0)
    pos, lnp, state = s.run_mcmc(mean + 0.1 * rng.standard_normal((32, 3)), 1500)
    assert s.chain.shape == (32, 1500, 3) and s.lnprobability.shape == (32, 1500) and s.iteration == 1500
    assert calls[0] == (32, 3) and set(calls[1:]) == {(16, 3)}      # W once, then half-ensembles
    flat = s.chain[:, 500:, :].reshape(-1, 3)
    assert np.allclose(flat.mean(axis=0), mean, atol=0.15)
    assert np.allclose(flat.std(axis=0), sigma, rtol=0.15)
    assert 0.2 < s.acceptance_fraction.mean() < 0.9
    with pytest.raises(RuntimeError):
        sampler.HostEnsembleSampler(4, 3, lnprob)                    # nwalkers < 2 ndim
    with pytest.raises(ValueError):
        s.compute_log_prob(np.full((2, 3), np.nan))


def test_host_sampler_error_semantics_and_resume():
    """emcee's contract as ``analysis/runner.py:416-419`` relies on it: ``-inf`` is a legal value (never
    accepted), NaN raises, non-finite coordinates raise with emcee's messages, and a run continued with the
    returned ``(pos, log_prob, random_state)`` equals one uninterrupted run."""
    def lnprob(theta):
        out = -0.5 * np.sum(theta ** 2, axis=1)
        out[theta[:, 0] > 1.0] = -np.inf                          # hard wall: proposals beyond it are rejected
        return out
    start = 0.3 * np.random.default_rng(1).standard_normal((12, 2)) - 0.5
    whole = sampler.HostEnsembleSampler(12, 2, lnprob, seed=5)
    whole.run_mcmc(start, 400)
    assert whole.chain[:, :, 0].max() <= 1.0 and np.isfinite(whole.lnprobability).all()
    assert whole.naccepted.sum() > 0 and np.all(whole.naccepted <= 400)

    parts = sampler.HostEnsembleSampler(12, 2, lnprob, seed=5)
    pos, lnp, state = parts.run_mcmc(start, 150)
    parts.run_mcmc(pos, 250, log_prob0=lnp, rstate0=state)
    assert np.array_equal(parts.chain, whole.chain) and np.array_equal(parts.lnprobability, whole.lnprobability)
    assert np.array_equal(parts.naccepted, whole.naccepted)
    assert parts.n_log_prob_calls == whole.n_log_prob_calls == 1 + 2 * 400

    odd = sampler.HostEnsembleSampler(13, 2, lnprob, seed=5)     # halves of 7 and 6 walkers
    odd.run_mcmc(np.vstack([start, [[-0.2, 0.1]]]), 50)
    assert odd.chain.shape == (13, 50, 2)

    def sometimes_nan(theta):
        out = lnprob(theta)
        if sometimes_nan.armed:
            out[-1] = np.nan
        return out
    sometimes_nan.armed = False
    bad = sampler.HostEnsembleSampler(12, 2, sometimes_nan, seed=5)
    bad.run_mcmc(start, 3)
    sometimes_nan.armed = True
    with pytest.raises(ValueError, match='returned NaN'):
        bad.run_mcmc(start, 3)
    with pytest.raises(ValueError, match='infinite'):
        whole.compute_log_prob(np.array([[0.0, np.inf]]))
    with pytest.raises(ValueError, match='NaN'):
        whole.compute_log_prob(np.array([[0.0, np.nan]]))
    huge = np.array([[1e200, -1e200]])                            # squares overflow, the entries are finite: accepted
    assert whole.compute_log_prob(huge).shape == (1,)
    with pytest.raises(ValueError, match='initial log_prob was NaN'):
        sampler.HostEnsembleSampler(12, 2, lnprob, seed=5).run_mcmc(start, 1, log_prob0=np.full(12, np.nan))


def test_chain_helpers_follow_the_reference(tmp_path):
    """``sample_chain`` (analysis/runner.py:820-850), the deprecated ``save_chain`` (:445-455) and the curve
    signatures the reference keeps on its model objects (constant.py:49-50, model.py:90-91): host-side only."""
    data, truth = synthetic.mock_cluster(40, seed=3)
    model = ModelFit(data, seed=7)
    model.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    model.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    assert list(model.rotation_parameters) == ['v_sys', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center', 'r_peak', 'kwargs']
    assert list(model.dispersion_parameters) == ['sigma_max', 'ra_center', 'dec_center', 'a', 'kwargs']
    assert list(ConstantFit(data).rotation_parameters) == ['v_sys', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center', 'kwargs']
    n_free = model.n_fitted_parameters
    chain = np.arange(5 * 9 * n_free, dtype=np.float64).reshape(5, 9, n_free)        # [walkers, steps, free]
    np.random.seed(11)
    samples = model.sample_chain(chain, n_burn=4, n_samples=6)
    np.random.seed(11)
    rows = chain[:, 4:].reshape(-1, n_free)[np.random.randint(0, 25, (6,))]
    assert len(samples) == 6
    for sample, row in zip(samples, rows):
        assert list(sample) == list(model.parameters)
        assert [sample[name].value for name in model.fitted_parameters] == list(row)
        assert sample['ra_center'].value == truth['ra_center'] and str(sample['a'].unit) == 'arcsec'

    class Stub(object):
        chain = np.zeros((2, 3, n_free))
        lnprobability = np.ones((2, 3))
    with pytest.warns(DeprecationWarning):
        model.save_chain(Stub(), filename=str(tmp_path / 'runchain.pkl'))
    assert np.array_equal(model.read_chain(str(tmp_path / 'run_chain.pkl')), Stub.chain)
    assert np.array_equal(model.read_chain(str(tmp_path / 'run_lnprob.pkl')), Stub.lnprobability)


def test_radial_bins_partition():
    data, truth = synthetic.mock_cluster(600, seed=2)
    data.make_radial_bins(truth['ra_center'], truth['dec_center'], nstars=50, dlogr=0.1)
    bins = np.asarray(data.data['bin'].value if hasattr(data.data['bin'], 'value') else data.data['bin'])
    assert bins.min() == 0 and len(bins) == 600
    sizes = np.bincount(bins.astype(int))
    assert np.all(sizes >= 25) and sizes.sum() == 600
    sub = data.fetch_radial_bin(1)
    assert sub.sample_size == sizes[1]


def test_parameter_edit_stamps_detect_every_assignment():
    """``Runner.pack`` compares the parameters' version stamps per call instead of rebuilding the full
    routing signature: reading never changes them, any assignment does."""
    from mcmc_dynamics_b200 import pack
    from mcmc_dynamics_b200.analysis import ModelFit
    parameters = ModelFit.default_parameters()
    before = pack.edit_stamps(parameters)
    signature = pack.routing_signature(parameters, ModelFit.MODEL_PARAMETERS)
    for par in parameters.values():
        par.value, par.unit, par.fixed, par.min, par.max          # reads
    assert pack.edit_stamps(parameters) == before
    assert pack.routing_signature(parameters, ModelFit.MODEL_PARAMETERS) == signature
    name = next(iter(parameters))
    parameters[name].set(fixed=not parameters[name].fixed)
    after = pack.edit_stamps(parameters)
    assert after != before
    parameters[name].max = 1e6
    assert pack.edit_stamps(parameters) != after
