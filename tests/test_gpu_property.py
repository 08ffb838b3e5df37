"""Property tests (hypothesis): random catalogues, random walker counts and random parameter vectors
inside the priors, every model variant, against the oracle (1e-9 relative), plus C-ABI error paths."""
import ctypes

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from common import RTOL, VARIANTS, build
from mcmc_dynamics_b200 import _native, synthetic
from oracle import harness

pytestmark = pytest.mark.gpu


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(variant=st.sampled_from(VARIANTS), free_centre=st.booleans(), math_mode=st.sampled_from(['fast', 'plain']),
       n_stars=st.integers(1, 700), n_walkers=st.integers(1, 70), seed=st.integers(0, 10_000),
       scale=st.floats(0.01, 0.6))
def test_random_inputs_match_oracle(variant, free_centre, math_mode, n_stars, n_walkers, seed, scale):
    model, oracle, theta, _ = build(variant, n_stars=n_stars, free_centre=free_centre, seed=seed, math_mode=math_mode)
    th = theta(n_walkers, seed=seed + 1, scale=scale)
    got = model.lnprob(th)
    want = oracle.lnprob_many(th)
    assert not np.any(np.isnan(got))
    assert np.array_equal(np.isinf(got), np.isinf(want))
    assert harness.relative_error(got, want) < RTOL


@settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck))
@given(seed=st.integers(0, 1000), v_sys=st.floats(-300, 300), n_walkers=st.integers(2, 40))
def test_offset_velocities_and_linearity_of_shards(seed, v_sys, n_walkers):
    """lnlike is additive over disjoint star sets, for any systemic velocity offset."""
    from mcmc_dynamics_b200.analysis import ConstantFit
    cols, truth = synthetic.mock_cluster(400, seed=seed, as_reader=False, v_sys=v_sys)

    def make(c):
        m = ConstantFit(synthetic.reader_from_columns(c))
        m.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
        m.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
        return m
    whole = make(cols)
    th = synthetic.initial_ball(truth, whole.fitted_parameters, n_walkers, seed=seed)
    a = make({k: v[:123] for k, v in cols.items()}).lnlike(th)
    b = make({k: v[123:] for k, v in cols.items()}).lnlike(th)
    assert harness.relative_error(a + b, whole.lnlike(th)) < 1e-12


def test_c_abi_error_paths(native_lib):
    lib = native_lib
    desc = _native.PackDesc()
    handle = ctypes.c_void_p()
    desc.rotation = 7
    assert lib.mcd_pack_create(ctypes.byref(desc), ctypes.byref(handle)) != 0
    assert b'rotation' in lib.mcd_last_error()
    desc.rotation = 0
    desc.n_theta = 99
    assert lib.mcd_pack_create(ctypes.byref(desc), ctypes.byref(handle)) != 0
    assert b'n_theta' in lib.mcd_last_error()
    desc.n_theta = 1
    desc.n_stars = 5                                     # columns missing
    for k in range(_native.NPARAM):
        desc.slot[k] = -1
    assert lib.mcd_pack_create(ctypes.byref(desc), ctypes.byref(handle)) != 0
    assert b'required' in lib.mcd_last_error()
    assert handle.value is None
    # a working handle rejects bad calls without crashing
    model, _, theta, _ = build('ModelFit', n_stars=64)
    packed = model.pack()
    out = np.zeros(4)
    th = theta(4)
    assert lib.mcd_lnprob(packed.handle, _native.as_double_ptr(th), -1, _native.as_double_ptr(out)) != 0
    assert lib.mcd_lnprob(packed.handle, None, 4, _native.as_double_ptr(out)) != 0
    assert lib.mcd_lnprob(packed.handle, _native.as_double_ptr(th), 0, _native.as_double_ptr(out)) == 0
    with pytest.raises(ValueError):
        packed.lnprob(th[:, :-1])
    with pytest.raises(_native.NativeError):
        packed.membership_per_star(th[0])                # no background component
    # an empty catalogue is legal: lnlike = 0 for accepted walkers
    empty = type(model)(synthetic.reader_from_columns({k: np.zeros(0) for k in ('ra', 'dec', 'v', 'verr')}))
    empty.parameters['ra_center'].set(value=10.0, fixed=True)
    empty.parameters['dec_center'].set(value=0.0, fixed=True)
    th[1, 1] = -1.0
    res = empty.lnprob(th)
    assert res[1] == -np.inf and np.all(np.delete(res, 1) == 0.0)


@settings(max_examples=30, deadline=None, suppress_health_check=list(HealthCheck))
@given(variant=st.sampled_from(VARIANTS), math_mode=st.sampled_from(['fast', 'plain']), seed=st.integers(0, 10_000),
       fixed_mask=st.integers(0, 2 ** 11 - 1), arcmin_units=st.booleans())
def test_random_routing_of_fixed_and_free_parameters(variant, math_mode, seed, fixed_mask, arcmin_units):
    """Any subset of parameters fixed (at arbitrary in-bounds values), a and r_peak optionally
    re-expressed in arcmin: the slot map / unit scales compiled at pack time must reproduce the oracle,
    which resolves parameters per call like the reference (runner.py:143-180)."""
    model, _, theta, truth = build(variant, n_stars=257, free_centre=True, seed=seed, math_mode=math_mode)
    names = list(model.parameters)
    start = theta(1, seed=seed + 3, scale=0.3)[0]
    for j, name in enumerate(model.fitted_parameters):
        if (fixed_mask >> names.index(name)) & 1:
            model.parameters[name].set(value=float(start[j]), fixed=True)
    if len(model.fitted_parameters) == 0:
        model.parameters[names[0]].fixed = False
    oracle = harness.oracle_for(model)
    free = model.fitted_parameters
    th = synthetic.initial_ball(dict(truth, v_back=5.0, sigma_back=55.0, f_back=0.3), free, 9, seed=seed + 5, scale=0.2)
    assert th.shape[1] == len(free)
    got = model.lnprob(th)
    want = oracle.lnprob_many(th)
    assert np.array_equal(np.isinf(got), np.isinf(want))
    assert harness.relative_error(got, want) < RTOL


def test_example_catalogue_of_the_reference():
    """BASELINE config 1: the reference's example/data/test.csv (6284 stars) placed on the sky
    (tests/golden/make_c1_fixture.py), ConstantFit with v_sys fixed as in bin/run.py:487-491."""
    import os
    from mcmc_dynamics_b200.analysis import ConstantFit
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'c1_example_catalogue.npz'))
    data = synthetic.reader_from_columns({k: d[k] for k in ('ra', 'dec', 'v', 'verr')})
    for mode in ('fast', 'plain'):
        m = ConstantFit(data, math_mode=mode)
        m.parameters['ra_center'].set(value=float(d['ra_center']), fixed=True)
        m.parameters['dec_center'].set(value=float(d['dec_center']), fixed=True)
        m.parameters['v_sys'].set(value=0.0, fixed=True)
        rng = np.random.default_rng(4)
        th = np.column_stack([rng.uniform(20, 60, 16), rng.normal(0, 5, 16), rng.normal(0, 5, 16)])
        got = m.lnprob(th)
        want = harness.oracle_for(m).lnprob_many(th)
        assert np.all(np.isfinite(want)) and harness.relative_error(got, want) < RTOL
    # the packed unit vectors reproduce the catalogue's own position angles: v_los = v_maxx sin - v_maxy cos
    theta_pa = d['theta']
    assert np.allclose(np.arctan2(np.sin(theta_pa), np.cos(theta_pa)), theta_pa)


def test_expr_constrained_parameter_is_evaluated_per_walker():
    """``a`` tied to a sampled parameter by an expression (the reference re-evaluates it through asteval
    on every call: analysis/runner.py:163-176, parameter.py:865-874).  The constrained model must return
    what the unconstrained one returns when handed the constrained values explicitly; the oracle agrees."""
    from common import RTOL, build
    from oracle import harness
    model, oracle, theta, truth = build('ModelFit', n_stars=1200)
    free = list(model.fitted_parameters)
    th = theta(24)
    th[:, free.index('a')] = 0.5 * th[:, free.index('r_peak')] + 1.0
    want = model.lnprob(th)
    assert harness.relative_error(want, oracle.lnprob_many(th)) < RTOL
    model.parameters['a'].set(expr='0.5 * r_peak + 1.0')
    assert 'a' not in model.fitted_parameters
    th_c = np.delete(th, free.index('a'), axis=1)
    got = model.lnprob(th_c)
    assert np.array_equal(got, want)
    assert isinstance(model.lnprob(th_c[0]), float) and model.lnprob(th_c[0]) == want[0]
    # the host sampler drives the same path (W >= 2 P with P = 5 sampled parameters)
    run = model(n_walkers=24, n_steps=5, pos=th_c, prefix=None, seed=3)
    assert run.chain.shape == (24, 5, 5) and np.all(np.isfinite(run.lnprobability))
    # the constrained value is bounds-checked like every other parameter (runner.py:207-216)
    model.parameters['a'].set(max=float(np.median(th[:, free.index('a')])))
    out = model.lnprob(th_c)
    inside = th[:, free.index('a')] <= model.parameters['a'].max
    assert np.array_equal(np.isfinite(out), inside) and np.array_equal(out[inside], want[inside])
    # the device-only paths say why they cannot evaluate host-side expressions
    model.parameters['a'].set(max=np.inf)
    import torch
    with pytest.raises(ValueError, match='box priors only'):
        model.lnprob_tensor(torch.as_tensor(th_c, device='cuda:0'))
    with pytest.raises(ValueError, match='box priors only'):
        model(n_walkers=24, n_steps=2, pos=th_c, sampler='device', prefix=None)


def test_inline_and_graph_host_calls_agree(monkeypatch):
    """Small host-buffer calls carry theta inside the kernel arguments and get the result through pinned memory
    and a flag (csrc/mcd_api.cu: host_call_inline); larger ones go through pinned staging and a replayed CUDA
    graph.  Same kernel, same launch geometry: bit-identical results, including -inf for rejected walkers,
    call after call (the completion flag carries a sequence number)."""
    from common import build
    model, oracle, theta, truth = build('ModelFitGB', n_stars=2500, free_centre=True)
    names = model.fitted_parameters
    for n_walkers in (1, 5, 34):                       # 34 x 11 = 374 <= 384 doubles: inline
        th = theta(n_walkers, seed=n_walkers)
        if n_walkers > 2:
            th[2, names.index('sigma_max')] = -1.0
        monkeypatch.delenv('MCD_HOST_CALL', raising=False)
        inline = [model.lnprob(th) for _ in range(4)]
        lnlike_inline = model.lnlike(th)
        monkeypatch.setenv('MCD_HOST_CALL', 'graph')       # staged copy-in -> kernel graph, result + flag in pinned memory
        graph = [model.lnprob(th) for _ in range(4)]
        lnlike_graph = model.lnlike(th)
        monkeypatch.setenv('MCD_HOST_CALL', 'sync')        # copy-in -> kernel -> copy-out graph, stream synchronisation
        sync = [model.lnprob(th) for _ in range(4)]
        for a, b, c in zip(inline, graph, sync):
            assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, inline[0])
        assert np.array_equal(lnlike_inline, lnlike_graph)
        if n_walkers > 2:
            assert inline[0][2] == -np.inf and np.isfinite(lnlike_inline[2])
        assert harness.relative_error(inline[0], oracle.lnprob_many(th)) < RTOL
    monkeypatch.delenv('MCD_HOST_CALL', raising=False)
    # interleaving the two paths and sizes on one handle keeps every result right
    big = theta(64, seed=3)
    small = theta(3, seed=4)
    want_big, want_small = oracle.lnprob_many(big), oracle.lnprob_many(small)
    for _ in range(3):
        assert harness.relative_error(model.lnprob(small), want_small) < RTOL
        assert harness.relative_error(model.lnprob(big), want_big) < RTOL


@pytest.mark.parametrize('variant', ['ModelFit', 'ConstantFit'])
def test_extreme_variances_are_right_or_nan(variant):
    """The FAST no-background kernels multiply the variances of four stars before folding the exponent
    (csrc/mcd_kernels.cu: Accum<BG_NONE, FAST>::end_group): variances within 2^+-250 are exact business as
    usual; beyond that the product may leave the normal range, which must surface as NaN -- never as a wrong
    finite number.  PLAIN arithmetic (the reference's formulas, runner.py:261-271) has no such limit."""
    from common import build
    model, oracle, theta, truth = build(variant, n_stars=1003)
    plain, _, _, _ = build(variant, n_stars=1003, math_mode='plain')
    names = model.fitted_parameters
    th = theta(6, seed=11)
    for row, sigma in enumerate([1e-30, 1e-20, 1e20, 1e35, 1e60, 1e-60]):
        th[row, names.index('sigma_max')] = sigma
    for m in (model, plain):       # the default bounds of sigma_max do not reach that far
        m.parameters['sigma_max'].set(min=0.0, max=1e80)
    want = oracle.lnlike_many(th)
    got = model.lnlike(th)
    assert harness.relative_error(plain.lnlike(th), want) < RTOL
    for g, w in zip(got, want):
        assert np.isnan(g) or abs(g - w) <= RTOL * abs(w)
    # sigma_max = 1e20, 1e35 are inside the documented range (variance 1e40, 1e70 < 2^250 = 1.8e75): exact
    assert np.all(np.isfinite(got[2:4])) and harness.relative_error(got[2:4], want[2:4]) < RTOL
