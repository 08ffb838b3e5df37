"""The oracle against the golden vectors produced by the reference's own source files
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare the CUDA
path with the oracle and, directly, with the same vectors (test_gpu_golden.py)."""
import numpy as np
import pytest

import golden_util
from oracle import reference_np as ref

GOLDEN = golden_util.load()
TOL = 1e-12        # restatement vs reference source, both float64 NumPy: rounding-order differences only


@pytest.mark.parametrize('case', GOLDEN['cases'], ids=[c['name'] for c in GOLDEN['cases']])
def test_oracle_reproduces_reference_outputs(case):
    oracle = golden_util.oracle_for_case(case)
    assert oracle.fitted_parameters == case['fitted_parameters']
    theta = np.asarray(case['theta'])
    exp = case['expected']
    with np.errstate(all='ignore'):
        for k, row in enumerate(theta):
            assert oracle.lnprior(row) == exp['lnprior'][k]
            if np.isfinite(exp['lnprior'][k]):
                assert oracle.lnlike(row) == pytest.approx(exp['lnlike'][k], rel=TOL, abs=0)
            got = oracle.lnprob(row)
            if np.isfinite(exp['lnprob'][k]):
                assert got == pytest.approx(exp['lnprob'][k], rel=TOL, abs=0)
            else:
                assert got == exp['lnprob'][k]
    assert any(np.isfinite(exp['lnprob']))
    if 'lnlike_background' in case:
        assert np.allclose(oracle.lnlike_background, case['lnlike_background'], rtol=1e-12, atol=0)
    if 'lnlike_per_star_theta0' in case:
        per_star = oracle.lnlike(theta[0], no_sum=True)
        assert np.allclose(per_star, case['lnlike_per_star_theta0'], rtol=1e-12, atol=1e-13)
    if 'membership_theta0' in case:
        with np.errstate(all='ignore'):
            assert np.allclose(oracle.membership(theta[0]), case['membership_theta0'], rtol=1e-11, atol=1e-15)


@pytest.mark.parametrize('case', GOLDEN['cases'], ids=[c['name'] for c in GOLDEN['cases']])
def test_oracle_model_curves_match_the_reference(case):
    """rotation_model / dispersion_model of the reference's own classes at theta[0] (constant.py:52-111,
    model.py:93-180), km/s per star."""
    oracle = golden_util.oracle_for_case(case)
    v_los, sigma_los = oracle._models(oracle.fetch_parameter_values(np.asarray(case['theta'][0])))
    want = case['model_curves_theta0']
    assert len(want['v_los']) == len(want['sigma_los']) == oracle.n_data
    assert np.allclose(v_los, want['v_los'], rtol=1e-12, atol=1e-13)
    assert np.allclose(sigma_los, want['sigma_los'], rtol=1e-12, atol=0)


def test_membership_vectors_cover_the_three_classes_that_define_them():
    have = {c['class'] for c in GOLDEN['cases'] if 'membership_theta0' in c}
    assert have == {'ConstantFitGB', 'ModelFitGB', 'ModelFitConstantBackground'}


def test_prior_rejections_are_present_in_the_vectors():
    rejected = [c['name'] for c in GOLDEN['cases'] if -np.inf in c['expected']['lnprob']]
    assert len(rejected) >= 2


def test_geometry_and_backgrounds():
    ex = GOLDEN['extras']
    g = ex['calc_xy_offset']
    dx, dy = ref.calc_xy_offset(np.asarray(g['ra']), np.asarray(g['dec']), g['ra_center'], g['dec_center'])
    assert np.allclose(dx, g['dx_arcmin'], rtol=1e-12, atol=1e-15)
    assert np.allclose(dy, g['dy_arcmin'], rtol=1e-12, atol=1e-15)
    s = ex['single_stars']
    v, verr = np.asarray(s['v']), np.asarray(s['verr'])
    assert np.allclose(ref.single_stars_background(np.asarray(s['v_bg']), v, verr), s['lnlike'], rtol=1e-13, atol=0)
    assert np.allclose(ref.single_stars_background(np.asarray(s['v_bg']), v, verr, sigma_int=3.0),
                       s['lnlike_sigma_int_3'], rtol=1e-13, atol=0)
    assert np.allclose(ref.gaussian_background(v, verr, ex['gaussian']['mean'], ex['gaussian']['sigma']),
                       ex['gaussian']['lnlike'], rtol=1e-13, atol=0)


def test_default_parameter_tables_as_loaded_by_the_reference():
    """config/*.json through the reference's own Parameters.load: names, order, defaults, units,
    bounds -- against the oracle's tables and the product's Parameters."""
    from mcmc_dynamics_b200 import analysis
    table_of = {'ConstantFit': 'constant', 'ConstantFitGB': 'constant_with_background', 'ModelFit': 'model',
                'ModelFitGB': 'model_with_background', 'ModelFitConstantBackground': 'model_with_background'}
    for cls_name, rows in GOLDEN['default_parameters'].items():
        oracle_rows = ref.default_params(table_of[cls_name])
        product = getattr(analysis, cls_name).default_parameters()
        assert [r[0] for r in rows] == [p.name for p in oracle_rows] == list(product)
        for row, o in zip(rows, oracle_rows):
            name, value, unit, fixed, lo, hi, initials = row
            p = product[name]
            assert o.value == value == p.value and o.fixed == fixed == p.fixed
            assert o.min == lo == p.min and o.max == hi == p.max
            assert (unit or '').replace(' ', '') == (o.unit or '') == ('' if p.unit is None else str(p.unit).replace(' ', ''))
            assert p.initials == initials


# ---------------------------------------------------------------------------------------------
# the callers either side of the path: radial binning and chain post-processing
# ---------------------------------------------------------------------------------------------
def test_make_radial_bins_matches_reference():
    from mcmc_dynamics_b200 import synthetic
    for case in GOLDEN['post']['make_radial_bins']:
        data, truth = synthetic.mock_cluster(case['n_stars'], seed=case['seed'])
        data.make_radial_bins(truth['ra_center'], truth['dec_center'], nstars=case['nstars'], dlogr=case['dlogr'])
        got = np.asarray(getattr(data.data['bin'], 'value', data.data['bin'])).astype(int)
        assert np.array_equal(got, case['labels']), (case['nstars'], case['dlogr'])
        sub = data.fetch_radial_bin(0)
        assert sub.sample_size == case['labels'].count(0)


def test_chain_post_processing_matches_reference():
    from mcmc_dynamics_b200 import synthetic
    from mcmc_dynamics_b200.analysis import ConstantFit
    c = GOLDEN['post']['chain_case']
    data, truth = synthetic.mock_cluster(60, seed=42)
    cf = ConstantFit(data)
    cf.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    cf.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    cf.parameters['v_sys'].set(value=1.5, fixed=True)
    assert cf.fitted_parameters == c['fitted_parameters']
    chain = np.asarray(c['chain'])
    assert np.allclose(cf.compute_percentiles(chain, n_burn=c['n_burn']), c['percentiles'], rtol=1e-14, atol=0)
    best = cf.compute_bestfit_values(chain, n_burn=c['n_burn'])
    for name, rows in c['bestfit'].items():
        got = [float(getattr(best.loc[r][name], 'value', best.loc[r][name])) for r in ('median', 'uperr', 'loerr')]
        assert np.allclose(got, rows, rtol=1e-13, atol=1e-15)
        assert cf.parameters[name].value == pytest.approx(rows[0])      # medians become the current values
    conv = cf.convert_to_parameters(chain, n_burn=c['n_burn'])
    assert {k: int(np.size(v)) for k, v in conv.items()} == c['convert_sizes']
    for name, head in c['convert_to_parameters'].items():
        assert np.allclose(np.asarray(conv[name])[:5], head, rtol=1e-14, atol=0)
    res = cf.compute_theta_vmax(chain, n_burn=c['n_burn'])
    for name, rows in c['theta_vmax'].items():
        got = [float(getattr(res.loc[r][name], 'value', res.loc[r][name])) for r in ('median', 'uperr', 'loerr')]
        assert np.allclose(got, rows, rtol=1e-12, atol=1e-14), name
    res2, v_max, theta, sig = cf.compute_theta_vmax(chain, n_burn=c['n_burn'], return_samples=True)
    assert v_max.shape == theta.shape == sig.shape == (12 * 30,)


def test_create_profiles_matches_reference(tmp_path):
    from mcmc_dynamics_b200 import synthetic
    from mcmc_dynamics_b200 import units as u
    from mcmc_dynamics_b200.analysis import ModelFit
    c = GOLDEN['post']['create_profiles']
    data, truth = synthetic.mock_cluster(60, seed=42)
    mf = ModelFit(data)
    mf.parameters['ra_center'].set(value=truth['ra_center'], fixed=True)
    mf.parameters['dec_center'].set(value=truth['dec_center'], fixed=True)
    assert mf.fitted_parameters == c['fitted_parameters']
    out = str(tmp_path / 'profile.csv')
    prof = mf.create_profiles(np.asarray(c['chain']), n_burn=c['n_burn'], radii=u.Quantity(c['radii_arcsec'], u.arcsec),
                              filename=out)
    assert prof.colnames == list(c['columns'])
    for name, want in c['columns'].items():
        assert np.allclose(np.asarray(prof[name].value), want, rtol=1e-12, atol=0), name
    # radii without unit are in the unit of r_peak (arcsec)
    prof2 = mf.create_profiles(np.asarray(c['chain']), n_burn=c['n_burn'], radii=c['radii_arcsec'])
    assert np.allclose(np.asarray(prof2['sigma'].value), c['columns']['sigma'], rtol=1e-12)
    assert np.loadtxt(out, delimiter=',').shape == (4, 11)
    assert len(mf.create_profiles(np.asarray(c['chain']), n_burn=c['n_burn'])) == 50


def test_parameters_json_written_by_the_reference_loads_here():
    """Config-format boundary: a Parameters object edited and serialised by the reference's own
    ``Parameters.dumps`` (with a stored ``random_state``) is read back with identical content.  The
    opposite direction (product ``dumps`` -> reference ``loads``) is asserted at generation time in
    tests/golden/make_golden.py."""
    from mcmc_dynamics_b200 import Parameters
    text = GOLDEN['post']['parameters_json_from_reference']
    pars = Parameters().loads(text)
    want = GOLDEN['post']['parameters_expected']
    assert list(pars) == [row[0] for row in want]
    for name, value, unit, fixed, lo, hi, initials, lnprior in want:
        p = pars[name]
        assert p.value == pytest.approx(value, rel=1e-15) and p.fixed == fixed and p.min == lo and p.max == hi
        assert (unit or '').replace(' ', '') == ('' if p.unit is None else str(p.unit).replace(' ', ''))
        assert p.initials == initials and p.lnprior == lnprior
    # the edited object behaves: unit conversion happened in the reference (0.5 arcmin -> 30 arcsec)
    assert pars['a'].value == pytest.approx(30.0) and pars['a'].max == 120.0
    assert pars['v_sys'].evaluate_lnprior(232.5) == pytest.approx(-np.log(2.0 * np.sqrt(2 * np.pi)))
    draws = pars['sigma_max'].evaluate_initials(50)
    assert draws.shape == (50,) and np.all(draws > 0)
