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
