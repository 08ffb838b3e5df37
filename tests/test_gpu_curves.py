"""``Runner._calculate_lnlike(v_los, sigma_los)`` for model curves computed by the caller
(``analysis/runner.py:240-286``; C ABI ``mcd_calculate_lnlike``) and its composition with the public
``rotation_model`` / ``dispersion_model`` curves -- the three pieces a user-defined model class on top of
``Runner`` is built from in the reference (``constant.py:137-154``)."""
import numpy as np
import pytest

import golden_util
from common import RTOL, build
from mcmc_dynamics_b200 import _native
from mcmc_dynamics_b200 import units as u

pytestmark = pytest.mark.gpu
GOLDEN = golden_util.load()
PLAIN_CASES = [c for c in GOLDEN['cases'] if c['class'] in ('ConstantFit', 'ModelFit')]


@pytest.mark.parametrize('case', PLAIN_CASES, ids=[c['name'] for c in PLAIN_CASES])
def test_calculate_lnlike_on_the_reference_curves(case):
    """The reference's own curves at theta[0] in, the reference's own lnlike(theta[0]) out."""
    model = golden_util.product_for_case(case)
    curves = case['model_curves_theta0']
    want = case['expected']['lnlike'][0]
    got = model._calculate_lnlike(v_los=np.asarray(curves['v_los']), sigma_los=np.asarray(curves['sigma_los']))
    assert isinstance(got, float) and got == pytest.approx(want, rel=RTOL, abs=0)
    # quantities are converted, and the hook composes with the curves of this package
    par = model.fetch_parameter_values(np.asarray(case['theta'][0]))
    v_los = model.rotation_model(**{k: v for k, v in par.items() if k in model.rotation_parameters})
    sigma_los = model.dispersion_model(**{k: v for k, v in par.items() if k in model.dispersion_parameters})
    assert model._calculate_lnlike(v_los, sigma_los) == pytest.approx(want, rel=RTOL, abs=0)
    in_m_s = u.Quantity(np.asarray(curves['v_los']) * 1e3, u.m_s)
    assert model._calculate_lnlike(in_m_s, sigma_los) == pytest.approx(want, rel=RTOL, abs=0)
    assert model._calculate_lnlike(v_los, sigma_los) == pytest.approx(model.lnlike(np.asarray(case['theta'][0])), rel=RTOL, abs=0)


@pytest.mark.parametrize('variant,n_stars', [('ConstantFit', 1), ('ConstantFit', 257), ('ModelFit', 5001),
                                             ('ConstantFit+bg', 3001), ('ModelFit+bg', 400001)])
def test_calculate_lnlike_for_a_user_defined_model(variant, n_stars):
    """Arbitrary caller-side curves (not one of the built-in models), ragged sizes, one block up to the
    grid-stride regime, against the oracle's restatement of runner.py:261-286."""
    model, oracle, _theta, _truth = build(variant, n_stars=n_stars, seed=9)
    rng = np.random.default_rng(n_stars)
    v_los = 3.0 * np.sin(np.linspace(0.0, 20.0, n_stars)) + 0.1 * rng.standard_normal(n_stars)
    sigma_los = 4.0 + 3.0 * rng.random(n_stars)
    want = oracle._calculate_lnlike(v_los=v_los, sigma_los=sigma_los)
    got = model._calculate_lnlike(v_los, sigma_los)
    assert np.isfinite(want) and got == pytest.approx(want, rel=RTOL, abs=0)
    assert model._calculate_lnlike(v_los, sigma_los) == got                     # fixed summation order
    assert model._calculate_lnlike(v_los, 5.0) == pytest.approx(                # a scalar dispersion broadcasts
        oracle._calculate_lnlike(v_los=v_los, sigma_los=np.full(n_stars, 5.0)), rel=RTOL, abs=0)


def test_calculate_lnlike_is_refused_where_the_reference_does_not_use_it():
    model, _oracle, _theta, _truth = build('ModelFitGB', n_stars=300)
    with pytest.raises(_native.NativeError, match='fitted background fraction'):
        model._calculate_lnlike(np.zeros(300), np.ones(300))
    model, _oracle, _theta, _truth = build('ConstantFit', n_stars=300)
    with pytest.raises(ValueError):
        model._calculate_lnlike(np.zeros(299), np.ones(300))                    # one value per star
