"""The CUDA path, through the model classes and the C ABI, directly against the golden vectors
produced by the reference's own source files (tolerance 1e-9 relative, north_star)."""
import numpy as np
import pytest

import golden_util

pytestmark = pytest.mark.gpu
GOLDEN = golden_util.load()


@pytest.mark.parametrize('math_mode', ['fast', 'plain'])
@pytest.mark.parametrize('case', GOLDEN['cases'], ids=[c['name'] for c in GOLDEN['cases']])
def test_cuda_reproduces_reference_outputs(case, math_mode):
    model = golden_util.product_for_case(case, math_mode)
    theta = np.asarray(case['theta'])
    exp = case['expected']
    lnprob = model.lnprob(theta)
    lnlike = model.lnlike(theta)
    lnprior = model.lnprior(theta)
    for k in range(len(theta)):
        assert lnprior[k] == exp['lnprior'][k]
        if np.isfinite(exp['lnprob'][k]):
            assert lnprob[k] == pytest.approx(exp['lnprob'][k], rel=1e-9, abs=0)
            assert lnlike[k] == pytest.approx(exp['lnlike'][k], rel=1e-9, abs=0)
            assert model.lnprob(theta[k]) == pytest.approx(exp['lnprob'][k], rel=1e-9, abs=0)   # scalar protocol
        else:
            assert lnprob[k] == exp['lnprob'][k] == -np.inf
    if 'lnlike_background' in case:
        assert np.allclose(np.asarray(model.lnlike_background), case['lnlike_background'], rtol=1e-11, atol=0)
    if 'lnlike_per_star_theta0' in case:
        got = model.lnlike(theta[0], no_sum=True)
        assert np.allclose(got, case['lnlike_per_star_theta0'], rtol=1e-9, atol=1e-12)
    if 'membership_theta0' in case:
        # calculate_membership_probabilities evaluates at the posterior median: a chain whose every
        # sample is theta[0] has that median
        chain = np.broadcast_to(theta[0], (4, 3, theta.shape[1])).copy()
        got = model.calculate_membership_probabilities(chain, n_burn=1)
        assert np.allclose(got, case['membership_theta0'], rtol=1e-9, atol=1e-13)
        assert np.all((got >= 0) & (got <= 1))


def _curve_kwargs(model, row):
    """name -> value for rotation_model / dispersion_model, routed as <Model>.lnlike routes them
    (constant.py:137-147): sampled parameters from `row`, fixed ones at their value, each with its unit."""
    par = model.fetch_parameter_values(row)
    rot = {k: v for k, v in par.items() if k in model.rotation_parameters}
    disp = {k: v for k, v in par.items() if k in model.dispersion_parameters}
    return rot, disp


@pytest.mark.parametrize('math_mode', ['fast', 'plain'])
@pytest.mark.parametrize('case', GOLDEN['cases'], ids=[c['name'] for c in GOLDEN['cases']])
def test_model_curves_reproduce_the_reference(case, math_mode):
    """`rotation_model` / `dispersion_model` (per-star kernel, `mcd_model_per_star`) against the curves the
    reference's own classes return for theta[0]."""
    model = golden_util.product_for_case(case, math_mode)
    rot, disp = _curve_kwargs(model, np.asarray(case['theta'][0]))
    want = case['model_curves_theta0']
    v_los = model.rotation_model(**rot)
    sigma_los = model.dispersion_model(**disp)
    assert str(v_los.unit) == str(sigma_los.unit) == 'km / s'
    assert np.allclose(v_los.value, want['v_los'], rtol=1e-9, atol=1e-11)
    assert np.allclose(sigma_los.value, want['sigma_los'], rtol=1e-9, atol=0)
    # plain numbers are taken to be in the parameter's own unit, as the reference takes them
    bare = {k: getattr(v, 'value', v) for k, v in rot.items()}
    assert np.array_equal(model.rotation_model(**bare).value, v_los.value)
    with pytest.raises(IOError, match='Unknown keyword argument'):
        model.rotation_model(nonsense=1.0, **rot)
    with pytest.raises(IOError, match='Unknown keyword argument'):
        model.dispersion_model(nonsense=1.0, **disp)
    if not case['free_centre']:
        moved = dict(rot, ra_center=getattr(rot['ra_center'], 'value', rot['ra_center']) + 0.01)
        with pytest.raises(ValueError, match='is fixed at'):
            model.rotation_model(**moved)
