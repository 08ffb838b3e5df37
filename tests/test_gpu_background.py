"""Background precompute kernels against the oracle (background/single_stars.py:42-77,
gaussian.py:23-28)."""
import numpy as np
import pytest

from mcmc_dynamics_b200.background import Gaussian, SingleStars
from oracle import reference_np as ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('m,n', [(1, 10), (7, 129), (300, 2000), (2000, 5000), (4099, 333)])
def test_single_stars_matches_oracle(m, n):
    rng = np.random.default_rng(m * 1000 + n)
    v_bg = np.concatenate([rng.normal(-20, 40, m // 2), rng.normal(30, 70, m - m // 2)])
    v = rng.normal(0, 60, n)
    verr = rng.uniform(0.3, 8.0, n)
    got = SingleStars(v_bg)(v, verr)
    want = ref.single_stars_background(v_bg, v, verr)
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) < 1e-12
    got2 = SingleStars(v_bg)(v, verr, sigma_int=3.0)
    want2 = ref.single_stars_background(v_bg, v, verr, sigma_int=3.0)
    assert np.max(np.abs(got2 - want2) / np.maximum(1.0, np.abs(want2))) < 1e-12


def test_single_star_background_is_a_gaussian():
    rng = np.random.default_rng(5)
    v = rng.normal(0, 30, 400)
    verr = rng.uniform(0.5, 5, 400)
    a = SingleStars([12.5])(v, verr)
    b = Gaussian(12.5, 0.0)(v, verr)
    assert np.allclose(a, b, rtol=1e-13, atol=1e-13)
    assert np.allclose(b, ref.gaussian_background(v, verr, 12.5, 0.0), rtol=1e-13, atol=1e-13)


def test_far_outlier_does_not_underflow():
    """log-sum-exp: a star hundreds of sigma from every background star still gets a finite value."""
    got = SingleStars([0.0, 1.0, 2.0])(np.array([5000.0]), np.array([1.0]))
    want = ref.single_stars_background(np.array([0.0, 1.0, 2.0]), np.array([5000.0]), np.array([1.0]))
    assert np.isfinite(got[0]) and abs(got[0] - want[0]) < 1e-9 * abs(want[0])


def test_gaussian_accepts_quantities():
    from mcmc_dynamics_b200 import units as u
    v = u.Quantity(np.array([1000.0, -2000.0]), u.m_s)       # m/s -> km/s
    verr = u.Quantity(np.array([500.0, 800.0]), u.m_s)
    got = Gaussian(u.Quantity(0.5, u.km_s), 2.0)(v, verr)
    want = ref.gaussian_background(np.array([1.0, -2.0]), np.array([0.5, 0.8]), 0.5, 2.0)
    assert np.allclose(got, want, rtol=1e-13)
