"""Known-answer tests of the NumPy oracle, hand-derived from the cited reference formulas
(SURVEY.md section 8c).  They pin the oracle independently of any GPU code."""
import numpy as np
import pytest

from oracle import reference_np as ref

R0 = 10800. / np.pi
CENTRE = (56.345, -26.675)


def _fixed(params, **values):
    for p in params:
        if p.name in values:
            p.fixed = True
            p.value = values[p.name]
    return params


def _data(ra, dec, v, verr, **extra):
    d = {'ra': np.atleast_1d(ra).astype(float), 'dec': np.atleast_1d(dec).astype(float),
         'v': np.atleast_1d(v).astype(float), 'verr': np.atleast_1d(verr).astype(float)}
    d.update({k: np.atleast_1d(x).astype(float) for k, x in extra.items()})
    return d


def test_kat1_single_star_unit_gaussian():
    """v = v_los, verr = 0, sigma = 1  =>  lnlike = -0.5 ln(2 pi)   (runner.py:269-271)"""
    data = _data(CENTRE[0], CENTRE[1] + 1. / 60., 3.0, 0.0)
    params = _fixed(ref.default_params('constant'), ra_center=CENTRE[0], dec_center=CENTRE[1])
    m = ref.OracleConstantFit(data, parameters=params)
    # theta = (v_sys, sigma_max, v_maxx, v_maxy); star due north => v_los = v_sys + v_maxx
    assert m.lnlike([1.0, 1.0, 2.0, 5.0]) == pytest.approx(-0.9189385332046727, abs=1e-13)


def test_kat2_geometry_north_and_west():
    """calc_xy_offset.py:30-31: a star 1' north has dx = 0, dy = r0 sin(1'); theta = pi/2 gives
    v_los = v_sys + v_maxx; a star with dx > 0, dy = 0 gives v_los = v_sys - v_maxy
    (constant.py:107-111)."""
    dx, dy = ref.calc_xy_offset(np.array([CENTRE[0]]), np.array([CENTRE[1] + 1. / 60.]), *CENTRE)
    assert abs(dx[0]) < 1e-12 and dy[0] == pytest.approx(R0 * np.sin(np.deg2rad(1. / 60.)), rel=1e-12)
    params = _fixed(ref.default_params('constant'), ra_center=CENTRE[0], dec_center=CENTRE[1])
    north = ref.OracleConstantFit(_data(CENTRE[0], CENTRE[1] + 1. / 60., 0., 1.), parameters=params)
    v_los, _ = north._models(north.fetch_parameter_values([1.5, 2.0, 3.0, 7.0]))
    assert v_los[0] == pytest.approx(1.5 + 3.0, abs=1e-12)
    # dx > 0 means ra < ra_center (dx = -r0 cos(dec) sin(ra - ra_c)); on the equator dy = 0 exactly
    params = _fixed(ref.default_params('constant'), ra_center=10.0, dec_center=0.0)
    west = ref.OracleConstantFit(_data(10.0 - 0.01, 0.0, 0., 1.), parameters=params)
    v_los, _ = west._models(west.fetch_parameter_values([1.5, 2.0, 3.0, 7.0]))
    assert v_los[0] == pytest.approx(1.5 - 7.0, abs=1e-12)
    # the star AT the centre: dx = -r0 cos(dec) sin(+0) = -0.0 and dy = +0.0, so atan2(+0, -0) = pi
    # (not 0) and v_los = v_sys + v_max sin(pi - theta_0) = v_sys + v_maxy
    at = ref.OracleConstantFit(_data(10.0, 0.0, 0., 1.), parameters=params)
    v_los, _ = at._models(at.fetch_parameter_values([1.5, 2.0, 3.0, 7.0]))
    assert v_los[0] == pytest.approx(1.5 + 7.0, abs=1e-12)


def test_kat3_modelfit_peak_and_scale_radius_with_unit_factor():
    """model.py:128,180: at r = r_peak and theta - theta_0 = pi/2, v_los = v_sys + v_max; at r = a,
    sigma = sigma_max 2^(-1/4) -- with a, r_peak in arcsec and r in arcmin (factor 60)."""
    r_peak_arcsec, a_arcsec = 60.0, 30.0
    params = _fixed(ref.default_params('model'), ra_center=10.0, dec_center=0.0)
    names = [p.name for p in params if not p.fixed]      # v_sys, sigma_max, a, v_maxx, v_maxy, r_peak

    def theta(**kw):
        return [kw[n] for n in names]
    # star on the equator one r_peak (= 1 arcmin) west of the centre: dx = +r, dy = 0, theta_i = 0
    sep_deg = np.rad2deg(np.arcsin(1.0 / R0))            # dx = r0 sin(sep) = 1 arcmin exactly
    m = ref.OracleModelFit(_data(10.0 - sep_deg, 0.0, 0., 1.), parameters=params)
    # theta_0 = -pi/2 (v_maxx = 0, v_maxy = -v_max) => sin(theta_i - theta_0) = 1
    v_los, sigma = m._models(m.fetch_parameter_values(theta(v_sys=2.0, sigma_max=8.0, a=a_arcsec, v_maxx=0.0,
                                                             v_maxy=-5.0, r_peak=r_peak_arcsec)))
    assert v_los[0] == pytest.approx(2.0 + 5.0, rel=1e-12)
    # same star, a = 60 arcsec = r  => sigma = sigma_max / 2^(1/4)
    v_los, sigma = m._models(m.fetch_parameter_values(theta(v_sys=2.0, sigma_max=8.0, a=60.0, v_maxx=0.0,
                                                             v_maxy=-5.0, r_peak=r_peak_arcsec)))
    assert sigma[0] == pytest.approx(8.0 * 2 ** -0.25, rel=1e-12)


def test_kat4_mixture_limits():
    """runner.py:280-286: p = 1 -> member term; p = 0 -> background; lm = lbg -> same value for any
    p; p = 1 with lm - lbg < -745 -> -inf (the reference's exp underflows)."""
    lm = np.array([-3.0, -3.0, -5.0, -2000.0])
    lbg = np.array([-7.0, -7.0, -5.0, -4.0])
    p = np.array([1.0, 0.0, 0.3, 1.0])
    with np.errstate(divide='ignore'):
        out = ref.OracleRunner._mixture(lm, lbg, p)
    assert out[0] == pytest.approx(-3.0, abs=1e-15)
    assert out[1] == pytest.approx(-7.0, abs=1e-15)
    assert out[2] == pytest.approx(-5.0, abs=1e-15)
    assert out[3] == -np.inf


def test_kat5_prior_bounds_inclusive_and_fixed_checked():
    """parameter.py:691-692, runner.py:207-216."""
    params = _fixed(ref.default_params('constant'), ra_center=10.0, dec_center=0.0)
    m = ref.OracleConstantFit(_data(10.0, 0.1, 0., 1.), parameters=params)
    assert m.lnprior([0.0, 0.0, 0.0, 0.0]) == 0            # sigma_max == min accepted
    assert m.lnprior([0.0, -1e-300, 0.0, 0.0]) == -np.inf
    m['v_sys'].min, m['v_sys'].max = -1.0, 1.0
    assert m.lnprior([1.0, 1.0, 0.0, 0.0]) == 0 and m.lnprior([-1.0, 1.0, 0.0, 0.0]) == 0
    assert m.lnprior([1.0 + 1e-12, 1.0, 0.0, 0.0]) == -np.inf
    m['dec_center'].min, m['dec_center'].max = 1.0, 2.0     # fixed value 0 now violates its bounds
    assert m.lnprior([0.0, 1.0, 0.0, 0.0]) == -np.inf
    assert m.lnprob([0.0, 1.0, 0.0, 0.0]) == -np.inf


def test_kat6_single_stars_with_one_star_is_a_gaussian():
    """single_stars.py:72-77 with M = 1 equals gaussian.py:25-28 with sigma = 0."""
    rng = np.random.default_rng(0)
    v = rng.normal(0, 30, 50)
    verr = rng.uniform(0.5, 5, 50)
    a = ref.single_stars_background(np.array([12.5]), v, verr)
    b = ref.gaussian_background(v, verr, 12.5, 0.0)
    assert np.allclose(a, b, rtol=1e-13, atol=1e-13)


def test_kat7_literal_vs_algebraic_identity():
    """v_max r sin(theta_i - theta_0) = dy v_maxx - dx v_maxy and sigma^2 = sigma_max^2 / sqrt(1 + r^2/a^2):
    the identities the CUDA kernels are built on, against the literal restatement (<= 1e-12)."""
    rng = np.random.default_rng(1)
    n = 5000
    ra = 56.345 + rng.normal(0, 0.03, n)
    dec = -26.675 + rng.normal(0, 0.03, n)
    v = rng.normal(0, 10, n)
    verr = rng.uniform(0.5, 3, n)
    params = ref.default_params('model')
    m = ref.OracleModelFit(_data(ra, dec, v, verr), parameters=params)
    names = [p.name for p in params]
    th = dict(v_sys=0.3, sigma_max=9.0, a=35.0, v_maxx=2.0, ra_center=56.3449, dec_center=-26.6752, v_maxy=-3.0,
              r_peak=70.0)
    literal = m.lnlike([th[k] for k in names])
    dx, dy = ref.calc_xy_offset(ra, dec, th['ra_center'], th['dec_center'])
    r2 = dx ** 2 + dy ** 2
    rp, a = th['r_peak'] / 60.0, th['a'] / 60.0
    v_los = th['v_sys'] + 2.0 / rp * (dy * th['v_maxx'] - dx * th['v_maxy']) / (1.0 + r2 / rp ** 2)
    sig2 = th['sigma_max'] ** 2 / np.sqrt(1.0 + r2 / a ** 2)
    norm = verr ** 2 + sig2
    algebraic = np.sum(-0.5 * np.log(2 * np.pi * norm) - 0.5 * (v - v_los) ** 2 / norm)
    assert abs(literal - algebraic) <= 1e-12 * abs(literal)


def test_default_value_rule():
    """parameter.py:794-798: value = (min + max) / 2 when both bounds are finite, else 0."""
    p = {q.name: q for q in ref.default_params('model_with_background')}
    assert p['ra_center'].value == 180.0 and p['dec_center'].value == 0.0 and p['f_back'].value == 0.5
    assert p['v_sys'].value == 0.0 and p['sigma_max'].value == 0.0
