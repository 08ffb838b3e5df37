"""Synthetic star catalogues for tests and benchmarks.

``mock_cluster`` scales up the only known-truth scenario the reference has, the mock generator of
``bin/run_tests.py:35-70`` (Lynden-Bell rotation + Plummer dispersion around a known centre, with
heteroscedastic uncertainties); ``add_background`` turns it into the contaminated catalogues of
BASELINE.json configs 3 (membership probabilities + Besancon-style field velocities).
"""
import numpy as np
from scipy import stats

from .data_reader import DataReader
from . import units as u


def offset_by(ra0_deg, dec0_deg, position_angle, separation):
    """Positions reached from (ra0, dec0) by moving `separation` [rad] along `position_angle` [rad,
    east of north] on the sphere."""
    ra0, dec0 = np.deg2rad(ra0_deg), np.deg2rad(dec0_deg)
    sin_dec = np.sin(dec0) * np.cos(separation) + np.cos(dec0) * np.sin(separation) * np.cos(position_angle)
    dec = np.arcsin(sin_dec)
    ra = ra0 + np.arctan2(np.sin(position_angle) * np.sin(separation) * np.cos(dec0),
                          np.cos(separation) - np.sin(dec0) * sin_dec)
    return np.rad2deg(ra) % 360.0, np.rad2deg(dec)


def mock_cluster(n_stars, seed=1, ra_center=56.345, dec_center=-26.675, v_sys=0.0, r_peak=60.0, a=30.0, rmax=5.0,
                 vsigma=0.5, errscale=0.1, sigma_max=None, theta_0=None, as_reader=True):
    """Mock cluster after ``bin/run_tests.py:35-70``.

    `r_peak`, `a` in arcsec; separations are truncated-normal(0, rmax/2 r_peak) cut at rmax r_peak.
    Returns ``(DataReader | dict of columns, truth dict)``.
    """
    rng = np.random.default_rng(seed)
    theta_0 = 2. * np.pi * rng.random() if theta_0 is None else theta_0
    sigma_max = 5. + 10. * rng.random() if sigma_max is None else sigma_max
    v_max = vsigma * sigma_max

    r_max = r_peak * rmax
    separation = stats.truncnorm.rvs(a=0, b=2.0, loc=0, scale=r_max / 2., size=n_stars, random_state=rng)   # arcsec
    position_angle = rng.uniform(-np.pi, np.pi, size=n_stars)
    ra, dec = offset_by(ra_center, dec_center, position_angle, np.deg2rad(separation / 3600.0))

    x_pa = separation * np.sin(position_angle + np.pi / 2. - theta_0)
    v_los = v_sys + 2. * (v_max / r_peak) * x_pa / (1. + (separation / r_peak) ** 2)
    sigma_los = sigma_max / (1. + (separation / a) ** 2) ** 0.25
    v = v_los + rng.normal(scale=sigma_los, size=n_stars)
    verr = errscale * sigma_los * rng.lognormal(0, 0.5, size=n_stars)
    v = v + rng.normal(scale=verr, size=n_stars)

    columns = {'ra': ra, 'dec': dec, 'v': v, 'verr': verr}
    truth = {'v_sys': v_sys, 'sigma_max': sigma_max, 'a': a, 'r_peak': r_peak, 'theta_0': theta_0, 'v_max': v_max,
             'v_maxx': v_max * np.cos(theta_0), 'v_maxy': v_max * np.sin(theta_0), 'ra_center': ra_center,
             'dec_center': dec_center, 'separation_arcsec': separation}
    if not as_reader:
        return columns, truth
    return reader_from_columns(columns), truth


def reader_from_columns(columns):
    units = {'ra': u.deg, 'dec': u.deg, 'v': u.km_s, 'verr': u.km_s}
    return DataReader({k: (u.Quantity(v, units[k]) if k in units else v) for k, v in columns.items()})


def add_background(columns, truth, contamination=0.3, seed=2, field=((-20.0, 40.0, 0.5), (30.0, 70.0, 0.5))):
    """Replace a fraction of the stars by field contaminants and attach ``pmember`` and ``density``.

    Contaminant velocities follow a two-Gaussian "Besancon-style" field distribution
    ``field = ((mean, sigma, weight), ...)``; ``pmember`` is a noisy membership probability
    correlated with the truth; ``density`` is the Plummer surface density (scale radius ``a``)
    normalised to 1 at the centre.  Returns the field-star velocity sampler for building a
    ``SingleStars`` background.
    """
    rng = np.random.default_rng(seed)
    n = columns['v'].size
    is_field = rng.random(n) < contamination
    weights = np.array([f[2] for f in field])
    comp = rng.choice(len(field), size=n, p=weights / weights.sum())
    means = np.array([f[0] for f in field])[comp]
    sigmas = np.array([f[1] for f in field])[comp]
    v_field = rng.normal(means, sigmas) + rng.normal(scale=columns['verr'])
    columns = dict(columns)
    columns['v'] = np.where(is_field, v_field, columns['v'])
    pm = np.where(is_field, rng.beta(1.5, 5.0, size=n), rng.beta(5.0, 1.5, size=n))
    columns['pmember'] = np.clip(pm, 1e-6, 1.0 - 1e-6)
    r = truth['separation_arcsec']
    columns['density'] = (1. + (r / truth['a']) ** 2) ** -2.0

    def sample_field(m, seed=3):
        g = np.random.default_rng(seed)
        c = g.choice(len(field), size=m, p=weights / weights.sum())
        return g.normal(np.array([f[0] for f in field])[c], np.array([f[1] for f in field])[c])

    return columns, sample_field


def initial_ball(truth, names, n_walkers, seed=5, scale=0.05):
    """Walker start positions: a small ball around the truth (relative `scale`, absolute for
    parameters whose truth is ~0), inside the default bounds."""
    rng = np.random.default_rng(seed)
    pos = np.empty((n_walkers, len(names)))
    for j, name in enumerate(names):
        centre = float(truth[name])
        if name in ('ra_center', 'dec_center'):
            width = 2.0 / 3600.0                     # 2 arcsec
        else:
            width = scale * max(abs(centre), 1.0)
        pos[:, j] = centre + width * rng.standard_normal(n_walkers)
        if name in ('sigma_max', 'a', 'r_peak', 'sigma_back'):
            pos[:, j] = np.abs(pos[:, j]) + 1e-3
        if name == 'f_back':
            pos[:, j] = np.clip(pos[:, j], 1e-3, 1 - 1e-3)
    return pos
