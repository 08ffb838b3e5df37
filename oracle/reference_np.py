"""Literal NumPy float64 restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Pinned against outputs of the reference's own source files (``tests/golden``, see
``oracle/__init__.py`` for how and for the one caveat about astropy's unit rules).

Every function follows the cited reference lines op for op, in the same order,
including the work the reference repeats (``calc_xy_offset`` is evaluated once
in the rotation model and once more in the dispersion model) and the
``max`` + two ``exp`` form of the two-component mixture.  What astropy does
implicitly -- unit conversion on ``+`` and inside trigonometric ufuncs -- is
written out as explicit factors.

All paths are relative to ``/root/reference/mcmc_dynamics``.
"""
import numpy as np

# --------------------------------------------------------------------------
# Units: the handful the reference's configs and drivers use.
# --------------------------------------------------------------------------
_ANGLE_TO_DEG = {
    'deg': 1.0,
    'rad': 180.0 / np.pi,
    'arcmin': 1.0 / 60.0,
    'arcsec': 1.0 / 3600.0,
    'mas': 1.0 / 3.6e6,
}
_VELOCITY_TO_KMS = {'km/s': 1.0, 'm/s': 1.0e-3}


def _norm_unit(unit):
    if unit is None:
        return None
    return str(unit).replace(' ', '')


def angle_factor(unit_from, unit_to):
    """Multiplicative factor converting an angle in `unit_from` to `unit_to`."""
    return _ANGLE_TO_DEG[_norm_unit(unit_from)] / _ANGLE_TO_DEG[_norm_unit(unit_to)]


def velocity_factor(unit_from, unit_to='km/s'):
    return _VELOCITY_TO_KMS[_norm_unit(unit_from)] / _VELOCITY_TO_KMS[_norm_unit(unit_to)]


# --------------------------------------------------------------------------
# Geometry -- utils/coordinates/calc_xy_offset.py:9-33
# --------------------------------------------------------------------------
def calc_xy_offset(ra, dec, ra_center, dec_center):
    """(dx, dy) in arcmin from coordinates in degrees (calc_xy_offset.py:11,30-31).

    ``np.cos(dec)`` on a degree Quantity converts to radians first; that is the
    ``np.deg2rad`` below.
    """
    r0 = 10800. / np.pi
    ra = np.asarray(ra, dtype=np.float64)
    dec = np.asarray(dec, dtype=np.float64)
    dec_rad = np.deg2rad(dec)
    dra_rad = np.deg2rad(ra - ra_center)
    dec_center_rad = np.deg2rad(dec_center)
    dx = -r0 * np.cos(dec_rad) * np.sin(dra_rad)
    dy = r0 * (np.sin(dec_rad) * np.cos(dec_center_rad)
               - np.cos(dec_rad) * np.sin(dec_center_rad) * np.cos(dra_rad))
    return dx, dy


# --------------------------------------------------------------------------
# Background populations -- background/gaussian.py:23-28, single_stars.py:42-77
# --------------------------------------------------------------------------
def gaussian_background(v, verr, mean, sigma):
    """background/gaussian.py:23-28."""
    norm = verr * verr + sigma * sigma
    exponent = -0.5 * np.power(v - mean, 2) / norm
    return -0.5 * np.log(2. * np.pi * norm) + exponent


def single_stars_background(v_bg, v, verr, sigma_int=0.0):
    """background/single_stars.py:70-77 (the reachable part; lines 78-88 are dead)."""
    v_bg = np.asarray(v_bg, dtype=np.float64)
    n_stars = v_bg.size
    norm = sigma_int ** 2 + verr ** 2
    exp_coeff = -(np.subtract.outer(v_bg, v)) ** 2 / (2. * norm)
    exp_coeff_max = np.max(exp_coeff, axis=0)
    lnlike = exp_coeff_max + np.log(np.sum(np.exp(exp_coeff - exp_coeff_max) /
                                           (np.sqrt(2. * np.pi * norm)), axis=0)) - np.log(n_stars)
    return lnlike


def single_stars_background_chunked(v_bg, v, verr, sigma_int=0.0, chunk=4096):
    """Same arithmetic as `single_stars_background`, evaluated in column chunks so
    that the M x N array (single_stars.py:73) never has to exist in full."""
    v = np.asarray(v, dtype=np.float64)
    verr = np.asarray(verr, dtype=np.float64)
    out = np.empty(v.shape, dtype=np.float64)
    for lo in range(0, v.size, chunk):
        hi = min(v.size, lo + chunk)
        out[lo:hi] = single_stars_background(v_bg, v[lo:hi], verr[lo:hi], sigma_int)
    return out


# --------------------------------------------------------------------------
# Parameter bookkeeping -- analysis/runner.py:143-217, parameter.py:684-705
# --------------------------------------------------------------------------
class OParam(object):
    """One row of a reference ``Parameters`` object (parameter.py:844-847)."""

    def __init__(self, name, value=None, unit=None, fixed=False, min=-np.inf, max=np.inf):
        self.name = name
        self.unit = _norm_unit(unit)
        self.fixed = bool(fixed)
        self.min = -np.inf if min is None else float(min)
        self.max = np.inf if max is None else float(max)
        # parameter.py:794-806
        if value is None:
            if np.isfinite(self.min) & np.isfinite(self.max):
                value = (self.min + self.max) / 2.
            else:
                value = 0.
        if self.min > self.max:
            self.min, self.max = self.max, self.min
        value = float(value)
        if value > self.max:
            value = self.max
        if value < self.min:
            value = self.min
        self.value = value

    def evaluate_lnprior(self, val):
        """parameter.py:691-705 with no expression prior (none is shipped)."""
        if val < self.min or val > self.max:
            return -np.inf
        return 0


#: default parameter tables, config/*.json (name, unit, min, max) in file order
DEFAULT_TABLES = {
    'constant': [
        ('v_sys', 'km/s', -np.inf, np.inf), ('sigma_max', 'km/s', 0.0, np.inf),
        ('v_maxx', 'km/s', -np.inf, np.inf), ('v_maxy', 'km/s', -np.inf, np.inf),
        ('ra_center', 'deg', 0.0, 360.0), ('dec_center', 'deg', -90.0, 90.0)],
    'model': [
        ('v_sys', 'km/s', -np.inf, np.inf), ('sigma_max', 'km/s', 0.0, np.inf),
        ('a', 'arcsec', 0.0, np.inf), ('v_maxx', 'km/s', -np.inf, np.inf),
        ('ra_center', 'deg', 0.0, 360.0), ('dec_center', 'deg', -90.0, 90.0),
        ('v_maxy', 'km/s', -np.inf, np.inf), ('r_peak', 'arcsec', 0.0, np.inf)],
}
_BACK = [('v_back', 'km/s', -np.inf, np.inf), ('sigma_back', 'km/s', 0.0, np.inf),
         ('f_back', None, 0.0, 1.0)]
DEFAULT_TABLES['constant_with_background'] = DEFAULT_TABLES['constant'] + _BACK
DEFAULT_TABLES['model_with_background'] = [
    DEFAULT_TABLES['model'][i] for i in (0, 1, 2, 3, 6, 7, 4, 5)] + _BACK


def default_params(table):
    return [OParam(name, None, unit, False, lo, hi) for name, unit, lo, hi in DEFAULT_TABLES[table]]


class OracleRunner(object):
    """analysis/runner.py:23-306 without astropy: parameters are `OParam` rows, star
    columns are float64 arrays in the reference's default units (km/s, deg)."""

    MODEL_PARAMETERS = []
    DEFAULT_TABLE = None

    def __init__(self, data, parameters=None, lnlike_background=None, pmember=None):
        if parameters is None:
            parameters = default_params(self.DEFAULT_TABLE)
        self.parameters = list(parameters)
        names = [p.name for p in self.parameters]
        missing = set(self.MODEL_PARAMETERS).difference(names)
        if missing:
            raise IOError("Missing required parameter(s): '{0}'".format(missing))   # runner.py:87-89
        self.v = np.asarray(data['v'], dtype=np.float64)
        self.verr = np.asarray(data['verr'], dtype=np.float64)
        self.ra = np.asarray(data['ra'], dtype=np.float64)
        self.dec = np.asarray(data['dec'], dtype=np.float64)
        self.density = None if 'density' not in data else np.asarray(data['density'], dtype=np.float64)
        self.n_data = self.v.size
        # runner.py:96-106
        self.lnlike_background = None if lnlike_background is None else np.asarray(lnlike_background, np.float64)
        self.pmember = None if pmember is None else np.asarray(pmember, np.float64)

    def __getitem__(self, name):
        for p in self.parameters:
            if p.name == name:
                return p
        raise KeyError(name)

    @property
    def fitted_parameters(self):
        return [p.name for p in self.parameters if not p.fixed]

    @property
    def n_fitted_parameters(self):
        return len(self.fitted_parameters)

    def fetch_parameter_values(self, values):
        """runner.py:160-180; values stay in each parameter's own unit."""
        current = {}
        i = 0
        for p in self.parameters:
            if p.fixed:
                current[p.name] = p.value
            else:
                current[p.name] = float(values[i])
                i += 1
        assert i == len(values), 'Not all parameters used.'
        return current

    def lnprior(self, values):
        """runner.py:206-217: every parameter, fixed ones included, is bounds-checked."""
        lnlike = 0
        for name, value in self.fetch_parameter_values(values).items():
            lnlike += self[name].evaluate_lnprior(value)
            if not np.isfinite(lnlike):
                return -np.inf
        return lnlike

    def lnlike(self, values):
        raise NotImplementedError

    def lnprob(self, values):
        """runner.py:303-306."""
        lp = self.lnprior(values)
        if not np.isfinite(lp):
            return -np.inf
        return self.lnlike(values) + lp

    # batched helpers: one reference call per walker, exactly how emcee drives it
    def lnprob_many(self, theta):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        with np.errstate(all='ignore'):
            return np.array([self.lnprob(row) for row in theta], dtype=np.float64)

    def lnlike_many(self, theta):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        with np.errstate(all='ignore'):
            return np.array([self.lnlike(row) for row in theta], dtype=np.float64)

    def lnprior_many(self, theta):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        return np.array([self.lnprior(row) for row in theta], dtype=np.float64)

    # ---- unit plumbing -----------------------------------------------------
    def _deg(self, name, value):
        return value * angle_factor(self[name].unit or 'deg', 'deg')

    def _kms(self, name, value):
        return value * velocity_factor(self[name].unit or 'km/s')

    def _arcmin_per_unit(self, name):
        return angle_factor(self[name].unit or 'arcsec', 'arcmin')

    # ---- likelihood kernels --------------------------------------------------
    def _calculate_lnlike(self, v_los, sigma_los):
        """runner.py:261-286."""
        norm = self.verr * self.verr + sigma_los * sigma_los
        exponent = -0.5 * np.power(self.v - v_los, 2) / norm

        if self.lnlike_background is None:
            sum1 = -0.5 * np.sum(np.log(2. * np.pi * norm))
            sum2 = np.sum(exponent)
            return sum1 + sum2
        else:
            lnlike_member = -0.5 * np.log(2. * np.pi * norm) + exponent
            max_lnlike = np.max([lnlike_member, self.lnlike_background], axis=0)
            lnlike = max_lnlike + np.log(self.pmember * np.exp(lnlike_member - max_lnlike) + (
                    1. - self.pmember) * np.exp(self.lnlike_background - max_lnlike))
            return lnlike.sum()

    def _cluster_gaussian(self, v_los, sigma_los):
        """constant.py:359-362 / model.py:447-450 / model.py:609-612."""
        norm = self.verr * self.verr + sigma_los * sigma_los
        exponent = -0.5 * np.power(self.v - v_los, 2) / norm
        return -0.5 * np.log(2. * np.pi * norm) + exponent

    def _fitted_background(self, v_back, sigma_back):
        """constant.py:333-336 / model.py:421-424."""
        norm = self.verr * self.verr + sigma_back * sigma_back
        exponent = -0.5 * np.power(self.v - v_back, 2) / norm
        return -0.5 * np.log(2. * np.pi * norm) + exponent

    @staticmethod
    def _membership(lnlike_cluster, lnlike_back, m):
        """constant.py:374 (max-shifted like model.py:507-510, 684-687), per star."""
        max_lnlike = np.max([lnlike_cluster, lnlike_back], axis=0)
        return m * np.exp(lnlike_cluster - max_lnlike) / (
            m * np.exp(lnlike_cluster - max_lnlike) + (1. - m) * np.exp(lnlike_back - max_lnlike))

    @staticmethod
    def _mixture(lnlike_cluster, lnlike_back, m):
        """constant.py:320-323 / model.py:452-454 / model.py:614-618, per star."""
        max_lnlike = np.max([lnlike_cluster, lnlike_back], axis=0)
        return max_lnlike + np.log(
            m * np.exp(lnlike_cluster - max_lnlike) + (1. - m) * np.exp(lnlike_back - max_lnlike))


class OracleConstantFit(OracleRunner):
    """analysis/constant.py:18-154."""
    MODEL_PARAMETERS = ['v_sys', 'sigma_max', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center']
    DEFAULT_TABLE = 'constant'

    def dispersion_model(self, sigma_max):
        """constant.py:74."""
        return sigma_max * np.ones(self.n_data, dtype=np.float64)

    def rotation_model(self, v_sys, v_maxx, v_maxy, ra_center, dec_center):
        """constant.py:106-111."""
        dx, dy = calc_xy_offset(ra=self.ra, dec=self.dec, ra_center=ra_center, dec_center=dec_center)
        theta = np.arctan2(dy, dx)

        v_max = np.sqrt(v_maxx ** 2 + v_maxy ** 2)
        theta_0 = np.arctan2(v_maxy, v_maxx)
        return v_sys + v_max * np.sin(theta - theta_0)

    def _models(self, par):
        v_los = self.rotation_model(
            v_sys=self._kms('v_sys', par['v_sys']), v_maxx=self._kms('v_maxx', par['v_maxx']),
            v_maxy=self._kms('v_maxy', par['v_maxy']),
            ra_center=self._deg('ra_center', par['ra_center']),
            dec_center=self._deg('dec_center', par['dec_center']))
        sigma_los = self.dispersion_model(sigma_max=self._kms('sigma_max', par['sigma_max']))
        return v_los, sigma_los

    def lnlike(self, values):
        """constant.py:136-154."""
        par = self.fetch_parameter_values(values)
        v_los, sigma_los = self._models(par)
        return self._calculate_lnlike(v_los=v_los, sigma_los=sigma_los)


class OracleConstantFitGB(OracleConstantFit):
    """analysis/constant.py:250-364."""
    MODEL_PARAMETERS = OracleConstantFit.MODEL_PARAMETERS + ['v_back', 'sigma_back', 'f_back']
    DEFAULT_TABLE = 'constant_with_background'

    def lnlike_per_star(self, values):
        par = self.fetch_parameter_values(values)
        lnlike_back = self._fitted_background(self._kms('v_back', par['v_back']),
                                              self._kms('sigma_back', par['sigma_back']))
        m = self.density / (self.density + par['f_back'])          # constant.py:339
        v_los, sigma_los = self._models(par)
        lnlike_cluster = self._cluster_gaussian(v_los, sigma_los)
        return self._mixture(lnlike_cluster, lnlike_back, m)

    def lnlike(self, values):
        """constant.py:316-324."""
        return self.lnlike_per_star(values).sum()

    def membership(self, values):
        """constant.py:366-374 at the parameter vector `values`."""
        par = self.fetch_parameter_values(values)
        lnlike_back = self._fitted_background(self._kms('v_back', par['v_back']),
                                              self._kms('sigma_back', par['sigma_back']))
        m = self.density / (self.density + par['f_back'])
        v_los, sigma_los = self._models(par)
        return self._membership(self._cluster_gaussian(v_los, sigma_los), lnlike_back, m)


class OracleModelFit(OracleRunner):
    """analysis/model.py:20-223."""
    MODEL_PARAMETERS = ['v_sys', 'v_maxx', 'v_maxy', 'r_peak', 'sigma_max', 'a', 'ra_center', 'dec_center']
    DEFAULT_TABLE = 'model'

    def dispersion_model(self, sigma_max, ra_center, dec_center, a, a_to_arcmin):
        """model.py:126-128.  `r` is in arcmin, `a` in its own unit; the `1. +` forces the
        ratio to an unscaled dimensionless number, i.e. (r_arcmin / a_arcmin)**2."""
        dx, dy = calc_xy_offset(ra=self.ra, dec=self.dec, ra_center=ra_center, dec_center=dec_center)
        r = np.sqrt(dx ** 2 + dy ** 2)
        scale = (1.0 / a_to_arcmin) ** 2            # arcmin2 / unit(a)2 -> dimensionless
        return sigma_max / (1. + (r ** 2 / a ** 2) * scale) ** 0.25

    def rotation_model(self, v_sys, v_maxx, v_maxy, ra_center, dec_center, r_peak, r_peak_to_arcmin):
        """model.py:171-180; the sum with `v_sys` converts km/s*arcmin/unit(r_peak) to km/s."""
        dx, dy = calc_xy_offset(ra=self.ra, dec=self.dec, ra_center=ra_center, dec_center=dec_center)
        r = np.sqrt(dx ** 2 + dy ** 2)

        v_max = np.sqrt(v_maxx ** 2 + v_maxy ** 2)
        theta_0 = np.arctan2(v_maxy, v_maxx)
        theta = np.arctan2(dy, dx)
        x_pa = r * np.sin(theta - theta_0)
        lin = 1.0 / r_peak_to_arcmin
        return v_sys + (2. * (v_max / r_peak) * x_pa / (1. + ((r / r_peak) ** 2) * lin ** 2)) * lin

    def _models(self, par):
        ra_c = self._deg('ra_center', par['ra_center'])
        dec_c = self._deg('dec_center', par['dec_center'])
        v_los = self.rotation_model(
            v_sys=self._kms('v_sys', par['v_sys']), v_maxx=self._kms('v_maxx', par['v_maxx']),
            v_maxy=self._kms('v_maxy', par['v_maxy']), ra_center=ra_c, dec_center=dec_c,
            r_peak=par['r_peak'], r_peak_to_arcmin=self._arcmin_per_unit('r_peak'))
        sigma_los = self.dispersion_model(
            sigma_max=self._kms('sigma_max', par['sigma_max']), ra_center=ra_c, dec_center=dec_c,
            a=par['a'], a_to_arcmin=self._arcmin_per_unit('a'))
        return v_los, sigma_los

    def lnlike(self, values):
        """model.py:205-223."""
        par = self.fetch_parameter_values(values)
        v_los, sigma_los = self._models(par)
        return self._calculate_lnlike(v_los=v_los, sigma_los=sigma_los)


class OracleModelFitGB(OracleModelFit):
    """analysis/model.py:338-456."""
    MODEL_PARAMETERS = OracleModelFit.MODEL_PARAMETERS + ['v_back', 'sigma_back', 'f_back']
    DEFAULT_TABLE = 'model_with_background'

    def lnlike_per_star(self, values):
        par = self.fetch_parameter_values(values)
        lnlike_back = self._fitted_background(self._kms('v_back', par['v_back']),
                                              self._kms('sigma_back', par['sigma_back']))
        m = self.density / (self.density + par['f_back'])          # model.py:427
        v_los, sigma_los = self._models(par)
        lnlike_cluster = self._cluster_gaussian(v_los, sigma_los)
        return self._mixture(lnlike_cluster, lnlike_back, m)

    def lnlike(self, values):
        """model.py:414-456."""
        return self.lnlike_per_star(values).sum()

    def membership(self, values):
        """model.py:458-510 at the parameter vector `values`."""
        par = self.fetch_parameter_values(values)
        lnlike_back = self._fitted_background(self._kms('v_back', par['v_back']),
                                              self._kms('sigma_back', par['sigma_back']))
        m = self.density / (self.density + par['f_back'])
        v_los, sigma_los = self._models(par)
        return self._membership(self._cluster_gaussian(v_los, sigma_los), lnlike_back, m)


class OracleModelFitConstantBackground(OracleModelFit):
    """analysis/model.py:513-623; `lnlike_background` is the precomputed column of
    model.py:562-563."""
    MODEL_PARAMETERS = OracleModelFit.MODEL_PARAMETERS + ['f_back', ]
    DEFAULT_TABLE = 'model_with_background'

    def lnlike(self, values, no_sum=False):
        par = self.fetch_parameter_values(values)
        m = self.density / (self.density + par['f_back'])          # model.py:589
        v_los, sigma_los = self._models(par)
        lnlike_cluster = self._cluster_gaussian(v_los, sigma_los)
        lnlike = self._mixture(lnlike_cluster, self.lnlike_background, m)
        if no_sum:
            return lnlike
        return lnlike.sum()

    def membership(self, values):
        """model.py:625-687 at the parameter vector `values`."""
        par = self.fetch_parameter_values(values)
        m = self.density / (self.density + par['f_back'])
        v_los, sigma_los = self._models(par)
        return self._membership(self._cluster_gaussian(v_los, sigma_los), self.lnlike_background, m)

    def _calculate_lnlike(self, v_los, sigma_los):   # not used by this class (model.py:565-623)
        raise NotImplementedError


ORACLE_CLASSES = {
    'ConstantFit': OracleConstantFit,
    'ConstantFitGB': OracleConstantFitGB,
    'ModelFit': OracleModelFit,
    'ModelFitGB': OracleModelFitGB,
    'ModelFitConstantBackground': OracleModelFitConstantBackground,
}
