"""Background populations (``mcmc_dynamics/background/``): callables ``bg(v, verr) -> lnlike[N]``
evaluated once per model object on the GPU (``csrc/mcd_background.cu``)."""
import logging

import numpy as np

from .. import _native
from .. import units as u

logger = logging.getLogger(__name__)


def _velocity(value, what):
    q = u.as_quantity(value)
    if q.unit.is_unity():
        logger.warning('Missing units for {0}. Assuming {1}.'.format(what, u.km_s))
        return np.asarray(q.value, dtype=np.float64)
    return np.asarray(q.to(u.km_s).value, dtype=np.float64)


class Gaussian(object):
    """Gaussian background in velocity (``background/gaussian.py:9-28``)."""

    def __init__(self, mean, sigma, device=0):
        self.mean = u.Quantity(float(_velocity(mean, 'parameter <mean>')), u.km_s)
        self.sigma = u.Quantity(float(_velocity(sigma, 'parameter <sigma>')), u.km_s)
        self.device = device

    def __call__(self, v, verr):
        v = _native.contiguous(u.strip(v, u.km_s))
        verr = _native.contiguous(u.strip(verr, u.km_s))
        out = np.empty_like(v)
        lib = _native.load_library()
        rc = lib.mcd_gaussian_lnlike(self.device, _native.as_double_ptr(v), _native.as_double_ptr(verr), v.size,
                                     float(self.mean.value), float(self.sigma.value), _native.as_double_ptr(out))
        _native.check(rc)
        return out


class SingleStars(object):
    """Background made of M individual stars with known velocities: log-mean-exp of M Gaussian
    kernels of width ``sqrt(verr_i^2 + sigma_int^2)`` (``background/single_stars.py:9-77``).  The
    reference materialises an M x N array; the kernel streams it."""

    def __init__(self, v, device=0):
        self.v = u.Quantity(_velocity(v, '<v> values'), u.km_s)
        self.n_stars = self.v.size
        self.device = device

    def __call__(self, v, verr, sigma_int=0.0):
        sigma_int = float(_velocity(sigma_int, 'parameter <sigma_int>'))
        v_bg = _native.contiguous(np.atleast_1d(self.v.value))
        v = _native.contiguous(u.strip(v, u.km_s))
        verr = _native.contiguous(u.strip(verr, u.km_s))
        out = np.empty_like(v)
        lib = _native.load_library()
        rc = lib.mcd_single_stars_lnlike(self.device, _native.as_double_ptr(v_bg), v_bg.size, _native.as_double_ptr(v),
                                         _native.as_double_ptr(verr), v.size, sigma_int, _native.as_double_ptr(out))
        _native.check(rc)
        return out


__all__ = ['Gaussian', 'SingleStars']
