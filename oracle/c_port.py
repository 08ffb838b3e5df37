"""ctypes front end of ``oracle/oracle_c.c`` (plain-C restatement of the reference).  TEST
INFRASTRUCTURE ONLY: second checker beside ``reference_np`` and compiled CPU baseline of bench.py.

``COracle(oracle)`` wraps a NumPy oracle object (``reference_np.Oracle*``) and evaluates the same
model through the C code: parameters, fixed values, units and star columns are taken from it."""
import ctypes
import os

import numpy as np

from . import reference_np as ref

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_build', 'liboracle_c.so')
_FIELDS = ('v_sys', 'sigma_max', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center', 'a', 'r_peak', 'v_back', 'sigma_back',
           'f_back')
_lib = None


class _Params(ctypes.Structure):
    _fields_ = [(name, ctypes.c_double) for name in _FIELDS]


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.oracle_lnlike.restype = None
        lib.oracle_lnlike.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_long, dp, dp, dp, dp, dp, dp, dp, ctypes.c_int,
                                      ctypes.POINTER(_Params), ctypes.c_double, ctypes.c_double, dp]
        lib.oracle_gaussian_background.restype = None
        lib.oracle_gaussian_background.argtypes = [ctypes.c_long, dp, dp, ctypes.c_double, ctypes.c_double, dp]
        lib.oracle_single_stars_background.restype = None
        lib.oracle_single_stars_background.argtypes = [ctypes.c_long, dp, ctypes.c_long, dp, dp, ctypes.c_double, dp]
        lib.oracle_set_threads.restype = ctypes.c_int
        lib.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


def set_threads(n):
    """OpenMP threads of the walker loops (0: leave as is); returns the number in force."""
    return int(load().oracle_set_threads(int(n)))


def _ptr(arr):
    return None if arr is None else arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


class COracle(object):
    def __init__(self, oracle):
        self.oracle = oracle
        self.lib = load()
        cont = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.float64)   # noqa: E731
        self.cols = {k: cont(getattr(oracle, k)) for k in ('ra', 'dec', 'v', 'verr', 'pmember', 'density',
                                                            'lnlike_background')}
        self.radial = int(isinstance(oracle, ref.OracleModelFit))
        if isinstance(oracle, (ref.OracleConstantFitGB, ref.OracleModelFitGB)):
            self.background = 3
        elif isinstance(oracle, ref.OracleModelFitConstantBackground):
            self.background = 2
        elif oracle.lnlike_background is not None:
            self.background = 1
        else:
            self.background = 0

    def _params(self, theta):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        out = (_Params * len(theta))()
        o = self.oracle
        names = [p.name for p in o.parameters]
        for k, row in enumerate(theta):
            par = o.fetch_parameter_values(row)
            for name in _FIELDS:
                if name not in names:
                    value = 0.0
                elif name in ('ra_center', 'dec_center'):
                    value = o._deg(name, par[name])
                elif name in ('a', 'r_peak', 'f_back'):
                    value = par[name]                      # own unit; factor passed separately
                else:
                    value = o._kms(name, par[name])
                setattr(out[k], name, float(value))
        return out, len(theta)

    def lnlike_many(self, theta):
        params, nw = self._params(theta)
        out = np.empty(nw, dtype=np.float64)
        o = self.oracle
        a_f = o._arcmin_per_unit('a') if self.radial else 1.0
        rp_f = o._arcmin_per_unit('r_peak') if self.radial else 1.0
        c = self.cols
        self.lib.oracle_lnlike(self.radial, self.background, o.n_data, _ptr(c['ra']), _ptr(c['dec']), _ptr(c['v']),
                               _ptr(c['verr']), _ptr(c['pmember']), _ptr(c['density']), _ptr(c['lnlike_background']), nw,
                               params, a_f, rp_f, _ptr(out))
        return out

    def lnprob_many(self, theta):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        prior = self.oracle.lnprior_many(theta)
        ok = np.isfinite(prior)
        out = np.full(len(theta), -np.inf)
        if np.any(ok):
            out[ok] = self.lnlike_many(theta[ok]) + prior[ok]
        return out


def gaussian_background(v, verr, mean, sigma):
    v = np.ascontiguousarray(v, dtype=np.float64)
    verr = np.ascontiguousarray(verr, dtype=np.float64)
    out = np.empty_like(v)
    load().oracle_gaussian_background(v.size, _ptr(v), _ptr(verr), float(mean), float(sigma), _ptr(out))
    return out


def single_stars_background(v_bg, v, verr, sigma_int=0.0):
    v_bg = np.ascontiguousarray(v_bg, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    verr = np.ascontiguousarray(verr, dtype=np.float64)
    out = np.empty_like(v)
    load().oracle_single_stars_background(v_bg.size, _ptr(v_bg), v.size, _ptr(v), _ptr(verr), float(sigma_int), _ptr(out))
    return out
