"""The binding a maintainer of mcmc-dynamics would add to route ``Runner.lnprob`` through libmcd_b200.so
(Option B).  This file IS that binding: drop it into the reference as
``mcmc_dynamics/analysis/_b200.py`` and call ``install()`` once (or apply the six-line patch shown in
INTEGRATION.md by hand).  It needs only ctypes, numpy and ``astropy.units`` -- nothing from
``mcmc_dynamics_b200``'s Python package.

    from mcmc_dynamics.analysis import _b200
    _b200.install()                         # Runner.lnprob -> one CUDA launch per (half-)ensemble
    sampler = emcee.EnsembleSampler(n_walkers, ndim, model.lnprob, vectorize=True)

What it replaces (paths relative to /root/reference/mcmc_dynamics/):

* ``Runner.lnprob`` / ``lnprior`` / ``<Model>.lnlike`` per walker (analysis/runner.py:182-217,288-306;
  constant.py:113-154,293-364; model.py:182-223,391-456,565-623) -> ``mcd_lnprob`` on ``[n, ndim]``;
* the per-call ``fetch_parameter_values`` + ``inspect`` routing + astropy unit conversions
  (runner.py:143-180, constant.py:140-147, model.py:208-215) -> ``describe()``, once per routing.

Executed by ``tests/test_binding_cpu.py`` (descriptor from the reference's real classes, in the
container that has /root/reference) and ``tests/test_gpu_binding.py`` (patched ``lnprob`` against the
unpatched reference's outputs, on the GPU).
"""
import ctypes
import os

import numpy as np

c_double_p = ctypes.POINTER(ctypes.c_double)

MCD_NPARAM = 11
MCD_MAX_THETA = 16


class mcd_pack_desc(ctypes.Structure):              # include/mcd_b200.h: struct mcd_pack_desc
    _fields_ = [('rotation', ctypes.c_int32), ('background', ctypes.c_int32), ('n_theta', ctypes.c_int32),
                ('math_mode', ctypes.c_int32), ('n_stars', ctypes.c_int64),
                ('ra', c_double_p), ('dec', c_double_p), ('v', c_double_p), ('verr', c_double_p),
                ('pmember', c_double_p), ('density', c_double_p), ('lnlike_background', c_double_p),
                ('slot', ctypes.c_int32 * MCD_NPARAM), ('fixed_value', ctypes.c_double * MCD_NPARAM),
                ('unit_scale', ctypes.c_double * MCD_NPARAM), ('lower', ctypes.c_double * MCD_MAX_THETA),
                ('upper', ctypes.c_double * MCD_MAX_THETA), ('fixed_prior_ok', ctypes.c_int32),
                ('device', ctypes.c_int32), ('n_stars_total', ctypes.c_int64), ('n_segments', ctypes.c_int32),
                ('segment_offsets', ctypes.POINTER(ctypes.c_int64))]


SLOTS = ('v_sys', 'sigma_max', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center', 'a', 'r_peak',
         'v_back', 'sigma_back', 'f_back')                                     # enum MCD_P_*
TARGET = dict(v_sys='km/s', sigma_max='km/s', v_maxx='km/s', v_maxy='km/s', ra_center='deg', dec_center='deg',
              a='arcmin', r_peak='arcmin', v_back='km/s', sigma_back='km/s', f_back='')

#: kernel variant of each reference class: (MCD_ROT_*, MCD_BG_*); the plain classes switch to the fixed-
#: background mixture (1) when a `background=` object was given (analysis/runner.py:96-103,272-286)
VARIANTS = {'ConstantFit': (0, 0), 'ConstantFitGB': (0, 3), 'ModelFit': (1, 0), 'ModelFitGB': (1, 3),
            'ModelFitConstantBackground': (1, 2)}

_lib = None


def library(path=None):
    global _lib
    if _lib is None or path is not None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = path or os.environ.get('MCD_B200_LIB') or os.path.join(
            os.path.dirname(here), 'mcmc_dynamics_b200', '_lib', 'libmcd_b200.so')
        lib = ctypes.CDLL(path)
        lib.mcd_pack_create.argtypes = [ctypes.POINTER(mcd_pack_desc), ctypes.POINTER(ctypes.c_void_p)]
        lib.mcd_pack_create.restype = ctypes.c_int
        lib.mcd_lnprob.argtypes = [ctypes.c_void_p, c_double_p, ctypes.c_int32, c_double_p]
        lib.mcd_lnprob.restype = ctypes.c_int
        lib.mcd_destroy.argtypes = [ctypes.c_void_p]
        lib.mcd_destroy.restype = None
        lib.mcd_last_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def _values(column, unit):
    """Plain float64 array of a data column: astropy Quantity/Column (converted to `unit`) or bare array."""
    if unit is not None and hasattr(column, 'to'):
        try:
            column = column.to(unit)
        except Exception:                      # unit-less column: taken to be in `unit` (runner.py:77-80)
            pass
    return np.ascontiguousarray(np.asarray(getattr(column, 'value', column), dtype=np.float64))


def variant_of(runner):
    rotation, background = VARIANTS[type(runner).__name__]
    if background == 0 and getattr(runner, 'lnlike_background', None) is not None:
        background = 1
    return rotation, background


def describe(runner, rotation=None, background=None, math_mode=0, device=0):
    """Runner -> (mcd_pack_desc, arrays it points into).  Pure host work; `summary()` of the result is what
    the tests compare.  Call again whenever a parameter is fixed/freed/moved or a bound changes."""
    from astropy import units as u
    if rotation is None or background is None:
        rotation, background = variant_of(runner)
    d = mcd_pack_desc(rotation=rotation, background=background, math_mode=math_mode, device=device)
    free = list(runner.fitted_parameters)
    if len(free) > MCD_MAX_THETA:
        raise ValueError('at most %d free parameters' % MCD_MAX_THETA)
    d.n_theta = len(free)
    for j in range(MCD_MAX_THETA):
        d.lower[j], d.upper[j] = -np.inf, np.inf
    for j, name in enumerate(free):
        d.lower[j], d.upper[j] = float(runner.parameters[name].min), float(runner.parameters[name].max)
    for name, p in runner.parameters.items():
        if getattr(p, 'expr', None) is not None:
            raise NotImplementedError("parameter '%s' is constrained by an expression: evaluate it per walker on the "
                                      "host (mcmc_dynamics_b200.analysis.Runner does) or keep the Python path" % name)
    # every parameter -- fixed ones too -- is bounds-checked by Runner.lnprior (runner.py:206-217)
    d.fixed_prior_ok = int(all(float(p.min) <= float(p.value) <= float(p.max)
                               for p in runner.parameters.values() if p.fixed))
    for k, name in enumerate(SLOTS):
        d.slot[k], d.unit_scale[k], d.fixed_value[k] = -1, 1.0, 0.0
        if name in runner.MODEL_PARAMETERS:
            p = runner.parameters[name]
            unit = p.unit if p.unit is not None else u.dimensionless_unscaled
            target = u.Unit(TARGET[name]) if TARGET[name] else u.dimensionless_unscaled
            # a parameter without a unit is taken to be in the kernel's unit of that slot
            d.unit_scale[k] = 1.0 if (p.unit is None and TARGET[name]) else float(unit.to(target))
            d.fixed_value[k] = float(p.value)
            if not p.fixed:
                d.slot[k] = free.index(name)
    keep = {'ra': _values(runner.ra, u.deg), 'dec': _values(runner.dec, u.deg),
            'v': _values(runner.v, u.km / u.s), 'verr': _values(runner.verr, u.km / u.s)}
    if background in (1, 2):
        keep['lnlike_background'] = _values(runner.lnlike_background, None)
    if background == 1:
        keep['pmember'] = _values(runner.pmember, None)
    if background in (2, 3):
        keep['density'] = _values(runner.density, None)
    d.n_stars = keep['v'].size
    for name, arr in keep.items():
        setattr(d, name, arr.ctypes.data_as(c_double_p))
    return d, keep


def summary(desc):
    """The routing part of a descriptor as plain Python values (for comparisons and fixtures)."""
    n = desc.n_theta
    return {'rotation': int(desc.rotation), 'background': int(desc.background), 'n_theta': int(n),
            'n_stars': int(desc.n_stars), 'slot': [int(x) for x in desc.slot],
            'fixed_value': [float(x) for x in desc.fixed_value], 'unit_scale': [float(x) for x in desc.unit_scale],
            'lower': [float(x) for x in desc.lower[:n]], 'upper': [float(x) for x in desc.upper[:n]],
            'fixed_prior_ok': int(desc.fixed_prior_ok)}


def pack(runner, rotation=None, background=None, math_mode=0, device=0):
    """Runner -> device handle (star columns uploaded, routing compiled)."""
    desc, keep = describe(runner, rotation, background, math_mode, device)
    handle = ctypes.c_void_p()
    lib = library()
    if lib.mcd_pack_create(ctypes.byref(desc), ctypes.byref(handle)) != 0:
        raise RuntimeError(lib.mcd_last_error().decode())
    del keep                                   # the library copied the columns
    return handle


def lnprob(handle, values):
    """One vector (the reference's call) or an [n_walkers, n_free] array (emcee with vectorize=True)."""
    theta = np.ascontiguousarray(np.atleast_2d(values), dtype=np.float64)
    out = np.empty(theta.shape[0])
    lib = library()
    if lib.mcd_lnprob(handle, theta.ctypes.data_as(c_double_p), theta.shape[0], out.ctypes.data_as(c_double_p)) != 0:
        raise RuntimeError(lib.mcd_last_error().decode())
    return out if np.ndim(values) == 2 else float(out[0])


def _routing_key(runner):
    return tuple((name, bool(p.fixed), float(p.value) if p.fixed else None, str(p.unit), float(p.min), float(p.max))
                 for name, p in runner.parameters.items())


def patched_lnprob(self, values):
    """Body of the patched ``Runner.lnprob`` (analysis/runner.py:288-306): box prior fused into the kernel,
    exactly -inf for rejected walkers; expression priors (none is shipped) stay on the host."""
    key = _routing_key(self)
    state = self.__dict__.get('_b200')
    if state is None or state[0] != key:
        if state is not None:
            library().mcd_destroy(state[1])
        state = (key, pack(self))
        self.__dict__['_b200'] = state
    out = lnprob(state[1], values)
    if any(getattr(p, 'lnprior', None) is not None for p in self.parameters.values()):
        rows = np.atleast_2d(values)
        extra = np.array([self.lnprior(row) for row in rows], dtype=np.float64)
        with np.errstate(invalid='ignore'):
            total = np.where(np.isfinite(extra), np.atleast_1d(out) + extra, -np.inf)
        return total if np.ndim(values) == 2 else float(total[0])
    return out


def install(runner_class=None):
    """Apply the patch: ``Runner.lnprob`` -> :func:`patched_lnprob`.  Returns the original method."""
    if runner_class is None:
        from mcmc_dynamics.analysis.runner import Runner as runner_class
    original = runner_class.lnprob
    runner_class.lnprob = patched_lnprob
    return original
