"""Compile a model object (class + ``Parameters`` + star columns) into the ``mcd_pack_desc`` the CUDA
library consumes, and own the resulting device handle.

This is everything the reference redoes on every ``lnprob`` call, hoisted out of the sampling loop:
``Runner.fetch_parameter_values`` (``analysis/runner.py:143-180``), the ``inspect``-based routing
of parameters to ``rotation_model`` / ``dispersion_model`` (``constant.py:140-147``,
``model.py:208-215``) and the implicit astropy unit conversions.
"""
import ctypes
import logging

import numpy as np

from . import _native
from . import units as u

logger = logging.getLogger(__name__)

#: unit every parameter slot is converted to before it reaches the kernel
SLOT_UNITS = {
    'v_sys': u.km_s, 'sigma_max': u.km_s, 'v_maxx': u.km_s, 'v_maxy': u.km_s, 'ra_center': u.deg,
    'dec_center': u.deg, 'a': u.arcmin, 'r_peak': u.arcmin, 'v_back': u.km_s, 'sigma_back': u.km_s,
    'f_back': u.dimensionless_unscaled,
}


class PackError(ValueError):
    pass


def _unit_scale(name, unit):
    """Factor from the parameter's own unit to the kernel's unit of that slot.  A parameter without a
    unit is taken to be in the kernel's unit already."""
    if unit is None or (unit.is_unity() and not SLOT_UNITS[name].is_unity()):
        return 1.0
    try:
        return unit.to(SLOT_UNITS[name])
    except u.UnitConversionError:
        raise PackError("Parameter '{0}' has unit '{1}', which cannot be converted to '{2}'.".format(
            name, unit, SLOT_UNITS[name]))


def edit_stamps(parameters):
    """Cheap change detector for a ``Parameters`` object: the version stamps of its members (every
    attribute assignment on a ``Parameter`` takes a new stamp).  Equal stamps = nothing was edited."""
    return tuple([par._version for par in parameters.values()])


def routing_signature(parameters, model_parameters):
    """Everything about a ``Parameters`` object the packed state depends on; a change triggers a
    re-pack before the next likelihood call (parameters are usually edited between construction and
    the run: ``bin/run_tests.py:88-93,137-148``)."""
    sig = []
    derived = set(derived_parameters(parameters))
    for name, par in parameters.items():
        # the value of a per-walker constrained parameter is not part of the packed state (it travels as
        # an extra column of theta)
        value = None if (not par.fixed or name in derived) else float(par.value)
        sig.append((name, bool(par.fixed), value, str(par.unit), float(par.min), float(par.max), par.expr))
    return tuple(sig), tuple(model_parameters)


def derived_parameters(parameters):
    """Names of the parameters whose ``expr`` constraint depends -- directly or through other constrained
    parameters -- on a sampled parameter, in insertion order.  The reference re-evaluates them through
    asteval on every call (``analysis/runner.py:163-176``, ``parameter.py:865-874``); here they are
    evaluated on the host once per walker and handed to the kernel as extra columns of theta."""
    free = {name for name, par in parameters.items() if not par.fixed}
    derived = []
    changed = True
    while changed:
        changed = False
        for name, par in parameters.items():
            if name in derived or par.expr is None:
                continue
            deps = set(getattr(par, '_expr_deps', []))
            if deps & free or deps & set(derived):
                derived.append(name)
                changed = True
    return [name for name in parameters if name in derived]


def build_descriptor(parameters, model_parameters, rotation, background, columns, math_mode=_native.MATH_FAST,
                     device=0, n_stars_total=0, segment_offsets=None, derived=()):
    """Fill a ``PackDesc`` from a ``Parameters`` object.

    Parameters
    ----------
    parameters : Parameters
        All parameters in insertion order; the free ones define the columns of theta.
    model_parameters : list of str
        ``MODEL_PARAMETERS`` of the class: which of them actually enter the likelihood.
    columns : dict name -> contiguous float64 array (``ra, dec, v, verr`` and optionally ``pmember``,
        ``density``, ``lnlike_background``)
    derived : names of per-walker constrained parameters (:func:`derived_parameters`): they occupy the
        columns of theta after the free ones, with their own bounds, and are filled in by the caller.

    Returns
    -------
    desc : PackDesc
    keep : list
        Arrays the descriptor points into (keep alive until ``mcd_pack_create`` returns).
    """
    desc = _native.PackDesc()
    desc.rotation = rotation
    desc.background = background
    desc.math_mode = math_mode
    desc.device = device
    desc.n_stars_total = n_stars_total

    derived = list(derived)
    free = [name for name, par in parameters.items() if not par.fixed] + derived
    if len(free) > _native.MAX_THETA:
        raise PackError('At most {0} free (plus per-walker constrained) parameters are supported, got {1}.'.format(
            _native.MAX_THETA, len(free)))
    desc.n_theta = len(free)
    for j in range(_native.MAX_THETA):
        desc.lower[j] = -np.inf
        desc.upper[j] = np.inf
    for j, name in enumerate(free):
        desc.lower[j] = float(parameters[name].min)
        desc.upper[j] = float(parameters[name].max)

    # every parameter -- fixed ones too -- is bounds-checked by Runner.lnprior (runner.py:206-217)
    fixed_ok = 1
    for name, par in parameters.items():
        if par.fixed and name not in derived:
            value = float(par.value)
            if value < par.min or value > par.max:
                fixed_ok = 0
    desc.fixed_prior_ok = fixed_ok

    for k, slot_name in enumerate(_native.PARAM_SLOTS):
        desc.slot[k] = -1
        desc.fixed_value[k] = 0.0
        desc.unit_scale[k] = 1.0
        if slot_name not in model_parameters:
            continue
        par = parameters[slot_name]
        desc.unit_scale[k] = _unit_scale(slot_name, par.unit)
        value = par.value
        desc.fixed_value[k] = 0.0 if value is None else float(value)
        if not par.fixed or slot_name in derived:
            desc.slot[k] = free.index(slot_name)

    keep = []
    n = None
    for name in ('ra', 'dec', 'v', 'verr', 'pmember', 'density', 'lnlike_background'):
        arr = columns.get(name)
        if arr is None:
            continue
        arr = _native.contiguous(arr)
        if n is None:
            n = arr.size
        elif arr.size != n:
            raise PackError("Column '{0}' has {1} rows, expected {2}.".format(name, arr.size, n))
        keep.append(arr)
        setattr(desc, name, _native.as_double_ptr(arr))
    desc.n_stars = 0 if n is None else n
    if segment_offsets is not None:
        offsets = np.ascontiguousarray(segment_offsets, dtype=np.int64)
        keep.append(offsets)
        desc.n_segments = offsets.size - 1
        desc.segment_offsets = offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
    return desc, keep


class PackedModel(object):
    """Owner of one ``mcd_handle``: star columns resident on one GPU plus the current routing."""

    def __init__(self, desc, keep):
        self._lib = _native.load_library()
        handle = ctypes.c_void_p()
        _native.check(self._lib.mcd_pack_create(ctypes.byref(desc), ctypes.byref(handle)))
        self._handle = handle
        self.n_theta = int(desc.n_theta)
        self.n_stars = int(desc.n_stars)
        self.n_segments = max(1, int(desc.n_segments))
        self.device = int(desc.device)
        del keep

    @property
    def handle(self):
        if self._handle is None:
            raise _native.NativeError('the packed model has been closed')
        return self._handle

    def reconfigure(self, desc):
        _native.check(self._lib.mcd_pack_reconfigure(self.handle, ctypes.byref(desc)))
        self.n_theta = int(desc.n_theta)

    def info(self):
        info = _native.Info()
        _native.check(self._lib.mcd_get_info(self.handle, ctypes.byref(info)))
        return {name: getattr(info, name) for name, _ in _native.Info._fields_}

    def _theta(self, theta):
        """[n_walkers, n_theta], or [n_segments, n_walkers, n_theta] for a segmented handle."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if self.n_segments > 1:
            if theta.ndim != 3 or theta.shape[0] != self.n_segments or theta.shape[2] != self.n_theta:
                raise ValueError('theta must have shape ({0}, n_walkers, {1}), got {2}'.format(
                    self.n_segments, self.n_theta, theta.shape))
            return theta, theta.shape[1], theta.shape[:2]
        if theta.ndim != 2 or theta.shape[1] != self.n_theta:
            raise ValueError('theta must have shape (n_walkers, {0}), got {1}'.format(self.n_theta, theta.shape))
        return theta, theta.shape[0], theta.shape[:1]

    def lnprob(self, theta):
        """Host buffers in, host buffer out: copy, one kernel launch, copy, synchronise."""
        theta, n_walkers, shape = self._theta(theta)
        out = np.empty(shape, dtype=np.float64)
        rc = self._lib.mcd_lnprob(self.handle, _native.address(theta), n_walkers, _native.address(out))
        if rc != 0:
            _native.check(rc)
        return out

    def lnlike(self, theta):
        theta, n_walkers, shape = self._theta(theta)
        out = np.empty(shape, dtype=np.float64)
        rc = self._lib.mcd_lnlike(self.handle, _native.address(theta), n_walkers, _native.address(out))
        if rc != 0:
            _native.check(rc)
        return out

    def lnlike_per_star(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
        if theta.size != self.n_theta:
            raise ValueError('theta must have {0} entries'.format(self.n_theta))
        out = np.empty(self.n_stars, dtype=np.float64)
        _native.check(self._lib.mcd_lnlike_per_star(self.handle, _native.as_double_ptr(theta),
                                                    _native.as_double_ptr(out)))
        return out

    def membership_per_star(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
        if theta.size != self.n_theta:
            raise ValueError('theta must have {0} entries'.format(self.n_theta))
        out = np.empty(self.n_stars, dtype=np.float64)
        _native.check(self._lib.mcd_membership_per_star(self.handle, _native.as_double_ptr(theta),
                                                        _native.as_double_ptr(out)))
        return out

    def model_per_star(self, theta):
        """``(v_los, sigma_los)`` of every star in km/s for ONE parameter vector (``mcd_model_per_star``)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
        if theta.size != self.n_theta:
            raise ValueError('theta must have {0} entries'.format(self.n_theta))
        v_los = np.empty(self.n_stars, dtype=np.float64)
        sigma_los = np.empty(self.n_stars, dtype=np.float64)
        _native.check(self._lib.mcd_model_per_star(self.handle, _native.as_double_ptr(theta),
                                                   _native.as_double_ptr(v_los), _native.as_double_ptr(sigma_los)))
        return v_los, sigma_los

    def calculate_lnlike(self, v_los, sigma_los):
        """``Runner._calculate_lnlike`` for curves computed by the caller, km/s per star (``mcd_calculate_lnlike``)."""
        v_los = np.ascontiguousarray(np.broadcast_to(np.asarray(v_los, dtype=np.float64), (self.n_stars,)))
        sigma_los = np.ascontiguousarray(np.broadcast_to(np.asarray(sigma_los, dtype=np.float64), (self.n_stars,)))
        out = ctypes.c_double(0.0)
        _native.check(self._lib.mcd_calculate_lnlike(self.handle, _native.as_double_ptr(v_los),
                                                     _native.as_double_ptr(sigma_los), ctypes.byref(out)))
        return out.value

    # ---- device tensors (torch operator library; current CUDA stream) ----------------------
    def lnprob_tensor(self, theta):
        return _native.load_torch_ops().lnprob(self.handle.value, theta)

    def lnlike_tensor(self, theta):
        return _native.load_torch_ops().lnlike(self.handle.value, theta)

    def lnprob_partial_tensor(self, theta):
        return _native.load_torch_ops().lnprob_partial(self.handle.value, theta)

    def close(self):
        if self._handle is not None:
            self._lib.mcd_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:       # interpreter shutdown
            pass
