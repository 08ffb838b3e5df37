"""Glue between product model objects and the NumPy oracle.  TEST INFRASTRUCTURE ONLY.

``oracle_for(model)`` builds the oracle class of the same name from a product model object by
duck typing (it reads ``.parameters``, ``.v/.verr/.ra/.dec[/.density]``, ``.lnlike_background``,
``.pmember``); nothing here imports the product package.
"""
import numpy as np

from . import reference_np as ref


def _values(column):
    return np.asarray(getattr(column, 'value', column), dtype=np.float64)


def oracle_parameters(parameters):
    rows = []
    for name, par in parameters.items():
        unit = None if par.unit is None else str(par.unit)
        if unit == '':
            unit = None
        rows.append(ref.OParam(name, value=par.value, unit=unit, fixed=par.fixed, min=par.min, max=par.max))
    return rows


def oracle_for(model, lnlike_background=None):
    """Oracle twin of a product model object.  `lnlike_background` overrides the model's own
    (GPU-computed) background column, e.g. with the oracle's `single_stars_background`."""
    cls = ref.ORACLE_CLASSES[type(model).__name__]
    data = {'v': _values(model.v), 'verr': _values(model.verr), 'ra': _values(model.ra), 'dec': _values(model.dec)}
    if getattr(model, 'density', None) is not None:
        data['density'] = _values(model.density)
    lbg = lnlike_background if lnlike_background is not None else getattr(model, 'lnlike_background', None)
    pmember = getattr(model, 'pmember', None)
    return cls(data, parameters=oracle_parameters(model.parameters),
               lnlike_background=None if lbg is None else _values(lbg),
               pmember=None if pmember is None else _values(pmember))


def relative_error(got, want):
    """max |got - want| / max(1, |want|), treating matching infinities as exact."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    same_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
    with np.errstate(invalid='ignore'):
        err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    err[same_inf] = 0.0
    return float(np.nanmax(err)) if not np.any(np.isnan(err[~same_inf])) else float('nan')
