"""Stand-ins with the attribute surface of the reference's ``Runner`` objects, built from the golden cases.

The Option-B binding (``tools/reference_binding.py``) reads ``parameters`` (objects with ``value, unit, fixed,
min, max, lnprior, expr``), ``fitted_parameters``, ``MODEL_PARAMETERS``, the data columns ``ra, dec, v, verr``
(astropy quantities), ``lnlike_background``, ``pmember`` and ``density`` -- nothing else -- of a reference model
object.  The reference's classes cannot be loaded where the GPU tests run (``/root/reference`` does not exist
there), so those tests hand the binding an object with exactly that surface, filled from the inputs stored in
``golden_reference.json``; ``tests/test_binding_cpu.py`` checks that the binding compiles the SAME descriptor
from such a stand-in as it did from the real reference object when the golden file was generated.
``astropy.units`` is the stand-in of ``tests/golden/ref_shims`` (real astropy is not installed).
"""
import collections
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def binding_module():
    """tools/reference_binding.py with ``astropy.units`` resolved to the stand-in of tests/golden/ref_shims.  The
    shim directory is taken off ``sys.path`` again at once: it also holds import-only stubs of emcee, pathos,
    ... that must not shadow anything for the rest of the test session."""
    tools = os.path.join(ROOT, 'tools')
    if tools not in sys.path:
        sys.path.insert(0, tools)
    if 'astropy.units' not in sys.modules:
        shims = os.path.join(HERE, 'golden', 'ref_shims')
        sys.path.insert(0, shims)
        try:
            import astropy.units  # noqa: F401
        finally:
            sys.path.remove(shims)
    import reference_binding
    return reference_binding


class Par(object):
    def __init__(self, name, value, unit, fixed, lo, hi):
        self.name, self.value, self.unit, self.fixed, self.min, self.max = name, value, unit, fixed, lo, hi
        self.lnprior = None
        self.expr = None


class RunnerStandIn(object):
    """Attribute surface of ``mcmc_dynamics.analysis.runner.Runner`` as far as the binding reads it."""
    MODEL_PARAMETERS = []

    @property
    def fitted_parameters(self):
        return [name for name, p in self.parameters.items() if not p.fixed]

    def lnprior(self, values):        # box prior only (no expression priors in the golden cases)
        raise AssertionError('the patched lnprob must not fall back to the Python prior here')

    def lnprob(self, values):         # replaced by reference_binding.install()
        raise AssertionError('unpatched')


def stand_in_for_case(case, golden):
    binding_module()
    from astropy import units as u
    cls = type(case['class'], (RunnerStandIn,), {'MODEL_PARAMETERS': list(case['model_parameters'])})
    obj = cls()
    obj.parameters = collections.OrderedDict()
    for name, value, unit, fixed, lo, hi, _initials in golden['default_parameters'][case['class']]:
        obj.parameters[name] = Par(name, value, None if unit is None else u.Unit(unit), fixed, lo, hi)
    for name, edit in case['parameter_edits'].items():
        p = obj.parameters[name]
        for key, val in edit.items():
            setattr(p, {'min': 'min', 'max': 'max'}.get(key, key), val)
    cols = {k: np.asarray(v, dtype=np.float64) for k, v in case['columns'].items()}
    obj.ra, obj.dec = u.Quantity(cols['ra'], u.deg), u.Quantity(cols['dec'], u.deg)
    obj.v, obj.verr = u.Quantity(cols['v'], u.km / u.s), u.Quantity(cols['verr'], u.km / u.s)
    obj.lnlike_background = np.asarray(case['lnlike_background']) if 'lnlike_background' in case else None
    obj.pmember = cols.get('pmember') if obj.lnlike_background is not None else None
    obj.density = u.Quantity(cols['density']) if 'density' in cols and 'density' in _observables(case) else None
    return obj


def _observables(case):
    # constant.py:257, model.py:355,527: the *GB / ConstantBackground classes read the density column
    return ('density',) if case['class'] in ('ConstantFitGB', 'ModelFitGB', 'ModelFitConstantBackground') else ()
