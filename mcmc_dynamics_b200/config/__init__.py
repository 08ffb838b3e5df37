"""Default parameter sets of the model classes.

The reference ships them as JSON files (``mcmc_dynamics/config/*.json``) that its classes read with
``Parameters().load(parameters_file)``.  Here the defaults are one Python table; ``default_file(name)``
serialises a set with the package's own ``Parameters.dumps`` into ``config/_generated/<name>.json`` the
first time it is asked for, so that ``Model.parameters_file`` / ``Model.default_parameters()`` keep
working with a real file in the reference's schema.  ``tests/test_golden_cpu.py`` checks every set
against what the reference's own ``Parameters.load`` makes of its files (names, order, units, bounds,
default values, ``initials`` expressions).
"""
import math
import os
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
_INF = math.inf
_NORMAL, _LOGNORMAL, _UNIFORM = 'rng.normal(size=n)', 'rng.lognormal(size=n)', 'rng.uniform(size=n)'

#: name -> (unit, min, max, LaTeX label, initials expression)
PARAMETERS = {
    'v_sys': ('km/s', -_INF, _INF, r'$v_{\rm sys}$', _NORMAL),
    'sigma_max': ('km/s', 0.0, _INF, r'$\sigma_{\rm max}$', _LOGNORMAL),
    'v_maxx': ('km/s', -_INF, _INF, r'$v_{\rm max,\,x}$', _NORMAL),
    'v_maxy': ('km/s', -_INF, _INF, r'$v_{\rm max,\,y}$', _NORMAL),
    'ra_center': ('deg', 0.0, 360.0, r'$\alpha_{\rm c}$', None),
    'dec_center': ('deg', -90.0, 90.0, r'$\delta_{\rm c}$', None),
    'a': ('arcsec', 0.0, _INF, r'$a$', _LOGNORMAL),
    'r_peak': ('arcsec', 0.0, _INF, r'$r_{\rm peak}$', _LOGNORMAL),
    'v_back': ('km/s', -_INF, _INF, r'$v_{\rm back}$', _NORMAL),
    'sigma_back': ('km/s', 0.0, _INF, r'$\sigma_{\rm back}$', _LOGNORMAL),
    'f_back': (None, 0.0, 1.0, r'$f_{\rm back}$', _UNIFORM),
}

#: parameter order of each set = column order of theta for the free ones
SETS = {
    'constant': ('v_sys', 'sigma_max', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center'),
    'constant_with_background': ('v_sys', 'sigma_max', 'v_maxx', 'v_maxy', 'ra_center', 'dec_center', 'v_back',
                                 'sigma_back', 'f_back'),
    'model': ('v_sys', 'sigma_max', 'a', 'v_maxx', 'ra_center', 'dec_center', 'v_maxy', 'r_peak'),
    'model_with_background': ('v_sys', 'sigma_max', 'a', 'v_maxx', 'v_maxy', 'r_peak', 'ra_center', 'dec_center',
                              'v_back', 'sigma_back', 'f_back'),
}


def build(name):
    """A fresh ``Parameters`` object holding the default set `name`."""
    from ..parameter import Parameters
    pars = Parameters()
    for key in SETS[name]:
        unit, lo, hi, label, initials = PARAMETERS[key]
        pars.add(key, value=None, unit=unit, fixed=False, min=lo, max=hi, label=label, initials=initials)
    return pars


def default_text(name):
    """JSON text of set `name` in the layout of the reference's ``config/*.json``: ``unique_symbols``
    (``rng_seed: null``) and ``params`` -- and, like those files, NO ``random_state`` entry, so that a
    ``Parameters`` object loaded from it keeps the unseeded generator it was constructed with
    (``mcmc_dynamics/parameter.py:73-74,207-209``; ``config/model.json:1-5``).  ``Parameters.dumps`` itself
    stores the generator state (``parameter.py:458-466``), which is right for a user's snapshot but would
    freeze the start positions drawn from the defaults in every process."""
    import json
    state = json.loads(build(name).dumps())
    state.pop('random_state', None)
    return json.dumps(state)


def default_file(name):
    """Path of the JSON serialisation of set `name` (written on first use, rewritten if stale)."""
    text = default_text(name)
    for out_dir in (os.path.join(_HERE, '_generated'),
                    os.path.join(tempfile.gettempdir(), 'mcmc_dynamics_b200_config_%d' % os.getuid())):
        path = os.path.join(out_dir, name + '.json')
        try:
            with open(path) as f:
                if f.read() == text:
                    return path
        except OSError:
            pass
        try:
            os.makedirs(out_dir, exist_ok=True)
            tmp = path + '.%d.tmp' % os.getpid()
            with open(tmp, 'w') as f:
                f.write(text)
            os.replace(tmp, path)
            return path
        except OSError:            # read-only installation: fall back to the temporary directory
            continue
    raise IOError('cannot write the default parameter file for {0}'.format(name))
