"""``Parameters`` / ``Parameter``: the parameter, prior and config boundary of the hot path.

Same public surface and JSON schema as the reference's lmfit-derived classes
(``mcmc_dynamics/parameter.py:30-555`` for the container, ``:558-1007`` for a
single parameter) so that user scripts and stored configs keep working, but
implemented without astropy / asteval / lmfit: units come from
:mod:`mcmc_dynamics_b200.units`, expressions from
:mod:`mcmc_dynamics_b200.expressions`.

JSON layout (``parameter.py:465-466,844-847``)::

    {"unique_symbols": {"rng_seed": null}, "random_state": {...}?,
     "params": [[name, value, unit, fixed, min, max, label, initials, lnprior, user_data, expr], ...]}

The GPU path never evaluates anything in here per star: `pack.py` compiles a
``Parameters`` object into a slot map, unit scale factors and bound vectors once.
"""
import json
import keyword
import logging
import pathlib
import re
from collections import OrderedDict
from copy import deepcopy

import itertools

import numpy as np
from scipy import stats

from . import expressions
from . import units as u

logger = logging.getLogger(__name__)

_NAME_PATTERN = re.compile(r'^[a-zA-Z_][a-zA-Z0-9_]*$')
#: scipy.stats distributions made available to expressions (parameter.py:19-21)
SCIPY_FUNCTIONS = {name: getattr(stats, name) for name in ('uniform', 'norm', 'lognorm')}

#: order of the fields in a serialised parameter (parameter.py:844-847)
STATE_FIELDS = ('name', 'value', 'unit', 'fixed', 'min', 'max', 'label', 'initials', 'lnprior', 'user_data',
                'expr')


def valid_symbol_name(name):
    return isinstance(name, str) and bool(_NAME_PATTERN.match(name)) and not keyword.iskeyword(name)


class Parameters(OrderedDict):
    """Ordered mapping name -> :class:`Parameter` with a shared expression symbol table.

    The order of insertion is the order in which free parameters appear in the
    vector emcee hands to ``lnprob`` (``analysis/runner.py:162-175``).
    """

    def __init__(self, usersyms=None, rng_seed=None, *args, **kwargs):
        super().__init__()
        self._builtin_symbols = set()
        self._symbols = expressions.default_symbols()
        self._symbols.update(SCIPY_FUNCTIONS)
        self._builtin_symbols.update(self._symbols)
        if usersyms is not None:
            self._symbols.update(usersyms)
        self._symbols['rng_seed'] = rng_seed
        self._symbols['rng'] = np.random.default_rng(rng_seed)

    # -- symbol table --------------------------------------------------------------
    @property
    def symtable(self):
        return self._symbols

    def user_defined_symbols(self):
        """Symbols that are neither built in nor parameter names; ``rng_seed`` and ``rng`` count
        as user-defined, as they do for asteval in the reference (parameter.py:448-457)."""
        return [k for k in self._symbols if k not in self._builtin_symbols and k not in self
                and k not in ('n', 'val')]

    def eval(self, expr):
        """Evaluate an expression in the symbol table (parameter.py:214-228)."""
        self._sync_symbols()
        return expressions.evaluate(expressions.parse(expr), self._symbols)

    def _sync_symbols(self):
        for name, par in self.items():
            if par._expr is None:
                self._symbols[name] = par._value

    # -- container protocol ----------------------------------------------------------
    def __setitem__(self, key, par):
        if key not in self and not valid_symbol_name(key):
            raise KeyError("'%s' is not a valid Parameters name" % key)
        if par is not None and not isinstance(par, Parameter):
            raise ValueError("'%s' is not a Parameter" % par)
        OrderedDict.__setitem__(self, key, par)
        par.name = key
        par._owner = self
        self._symbols[key] = par._value

    def __delitem__(self, key):
        OrderedDict.__delitem__(self, key)
        self._symbols.pop(key, None)

    def copy(self):
        return self.__deepcopy__(None)

    def __copy__(self):
        return self.__deepcopy__(None)

    def __deepcopy__(self, memo):
        other = Parameters()
        for key in self.user_defined_symbols():
            if key == 'rng':
                continue
            other._symbols[key] = deepcopy(self._symbols[key])
        other._symbols['rng'] = deepcopy(self._symbols['rng'])
        other.add_many(*[Parameter(**par._as_kwargs()) for par in self.values()])
        return other

    def update(self, other):
        if not isinstance(other, Parameters):
            raise ValueError("'%s' is not a Parameters object" % other)
        self.add_many(*other.values())
        for sym in other.user_defined_symbols():
            self._symbols[sym] = other._symbols[sym]
        return self

    def __add__(self, other):
        if not isinstance(other, Parameters):
            raise ValueError("'%s' is not a Parameters object" % other)
        out = deepcopy(self)
        out.add_many(*[Parameter(**par._as_kwargs()) for par in other.values()])
        for sym in other.user_defined_symbols():
            if sym not in out._symbols:
                out._symbols[sym] = other._symbols[sym]
        return out

    def __iadd__(self, other):
        return self.update(other)

    def __array__(self, dtype=None, copy=None):
        return np.array([float(k) for k in self.values()], dtype=dtype)

    def __reduce__(self):
        symbols = {k: deepcopy(self._symbols[k]) for k in self.user_defined_symbols()}
        return self.__class__, (), {'unique_symbols': symbols, 'params': [self[k] for k in self]}

    def __setstate__(self, state):
        for key, val in state['unique_symbols'].items():
            self._symbols[key] = val
        if state.get('random_state') is not None:
            self._symbols['rng'].bit_generator.state = state['random_state']
        self.add_many(*state['params'])

    # -- construction ------------------------------------------------------------------
    def add(self, name, value=None, unit=None, fixed=False, min=-np.inf, max=np.inf, label=None, initials=None,
            lnprior=None, expr=None):
        """Add one parameter (parameter.py:332-378)."""
        if isinstance(name, Parameter):
            self[name.name] = name
        else:
            self[name] = Parameter(name=name, value=value, unit=unit, fixed=fixed, min=min, max=max, label=label,
                                   initials=initials, lnprior=lnprior, expr=expr)

    def add_many(self, *parlist):
        """Add parameters given as `Parameter` objects or as tuples in constructor order."""
        for par in parlist:
            if not isinstance(par, Parameter):
                par = Parameter(*par)
            self[par.name] = par

    def valuesdict(self):
        return OrderedDict((p.name, p.value) for p in self.values())

    # -- (de)serialisation -------------------------------------------------------------
    def dumps(self, **kws):
        """JSON string in the reference's schema (parameter.py:445-466)."""
        params = []
        for par in self.values():
            state = list(par.__getstate__())
            unit = state[2]
            state[2] = None if unit is None or unit.is_unity() else unit.name.replace(' ', '')
            params.append(state)
        symbols = {}
        for key in self.user_defined_symbols():
            if key == 'rng':
                continue
            value = self._symbols[key]
            try:
                json.dumps(value)
            except TypeError:
                logger.error("Cannot encode user-defined symbol '{0}' as JSON object".format(key))
            else:
                symbols[key] = value
        random_state = _jsonable(self._symbols['rng'].bit_generator.state)
        return json.dumps({'unique_symbols': symbols, 'random_state': random_state, 'params': params}, **kws)

    def loads(self, s, **kws):
        """Replace the content by what a JSON string describes (parameter.py:493-507)."""
        self.clear()
        for name in [k for k in self._symbols if k not in self._builtin_symbols and k not in ('rng', 'rng_seed')]:
            del self._symbols[name]
        tmp = json.loads(s, **kws)
        state = {'unique_symbols': dict(tmp.get('unique_symbols', {})),
                 'random_state': tmp.get('random_state'), 'params': []}
        for parstate in tmp['params']:
            par = Parameter(name='')
            par.__setstate__(parstate)
            state['params'].append(par)
        self.__setstate__(state)
        return self

    def dump(self, fp, **kws):
        return fp.write(self.dumps(**kws))

    def load(self, fp, **kws):
        """Load from an open file or a path-like object (parameter.py:552-555)."""
        if isinstance(fp, (str, pathlib.PurePath)) or hasattr(fp, 'read_text'):
            return self.loads(pathlib.Path(fp).read_text() if isinstance(fp, str) else fp.read_text(), **kws)
        return self.loads(fp.read(), **kws)

    # -- display -------------------------------------------------------------------------
    def pretty_repr(self, oneline=False):
        if oneline:
            return super().__repr__()
        return 'Parameters({\n' + ''.join("    '%s': %s, \n" % (k, self[k]) for k in self) + '    })\n'

    def pretty_print(self, oneline=False, colwidth=8, precision=4, fmt='g', columns=None):
        """Tabulate the parameters, sorted by name (parameter.py:299-326)."""
        if columns is None:
            columns = ['value', 'unit', 'min', 'max', 'fixed', 'initials', 'lnprior']
        if oneline:
            print(self.pretty_repr(oneline=True))
            return
        width = max(len(name) for name in self)
        print(('{:{w}} '.format('Name', w=width)
               + ''.join(' {:>{n}}'.format(c.title(), n=colwidth) for c in columns)))
        for name, par in sorted(self.items()):
            cells = []
            for column in columns:
                cell = getattr(par, column)
                if isinstance(cell, (float, np.floating)) and not isinstance(cell, bool):
                    cells.append('{0:>{n}.{p}{f}}'.format(cell, n=colwidth, p=precision, f=fmt))
                else:
                    cells.append('{0!s:>{n}}'.format(cell, n=colwidth))
            print('{0:<{w}}  '.format(name, w=width) + ' '.join(cells))

    def _repr_html_(self):
        rows = ''.join('<tr><td>{0}</td><td>{1}</td><td>{2}</td><td>{3}</td><td>{4}</td><td>{5}</td></tr>'.format(
            p.name, p.value, p.unit, p.min, p.max, p.fixed) for p in self.values())
        return ('<table><tr><th>name</th><th>value</th><th>unit</th><th>min</th><th>max</th><th>fixed</th></tr>'
                + rows + '</table>')


def _jsonable(obj):
    if isinstance(obj, dict):
        return {k: _jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_jsonable(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    return obj


class Parameter(object):
    """One model parameter: value, unit, bounds, fixed flag and the optional ``initials`` /
    ``lnprior`` / ``expr`` expression strings (parameter.py:558-587)."""

    #: source of ``_version`` stamps: every attribute assignment on any parameter takes a new one, so
    #: ``tuple(p._version for p in parameters.values())`` changes iff something was edited (what the
    #: likelihood entry points compare per call instead of rebuilding the full routing signature)
    _stamps = itertools.count(1)

    def __setattr__(self, key, value):
        object.__setattr__(self, key, value)
        object.__setattr__(self, '_version', next(Parameter._stamps))

    def __init__(self, name, value=None, unit=None, fixed=False, min=-np.inf, max=np.inf, label=None,
                 initials=None, lnprior=None, expr=None, user_data=None):
        self.name = name
        self.fixed = fixed
        self.min = min
        self.max = max
        self.user_data = user_data
        self._label = label
        self._owner = None
        self._value = None
        self.unit = None
        self._initials = self._lnprior = self._expr = None
        self._initials_ast = self._lnprior_ast = self._expr_ast = None
        self._expr_deps = []

        self._set_unit(unit)             # before the value, so that a quantity is converted (parameter.py:580-585)
        self._set_value(value)
        self._init_bounds()
        self._set_expression('initials', initials)
        self._set_expression('lnprior', lnprior)
        self._set_expression('expr', expr)

    def _as_kwargs(self):
        return dict(name=self.name, value=self._value if self._expr is None else self.value, unit=self.unit,
                    fixed=self.fixed, min=self.min, max=self.max, label=self._label, initials=self._initials,
                    lnprior=self._lnprior, expr=self._expr, user_data=self.user_data)

    def set(self, value=None, unit=None, fixed=None, min=None, max=None, label=None, initials=None, lnprior=None,
            expr=None):
        """Change any subset of the attributes; ``None`` leaves one untouched (parameter.py:589-619)."""
        if unit is not None:
            self._set_unit(unit)
        if value is not None:
            self._set_value(value)
        if fixed is not None:
            self.fixed = fixed
        if min is not None:
            self.min = min
        if max is not None:
            self.max = max
        self._init_bounds()
        if initials is not None:
            self._set_expression('initials', initials)
        if lnprior is not None:
            self._set_expression('lnprior', lnprior)
        if expr is not None:
            self._set_expression('expr', expr)
        if label is not None:
            self._label = label

    # -- expressions -----------------------------------------------------------------------
    def _set_expression(self, kind, text):
        if text == '':
            text = None
        setattr(self, '_' + kind, text)
        setattr(self, '_{0}_ast'.format(kind), None if text is None else expressions.parse(text))
        if kind == 'expr':
            if text is not None:
                self.fixed = True        # a constrained parameter is never sampled (parameter.py:725-726)
            self._expr_deps = [] if text is None else expressions.names(self._expr_ast)

    def _symbols(self):
        if self._owner is None:
            return None
        self._owner._sync_symbols()
        return self._owner._symbols

    initials = property(lambda self: self._initials, lambda self, val: self._set_expression('initials', val))
    lnprior = property(lambda self: self._lnprior, lambda self, val: self._set_expression('lnprior', val))
    expr = property(lambda self: self._expr, lambda self, val: self._set_expression('expr', val))

    def evaluate_initials(self, n):
        """`n` start values: the ``initials`` expression with ``n`` and ``rng`` in scope, else draws
        from a unit-width (truncated) normal around the value (parameter.py:642-661)."""
        if self._initials is not None:
            symbols = self._symbols()
            if symbols is None:
                raise IOError("Cannot evaluate 'initials' expression: '{0}'".format(self._initials))
            symbols['n'] = int(n)
            return expressions.evaluate(self._initials_ast, symbols)
        loc = self.value
        scale = 1
        if self.min == -np.inf and self.max == np.inf:
            fct = stats.norm(loc=loc, scale=scale)
        else:
            fct = stats.truncnorm((self.min - loc) / scale, (self.max - loc) / scale, loc=loc, scale=scale)
        return fct.rvs(n)

    def evaluate_lnprior(self, val):
        """Box prior with inclusive bounds, then the optional expression in ``val``
        (parameter.py:684-705).  The reference injects ``val`` through ``'{:f}'`` -- six decimals --
        which is reproduced so that expression priors agree."""
        if u.is_quantity(val):
            val = u.strip(val, self.unit) if self.unit is not None else np.asarray(val.value, dtype=np.float64)
        if val < self.min or val > self.max:
            return -np.inf
        if self._lnprior is None:
            return 0
        symbols = self._symbols()
        if symbols is None:
            raise IOError("Cannot evaluate expression: '{0}'".format(self._lnprior))
        symbols['val'] = float('{0:f}'.format(float(val)))
        return expressions.evaluate(self._lnprior_ast, symbols)

    # -- value, unit, bounds ---------------------------------------------------------------
    def _set_value(self, val):
        if u.is_quantity(val):
            quantity = u.as_quantity(val)
            if self.unit is not None:
                try:
                    val = quantity.to(self.unit).value
                except u.UnitConversionError:
                    raise IOError("Unit '{0}' of new value incompatible with existing unit '{1}'.".format(
                        quantity.unit, self.unit))
            else:
                self._set_unit(quantity.unit)
                val = quantity.value
        if val is not None and np.ndim(val) == 0 and not isinstance(val, (bool, str)):
            val = float(val)
        self._value = val
        if self._owner is not None:
            self._owner._symbols[self.name] = self._value

    def _set_unit(self, unit):
        if unit is None:
            return
        unit = u.as_unit(unit if isinstance(unit, (str, u.Unit)) else str(unit))
        if self.unit is None:
            self.unit = unit
        elif unit != self.unit:
            logger.error("Cannot change unit from '{0}' to '{1}'.".format(self.unit, unit))

    def _init_bounds(self):
        """Keep min/max/value mutually consistent (parameter.py:773-806)."""
        if self.max is None:
            self.max = np.inf
        if self.min is None:
            self.min = -np.inf
        for side in ('min', 'max'):
            bound = getattr(self, side)
            if u.is_quantity(bound):
                bound = u.as_quantity(bound)
                if self.unit is None:
                    self.unit = bound.unit
                try:
                    setattr(self, side, float(bound.to(self.unit).value))
                except u.UnitConversionError:
                    raise IOError("Incompatible units provided for '{0}' of parameter '{1}'.".format(
                        side, self.name))
        if self._value is None and self._expr is None:
            if np.isfinite(self.min) & np.isfinite(self.max):
                self._value = (self.min + self.max) / 2.
            else:
                self._value = 0.
        if self.min > self.max:
            self.min, self.max = self.max, self.min
        if np.isclose(self.min, self.max, atol=1e-13, rtol=1e-13):
            raise ValueError("Parameter '%s' has min == max" % self.name)
        if self._expr is None:
            if self._value > self.max:
                self._value = self.max
            if self._value < self.min:
                self._value = self.min

    @property
    def value(self):
        """Current value; a constrained parameter is re-evaluated from its expression
        (parameter.py:865-874)."""
        if self._expr is not None:
            symbols = self._symbols()
            if symbols is not None:
                self._value = expressions.evaluate(self._expr_ast, symbols)
        return self._value

    @value.setter
    def value(self, val):
        self._set_value(val)

    @property
    def label(self, format='latex_inline'):
        text = self._label if self._label is not None else r"${{\rm {0}}}$".format(self.name)
        if self.unit is not None:
            text += "/" + self.unit.to_string(format)
        return text

    @label.setter
    def label(self, val):
        self._label = val

    # -- pickling / JSON -------------------------------------------------------------------
    def __getstate__(self):
        return (self.name, self.value, self.unit, self.fixed, self.min, self.max, self._label, self.initials,
                self.lnprior, self.user_data, self.expr)

    def __setstate__(self, state):
        (name, value, unit, fixed, lo, hi, label, initials, lnprior, user_data, expr) = state
        self.__init__(name=name, value=value, unit=unit, fixed=fixed, min=lo, max=hi, label=label,
                      initials=initials, lnprior=lnprior, expr=expr, user_data=user_data)

    def __repr__(self):
        parts = ["value=%s" % repr(self.value) + (" (fixed)" if self.fixed and self._expr is None else "")
                 + (" unit={0}".format(self.unit) if self.unit is not None else "")]
        parts.append("bounds=[%s:%s]" % (repr(self.min), repr(self.max)))
        if self._initials is not None:
            parts.append("initials='%s'" % self.initials)
        if self._expr is not None:
            parts.append("expr='%s'" % self.expr)
        if self._lnprior is not None:
            parts.append("lnprior=%s" % self.lnprior)
        return "<Parameter '%s', %s>" % (self.name, ', '.join(parts))

    __str__ = __repr__

    # -- numeric protocol: a Parameter behaves like its value (parameter.py:886-1007) --------
    def __array__(self, dtype=None, copy=None):
        return np.array(float(self.value), dtype=dtype)

    def __float__(self):
        return float(self.value)

    def __int__(self):
        return int(self.value)

    def __bool__(self):
        return self.value != 0

    def __abs__(self):
        return abs(self.value)

    def __neg__(self):
        return -self.value

    def __pos__(self):
        return +self.value

    def __trunc__(self):
        return self.value.__trunc__()


def _forward(op, reflected=False):
    import operator
    fn = getattr(operator, op) if hasattr(operator, op) else divmod
    if reflected:
        return lambda self, other: fn(other, self.value)
    return lambda self, other: fn(self.value, other)


for _op in ('add', 'sub', 'mul', 'truediv', 'floordiv', 'mod', 'pow', 'divmod'):
    setattr(Parameter, '__{0}__'.format(_op), _forward(_op))
    setattr(Parameter, '__r{0}__'.format(_op), _forward(_op, reflected=True))
for _op in ('lt', 'le', 'gt', 'ge', 'eq', 'ne'):
    setattr(Parameter, '__{0}__'.format(_op), _forward(_op))
Parameter.__hash__ = object.__hash__
