"""A very small stand-in for the part of ``astropy.units`` the hot path touches.

The reference attaches astropy units to every parameter and data column
(``parameter.py:762-771``, ``analysis/runner.py:75-81``) and relies on astropy to
reconcile arcmin with arcsec and degrees with radians inside the likelihood
(SURVEY.md section 3.3).  astropy is not a dependency of this package: the GPU
kernels work on plain float64 columns in fixed units, and the unit system is
needed only at the boundary -- to parse the unit strings of the JSON configs,
to convert user-supplied quantities once at pack time, and to hand unit scale
factors to the kernel.

Objects coming from a real astropy installation are accepted wherever a
quantity or unit is expected: anything with ``.value`` and ``.unit`` is treated
as a quantity and its unit is looked up by ``str(unit)``.
"""
import numpy as np

__all__ = ['Unit', 'Quantity', 'UnitConversionError', 'deg', 'rad', 'arcmin', 'arcsec', 'mas', 'km', 's',
           'dimensionless_unscaled', 'one', 'as_unit', 'as_quantity', 'is_quantity', 'strip']


class UnitConversionError(ValueError):
    pass


class Unit(object):
    """A named unit: a physical dimension plus the factor to that dimension's base unit."""

    _registry = {}

    def __new__(cls, name=None, dim=None, scale=None, aliases=(), latex=None):
        # Unit('km/s') or Unit(existing) looks an existing unit up, like astropy's u.Unit(str)
        if dim is None:
            return as_unit(name)
        self = super(Unit, cls).__new__(cls)
        self.name = name
        self.dim = dim
        self.scale = float(scale)
        self.latex = latex if latex is not None else name
        for key in (name,) + tuple(aliases):
            cls._registry[_key(key)] = self
        return self

    def __reduce__(self):
        return (as_unit, (self.name,))

    def to(self, other, value=1.0):
        """Factor (or converted value) from this unit to `other`."""
        other = as_unit(other)
        if other.dim != self.dim:
            raise UnitConversionError("'{0}' and '{1}' are not convertible".format(self.name, other.name))
        return value * (self.scale / other.scale)

    def is_unity(self):
        return self.dim == 'dimensionless' and self.scale == 1.0

    def to_string(self, format=None):
        if format in ('latex', 'latex_inline'):
            return self.latex
        return self.name

    def __str__(self):
        return self.name

    def __repr__(self):
        return 'Unit("{0}")'.format(self.name)

    def __eq__(self, other):
        try:
            other = as_unit(other)
        except (ValueError, TypeError):
            return False
        if other is None:
            return False
        return self.dim == other.dim and self.scale == other.scale

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((self.dim, self.scale))

    # value * unit -> Quantity
    def __rmul__(self, value):
        return Quantity(value, self)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return _compose(self, other, +1)
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return _compose(self, other, -1)
        return Quantity(1.0 / np.asarray(other, dtype=np.float64), self)

    def __rtruediv__(self, value):
        raise UnitConversionError('inverse units are not supported by this minimal unit system')


def _key(name):
    return str(name).replace(' ', '').replace('**', '^')


def _compose(a, b, sign):
    """Only the products the reference spells as arithmetic are known: ``u.km / u.s``."""
    if sign < 0 and a.dim == 'length' and b.dim == 'time':
        scale = a.scale / b.scale          # in km/s
        for unit in set(Unit._registry.values()):
            if unit.dim == 'velocity' and np.isclose(unit.scale, scale, rtol=1e-14):
                return unit
    if b.is_unity():
        return a
    if a.is_unity() and sign > 0:
        return b
    raise UnitConversionError("cannot combine '{0}' and '{1}'".format(a.name, b.name))


def as_unit(unit):
    """Look a unit up by object, by name, or by the string form of a foreign (astropy) unit.

    ``None`` and the empty string mean "no unit" and return ``None``; the caller
    decides whether that is dimensionless (parameter.py:764-765 keeps ``None``).
    """
    if unit is None:
        return None
    if isinstance(unit, Unit):
        return unit
    key = _key(unit)
    if key == '':
        return dimensionless_unscaled
    try:
        return Unit._registry[key]
    except KeyError:
        raise ValueError("Unknown unit '{0}'. Known units: {1}".format(
            unit, ', '.join(sorted(set(u.name for u in Unit._registry.values())))))


# base units: angle -> deg, velocity -> km/s, length -> km, time -> s
dimensionless_unscaled = Unit('', 'dimensionless', 1.0, aliases=('dimensionless', '1', 'None'), latex='')
one = dimensionless_unscaled
deg = Unit('deg', 'angle', 1.0, aliases=('degree', 'degrees'), latex=r'$\mathrm{{}^{\circ}}$')
rad = Unit('rad', 'angle', 180.0 / np.pi, aliases=('radian', 'radians'), latex=r'$\mathrm{rad}$')
arcmin = Unit('arcmin', 'angle', 1.0 / 60.0, aliases=('arcminute',), latex=r'$\mathrm{{}^{\prime}}$')
arcsec = Unit('arcsec', 'angle', 1.0 / 3600.0, aliases=('arcsecond',), latex=r'$\mathrm{{}^{\prime\prime}}$')
mas = Unit('mas', 'angle', 1.0 / 3.6e6, aliases=('milliarcsecond',), latex=r'$\mathrm{mas}$')
km = Unit('km', 'length', 1.0, latex=r'$\mathrm{km}$')
m = Unit('m', 'length', 1.0e-3, latex=r'$\mathrm{m}$')
pc = Unit('pc', 'length', 3.0856775814913674e13, latex=r'$\mathrm{pc}$')
kpc = Unit('kpc', 'length', 3.0856775814913674e16, latex=r'$\mathrm{kpc}$')
s = Unit('s', 'time', 1.0, latex=r'$\mathrm{s}$')
yr = Unit('yr', 'time', 31557600.0, latex=r'$\mathrm{yr}$')
km_s = Unit('km / s', 'velocity', 1.0, aliases=('km/s', 'kms-1', 'km.s-1', 'kms^-1'),
            latex=r'$\mathrm{km\,s^{-1}}$')
m_s = Unit('m / s', 'velocity', 1.0e-3, aliases=('m/s', 'ms-1'), latex=r'$\mathrm{m\,s^{-1}}$')
mas_yr = Unit('mas / yr', 'proper_motion', 1.0, aliases=('mas/yr',), latex=r'$\mathrm{mas\,yr^{-1}}$')


def is_quantity(obj):
    """True for our `Quantity` and for foreign (astropy) quantities."""
    return isinstance(obj, Quantity) or (hasattr(obj, 'unit') and hasattr(obj, 'value')
                                         and not isinstance(obj, (Unit, type)))


def as_quantity(obj, default_unit=None):
    """`u.Quantity(obj)` of the reference: wrap plain numbers as dimensionless (or `default_unit`)."""
    if isinstance(obj, Quantity):
        return obj
    if is_quantity(obj):
        return Quantity(np.asarray(obj.value, dtype=np.float64), as_unit(str(obj.unit)))
    unit = as_unit(default_unit) if default_unit is not None else dimensionless_unscaled
    return Quantity(obj, unit)


def strip(obj, unit):
    """Plain float64 value(s) of `obj` expressed in `unit`.

    Unit-less input is assumed to be in `unit` already, which is what the
    reference does (with a warning) at ``analysis/runner.py:77-80``.
    """
    if is_quantity(obj):
        q = as_quantity(obj)
        if unit is None or q.unit is None or q.unit.is_unity() and not as_unit(unit).is_unity():
            return np.asarray(q.value, dtype=np.float64)
        return np.asarray(q.value, dtype=np.float64) * q.unit.to(unit)
    return np.asarray(obj, dtype=np.float64)


class Quantity(object):
    """Value(s) with a unit.  Deliberately not an ndarray subclass: the hot path never sees it."""

    __array_priority__ = 1000

    def __init__(self, value, unit=None):
        if isinstance(value, Quantity) or is_quantity(value):
            src = as_quantity(value)
            if unit is None:
                unit = src.unit
                value = src.value
            else:
                value = src.to(unit).value
        self.unit = as_unit(unit) if unit is not None else dimensionless_unscaled
        arr = np.asarray(value, dtype=np.float64)
        self.value = float(arr) if arr.ndim == 0 else arr

    def to(self, unit):
        unit = as_unit(unit)
        return Quantity(np.asarray(self.value) * self.unit.to(unit), unit)

    def to_value(self, unit):
        return self.to(unit).value

    @property
    def size(self):
        return np.size(self.value)

    @property
    def shape(self):
        return np.shape(self.value)

    def __len__(self):
        return len(self.value)

    def __getitem__(self, item):
        return Quantity(np.asarray(self.value)[item], self.unit)

    def __iter__(self):
        for x in np.asarray(self.value):
            yield Quantity(x, self.unit)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def __float__(self):
        return float(self.value)

    def __repr__(self):
        return '<Quantity {0} {1}>'.format(self.value, self.unit.name)

    def __str__(self):
        return '{0} {1}'.format(self.value, self.unit.name)

    def _coerce(self, other):
        if is_quantity(other):
            return as_quantity(other).to(self.unit).value
        if self.unit.is_unity():
            return np.asarray(other, dtype=np.float64)
        raise UnitConversionError("Can only combine '{0}' quantities with quantities".format(self.unit.name))

    def __add__(self, other):
        return Quantity(np.asarray(self.value) + self._coerce(other), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        return Quantity(np.asarray(self.value) - self._coerce(other), self.unit)

    def __rsub__(self, other):
        return Quantity(self._coerce(other) - np.asarray(self.value), self.unit)

    def __neg__(self):
        return Quantity(-np.asarray(self.value), self.unit)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, _compose(self.unit, other, +1))
        if is_quantity(other):
            other = as_quantity(other)
            return Quantity(np.asarray(self.value) * other.value, _compose(self.unit, other.unit, +1))
        return Quantity(np.asarray(self.value) * np.asarray(other, dtype=np.float64), self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Quantity(self.value, _compose(self.unit, other, -1))
        if is_quantity(other):
            other = as_quantity(other)
            if other.unit.dim == self.unit.dim:
                return Quantity(np.asarray(self.value) / other.to(self.unit).value, dimensionless_unscaled)
            return Quantity(np.asarray(self.value) / other.value, _compose(self.unit, other.unit, -1))
        return Quantity(np.asarray(self.value) / np.asarray(other, dtype=np.float64), self.unit)

    def _cmp(self, other, op):
        return op(np.asarray(self.value), self._coerce(other))

    def __lt__(self, other):
        return self._cmp(other, np.less)

    def __le__(self, other):
        return self._cmp(other, np.less_equal)

    def __gt__(self, other):
        return self._cmp(other, np.greater)

    def __ge__(self, other):
        return self._cmp(other, np.greater_equal)

    def __eq__(self, other):
        try:
            return self._cmp(other, np.equal)
        except UnitConversionError:
            return False

    def __ne__(self, other):
        return np.logical_not(self.__eq__(other))

    __hash__ = None

    def min(self):
        return Quantity(np.min(self.value), self.unit)

    def max(self):
        return Quantity(np.max(self.value), self.unit)

    def mean(self):
        return Quantity(np.mean(self.value), self.unit)
