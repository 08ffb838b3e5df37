"""Stand-in for the part of astropy.units the reference's hot path uses (see ../../README.md)."""
import numpy as np


class UnitConversionError(ValueError):
    pass


class UnitTypeError(TypeError):
    pass


class UnitsError(ValueError):
    pass


class _Core(object):
    UnitConversionError = UnitConversionError
    UnitTypeError = UnitTypeError


core = _Core()
_BASES = ('km', 's', 'rad')


class UnitBase(object):
    """scale * km^a s^b rad^c."""

    def __init__(self, scale, powers, name=None):
        self.scale = float(scale)
        self.powers = tuple(float(p) for p in powers)
        self._name = name

    # -- algebra -------------------------------------------------------------------------------
    def _combine(self, other, sign):
        return UnitBase(self.scale * other.scale ** sign, [a + sign * b for a, b in zip(self.powers, other.powers)])

    def __mul__(self, other):
        if isinstance(other, UnitBase):
            return self._combine(other, 1)
        return Quantity(other, self)

    def __rmul__(self, other):
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, UnitBase):
            return self._combine(other, -1)
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        return Quantity(other, UnitBase(1.0 / self.scale, [-p for p in self.powers]))

    def __pow__(self, p):
        return UnitBase(self.scale ** p, [q * p for q in self.powers])

    # -- queries -------------------------------------------------------------------------------
    @property
    def is_dimensionless(self):
        return all(p == 0 for p in self.powers)

    def is_unity(self):
        return self.is_dimensionless and self.scale == 1.0

    def is_equivalent(self, other):
        return self.powers == Unit(other).powers

    def to(self, other, value=1.0):
        other = Unit(other)
        if self.powers != other.powers:
            raise UnitConversionError("'{0}' and '{1}' are not convertible".format(self, other))
        return value * (self.scale / other.scale)

    def __eq__(self, other):
        try:
            other = Unit(other)
        except Exception:
            return False
        return self.powers == other.powers and np.isclose(self.scale, other.scale, rtol=1e-14, atol=0)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((round(np.log(self.scale), 10), self.powers))

    def to_string(self, format=None):
        if self._name is not None:
            return self._name
        for name, unit in _NAMED.items():
            if unit == self:
                return name
        parts = ['{0}{1:g}'.format(b, p) for b, p in zip(_BASES, self.powers) if p]
        return '{0:g} {1}'.format(self.scale, ' '.join(parts))

    __str__ = to_string

    def __repr__(self):
        return 'Unit("{0}")'.format(self.to_string())


dimensionless_unscaled = UnitBase(1.0, (0, 0, 0), '')
one = dimensionless_unscaled
km = UnitBase(1.0, (1, 0, 0), 'km')
m = UnitBase(1e-3, (1, 0, 0), 'm')
s = UnitBase(1.0, (0, 1, 0), 's')
yr = UnitBase(31557600.0, (0, 1, 0), 'yr')
rad = UnitBase(1.0, (0, 0, 1), 'rad')
deg = UnitBase(np.pi / 180.0, (0, 0, 1), 'deg')
arcmin = UnitBase(np.pi / 180.0 / 60.0, (0, 0, 1), 'arcmin')
arcsec = UnitBase(np.pi / 180.0 / 3600.0, (0, 0, 1), 'arcsec')
mas = UnitBase(np.pi / 180.0 / 3.6e6, (0, 0, 1), 'mas')
pc = UnitBase(3.0856775814913674e13, (1, 0, 0), 'pc')
kpc = UnitBase(3.0856775814913674e16, (1, 0, 0), 'kpc')
_NAMED = {'': dimensionless_unscaled, 'km': km, 'm': m, 's': s, 'yr': yr, 'rad': rad, 'deg': deg, 'arcmin': arcmin,
          'arcsec': arcsec, 'mas': mas, 'pc': pc, 'kpc': kpc}
_km_s = km / s
_km_s._name = 'km / s'
_NAMED['km / s'] = _km_s


def Unit(x):
    if isinstance(x, UnitBase):
        return x
    if isinstance(x, Quantity):
        raise UnitTypeError('a quantity is not a unit')
    if x is None:
        raise TypeError('None is not a valid Unit')
    text = str(x).strip()
    if text in ('', 'dimensionless'):
        return dimensionless_unscaled
    if text.replace(' ', '') == 'km/s':
        return _km_s
    out = dimensionless_unscaled
    sign = 1
    for token in text.replace('/', ' / ').split():
        if token == '/':
            sign = -1
            continue
        if token not in _NAMED:
            raise ValueError("unit '{0}' is not known to the astropy stand-in".format(token))
        out = out._combine(_NAMED[token], sign)
    return out


class Dex(object):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('Dex is not needed on the hot path')


def _unit_of(x):
    return x.unit if isinstance(x, Quantity) else dimensionless_unscaled


def _plain(x):
    return x.view(np.ndarray) if isinstance(x, Quantity) else np.asarray(x)


def _as_unit_value(x, unit):
    """Values of `x` expressed in `unit`; a bare number counts as dimensionless (astropy)."""
    if isinstance(x, Quantity):
        return _plain(x) * x.unit.to(unit)
    if not unit.is_dimensionless:
        # astropy allows 0, inf and nan to be combined with any unit
        arr = np.asarray(x)
        if arr.dtype.kind in 'fiu' and np.all((arr == 0) | ~np.isfinite(arr)):
            return arr
        raise UnitConversionError("Can only apply this function to quantities with compatible dimensions")
    return np.asarray(x) * dimensionless_unscaled.to(unit)


_SAME_UNIT = {np.add, np.subtract, np.maximum, np.minimum, np.fmax, np.fmin, np.hypot}
_COMPARE = {np.less, np.less_equal, np.greater, np.greater_equal, np.equal, np.not_equal}
_KEEP = {np.negative, np.positive, np.absolute, np.fabs, np.rint, np.floor, np.ceil}
_PLAIN_OUT = {np.isfinite, np.isnan, np.isinf, np.signbit, np.sign}
_TRIG = {np.sin, np.cos, np.tan}
_INV_TRIG = {np.arcsin, np.arccos, np.arctan}
_DIMLESS = {np.exp, np.exp2, np.expm1, np.log, np.log2, np.log10, np.log1p}


class Quantity(np.ndarray):
    __array_priority__ = 10000

    def __new__(cls, value, unit=None, dtype=None, copy=True, **kwargs):
        if isinstance(value, Quantity):
            if unit is None:
                unit, arr = value.unit, np.array(_plain(value), dtype=float)
            else:
                unit = Unit(unit)
                arr = np.array(_plain(value), dtype=float) * value.unit.to(unit)
        else:
            if isinstance(value, (list, tuple)) and len(value) and all(isinstance(v, Quantity) for v in value):
                first = value[0].unit
                value = [_plain(v) * v.unit.to(first) for v in value]
                unit = first if unit is None else unit
            arr = np.array(value, dtype=float)
            unit = dimensionless_unscaled if unit is None else Unit(unit)
        obj = arr.view(cls)
        obj._unit = unit
        return obj

    def __array_finalize__(self, obj):
        self._unit = getattr(obj, '_unit', dimensionless_unscaled)

    @property
    def unit(self):
        return self._unit

    @property
    def value(self):
        arr = self.view(np.ndarray)
        return arr[()] if arr.ndim == 0 else arr

    def to(self, unit, equivalencies=None):
        unit = Unit(unit)
        return Quantity(_plain(self) * self.unit.to(unit), unit)

    def to_value(self, unit=None):
        return self.value if unit is None else self.to(unit).value

    @property
    def si(self):
        return self

    def decompose(self):
        return Quantity(_plain(self) * self.unit.scale, UnitBase(1.0, self.unit.powers))

    def __float__(self):
        if not self.unit.is_dimensionless:
            raise TypeError('only dimensionless scalar quantities can be converted to Python scalars')
        return float(_plain(self) * self.unit.scale)

    def __repr__(self):
        return '<Quantity {0} {1}>'.format(_plain(self), self.unit)

    __str__ = __repr__

    def __reduce__(self):
        return (Quantity, (np.array(_plain(self)), self.unit))

    def __deepcopy__(self, memo):
        return Quantity(np.array(_plain(self)), self.unit)

    def copy(self, order='C'):
        return Quantity(np.array(_plain(self)), self.unit)

    # -- unit operands -------------------------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, UnitBase):
            return Quantity(np.array(_plain(self)), self.unit * other)
        return np.multiply(self, other)

    def __rmul__(self, other):
        if isinstance(other, UnitBase):
            return Quantity(np.array(_plain(self)), other * self.unit)
        return np.multiply(other, self)

    def __imul__(self, other):
        if isinstance(other, UnitBase):
            self._unit = self.unit * other
            return self
        res = np.multiply(self, other)
        return res

    def __truediv__(self, other):
        if isinstance(other, UnitBase):
            return Quantity(np.array(_plain(self)), self.unit / other)
        return np.true_divide(self, other)

    def __itruediv__(self, other):
        if isinstance(other, UnitBase):
            self._unit = self.unit / other
            return self
        return np.true_divide(self, other)

    def __iadd__(self, other):
        return np.add(self, other)

    def __isub__(self, other):
        return np.subtract(self, other)

    # -- ufuncs --------------------------------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        kwargs.pop('out', None)
        if method == 'reduce':
            (arr,) = inputs
            unit = _unit_of(arr)
            res = getattr(ufunc, method)(_plain(arr), **kwargs)
            if ufunc in (np.add, np.maximum, np.minimum, np.fmax, np.fmin):
                return Quantity(res, unit)
            if ufunc is np.multiply:
                raise NotImplementedError
            return res
        if method == 'outer':
            a, b = inputs
            if ufunc in (np.add, np.subtract):
                unit = _unit_of(a) if isinstance(a, Quantity) else _unit_of(b)
                return Quantity(ufunc.outer(_as_unit_value(a, unit), _as_unit_value(b, unit), **kwargs), unit)
            if ufunc is np.multiply:
                return Quantity(ufunc.outer(_plain(a), _plain(b), **kwargs), _unit_of(a) * _unit_of(b))
            raise NotImplementedError
        if method != '__call__':
            return NotImplemented
        if ufunc in _SAME_UNIT or ufunc in _COMPARE or ufunc is np.arctan2:
            a, b = inputs
            unit = _unit_of(a) if isinstance(a, Quantity) else _unit_of(b)
            if not isinstance(a, Quantity) or not isinstance(b, Quantity):
                # quantity <op> bare number: only dimensionless quantities (converted to unscaled),
                # or the special values 0 / inf / nan
                q = a if isinstance(a, Quantity) else b
                if q.unit.is_dimensionless:
                    unit = dimensionless_unscaled
            res = ufunc(_as_unit_value(a, unit), _as_unit_value(b, unit), **kwargs)
            if ufunc in _COMPARE:
                return res
            if ufunc is np.arctan2:
                return Quantity(res, rad)
            return Quantity(res, unit)
        if ufunc is np.multiply:
            a, b = inputs
            return Quantity(ufunc(_plain(a), _plain(b), **kwargs), _unit_of(a) * _unit_of(b))
        if ufunc in (np.true_divide, np.divide):
            a, b = inputs
            return Quantity(ufunc(_plain(a), _plain(b), **kwargs), _unit_of(a) / _unit_of(b))
        if ufunc in (np.power, np.float_power):
            a, p = inputs
            if isinstance(p, Quantity):
                p = float(p)
            if not np.isscalar(p) and np.ndim(p) != 0:
                raise UnitsError('can only raise a quantity to a scalar power')
            return Quantity(ufunc(_plain(a), p, **kwargs), _unit_of(a) ** float(p))
        if ufunc is np.sqrt:
            return Quantity(ufunc(_plain(inputs[0]), **kwargs), inputs[0].unit ** 0.5)
        if ufunc is np.square:
            return Quantity(ufunc(_plain(inputs[0]), **kwargs), inputs[0].unit ** 2)
        if ufunc is np.reciprocal:
            return Quantity(ufunc(_plain(inputs[0]), **kwargs), inputs[0].unit ** -1)
        if ufunc in _KEEP:
            return Quantity(ufunc(_plain(inputs[0]), **kwargs), inputs[0].unit)
        if ufunc in _PLAIN_OUT:
            return ufunc(_plain(inputs[0]), **kwargs)
        if ufunc in _TRIG:
            (a,) = inputs
            if a.unit.powers != rad.powers:
                raise UnitTypeError("Can only apply '{0}' function to quantities with angle units".format(ufunc.__name__))
            return Quantity(ufunc(_plain(a) * a.unit.to(rad), **kwargs), dimensionless_unscaled)
        if ufunc in _INV_TRIG:
            (a,) = inputs
            return Quantity(ufunc(_as_unit_value(a, dimensionless_unscaled), **kwargs), rad)
        if ufunc in _DIMLESS:
            (a,) = inputs
            if not a.unit.is_dimensionless:
                raise UnitTypeError("Can only apply '{0}' function to dimensionless quantities".format(ufunc.__name__))
            return Quantity(ufunc(_plain(a) * a.unit.scale, **kwargs), dimensionless_unscaled)
        raise NotImplementedError('ufunc {0} is not supported by the astropy stand-in'.format(ufunc.__name__))
