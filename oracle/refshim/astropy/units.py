"""Minimal dimensional-analysis stand-in for astropy.units (see ../README.md)."""
import re
import numpy as np

_DIMS = 5  # length, mass, time, temperature, angle/other


class Unit:
    def __init__(self, scale, dims, name=None):
        self.scale = float(scale)
        self.dims = tuple(dims)
        self.name = name

    # algebra -------------------------------------------------------------------------------
    def _coerce(self, other):
        if isinstance(other, Unit):
            return other
        if isinstance(other, Quantity) and np.ndim(other) == 0:
            return Unit(float(np.asarray(other)) * other.unit.scale, other.unit.dims)
        return None

    def __mul__(self, other):
        o = self._coerce(other)
        if o is not None:
            return Unit(self.scale * o.scale, [a + b for a, b in zip(self.dims, o.dims)])
        return Quantity(other, self)

    def __rmul__(self, other):
        return Quantity(other, self)

    def __truediv__(self, other):
        o = self._coerce(other)
        if o is not None:
            return Unit(self.scale / o.scale, [a - b for a, b in zip(self.dims, o.dims)])
        return Quantity(1. / np.asarray(other, float), self)

    def __rtruediv__(self, other):
        return Quantity(other, self ** -1)

    def __pow__(self, p):
        return Unit(self.scale ** p, [a * p for a in self.dims])

    def to(self, other, value=1.):
        other = as_unit(other)
        if tuple(np.round(self.dims, 9)) != tuple(np.round(other.dims, 9)):
            raise ValueError('incompatible units %s -> %s' % (self.dims, other.dims))
        return value * self.scale / other.scale

    def __repr__(self):
        return 'Unit(%r, %r)' % (self.scale, self.dims)

    def __format__(self, spec):
        return self.name or repr(self)


def _u(scale, L=0, M=0, T=0, K=0, A=0, name=None):
    return Unit(scale, (L, M, T, K, A), name)


dimensionless_unscaled = _u(1., name='')
m = _u(1., L=1, name='m')
cm = _u(1e-2, L=1, name='cm')
nm = _u(1e-9, L=1, name='nm')
angstrom = AA = _u(1e-10, L=1, name='angstrom')
au = _u(1.495978707e11, L=1, name='au')
pc = _u(1.495978707e11 * 648000. / np.pi, L=1, name='pc')
Mpc = _u(1e6 * pc.scale, L=1, name='Mpc')
Rsun = _u(6.957e8, L=1, name='Rsun')
kg = _u(1., M=1, name='kg')
g = _u(1e-3, M=1, name='g')
Msun = _u(1.988409870698051e30, M=1, name='Msun')
s = _u(1., T=1, name='s')
d = day = _u(86400., T=1, name='d')
Hz = _u(1., T=-1, name='Hz')
THz = _u(1e12, T=-1, name='THz')
K = _u(1., K=1, name='K')
kK = _u(1e3, K=1, name='kK')
J = _u(1., L=2, M=1, T=-2, name='J')
erg = _u(1e-7, L=2, M=1, T=-2, name='erg')
eV = _u(1.602176634e-19, L=2, M=1, T=-2, name='eV')
W = _u(1., L=2, M=1, T=-3, name='W')
mag = _u(1., A=1, name='mag')
deg = _u(np.pi / 180., name='deg')
_NAMES = {k: v for k, v in list(globals().items()) if isinstance(v, Unit)}


def def_unit(name, represents=None, format=None, **kw):
    base = as_unit(represents)
    return Unit(base.scale, base.dims, name)


def as_unit(x):
    if isinstance(x, Unit):
        return x
    if isinstance(x, Quantity):
        return Unit(float(np.asarray(x)) * x.unit.scale, x.unit.dims)
    if isinstance(x, str):
        return _parse(x)
    raise TypeError('not a unit: %r' % (x,))


def _parse(text):
    """'eV / kK', 'erg s-1 Rsun-2 kK-4'"""
    out = dimensionless_unscaled
    sign = 1
    for tok in text.replace('/', ' / ').split():
        if tok == '/':
            sign = -1
            continue
        mt = re.match(r'^([A-Za-z]+)(-?\d+)?$', tok)
        base = _NAMES[mt.group(1)]
        power = int(mt.group(2)) if mt.group(2) else 1
        out = out * base ** (sign * power)
        sign = 1 if sign == -1 and False else sign  # 'a / b c' keeps dividing only b in astropy's generic format
        if sign == -1:
            sign = 1
    return out


class Quantity(np.ndarray):
    """ndarray carrying a unit; * / ** propagate units exactly, everything else keeps the left operand's."""
    __array_priority__ = 1000

    def __new__(cls, value, unit=None):
        if isinstance(value, Quantity) and unit is None:
            unit = value.unit
        obj = np.asarray(value, dtype=float).view(cls)
        obj.unit = dimensionless_unscaled if unit is None else as_unit(unit)
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, 'unit', dimensionless_unscaled)

    def __array_wrap__(self, obj, context=None, return_scalar=False):
        out = np.asarray(obj).view(type(self))
        out.unit = self.unit
        return out

    @property
    def value(self):
        a = np.asarray(self)
        return float(a) if a.ndim == 0 else a

    @property
    def data(self):
        return np.asarray(self)

    @property
    def quantity(self):
        return Quantity(np.asarray(self), self.unit)

    def to(self, unit):
        unit = as_unit(unit)
        return Quantity(np.asarray(self) * self.unit.to(unit), unit)

    def _other(self, other):
        if isinstance(other, Unit):
            return 1., other
        if isinstance(other, Quantity):
            return np.asarray(other), other.unit
        return np.asarray(other), dimensionless_unscaled

    def __mul__(self, other):
        v, uo = self._other(other)
        return type(self)(np.asarray(self) * v, self.unit * uo)

    __rmul__ = __mul__

    def __truediv__(self, other):
        v, uo = self._other(other)
        return type(self)(np.asarray(self) / v, self.unit / uo)

    def __rtruediv__(self, other):
        v, uo = self._other(other)
        return type(self)(v / np.asarray(self), uo / self.unit)

    def __itruediv__(self, other):
        v, uo = self._other(other)
        np.asarray(self).__itruediv__(v)
        self.unit = self.unit / uo
        return self

    def __pow__(self, p):
        return type(self)(np.asarray(self) ** p, self.unit ** p)

    def __getitem__(self, key):
        out = np.ndarray.__getitem__(self, key)
        if isinstance(out, np.ndarray):
            out = out.view(type(self))
            out.unit = self.unit
            return out
        return out
