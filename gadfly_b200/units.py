"""
A very small physical-units shim so that the reference's call style
(``1 * u.min``, ``5777 * u.K``, ``t * u.d``, ``u.cds.ppm``) keeps working
without astropy, which is not available on the GPU box.

The reference strips units at its boundary with celerite2 and works in
time = 1/uHz (= 1e6 s), frequency = uHz, flux = ppm
(reference: gadfly/gp.py:61-126, gadfly/core.py:392).  This module only has to
perform those conversions; it is not a general unit system.

astropy Quantities (if a caller has astropy) are accepted everywhere a
Quantity is: see :func:`to_value`.
"""
import numpy as np

__all__ = [
    "Unit", "Quantity", "to_value", "s", "min", "h", "hour", "d", "day", "yr", "Hz", "uHz", "mHz",
    "K", "g", "kg", "m", "km", "nm", "um", "AA", "M_sun", "R_sun", "L_sun", "W", "ppm", "cds",
    "dimensionless_unscaled", "electron",
]

# base dimensions: time, mass, length, temperature, flux-fraction, electrons
_NDIM = 6


class Unit:
    """scale * prod(base_i ** dims_i)"""
    __array_priority__ = 1000

    def __init__(self, scale, dims, name=None):
        self.scale = float(scale)
        self.dims = tuple(dims)
        self.name = name

    # -- algebra -----------------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale * other.scale,
                        [a + b for a, b in zip(self.dims, other.dims)])
        if isinstance(other, Quantity):
            return Quantity(other.value, self * other.unit)
        return Quantity(other, self)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale / other.scale,
                        [a - b for a, b in zip(self.dims, other.dims)])
        if isinstance(other, Quantity):
            return Quantity(1.0 / np.asarray(other.value), self / other.unit)
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        inv = Unit(1.0 / self.scale, [-a for a in self.dims])
        if isinstance(other, Quantity):
            return Quantity(other.value, other.unit * inv)
        if np.isscalar(other) and other == 1:
            return inv
        return Quantity(other, inv)

    def __pow__(self, p):
        return Unit(self.scale ** p, [a * p for a in self.dims])

    def is_equivalent(self, other):
        return np.allclose(self.dims, other.dims)

    def __eq__(self, other):
        return isinstance(other, Unit) and self.is_equivalent(other) and \
            np.isclose(self.scale, other.scale, rtol=1e-14)

    def __hash__(self):
        return hash((round(self.scale, 12), self.dims))

    def __repr__(self):
        if self.name:
            return self.name
        return f"Unit({self.scale:g}, dims={self.dims})"


def _base(i, scale=1.0, name=None):
    dims = [0] * _NDIM
    dims[i] = 1
    return Unit(scale, dims, name)


class Quantity:
    """A value (scalar or ndarray) with a :class:`Unit`."""
    __array_priority__ = 2000

    def __init__(self, value, unit=None):
        if isinstance(value, Quantity):
            unit = unit if unit is not None else value.unit
            value = value.to(unit).value
        self.value = np.asarray(value, dtype=float) if not np.isscalar(value) else float(value)
        self.unit = unit if unit is not None else dimensionless_unscaled

    def to(self, unit):
        if not self.unit.is_equivalent(unit):
            raise ValueError(f"cannot convert {self.unit!r} to {unit!r}")
        return Quantity(np.multiply(self.value, self.unit.scale / unit.scale), unit)

    def to_value(self, unit=None):
        return self.value if unit is None else self.to(unit).value

    # numpy-ish helpers the reference uses on Quantities
    @property
    def shape(self):
        return np.shape(self.value)

    @property
    def ndim(self):
        return np.ndim(self.value)

    def __len__(self):
        return len(self.value)

    def __getitem__(self, item):
        return Quantity(self.value[item], self.unit)

    def mean(self, *a, **k):
        return Quantity(np.mean(self.value, *a, **k), self.unit)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def _coerce(self, other):
        if isinstance(other, Quantity):
            return other.to(self.unit).value
        if isinstance(other, Unit):
            raise TypeError("cannot add a bare unit")
        if all(abs(x) < 1e-15 for x in self.unit.dims):
            return np.asarray(other) / self.unit.scale
        raise TypeError("cannot combine a Quantity with a bare number")

    def __add__(self, other):
        return Quantity(self.value + self._coerce(other), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        return Quantity(self.value - self._coerce(other), self.unit)

    def __rsub__(self, other):
        return Quantity(self._coerce(other) - self.value, self.unit)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __mul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.value * other.value, self.unit * other.unit)
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit * other)
        return Quantity(self.value * other, self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.value / other.value, self.unit / other.unit)
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit / other)
        return Quantity(self.value / other, self.unit)

    def __rtruediv__(self, other):
        return Quantity(other / self.value, 1 / self.unit)

    def __pow__(self, p):
        return Quantity(self.value ** p, self.unit ** p)

    def __float__(self):
        if not all(abs(x) < 1e-15 for x in self.unit.dims):
            raise TypeError("only dimensionless quantities convert to float")
        return float(self.value * self.unit.scale)

    def _cmp(self, other, op):
        return op(self.value, self._coerce(other))

    def __lt__(self, o):
        return self._cmp(o, np.less)

    def __le__(self, o):
        return self._cmp(o, np.less_equal)

    def __gt__(self, o):
        return self._cmp(o, np.greater)

    def __ge__(self, o):
        return self._cmp(o, np.greater_equal)

    def __repr__(self):
        return f"<Quantity {self.value!r} {self.unit!r}>"


dimensionless_unscaled = Unit(1.0, [0] * _NDIM, "")

# time (base: second) and frequency
s = _base(0, 1.0, "s")
min = _base(0, 60.0, "min")  # noqa: A001  (mirrors astropy.units.min)
h = hour = _base(0, 3600.0, "h")
d = day = _base(0, 86400.0, "d")
yr = _base(0, 365.25 * 86400.0, "yr")
Hz = Unit(1.0, (1 / s).dims, "Hz")
mHz = Unit(1e-3, Hz.dims, "mHz")
uHz = Unit(1e-6, Hz.dims, "uHz")
# mass (base: kg), length (base: m), temperature (K)
kg = _base(1, 1.0, "kg")
g = _base(1, 1e-3, "g")
M_sun = _base(1, 1.988409870698051e30, "solMass")
m = _base(2, 1.0, "m")
km = _base(2, 1e3, "km")
um = _base(2, 1e-6, "um")
nm = _base(2, 1e-9, "nm")
AA = _base(2, 1e-10, "Angstrom")
R_sun = _base(2, 6.957e8, "solRad")
K = _base(3, 1.0, "K")
W = Unit(1.0, (kg * m ** 2 / s ** 3).dims, "W")
L_sun = Unit(3.828e26, W.dims, "solLum")
# relative flux
ppm = _base(4, 1.0, "ppm")
electron = _base(5, 1.0, "electron")


class _CDS:
    ppm = ppm


cds = _CDS()

_ASTROPY_NAMES = {
    "solMass": "solMass", "solRad": "solRad", "solLum": "solLum", "K": "K", "s": "s",
    "uHz": "uHz", "ppm": "cds.ppm", "um": "um", "nm": "nm",
}


def to_value(x, unit, assume=None):
    """Return ``x`` as plain float(s) in ``unit``.

    * our :class:`Quantity` -> converted;
    * an astropy Quantity/Time-like (has ``.unit`` and ``.to``) -> converted through astropy;
    * a bare number/array -> assumed to be already in ``assume`` (default ``unit``), the
      reference's convention for ndarray inputs (gadfly/gp.py:82-84,111-113).
    """
    if isinstance(x, Quantity):
        return x.to(unit).value
    if hasattr(x, "unit") and hasattr(x, "to"):  # astropy Quantity (duck-typed)
        import astropy.units as au  # only reachable if the caller has astropy
        from astropy.units import cds as _cds  # noqa: F401
        if unit.name == "1/uHz" or (unit.dims == s.dims and unit.scale == 1e6):
            return x.to(1 / au.uHz).value
        name = _ASTROPY_NAMES.get(unit.name, unit.name)
        if name == "cds.ppm":
            return x.to(_cds.ppm).value
        return x.to(au.Unit(name)).value
    if hasattr(x, "jd"):  # astropy Time-like
        return np.asarray(x.jd, dtype=float) * (86400.0 / unit.scale)
    arr = np.asarray(x, dtype=float) if not np.isscalar(x) else float(x)
    if assume is not None and assume is not unit:
        return arr * (assume.scale / unit.scale)
    return arr


inv_uHz = Unit(1e6, s.dims, "1/uHz")
