"""Angle-unit helpers replacing the `astropy.units` calls on the pointing-search path.

Reference call sites: `hdrshift/alignment.py:819-837` (lag unit -> header unit),
`utils/Util.py:174-208` (arcsec -> CUNIT), `hdrshift/AlignmentResults.py:69-82`.
"""
import math

import numpy as np

# scale of each unit in radians, composed the way astropy defines them
# (degree = pi/180 rad, arcmin = degree/60, arcsec = arcmin/60).
_RAD = {
    "rad": 1.0,
    "deg": math.pi / 180.0,
    "arcmin": math.pi / 180.0 / 60.0,
    "arcsec": math.pi / 180.0 / 3600.0,
    "mas": math.pi / 180.0 / 3600.0e3,
}
_ALIASES = {
    "degree": "deg", "degrees": "deg", "radian": "rad", "radians": "rad",
    "arcseconds": "arcsec", "arcsecond": "arcsec", "arcminute": "arcmin", "arcminutes": "arcmin",
    "asec": "arcsec", "amin": "arcmin",
}


def canon(unit) -> str:
    """Canonical unit name; accepts str or any object whose str() is a unit name."""
    s = str(unit).strip()
    s = _ALIASES.get(s.lower(), s)
    if s not in _RAD:
        raise ValueError(f"unsupported angular unit {unit!r}")
    return s


def factor(src, dst) -> float:
    """Multiplicative factor converting `src` values to `dst` values."""
    a, b = canon(src), canon(dst)
    if a == b:
        return 1.0
    # exact small-integer ratios where they exist (arcsec<->deg etc.)
    table = {("arcsec", "deg"): 1.0 / 3600.0, ("deg", "arcsec"): 3600.0,
             ("arcmin", "deg"): 1.0 / 60.0, ("deg", "arcmin"): 60.0,
             ("arcsec", "arcmin"): 1.0 / 60.0, ("arcmin", "arcsec"): 60.0}
    if (a, b) in table:
        return table[(a, b)]
    return _RAD[a] / _RAD[b]


def convert(value, src, dst):
    """`u.Quantity(value, src).to(dst).value`"""
    f = factor(src, dst)
    if f == 1.0:
        return np.asarray(value, dtype=np.float64) if not np.isscalar(value) else float(value)
    if np.isscalar(value):
        return float(value) * f
    return np.asarray(value, dtype=np.float64) * f


def strip(value, default_unit="arcsec"):
    """Accept plain numbers/arrays or astropy-like quantities; return (ndarray|float, unit)."""
    if hasattr(value, "unit") and hasattr(value, "value"):
        return value.value, canon(value.unit)
    return value, canon(default_unit)


def ang2pipi_deg(ang):
    """`AlignCommonUtil.ang2pipi` for values in degrees (`utils/Util.py:76-80`):
    ``-((-ang + 180) % 360 - 180)`` -> (-180, 180]."""
    ang = np.asarray(ang, dtype=np.float64)
    return -((-ang + 180.0) % 360.0 - 180.0)


def ang2pipi(ang, unit):
    """ang2pipi on a value expressed in `unit` (the reference does the arithmetic in the
    quantity's own unit: pi = Quantity(180,'deg') is converted to that unit first)."""
    f = factor("deg", unit)
    pi = 180.0 * f
    ang = np.asarray(ang, dtype=np.float64)
    return -((-ang + pi) % (2 * pi) - pi)
