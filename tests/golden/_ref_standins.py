"""Stand-ins that let the REFERENCE package import and run in the build container, where astropy / matplotlib /
multiprocess are not installed (used only by the golden-vector generators in this directory, never at test time).

What is replaced, and by what:
* `astropy.units`        -> `Quantity` below: angular / time unit conversion and the five operators `ang2pipi` uses.
* `astropy.io.fits`      -> the product's pure-Python FITS reader (`_compat/fits_lite.py`) + a dict-equality `HeaderDiff`.
* `astropy.wcs.WCS`      -> the oracle's wcslib-structured restatements (`oracle/wcs_tan.py`, `oracle/wcs_car.py`). This is
                            the one arithmetic boundary that stays UNPINNED: the goldens made here check everything the
                            reference does AROUND the WCS calls, not wcslib itself.
* `astropy.time.Time`    -> ISO string -> datetime; differences in days (`.value`) or seconds (`.to("s").value`).
* `astropy.constants`    -> R_sun = 695 700 000 m (IAU 2015 nominal, astropy's value).
* `multiprocess.shared_memory` -> the standard library's `multiprocessing.shared_memory` (same API).
* matplotlib, astropy.visualization, astropy.coordinates, ... -> inert `MagicMock` modules (import-only).
"""
import datetime
import sys
import types
from unittest import mock

import numpy as np

_TO_ARCSEC = {"arcsec": 1.0, "deg": 3600.0, "arcmin": 60.0, "rad": 3600.0 * 180.0 / np.pi}


def _factor(src, dst):
    src, dst = str(src), str(dst)
    if src == dst:
        return 1.0
    if src in ("s", "m") or dst in ("s", "m"):
        raise ValueError(f"no conversion {src} -> {dst}")
    table = {("arcsec", "deg"): 1.0 / 3600.0, ("deg", "arcsec"): 3600.0, ("arcmin", "deg"): 1.0 / 60.0,
             ("deg", "arcmin"): 60.0, ("arcsec", "arcmin"): 1.0 / 60.0, ("arcmin", "arcsec"): 60.0}
    return table.get((src, dst), _TO_ARCSEC[src] / _TO_ARCSEC[dst])


class Quantity:
    __array_priority__ = 1000

    def __init__(self, value, unit=None):
        if isinstance(value, Quantity):
            unit0 = value.unit
            value = value.value if unit is None or str(unit) == unit0 else value.value * _factor(unit0, unit)
            unit = unit0 if unit is None else unit
        self.value = np.asarray(value, dtype=np.float64) if not np.isscalar(value) else value
        self.unit = str(unit)

    @property
    def shape(self):
        return np.shape(self.value)

    def to(self, unit):
        f = _factor(self.unit, unit)
        return Quantity(self.value if f == 1.0 else self.value * f, str(unit))

    def _other(self, o):
        return o.to(self.unit).value if isinstance(o, Quantity) else o

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __add__(self, o):
        return Quantity(self.value + self._other(o), self.unit)

    def __sub__(self, o):
        return Quantity(self.value - self._other(o), self.unit)

    def __mod__(self, o):
        return Quantity(self.value % self._other(o), self.unit)

    def __mul__(self, o):
        return Quantity(self.value * o, self.unit)

    __rmul__ = __mul__

    def __getitem__(self, i):
        return Quantity(self.value[i], self.unit)

    def __gt__(self, o):
        return self.value > self._other(o)

    def __lt__(self, o):
        return self.value < self._other(o)

    def __ge__(self, o):
        return self.value >= self._other(o)

    def __le__(self, o):
        return self.value <= self._other(o)


class Unit(str):
    __array_ufunc__ = None      # ndarray * u.deg -> Unit.__rmul__

    def __rmul__(self, value):
        return Quantity(value, str(self))


class _TimeDelta:
    def __init__(self, seconds):
        self.seconds = seconds
        self.value = seconds / 86400.0      # astropy TimeDelta.value: days

    def to(self, unit):
        assert str(unit) == "s"
        return Quantity(self.seconds, "s")


class Time:
    def __init__(self, iso):
        self.t = iso.t if isinstance(iso, Time) else datetime.datetime.fromisoformat(str(iso))

    def __sub__(self, other):
        return _TimeDelta((self.t - other.t).total_seconds())


class FITSFixedWarning(Warning):
    pass


class SkyCoord:
    pass


class _WcsPrm:
    def __init__(self, ctype):
        self.ctype = ctype


class WCS:
    """`astropy.wcs.WCS(header)` for 2-axis -TAN / -CAR headers, arithmetic from the oracle restatements."""

    def __init__(self, hdr):
        from oracle import wcs_car, wcs_tan
        h = dict(hdr.items()) if hasattr(hdr, "items") else dict(hdr)
        impl = wcs_car.WcsCar if str(h["CTYPE1"]).endswith("CAR") else wcs_tan.WcsTan
        self._w = impl(h)
        self.wcs = _WcsPrm([str(h["CTYPE1"]), str(h["CTYPE2"])])
        self.pixel_shape = self._w.pixel_shape

    def pixel_to_world(self, x, y):
        lng, lat = self._w.pixel_to_world(x, y)
        return [Quantity(lng, "deg"), Quantity(lat, "deg")]

    def world_to_pixel(self, lng, lat):
        lng = lng.to("deg").value if isinstance(lng, Quantity) else lng
        lat = lat.to("deg").value if isinstance(lat, Quantity) else lat
        return self._w.world_to_pixel(lng, lat)


class _HeaderDiff:
    def __init__(self, a, b):
        self.identical = dict(a.items()) == dict(b.items())


def _plain_function():
    pass


_plain_function.__module__ = "astropy.wcs.utils"


def install(repo_root):
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    from euispice_coreg_b200._compat import fits_lite
    import multiprocessing.shared_memory as std_shm
    units = types.ModuleType("astropy.units")
    units.Quantity = Quantity
    units.__getattr__ = lambda name: Unit(name)
    time_mod = types.ModuleType("astropy.time")
    time_mod.Time = Time
    time_mod.TimeDelta = _TimeDelta
    fits = types.ModuleType("astropy.io.fits")
    fits.open = fits_lite.open
    fits.HeaderDiff = _HeaderDiff
    fits.PrimaryHDU, fits.HDUList, fits.Header = fits_lite.PrimaryHDU, fits_lite.HDUList, fits_lite.Header
    wcs = types.ModuleType("astropy.wcs")
    wcs.WCS, wcs.FITSFixedWarning = WCS, FITSFixedWarning
    wcs_utils = types.ModuleType("astropy.wcs.utils")
    wcs_utils.WCS_FRAME_MAPPINGS = [[_plain_function]]
    wcs_utils.FRAME_WCS_MAPPINGS = [[_plain_function]]
    wcs.utils = wcs_utils
    coords = types.ModuleType("astropy.coordinates")
    coords.SkyCoord = SkyCoord
    consts = types.ModuleType("astropy.constants")
    consts.R_sun = Quantity(695700000.0, "m")
    shm = types.ModuleType("multiprocess.shared_memory")
    shm.SharedMemory = std_shm.SharedMemory
    inert = ["astropy", "astropy.io", "astropy.io.ascii", "astropy.visualization", "matplotlib", "matplotlib.pyplot",
             "matplotlib.collections", "matplotlib.gridspec", "matplotlib.patches", "matplotlib.colors",
             "matplotlib.backends", "matplotlib.backends.backend_pdf", "mpl_toolkits", "mpl_toolkits.axes_grid1",
             "multiprocess"]
    for n in inert:
        sys.modules.setdefault(n, mock.MagicMock(name=n))
    real = {"astropy.units": units, "astropy.time": time_mod, "astropy.io.fits": fits, "astropy.wcs": wcs,
            "astropy.wcs.utils": wcs_utils, "astropy.coordinates": coords, "astropy.constants": consts,
            "multiprocess.shared_memory": shm}
    sys.modules.update(real)
    for name in sorted(list(real) + inert):       # `import a.b.c as m` walks attributes: link every child to its parent
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], leaf, sys.modules[name])
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
