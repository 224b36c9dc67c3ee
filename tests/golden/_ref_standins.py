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


_EPOCH = datetime.datetime(2000, 1, 1, 12, 0, 0)


def _iso_to_seconds(iso):
    d = datetime.datetime.fromisoformat(str(iso).replace("Z", ""))
    whole = d.replace(microsecond=0) - _EPOCH
    return whole.days * 86400.0 + whole.seconds + d.microsecond * 1e-6


class Time:
    """Scalar or array of UTC instants held as float seconds since J2000 (what the product and the oracle use; astropy
    keeps two doubles, the 1e-7 s difference only shows in which imager frame is 'closest' at exact ties)."""

    def __init__(self, value):
        if isinstance(value, Time):
            self.s = value.s
        elif isinstance(value, (float, np.floating, np.ndarray)):
            self.s = value
        else:
            self.s = _iso_to_seconds(value)

    def __sub__(self, other):
        if isinstance(other, Time):
            return _TimeDelta(self.s - other.s)
        if isinstance(other, Quantity):
            assert other.unit == "s"
            return Time(self.s - other.value)
        return NotImplemented

    def __getitem__(self, i):
        return Time(self.s[i])

    def __len__(self):
        return len(self.s)

    def __iter__(self):
        return (Time(v) for v in self.s)

    @property
    def shape(self):
        return np.shape(self.s)

    @property
    def fits(self):
        whole = int(np.floor(self.s))
        ms = int(round((self.s - whole) * 1000.0))
        if ms == 1000:
            whole, ms = whole + 1, 0
        return (_EPOCH + datetime.timedelta(seconds=whole)).strftime("%Y-%m-%dT%H:%M:%S") + ".%03d" % ms


class FITSFixedWarning(Warning):
    pass


class SkyCoord:
    pass


class _WcsPrm:
    def __init__(self, ctype, pc=None, crpix=None, cdelt=None, crval=None, cunit=None):
        self.ctype, self.pc, self.crpix, self.cdelt, self.crval, self.cunit = ctype, pc, crpix, cdelt, crval, cunit


_TIME_SCALE = {"s": 1.0, "min": 60.0, "h": 3600.0, "d": 86400.0}


class WCS:
    """`astropy.wcs.WCS(header)`: a celestial -TAN / -CAR pair (arithmetic from the oracle restatements) plus linear
    spectral / time axes, with the handful of methods the reference calls: `dropaxis`, `deepcopy`, `sub(['spectral'])`,
    `pixel_to_world`, `world_to_pixel`, `to_header`, `.wcs.ctype`, `.wcs.pc`, `.pixel_shape`."""

    def __init__(self, hdr):
        import copy
        h = dict(hdr.items()) if hasattr(hdr, "items") else dict(hdr)
        n = int(h.get("WCSAXES", h.get("NAXIS", 2)))
        rng = range(1, n + 1)
        pc = np.eye(n)
        for i in rng:
            for j in rng:
                if f"PC{i}_{j}" in h:
                    pc[i - 1, j - 1] = float(h[f"PC{i}_{j}"])
        self.wcs = _WcsPrm([str(h.get(f"CTYPE{i}", "")) for i in rng], pc,
                           [float(h.get(f"CRPIX{i}", 0.0)) for i in rng], [float(h.get(f"CDELT{i}", 1.0)) for i in rng],
                           [float(h.get(f"CRVAL{i}", 0.0)) for i in rng], [str(h.get(f"CUNIT{i}", "")).strip() for i in rng])
        shape = [h.get(f"ZNAXIS{i}", h.get(f"NAXIS{i}")) for i in rng]
        self._shape = [None if v is None else int(v) for v in shape]
        self._has_pc = any(f"PC{i}_{j}" in h for i in rng for j in rng)
        self._rota = h.get("CROTA2", h.get("CROTA")) if not self._has_pc else None
        self._extra = {k: copy.copy(h[k]) for k in ("LONPOLE", "LATPOLE", "DATEREF", "DATE-BEG", "DATE-OBS", "DATE-AVG",
                                                    "DATE-END", "RSUN_REF", "DSUN_OBS") if k in h}

    @property
    def naxis(self):
        return len(self.wcs.ctype)

    @property
    def pixel_shape(self):
        return None if any(v is None for v in self._shape) else tuple(self._shape)

    def deepcopy(self):
        import copy
        return copy.deepcopy(self)

    def _keep(self, keep):
        w = self.deepcopy()
        w.wcs = _WcsPrm([self.wcs.ctype[i] for i in keep], self.wcs.pc[np.ix_(keep, keep)].copy(),
                        [self.wcs.crpix[i] for i in keep], [self.wcs.cdelt[i] for i in keep],
                        [self.wcs.crval[i] for i in keep], [self.wcs.cunit[i] for i in keep])
        w._shape = [self._shape[i] for i in keep]
        return w

    def dropaxis(self, axis):
        return self._keep([i for i in range(self.naxis) if i != axis])

    def sub(self, axes):
        assert list(axes) == ["spectral"]
        return self._keep([i for i, c in enumerate(self.wcs.ctype) if c[:4] in ("WAVE", "AWAV", "FREQ")])

    def _celestial(self):
        idx = [i for i, c in enumerate(self.wcs.ctype) if c[:4] in ("HPLN", "HPLT", "CRLN", "CRLT")]
        return idx if len(idx) == 2 else None

    def _celestial_header(self):
        a, b = self._celestial()
        w = self.wcs
        for i in (a, b):
            for j in range(self.naxis):
                assert j in (a, b) or w.pc[i, j] == 0.0, "celestial axis coupled to a non-celestial one"
        h = {"CTYPE1": w.ctype[a], "CTYPE2": w.ctype[b], "CUNIT1": w.cunit[a] or "deg", "CUNIT2": w.cunit[b] or "deg",
             "CRPIX1": w.crpix[a], "CRPIX2": w.crpix[b], "CDELT1": w.cdelt[a], "CDELT2": w.cdelt[b],
             "CRVAL1": w.crval[a], "CRVAL2": w.crval[b], "NAXIS1": self._shape[a] or 0, "NAXIS2": self._shape[b] or 0}
        if self._has_pc:
            h.update(PC1_1=w.pc[a, a], PC1_2=w.pc[a, b], PC2_1=w.pc[b, a], PC2_2=w.pc[b, b])
        elif self._rota is not None:
            h["CROTA2"] = self._rota
        for k in ("LONPOLE", "LATPOLE"):
            if k in self._extra:
                h[k] = self._extra[k]
        return h

    def _impl(self):
        from oracle import wcs_car, wcs_tan
        h = self._celestial_header()
        return (wcs_car.WcsCar if str(h["CTYPE1"]).endswith("CAR") else wcs_tan.WcsTan)(h)

    def pixel_to_world(self, *pix):
        pix = [np.asarray(p, dtype=np.float64) for p in pix]
        assert len(pix) == self.naxis
        cel = self._celestial()
        out = [None] * self.naxis
        if cel is not None:
            lng, lat = self._impl().pixel_to_world(pix[cel[0]], pix[cel[1]])
            out[cel[0]], out[cel[1]] = Quantity(lng, "deg"), Quantity(lat, "deg")
        w = self.wcs
        for k in range(self.naxis):
            if out[k] is not None:
                continue
            lin = sum(w.pc[k, j] * (pix[j] + 1.0 - w.crpix[j]) for j in range(self.naxis))
            world = w.crval[k] + w.cdelt[k] * lin
            if w.ctype[k] in ("UTC", "TIME", "TAI", "TT"):
                ref = self._extra.get("DATEREF", self._extra.get("DATE-BEG", self._extra.get("DATE-OBS")))
                out[k] = Time(_iso_to_seconds(ref) + world * _TIME_SCALE.get(w.cunit[k], 1.0))
            else:
                out[k] = Quantity(world, w.cunit[k] or "")
        return out if self.naxis > 1 else out[0]

    def world_to_pixel(self, lng, lat):
        assert self.naxis == 2
        lng = lng.to("deg").value if isinstance(lng, Quantity) else lng
        lat = lat.to("deg").value if isinstance(lat, Quantity) else lat
        return self._impl().world_to_pixel(lng, lat)

    def to_header(self):
        """Two celestial axes: degrees, PCi_j only where they differ from the identity (what wcslib's wcshdo writes)."""
        from euispice_coreg_b200._compat import fits_lite
        assert self.naxis == 2
        impl = self._impl()
        h = fits_lite.Header()
        h["WCSAXES"] = 2
        h["CRPIX1"], h["CRPIX2"] = impl.crpix
        for key, val, default in (("PC1_1", impl.pc[0][0], 1.0), ("PC1_2", impl.pc[0][1], 0.0),
                                  ("PC2_1", impl.pc[1][0], 0.0), ("PC2_2", impl.pc[1][1], 1.0)):
            if val != default:
                h[key] = float(val)
        h["CDELT1"], h["CDELT2"] = impl.cdelt
        h["CUNIT1"], h["CUNIT2"] = "deg", "deg"
        h["CTYPE1"], h["CTYPE2"] = self.wcs.ctype
        h["CRVAL1"], h["CRVAL2"] = impl.crval
        h["LONPOLE"] = float(impl.lonpole)
        h["LATPOLE"] = float(self._extra.get("LATPOLE", impl.crval[1]))
        for k in ("DATE-OBS", "DATE-BEG", "DATE-AVG", "DATE-END", "RSUN_REF", "DSUN_OBS"):
            if k in self._extra:
                h[k] = self._extra[k]
        return h


class _HeaderDiff:
    def __init__(self, a, b):
        self.identical = dict(a.items()) == dict(b.items())


def _plain_function():
    pass


_plain_function.__module__ = "astropy.wcs.utils"


def real_astropy():
    """True when the real astropy (with its bundled wcslib) is importable: the probe every generator and
    tests/test_wcs_pin.py run, so that the wcslib boundary gets pinned on the first box that has it."""
    import importlib.util
    try:
        return importlib.util.find_spec("astropy") is not None and importlib.util.find_spec("astropy.wcs") is not None
    except (ImportError, ValueError):
        return False


def install(repo_root):
    """Make `import euispice_coreg` work. Returns "astropy" when the real astropy answered (goldens generated in that
    process are pinned to wcslib / astropy.units / astropy.io.fits themselves) or "stand-ins" otherwise."""
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    if real_astropy():
        import importlib.util
        for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.gridspec",
                  "matplotlib.patches", "matplotlib.colors", "matplotlib.backends", "matplotlib.backends.backend_pdf",
                  "mpl_toolkits", "mpl_toolkits.axes_grid1", "multiprocess", "multiprocess.shared_memory"):
            top = n.split(".")[0]
            if importlib.util.find_spec(top) is None:
                sys.modules.setdefault(n, mock.MagicMock(name=n))
        if "multiprocess.shared_memory" in sys.modules and isinstance(sys.modules["multiprocess.shared_memory"],
                                                                       mock.MagicMock):
            import multiprocessing.shared_memory as std_shm
            shm = types.ModuleType("multiprocess.shared_memory")
            shm.SharedMemory = std_shm.SharedMemory
            sys.modules["multiprocess.shared_memory"] = shm
            sys.modules["multiprocess"].shared_memory = shm
        if "/root/reference" not in sys.path:
            sys.path.insert(0, "/root/reference")
        return "astropy"
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
    return "stand-ins"
