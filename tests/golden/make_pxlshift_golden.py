"""Generate tests/golden/pxlshift_golden.npz by running the REFERENCE's own `pxlshift.AlignmentPixels`
(`/root/reference/euispice_coreg/pxlshift/alignment_pixels.py`) on seeded inputs.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_pxlshift_golden.py
The reference package imports astropy / matplotlib / multiprocess at module level; none of them is installed here.
They are replaced by inert stand-ins for the duration of this script. The only stand-in code that takes part in the
arithmetic is `astropy.units.Quantity(v, unit).to(unit2).value` (angular unit conversion) and, for the
solar-rotation case, `astropy.time.Time` differences in seconds -- both written below. Everything else that produces
the numbers is the reference's code as it lies under /root/reference: `find_best_parameters`, `_sub_resolution_large_fov`,
`_shift_large_fov`, `_step`, `matrix_transform.polar_transform`, `rectify.interpol2d` (scipy), `Util.diff_rot`, and
the numba `pxlshift/c_correlate.py` with its float32 numerator.
"""
import datetime
import importlib
import os
import sys
import types
import warnings
from unittest import mock

import numpy as np

REF_ROOT = "/root/reference"
_TO_ARCSEC = {"arcsec": 1.0, "deg": 3600.0, "arcmin": 60.0, "rad": 3600.0 * 180.0 / np.pi, "s": 1.0}


class Quantity:
    def __init__(self, value, unit):
        self.value, self.unit = value, str(unit)

    def to(self, unit):
        unit = str(unit)
        if unit == self.unit:
            return Quantity(self.value, unit)
        if (self.unit, unit) == ("arcsec", "deg"):
            return Quantity(self.value * (1.0 / 3600.0), unit)      # astropy's arcsec -> deg factor
        return Quantity(self.value * (_TO_ARCSEC[self.unit] / _TO_ARCSEC[unit]), unit)


class Unit(str):
    def __rmul__(self, value):
        return Quantity(value, str(self))


class Time:
    def __init__(self, iso):
        self.t = datetime.datetime.fromisoformat(str(iso))

    def __sub__(self, other):
        return Quantity((self.t - other.t).total_seconds(), "s")


def _install_stubs():
    units = types.ModuleType("astropy.units")
    units.Quantity = Quantity
    units.__getattr__ = lambda name: Unit(name)      # u.s, u.arcsec, u.deg ... (default arguments of the reference)
    time_mod = types.ModuleType("astropy.time")
    time_mod.Time = Time
    time_mod.TimeDelta = mock.MagicMock()
    names = ["astropy", "astropy.io", "astropy.io.fits", "astropy.io.ascii", "astropy.constants", "astropy.wcs",
             "astropy.wcs.utils", "astropy.visualization", "astropy.coordinates", "matplotlib", "matplotlib.pyplot",
             "matplotlib.collections", "matplotlib.gridspec", "matplotlib.patches", "matplotlib.colors",
             "matplotlib.backends", "matplotlib.backends.backend_pdf", "mpl_toolkits", "mpl_toolkits.axes_grid1",
             "multiprocess", "multiprocess.shared_memory"]
    for n in names:
        sys.modules.setdefault(n, mock.MagicMock(name=n))
    sys.modules["astropy.units"] = units
    sys.modules["astropy.time"] = time_mod
    sys.modules["astropy"].units = units
    sys.modules["astropy"].time = time_mod


def load_reference_class():
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("euispice_coreg.pxlshift.alignment_pixels").AlignmentPixels


def make_inputs(case):
    """Seeded small / large images and headers. case 0: plain; case 1: NaN holes in both images, rectangular pixels;
    case 2: solar-rotation shift of the large image (CROTA present)."""
    rng = np.random.default_rng(4200 + case)
    from scipy.ndimage import gaussian_filter
    H, W = (64, 72) if case != 1 else (58, 66)
    sky = np.exp(0.6 * gaussian_filter(rng.standard_normal((H * 4, W * 4)), 3.0) * 6.0) * 300.0
    large = sky[::4, ::4] + rng.normal(0, 2.0, (H, W))
    ratio = (2.5, 2.5) if case != 1 else (2.0, 3.2)          # large pixel / small pixel along x, y
    hdr_large = {"CDELT1": 10.0, "CDELT2": 10.0 if case != 1 else 16.0, "CUNIT1": "arcsec", "CUNIT2": "arcsec",
                 "WAVELNTH": 174, "SOLAR_B0": 3.1, "RSUN_REF": 6.957e8, "DSUN_OBS": 7.1e10,
                 "DATE-AVG": "2022-03-17T09:50:45.000"}
    hdr_small = {"CDELT1": hdr_large["CDELT1"] / ratio[0] / 3600.0, "CDELT2": hdr_large["CDELT2"] / ratio[1] / 3600.0,
                 "CUNIT1": "deg", "CUNIT2": "deg", "DATE-AVG": "2022-03-17T10:20:45.000"}
    if case == 2:
        hdr_large["CROTA"] = 4.0
    sh, sw = (44, 60) if case != 1 else (37, 51)
    from scipy.ndimage import map_coordinates
    # the small image: the same sky at the small resolution, offset by (+3, -2) small pixels from the centred slice
    y0 = (H * ratio[1] - sh - 1) / 2 - 2 + 0.3
    x0 = (W * ratio[0] - sw - 1) / 2 + 3 - 0.2
    yy, xx = np.meshgrid(np.arange(sh), np.arange(sw), indexing="ij")
    small = map_coordinates(sky, [(yy + y0) * 4 / ratio[1], (xx + x0) * 4 / ratio[0]], order=1, mode="nearest")
    small = small * 0.8 + rng.normal(0, 3.0, small.shape) + 20.0
    if case == 1:
        small[5:9, 10:14] = np.nan
        small[30, :] = np.nan
        large[20:23, 30:33] = np.nan
    return large.astype(np.float64), hdr_large, small.astype(np.float64), hdr_small


LAGS = {0: (np.arange(-5, 8), np.arange(-6, 5), np.array([0.0])),
        1: (np.arange(-4, 5, 2), np.arange(-3, 4), np.array([-2.0, 0.0, 1.5])),
        2: (np.arange(-3, 6), np.arange(-5, 3), np.array([0.0, 0.7]))}


def main():
    warnings.simplefilter("ignore")
    cls = load_reference_class()
    out = {}
    for case in (0, 1, 2):
        large, hl, small, hs = make_inputs(case)
        obj = cls.__new__(cls)            # __init__ only reads the two FITS files
        obj.hdr_large, obj.data_large = dict(hl), large.copy()
        obj.hdr_small, obj.data_small = dict(hs), small.copy()
        obj.slc_small_ref = obj.x_large = obj.y_large = None
        dx, dy, drot = LAGS[case]
        corr = obj.find_best_parameters(dx, dy, drot, unit_rot="degree", shift_solar_rotation_dx_large=(case == 2))
        out[f"corr{case}"] = np.asarray(corr, dtype=np.float64)
        out[f"large_sub{case}"] = np.asarray(obj.data_large, dtype=np.float64)      # after _sub_resolution_large_fov
        out[f"rot_last{case}"] = np.asarray(obj.data_small_rotated, dtype=np.float64)
        i = np.unravel_index(np.nanargmax(corr), corr.shape)
        print("case", case, corr.shape, "max", np.nanmax(corr), "at dx, dy, drot =", dx[i[0]], dy[i[1]], drot[i[2]])
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pxlshift_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
