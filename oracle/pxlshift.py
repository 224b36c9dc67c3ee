"""Pixel-shift co-alignment, CPU restatement (TEST INFRASTRUCTURE).

Follows `pxlshift/alignment_pixels.py:14-156` (`AlignmentPixels`): the large image is brought to the small image's
pixel size (`_sub_resolution_large_fov`, order-1 `map_coordinates`, `:126-143`), optionally shifted along x by the
solar rotation between the two exposures (`_shift_large_fov`, `:86-108`), the small image is rotated about its centre
per rotation lag (`matrix_transform.polar_transform`, `utils/matrix_transform.py:77-106`, order-1 sampling,
`:72-79`), and every integer (dx, dy) lag scores the centred slice of the large image displaced by the lag against
the small image with a NaN-masked Pearson coefficient (`_step`, `:41-58`; `pxlshift/c_correlate.py:41-62`, whose
numerator is stored in a float32 array).

PINNED: `tests/golden/pxlshift_golden.npz` holds the output of the reference's own `find_best_parameters` run in
the build container on seeded inputs (`tests/golden/make_pxlshift_golden.py`); `tests/test_pxlshift.py` checks this
restatement against it bit for bit.
"""
from __future__ import annotations

import datetime

import numpy as np
from scipy.ndimage import map_coordinates

try:
    from numba import njit
except Exception:  # pragma: no cover
    njit = None

_TO_ARCSEC = {"arcsec": 1.0, "deg": 3600.0, "arcmin": 60.0, "rad": 3600.0 * 180.0 / np.pi}


def _convert(value, src, dst):
    src, dst = str(src).strip(), str(dst).strip()
    if src == dst:
        return value
    if (src, dst) == ("arcsec", "deg"):
        return value * (1.0 / 3600.0)
    return value * (_TO_ARCSEC[src] / _TO_ARCSEC[dst])


def _pearson_f32num_py(a, b):
    n = a.shape[0]
    s1 = 0.0
    for i in range(n):
        s1 += a[i]
    s2 = 0.0
    for i in range(n):
        s2 += b[i]
    m1 = s1 / n
    m2 = s2 / n
    sab = 0.0
    for i in range(n):
        sab += (a[i] - m1) * (b[i] - m2)
    saa = 0.0
    for i in range(n):
        d = a[i] - m1
        saa += d * d
    sbb = 0.0
    for i in range(n):
        d = b[i] - m2
        sbb += d * d
    return sab, saa, sbb


_pearson_f32num_jit = njit(cache=False, error_model="numpy")(_pearson_f32num_py) if njit is not None else None


def pearson_f32num(a, b) -> float:
    """`pxlshift/c_correlate.py:41-62` at lag 0: sequential float64 sums, the numerator rounded to float32 (it is stored
    in `np.zeros(len(lags), dtype="float32")`), then divided in float64."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    if a.size == 0:
        return float("nan")
    if _pearson_f32num_jit is not None:
        sab, saa, sbb = _pearson_f32num_jit(a, b)
    else:  # pragma: no cover
        ac, bc = a - np.cumsum(a)[-1] / a.size, b - np.cumsum(b)[-1] / b.size
        sab, saa, sbb = np.cumsum(ac * bc)[-1], np.cumsum(ac * ac)[-1], np.cumsum(bc * bc)[-1]
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        return float(np.float64(np.float32(sab)) / np.sqrt(saa * sbb))


def interpol2d(image, x, y, order=1, fill=0.0):
    """`rectify.interpol2d` / `AlignCommonUtil.interpol2d` with `dst=None`: output in the image's dtype."""
    x = np.asarray(x)
    y = np.asarray(y)
    dst = np.empty(x.shape, dtype=image.dtype)
    out = dst.ravel()
    map_coordinates(image, np.stack((y.ravel(), x.ravel()), axis=0), order=order, mode="constant", cval=fill,
                    output=out, prefilter=False)
    return out.reshape(x.shape)


def polar_transform(xx, yy, theta, units="radian"):
    """`MatrixTransform.polar_transform` with two arguments (`utils/matrix_transform.py:77-106`): rotation by theta
    about the pixel (round(H / 2), round(W / 2)) -- Python's round, half to even."""
    if units == "degree":
        theta = np.radians(theta)
    xc = xx[round(xx.shape[0] / 2), round(xx.shape[1] / 2)]
    yc = yy[round(xx.shape[0] / 2), round(xx.shape[1] / 2)]
    nr = np.sqrt(np.power(xx - xc, 2) + np.power(yy - yc, 2))
    ntheta = np.arctan2(yy - yc, xx - xc)
    ntheta[np.isnan(ntheta)] = 0
    ntheta = ntheta + theta
    return np.multiply(nr, np.cos(ntheta)) + xc, np.multiply(nr, np.sin(ntheta)) + yc


def diff_rot(lat, wvl="default"):
    p = {"EIT 171": (14.56, -2.65, 0.96), "EIT 195": (14.50, -2.14, 0.66), "EIT 284": (14.60, -0.71, -1.18),
         "EIT 304": (14.51, -3.12, 0.34)}
    p["default"] = p["EIT 195"]
    a, b, c = p[wvl]
    corr = a - 360 / 25.38 + b * np.sin(lat) ** 2 + c * np.sin(lat) ** 4
    return np.deg2rad(corr / 86400)


def _seconds_between(iso_a, iso_b):
    fa = datetime.datetime.fromisoformat(str(iso_a))
    fb = datetime.datetime.fromisoformat(str(iso_b))
    return (fa - fb).total_seconds()


class PixelShiftSearch:
    """State of `AlignmentPixels` given the two images and headers (what `__init__` reads from the FITS files)."""

    def __init__(self, data_large, hdr_large, data_small, hdr_small):
        self.data_large = np.array(data_large, dtype=np.float64)
        self.hdr_large = dict(hdr_large)
        self.data_small = np.array(data_small, dtype=np.float64)
        self.hdr_small = dict(hdr_small)

    # `_return_shift_large_fov_solar_rotation` (:110-124), arcsec
    def solar_rotation_shift_arcsec(self):
        band = self.hdr_large["WAVELNTH"]
        b0 = np.deg2rad(self.hdr_large["SOLAR_B0"])
        omega_car = np.deg2rad(360 / 25.38 / 86400)
        if band == 174:
            band = 171
        omega = omega_car + diff_rot(b0, f"EIT {band}")
        rsun, dsun = self.hdr_large["RSUN_REF"], self.hdr_large["DSUN_OBS"]
        phi = omega * rsun / (dsun - rsun)
        phi = np.rad2deg(phi) * 3600
        dt = _seconds_between(self.hdr_small["DATE-AVG"], self.hdr_large["DATE-AVG"])
        return dt * phi

    # `_shift_large_fov` (:86-108)
    def shift_large_fov(self):
        xx, yy = np.meshgrid(np.arange(self.data_large.shape[1]), np.arange(self.data_large.shape[0]))
        data_large = interpol2d(self.data_large, xx, yy, fill=-32762, order=1)
        data_large = np.where(data_large == -32762, np.nan, data_large)
        dcrval = self.solar_rotation_shift_arcsec()
        h = self.hdr_large
        if "CROTA" in h:
            theta = np.deg2rad(h["CROTA"])
            dx = (_convert(dcrval, "arcsec", h["CUNIT1"]) / h["CDELT1"]) * np.cos(-theta)
            dy = (_convert(dcrval, "arcsec", h["CUNIT2"]) / h["CDELT2"]) * np.sin(-theta)
        else:
            dx = _convert(dcrval, "arcsec", h["CUNIT1"]) / h["CDELT1"]
            dy = 0
        mat = np.array([[1, 0, dx], [0, 1, dy], [0, 0, 1]])
        xyz = np.stack((xx.ravel(), yy.ravel(), np.ones(xx.shape).ravel()))
        nx, ny, _ = np.matmul(mat, xyz)
        data_large = interpol2d(data_large, nx.reshape(xx.shape), ny.reshape(yy.shape), fill=-32762, order=1)
        self.data_large = np.where(data_large == -32762, np.nan, data_large)
        return dx, dy

    # `_sub_resolution_large_fov` (:126-143)
    def sub_resolution_large_fov(self):
        hs, hl = self.hdr_small, self.hdr_large
        ratio1 = _convert(hs["CDELT1"], hs["CUNIT1"], hl["CUNIT1"]) / hl["CDELT1"]
        ratio2 = _convert(hs["CDELT2"], hs["CUNIT2"], hl["CUNIT2"]) / hl["CDELT2"]
        x, y = np.meshgrid(np.arange(0, self.data_large.shape[1], ratio1),
                           np.arange(0, self.data_large.shape[0], ratio2))
        self.data_large = interpol2d(self.data_large, x=x, y=y, order=1, fill=-32768)
        self.data_large[self.data_large == -32768] = np.nan

    def rotated_small(self, drot, unit_rot="degree"):
        if drot != 0:
            xx, yy = np.meshgrid(np.arange(self.data_small.shape[1]), np.arange(self.data_small.shape[0]))
            nx, ny = polar_transform(xx, yy, theta=drot, units=unit_rot)
            rot = interpol2d(self.data_small.copy(), x=nx, y=ny, fill=-32762, order=1)
            rot[rot == -32762] = np.nan
            return rot
        return self.data_small.copy()

    def step(self, small_rot, dx, dy):
        slc = (slice(self.slc[0].start + dy, self.slc[0].stop + dy), slice(self.slc[1].start + dx, self.slc[1].stop + dx))
        for n in range(2):
            if slc[n].start < 0 or slc[n].stop > self.data_large.shape[n]:
                raise ValueError("too large shift : outside FSI")
        win = self.data_large[slc[0], slc[1]]
        if win.shape != small_rot.shape:
            raise ValueError("shapes not similar")
        is_nan = np.isnan(win.ravel()) | np.isnan(small_rot.ravel())
        return pearson_f32num(small_rot.ravel()[~is_nan], win.ravel()[~is_nan])

    def find_best_parameters(self, lag_dx, lag_dy, lag_drot, unit_rot="degree", shift_solar_rotation_dx_large=False):
        if shift_solar_rotation_dx_large:
            self.shift_large_fov()
        self.sub_resolution_large_fov()
        lo = [int((self.data_large.shape[n] - self.data_small.shape[n] - 1) / 2) for n in range(2)]
        self.slc = (slice(lo[0], lo[0] + self.data_small.shape[0]), slice(lo[1], lo[1] + self.data_small.shape[1]))
        corr = np.zeros((len(lag_dx), len(lag_dy), len(lag_drot)), dtype=np.float64)
        for kk, drot in enumerate(lag_drot):
            self.data_small_rotated = self.rotated_small(drot, unit_rot)
            for ii, dx in enumerate(lag_dx):
                for jj, dy in enumerate(lag_dy):
                    corr[ii, jj, kk] = self.step(self.data_small_rotated, int(dx), int(dy))
        return corr
