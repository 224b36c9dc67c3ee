"""Carrington-grid ("fa") lag search, CPU restatement (TEST INFRASTRUCTURE).

Follows `Alignment.align_using_carrington(method_carrington_reprojection="fa")`:
`hdrshift/alignment.py:144-261, 889-901` -> `utils/rectify.py` `CarringtonTransform.__init__` `:377-423`,
`DifferentialRotationTransform.forward` `:304-311`, `SphericalTransform.forward` `:340-363`,
`Rectifier.__call__` `:865-888`, `interpol2d` `:22-56`.

NumPy >= 2 dtype trail reproduced by construction (the same numpy expressions on the same dtypes): the lon/lat
grid is float32, `np.radians(lat)` / `sin(lat)` / `cos(lat)` stay float32, longitude becomes float64 when
`radians(CRLN_OBS)` (an np.float64 scalar) is subtracted. `rate_wave` is always None in the reference (the int
WAVELNTH is compared with string keys, `alignment.py:891-894`), so the differential-rotation
coefficients are (14.18, 0, 0) and dx == 0 for any delta_t; `delta_t` is a Python float (scalar TimeDelta.value).
"""
from __future__ import annotations

import numpy as np

from .hpc import LagKillsWorker, Refs, check_and_create_pcij, shift_header, threshold_to_nan
from .pearson import masked_pearson
from .resample import interpol2d

R_SUN = 695700000.0  # astropy.constants.R_sun.value


def _iso_to_days(s):
    d = np.datetime64(str(s).replace("Z", ""), "ms")
    return float((d - np.datetime64("2000-01-01T12:00:00", "ms")) / np.timedelta64(1, "ms")) / 86400000.0


class CarringtonTransform:
    """`rectify.CarringtonTransform` (`utils/rectify.py:377-423`) as one callable."""

    def __init__(self, hdr, radius_correction=1.0, reference_date=None, rate_wave=None):
        if "CROTA" in hdr:
            roll = hdr["CROTA"]
        elif "CROTA2" in hdr:
            roll = hdr["CROTA2"]
        else:
            raise ValueError("No roll value found in header")
        cos = np.cos(np.radians(roll))
        sin = np.sin(np.radians(roll))
        dx = cos * hdr["CRVAL1"] + sin * hdr["CRVAL2"]
        dy = -sin * hdr["CRVAL1"] + cos * hdr["CRVAL2"]
        self.x = (hdr["CRPIX1"] - 1) - dx / hdr["CDELT1"]
        self.y = (hdr["CRPIX2"] - 1) - dy / hdr["CDELT2"]
        self.dist = hdr["DSUN_OBS"] / (radius_correction * R_SUN)
        self.lon = np.radians(hdr["CRLN_OBS"])
        self.lat = np.radians(hdr["CRLT_OBS"])
        self.roll = np.radians(roll)
        self.cdelt1 = hdr["CDELT1"]
        self.cdelt2 = hdr["CDELT2"]
        if reference_date is None:
            reference_date = hdr["DATE-OBS"]
        self.delta_t = float(_iso_to_days(hdr["DATE-OBS"]) - _iso_to_days(reference_date))
        self.carrington_rate = 14.18
        self.coeffs = (self.carrington_rate, 0, 0) if rate_wave is None else rate_wave

    def __call__(self, x, y):
        # DifferentialRotationTransform.forward
        siny2 = np.sin(np.radians(y)) ** 2
        dx = self.delta_t * (self.coeffs[0] + siny2 * (self.coeffs[1] + self.coeffs[2] * siny2) - self.carrington_rate)
        x = x - dx
        # SphericalTransform.forward
        lon = np.radians(x) - self.lon
        lat = np.radians(y)
        X = np.cos(lat) * np.sin(lon)
        Y = np.sin(lat)
        Z = np.cos(lat) * np.cos(lon)
        zz = Z * np.cos(self.lat) + Y * np.sin(self.lat)
        yy = Y * np.cos(self.lat) - Z * np.sin(self.lat)
        gd = zz >= 0
        y2 = yy[gd] * np.cos(self.roll) - X[gd] * np.sin(self.roll)
        x2 = X[gd] * np.cos(self.roll) + yy[gd] * np.sin(self.roll)
        z2 = self.dist - zz[gd]
        nx = np.full_like(lon, np.nan)
        ny = np.full_like(lon, np.nan)
        nx[gd] = self.x + np.degrees(np.arctan(x2 / z2)) * 3600 / self.cdelt1
        ny[gd] = self.y + np.degrees(np.arctan(y2 / z2)) * 3600 / self.cdelt2
        return nx, ny


def rectify(image, transform, shape, lonlims, latlims, order, fill):
    """`Rectifier.__call__` (`utils/rectify.py:865-888`) with dtype=float32 grid."""
    x, y = np.meshgrid(np.linspace(lonlims[0], lonlims[1], shape[0], dtype=np.float32),
                       np.linspace(latlims[0], latlims[1], shape[1], dtype=np.float32))
    nx, ny = transform(x, y)
    dst = np.empty(nx.shape, dtype=image.dtype)
    interpol2d(image, nx, ny, order=order, fill=fill, dst=dst)
    return dst


def carrington_transform_fa(data, hdr, d_solar_r, reference_date, shape, lonlims, latlims, order):
    """`Alignment._carrington_transform_fa` (`hdrshift/alignment.py:889-901`)."""
    t = CarringtonTransform(hdr, radius_correction=d_solar_r, reference_date=reference_date, rate_wave=None)
    image = rectify(data, t, shape, lonlims, latlims, order, -32762)
    return np.where(image == -32762, np.nan, image)


class CarringtonSearch:
    def __init__(self, data_large, hdr_large, data_small, hdr_small, lag_crval1, lag_crval2, lag_cdelt1, lag_cdelt2,
                 lag_crota, lonlims, latlims, shape, lag_solar_r=None, reference_date=None, order=2,
                 small_fov_value_min=None, small_fov_value_max=None, unit_lag="arcsec", force_crota_0=False):
        import warnings
        self.order = order
        self.hdr_small = dict(hdr_small)
        self.hdr_large = dict(hdr_large)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            check_and_create_pcij(self.hdr_small, force_crota_0)
            check_and_create_pcij(self.hdr_large, force_crota_0)
        self.reference_date = self.hdr_large["DATE-AVG"] if reference_date is None else reference_date
        self.lonlims, self.latlims, self.shape = lonlims, latlims, shape
        self.data_small = np.array(data_small, dtype=np.float64)
        threshold_to_nan(self.data_small, small_fov_value_min, small_fov_value_max)
        self.refs = Refs(self.hdr_small, lag_crval1, lag_crval2, lag_cdelt1, lag_cdelt2, lag_crota, lag_solar_r,
                         unit_lag=unit_lag)
        if len(self.refs.lag_solar_r) != 1:
            raise ValueError("the reference only works with one lag_solar_r value (SURVEY App. B4)")
        self.d_solar_r = self.refs.lag_solar_r[0]
        self.data_large = carrington_transform_fa(np.array(data_large, dtype=np.float64), self.hdr_large,
                                                  self.d_solar_r, self.reference_date, shape, lonlims, latlims, order)

    @property
    def cube_shape(self):
        r = self.refs
        return (len(r.lag_crval1), len(r.lag_crval2), len(r.lag_cdelt1), len(r.lag_cdelt2), len(r.lag_crota), 1)

    def reprojected(self, d1, d2, d3, d4, d5):
        hdr = dict(self.hdr_small)
        shift_header(hdr, self.refs, d1, d2, d3, d4, d5, "reference")
        return carrington_transform_fa(self.data_small, hdr, self.d_solar_r, self.reference_date, self.shape,
                                       self.lonlims, self.latlims, self.order)

    def step(self, d1, d2, d3, d4, d5):
        try:
            interp = self.reprojected(d1, d2, d3, d4, d5)
        except LagKillsWorker:
            return 0.0
        return masked_pearson(self.data_large, interp)

    def cube(self):
        r = self.refs
        g = np.meshgrid(r.lag_crval1, r.lag_crval2, r.lag_cdelt1, r.lag_cdelt2, r.lag_crota, indexing="ij")
        flat = [a.ravel() for a in g]
        out = np.zeros(flat[0].size, dtype=np.float64)
        for i in range(out.size):
            out[i] = self.step(*(f[i] for f in flat))
        return out.reshape(self.cube_shape)
