"""Carrington search with `method_carrington_reprojection="sunpy"`, CPU restatement (TEST INFRASTRUCTURE).

PARITY UNPINNED. The reference delegates this path to third-party code that is absent from the image -- sunpy 6.1.2,
reproject 0.14.1, astropy 7.2.0 in the reference's poetry.lock (pyproject: sunpy ^6.0.4): `hdrshift/alignment.py:939-985`

    with propagate_with_solar_surface():
        map_ref_rep = map_ref.reproject_to(map_to_align.wcs)            # once: large -> grid of the small image
    ...
    with propagate_with_solar_surface():
        map_to_align_rep = Map(data_small, hdr_shifted).reproject_to(w_large)   # per lag, w_large = WCS(hdr_small)

What is restated here is the published algorithm of those calls:

* `Map.reproject_to(wcs)` = `reproject.reproject_interp(map, wcs, shape_out)`: for every output pixel, world coordinates
  through the output WCS, transformed to the input map's coordinate frame, input pixel through the input WCS, BILINEAR
  interpolation (`order='bilinear'`), NaN outside. reproject's `map_coordinates` wrapper pads the input by one
  replicated pixel and accepts coordinates within half a pixel of the array edge (-0.5 <= x <= n - 0.5)
  (`bilinear_edge`); `roundtrip_coords=True` blanks output pixels whose input position does not map back onto them
  (here: surface points the input observer cannot see, `visible`).
* helioprojective -> helioprojective between two observers (sunpy `hpc_to_hpc`): the 2-D direction is put on the
  sphere of radius rsun (`make_3d`; off-disc -> NaN), helioprojective -> heliocentric-cartesian of the first
  observer -> Stonyhurst -> heliocentric-cartesian of the second observer -> helioprojective. Equal observers and
  equal times short-circuit to the identity (the per-lag call: both frames come from the small image's header).
* `propagate_with_solar_surface()` (rotation_model='howard') = `transform_with_sun_center()` + differential rotation of
  the Stonyhurst longitude over the time between the two frames: `sunpy.sun.models.differential_rotation(dt, lat,
  model='howard', frame_time='synodic')` = (2.894 - 0.428 sin^2 lat - 0.370 sin^4 lat) urad/s * dt - 0.9856 deg/day * dt.
* observer and time of a frame: HGLN_OBS / HGLT_OBS / DSUN_OBS and DATE-AVG (else DATE-OBS) of the header
  (sunpy `wcs_utils`), rsun = `d_solar_r` * R_sun (`alignment.py:940`, RSUN_REF set on both maps).

Consequence worth stating: the per-lag reprojection is the helioprojective search's own geometry (the unshifted small
grid sampled through the shifted header) with bilinear interpolation and reproject's edge rule; nothing Carrington is
left in it. The correlation is the masked Pearson of `hdrshift/c_correlate.py` over pixels finite in both images
(`alignment.py:524-531`).
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import map_coordinates

from . import wcs_tan
from .hpc import Refs, LagKillsWorker, check_and_create_pcij, shift_header
from .pearson import masked_pearson

R_SUN_M = 695700000.0                       # astropy.constants.R_sun (IAU 2015 nominal)
HOWARD_URAD_S = (2.894, -0.428, -0.370)     # sunpy.sun.models.differential_rotation, model='howard'
SYNODIC_DEG_DAY = 0.9856


def _seconds(date):
    import datetime as dt
    s = str(date).strip().rstrip("Z")
    fmt = "%Y-%m-%dT%H:%M:%S.%f" if "." in s else "%Y-%m-%dT%H:%M:%S"
    return (dt.datetime.strptime(s, fmt) - dt.datetime(2000, 1, 1, 12)).total_seconds()


def frame_of(hdr, rsun_m):
    """Observer (Stonyhurst lon, lat [rad], distance [m]), observation time [s] and rsun [m] of a header's frame."""
    for k in ("HGLN_OBS", "HGLT_OBS", "DSUN_OBS"):
        if k not in hdr:
            raise ValueError(f"the sunpy reprojection needs {k} in the header")
    date = hdr["DATE-AVG"] if "DATE-AVG" in hdr else hdr["DATE-OBS"]
    return dict(lon=np.deg2rad(float(hdr["HGLN_OBS"])), lat=np.deg2rad(float(hdr["HGLT_OBS"])),
                dsun=float(hdr["DSUN_OBS"]), t=_seconds(date), rsun=float(rsun_m))


def differential_rotation_deg(dt_days, lat_rad):
    a, b, c = HOWARD_URAD_S
    s2 = np.sin(lat_rad) ** 2
    rate = (a + b * s2 + c * s2 * s2) * 1e-6 * 86400.0          # rad / day, sidereal
    return np.rad2deg(rate * dt_days) - SYNODIC_DEG_DAY * dt_days


def _basis(lon, lat):
    """Heliocentric-cartesian axes (x west, y north, z to the observer) of an observer at Stonyhurst (lon, lat)."""
    ez = np.array([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
    ex = np.array([-np.sin(lon), np.cos(lon), 0.0])
    ey = np.cross(ez, ex)
    return ex, ey, ez


def hpc_to_hpc_on_surface(tx_deg, ty_deg, f_from, f_to):
    """(Tx, Ty) [deg] of frame `f_from` -> (Tx, Ty) [deg] of frame `f_to` for points on the solar surface, rotated
    differentially over the time between the frames; `visible`: the point faces the second observer. NaN off-disc."""
    if (f_from["lon"], f_from["lat"], f_from["dsun"], f_from["t"]) == (f_to["lon"], f_to["lat"], f_to["dsun"], f_to["t"]):
        return np.asarray(tx_deg, dtype=np.float64), np.asarray(ty_deg, dtype=np.float64), \
            np.ones(np.shape(tx_deg), dtype=bool)
    tx, ty = np.deg2rad(tx_deg), np.deg2rad(ty_deg)
    D, R = f_from["dsun"], f_from["rsun"]
    cosa = np.cos(ty) * np.cos(tx)
    with np.errstate(invalid="ignore"):
        d = D * cosa - np.sqrt(D * D * cosa * cosa - D * D + R * R)      # near intersection with the sphere (make_3d)
    x = d * np.cos(ty) * np.sin(tx)
    y = d * np.sin(ty)
    z = D - d * np.cos(ty) * np.cos(tx)
    ex, ey, ez = _basis(f_from["lon"], f_from["lat"])
    p = x[..., None] * ex + y[..., None] * ey + z[..., None] * ez        # Stonyhurst cartesian at the first time
    r = np.sqrt((p * p).sum(-1))
    lon = np.arctan2(p[..., 1], p[..., 0])
    lat = np.arcsin(p[..., 2] / r)
    lon = lon + np.deg2rad(differential_rotation_deg((f_to["t"] - f_from["t"]) / 86400.0, lat))
    q = np.stack([r * np.cos(lat) * np.cos(lon), r * np.cos(lat) * np.sin(lon), r * np.sin(lat)], -1)
    ex, ey, ez = _basis(f_to["lon"], f_to["lat"])
    x2, y2, z2 = q @ ex, q @ ey, q @ ez
    D2 = f_to["dsun"]
    dist = np.sqrt(x2 * x2 + y2 * y2 + (D2 - z2) ** 2)
    tx2 = np.rad2deg(np.arctan2(x2, D2 - z2))
    ty2 = np.rad2deg(np.arcsin(y2 / dist))
    with np.errstate(invalid="ignore"):
        visible = z2 * D2 > r * r
    return tx2, ty2, visible


def bilinear_edge(image, x, y):
    """reproject's `map_coordinates(image, coords, order=1, cval=nan, mode='constant')`: one replicated pixel of padding,
    coordinates beyond half a pixel from the array edge (and NaN coordinates) -> NaN."""
    image = np.asarray(image, dtype=np.float64)
    ny, nx = image.shape
    padded = np.pad(image, 1, mode="edge")
    xx, yy = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        reset = ~((xx >= -0.5) & (xx <= nx - 0.5) & (yy >= -0.5) & (yy <= ny - 0.5))
    out = map_coordinates(padded, [np.where(reset, 0.0, yy).ravel() + 1.0, np.where(reset, 0.0, xx).ravel() + 1.0],
                          order=1, mode="constant", cval=np.nan, prefilter=False).reshape(xx.shape)
    out[reset] = np.nan
    return out


def reproject_to(data, hdr_in, hdr_out, rsun_m):
    """`Map(data, hdr_in).reproject_to(WCS(hdr_out))` under `propagate_with_solar_surface()`; RSUN_REF = rsun_m on both."""
    w_in, w_out = wcs_tan.WcsTan(hdr_in), wcs_tan.WcsTan(hdr_out)
    nx, ny = int(hdr_out["NAXIS1"]), int(hdr_out["NAXIS2"])
    gx, gy = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))
    lon, lat = w_out.pixel_to_world(gx, gy)
    tx, ty, visible = hpc_to_hpc_on_surface(lon, lat, frame_of(hdr_out, rsun_m), frame_of(hdr_in, rsun_m))
    x, y = w_in.world_to_pixel(tx, ty)
    x = np.where(visible, x, np.nan)
    return bilinear_edge(data, x, y)


class SurfaceSearch:
    """The lag cube of `align_using_carrington(method_carrington_reprojection="sunpy")` (`alignment.py:144-261, 613-797`)."""

    def __init__(self, data_large, hdr_large, data_small, hdr_small, lag_crval1, lag_crval2, lag_cdelt1, lag_cdelt2,
                 lag_crota, lag_solar_r=(1.004,), unit_lag="arcsec", cdelt_mode="reference"):
        self.hdr_large, self.hdr_small = dict(hdr_large), dict(hdr_small)
        check_and_create_pcij(self.hdr_small)
        check_and_create_pcij(self.hdr_large)
        self.data_small = np.array(data_small, dtype=np.float64)
        self.refs = Refs(self.hdr_small, lag_crval1, lag_crval2, lag_cdelt1, lag_cdelt2, lag_crota, lag_solar_r,
                         unit_lag=unit_lag, ang2pipi=True)
        self.cdelt_mode = cdelt_mode
        self.rsun = float(self.refs.lag_solar_r[0]) * R_SUN_M
        self.ref = reproject_to(np.array(data_large, dtype=np.float64), self.hdr_large, self.hdr_small, self.rsun)

    def step(self, d1, d2, d3, d4, d5):
        hdr = dict(self.hdr_small)
        shift_header(hdr, self.refs, d1, d2, d3, d4, d5, cdelt_mode=self.cdelt_mode)
        b = reproject_to(self.data_small, hdr, self.hdr_small, self.rsun)
        return masked_pearson(self.ref, b)

    def cube(self):
        r = self.refs
        shape = tuple(len(v) for v in (r.lag_crval1, r.lag_crval2, r.lag_cdelt1, r.lag_cdelt2, r.lag_crota))
        out = np.zeros(shape + (1,))
        for idx in np.ndindex(*shape):
            try:
                out[idx + (0,)] = self.step(r.lag_crval1[idx[0]], r.lag_crval2[idx[1]], r.lag_cdelt1[idx[2]],
                                            r.lag_cdelt2[idx[3]], r.lag_crota[idx[4]])
            except LagKillsWorker:
                pass
        return out
