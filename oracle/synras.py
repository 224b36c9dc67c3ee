"""Synthetic raster + SPICE L2 preparation, CPU restatement (TEST INFRASTRUCTURE).

* `build_synras`  <- `SPICEComposedMapBuilder._create_map_from_hdu` / `_prepare_spectro_data` (level 2, raster grid)
                     `synras/map_builder.py:87-131, 249-294`: per raster column, the imager frame closest in time
                     to the column's mean exposure time, `world_to_pixel` of the slit's sky positions in that
                     frame, `interpol2d(order=2, fill=nan)` (float64 out, the imager's dtype promoted by scipy).
* `spice_l2_image` <- `AlignmentSpice._prepare_spice_from_l2` `hdrshift/alignment_spice.py:250-323`.
* `xy_header`     <- what `w_xy.to_header()` holds for the celestial pair (degrees; wcslib/astropy not available
                     here => restated, "parity unpinned" like oracle/wcs_tan.py).
Times are seconds since J2000 (UTC without leap seconds); headers are plain dicts.
"""
from __future__ import annotations

import numpy as np

from . import wcs_tan
from .resample import interpol2d

_UNIT_TO_DEG = {"deg": 1.0, "arcsec": 1.0 / 3600.0, "arcmin": 1.0 / 60.0}


def _to_s(iso):
    d = np.datetime64(str(iso).replace("Z", ""), "ms")
    return float((d - np.datetime64("2000-01-01T12:00:00", "ms")) / np.timedelta64(1, "ms")) / 1000.0


def xy_header(h4):
    """2-D celestial header of a 4-axis SPICE header, as `WCS(h4).dropaxis(2)...dropaxis(2).to_header()`."""
    s1, s2 = _UNIT_TO_DEG[h4["CUNIT1"].strip()], _UNIT_TO_DEG[h4["CUNIT2"].strip()]
    out = {"WCSAXES": 2, "CRPIX1": float(h4["CRPIX1"]), "CRPIX2": float(h4["CRPIX2"])}
    for k, default in (("PC1_1", 1.0), ("PC1_2", 0.0), ("PC2_1", 0.0), ("PC2_2", 1.0)):
        v = float(h4[k]) if k in h4 else default
        if v != default:
            out[k] = v
    out.update(CDELT1=float(h4["CDELT1"]) * s1, CDELT2=float(h4["CDELT2"]) * s2, CUNIT1="deg", CUNIT2="deg",
               CTYPE1=h4["CTYPE1"], CTYPE2=h4["CTYPE2"], CRVAL1=float(h4["CRVAL1"]) * s1,
               CRVAL2=float(h4["CRVAL2"]) * s2, LONPOLE=float(h4["LONPOLE"]) if "LONPOLE" in h4 else 180.0)
    return out


def slit_limits(h):
    """`AlignSpiceUtil.vertical_edges_limits` (`utils/Util.py:430-455`)."""
    ybin = h["NBIN2"]
    h_detector = 1024 / ybin
    h_slit = {"SW": 600, "LW": 626}[h["DETECTOR"]] / ybin
    beg = (h_detector - h_slit) / 2
    end = h_detector - beg
    beg = int(np.ceil(beg - h["PXBEG2"] / ybin + 1)) + int(20 / ybin)
    end = int(np.floor(end - h["PXBEG2"] / ybin + 1)) - int(20 / ybin)
    return beg, end


def spice_l2_image(data4, h4, wave_interval="all", sub_fov_arcsec=None):
    """2-D image + 2-D header of an L2 cube [1, n_lambda, ny, nx]. `sub_fov_arcsec` = (lon_min, lon_max, lat_min,
    lat_max): pixels outside are set to NaN (`alignment_spice.py:289-312`; longitudes as wcslib reports them)."""
    data = np.array(data4, dtype=np.float64)
    ymin, ymax = slit_limits(h4)
    data[:, :, :ymin, :] = np.nan
    data[:, :, ymax:, :] = np.nan
    if isinstance(wave_interval, str):
        img = np.nansum(data[0, :, :, :], axis=0)
    else:
        z = np.arange(data.shape[1])
        wave = h4["CRVAL3"] + h4["CDELT3"] * (h4["PC3_3"] if "PC3_3" in h4 else 1.0) * (z + 1 - h4["CRPIX3"])
        sel = (wave >= wave_interval[0]) & (wave <= wave_interval[1])
        img = np.nansum(data[0, sel, :, :], axis=0)
    img[:ymin, :] = np.nan
    img[ymax:, :] = np.nan
    if sub_fov_arcsec is not None:
        ny, nx = img.shape
        w = wcs_tan.WcsTan(dict(xy_header(h4), NAXIS1=nx, NAXIS2=ny))
        lon, lat = w.pixel_to_world(*np.meshgrid(np.arange(nx), np.arange(ny)))
        lims = [v * (1.0 / 3600.0) for v in sub_fov_arcsec]
        sel = (lon >= lims[0]) & (lon <= lims[1]) & (lat >= lims[2]) & (lat <= lims[3])
        img[~sel] = np.nan
    hdr = xy_header(h4)
    for k in ("SOLAR_B0", "RSUN_REF", "DSUN_OBS", "CROTA"):
        hdr[k] = h4[k]
    hdr["NAXIS1"], hdr["NAXIS2"] = img.shape[1], img.shape[0]
    return img, hdr


def build_synras(h4, imager_frames, imager_headers, threshold_s, order=2, keep_original_imager_pixel_size=False):
    """-> (data_composed float64 [rows, columns], frame index per column). With `keep_original_imager_pixel_size` the
    grid is `np.arange(0, NAXIS, CDELT_imager / CDELT_raster)` along each axis (`map_builder.py:259-275`)."""
    nx, ny = int(h4["NAXIS1"]), int(h4["NAXIS2"])
    hxy = dict(xy_header(h4), NAXIS1=nx, NAXIS2=ny)
    w = wcs_tan.WcsTan(hxy)
    if keep_original_imager_pixel_size:
        h_im = imager_headers[0]
        x, y = np.meshgrid(np.arange(0, nx, h_im["CDELT1"] / h4["CDELT1"]), np.arange(0, ny, h_im["CDELT2"] / h4["CDELT2"]))
        ny, nx = x.shape
    else:
        x, y = np.meshgrid(np.arange(nx), np.arange(ny))
    lng, lat = w.pixel_to_world(x, y)          # no ang2pipi here (map_builder.py:294)
    t_ref = _to_s(h4["DATEREF"] if "DATEREF" in h4 else h4["DATE-BEG"])
    pc41 = h4["PC4_1"] if "PC4_1" in h4 else 0.0
    pc42 = h4["PC4_2"] if "PC4_2" in h4 else 0.0
    pc44 = h4["PC4_4"] if "PC4_4" in h4 else 1.0
    t = t_ref + h4["CRVAL4"] + h4["CDELT4"] * (pc41 * (x + 1 - h4["CRPIX1"]) + pc42 * (y + 1 - h4["CRPIX2"])
                                               + pc44 * (0 + 1 - h4["CRPIX4"]))
    dates = np.array([_to_s(h["DATE-AVG"]) for h in imager_headers])
    out = np.empty((ny, nx), dtype=np.float64)
    chosen = np.empty(nx, dtype=np.int64)
    for ii in range(nx):
        col = t[:, ii]
        utc_slit = col[0] - np.mean(col[0] - col)
        delta = np.abs(utc_slit - dates)
        k = int(delta.argmin())
        if delta.min() > threshold_s:
            raise ValueError("Could not find imager sufficiently close in time")
        chosen[ii] = k
        xf, yf = wcs_tan.WcsTan(imager_headers[k]).world_to_pixel(lng[:, ii], lat[:, ii])
        out[:, ii] = interpol2d(np.asarray(imager_frames[k]), x=xf, y=yf, order=order, fill=np.nan)
    return out, chosen
