"""Synthetic pair of Carrington maps (CRLN-CAR / CRLT-CAR) with a known pointing error, for
`Alignment.align_using_initial_carrington` (`hdrshift/alignment.py:344-399`).

A positive log-normal field is defined on a master plate-carree grid; the small and the large map are rendered by
sampling it at their true Carrington coordinates; the small map's header is then written with a CRVAL error so that the
search must recover `true_shift` (degrees). Everything is seeded.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
from scipy.ndimage import gaussian_filter, map_coordinates

from .._compat import fits_lite
from .._compat.wcs import CarWcs


@dataclass
class CarPairSpec:
    small_n: tuple = (96, 64)           # NAXIS1, NAXIS2
    large_n: tuple = (160, 120)
    small_cdelt: float = 0.02           # deg / px
    large_cdelt: float = 0.05
    master_n: int = 768
    master_cdelt: float = 0.008
    center: tuple = (250.0, 0.0)        # deg: CRVAL of the large and the master map
    true_crval: tuple = (250.3, 0.2)    # deg: true CRVAL of the small map
    true_shift: tuple = (0.12, -0.06)   # deg: correction the search must find
    crota: float = 0.0                  # deg, small map only
    seed: int = 304
    date: str = "2022-03-17T09:50:45.000"
    lonpole: bool = False               # write LONPOLE / LATPOLE keywords (astropy's to_header does)


def car_header(n1, n2, cdelt, crval, crota=0.0, date="2022-03-17T09:50:45.000", lonpole=False):
    rho = np.deg2rad(crota)
    h = fits_lite.Header()
    h["NAXIS1"], h["NAXIS2"] = int(n1), int(n2)
    h["CTYPE1"], h["CTYPE2"] = "CRLN-CAR", "CRLT-CAR"
    h["CUNIT1"], h["CUNIT2"] = "deg", "deg"
    h["CRPIX1"], h["CRPIX2"] = (n1 + 1) / 2.0, (n2 + 1) / 2.0
    h["CDELT1"], h["CDELT2"] = float(cdelt), float(cdelt)
    h["CRVAL1"], h["CRVAL2"] = float(crval[0]), float(crval[1])
    h["CROTA"] = float(crota)
    h["PC1_1"], h["PC1_2"] = float(np.cos(rho)), float(-np.sin(rho))
    h["PC2_1"], h["PC2_2"] = float(np.sin(rho)), float(np.cos(rho))
    if lonpole:
        h["LONPOLE"] = 0.0 if crval[1] >= 0.0 else 180.0
        h["LATPOLE"] = 90.0
    h["DATE-OBS"] = date
    h["DATE-AVG"] = date
    h["WAVELNTH"] = 304
    h["BUNIT"] = "DN/s"
    return h


def master_map(spec: CarPairSpec):
    rng = np.random.default_rng(spec.seed)
    n = spec.master_n
    white = rng.standard_normal((n, n)).astype(np.float32)
    field = np.zeros((n, n), dtype=np.float32)
    for sig_deg, amp in ((0.03, 0.5), (0.12, 0.8), (0.5, 1.0)):
        g = gaussian_filter(white, sig_deg / spec.master_cdelt, mode="wrap")
        field += amp * g / g.std()
    sky = np.exp(0.55 * field / field.std()).astype(np.float64)
    sky = (sky - sky.mean()) / sky.std() * 300.0 + 500.0
    return np.maximum(sky, 5.0)


def _render(sky, w_master: CarWcs, w_img: CarWcs, n1, n2, floor):
    x, y = np.meshgrid(np.arange(n1, dtype=np.float64), np.arange(n2, dtype=np.float64))
    lon, lat = w_img.pixel_to_world(x, y)
    mx, my = w_master.world_to_pixel(lon, lat)
    return map_coordinates(sky, np.stack((my.ravel(), mx.ravel())), order=1, mode="constant",
                           cval=floor).reshape(n2, n1)


def make_car_pair(out_dir, spec: CarPairSpec | None = None, tag="car"):
    """Write `<out_dir>/<tag>_small.fits` and `<tag>_large.fits`; returns (path_large, path_small, spec)."""
    spec = spec or CarPairSpec()
    os.makedirs(out_dir, exist_ok=True)
    p_small = os.path.join(out_dir, f"{tag}_small.fits")
    p_large = os.path.join(out_dir, f"{tag}_large.fits")
    sky = master_map(spec)
    w_master = CarWcs.from_header(car_header(spec.master_n, spec.master_n, spec.master_cdelt, spec.center))
    h_true = car_header(spec.small_n[0], spec.small_n[1], spec.small_cdelt, spec.true_crval, spec.crota, spec.date,
                        spec.lonpole)
    small = _render(sky, w_master, CarWcs.from_header(h_true), spec.small_n[0], spec.small_n[1], 50.0)
    rng = np.random.default_rng(spec.seed + 1)
    small = small + rng.normal(0.0, 3.0, small.shape)
    h_small = h_true.copy()
    h_small["CRVAL1"] = spec.true_crval[0] - spec.true_shift[0]
    h_small["CRVAL2"] = spec.true_crval[1] - spec.true_shift[1]
    if spec.lonpole:
        h_small["LONPOLE"] = 0.0 if h_small["CRVAL2"] >= 0.0 else 180.0
    fits_lite.writeto(p_small, [fits_lite.PrimaryHDU(small.astype(np.float32), h_small)], overwrite=True)
    h_large = car_header(spec.large_n[0], spec.large_n[1], spec.large_cdelt, spec.center, 0.0, spec.date, spec.lonpole)
    blur = gaussian_filter(sky, 0.5 * spec.large_cdelt / spec.master_cdelt / 1.2, mode="nearest")
    large = _render(blur, w_master, CarWcs.from_header(h_large), spec.large_n[0], spec.large_n[1], 50.0)
    rng = np.random.default_rng(spec.seed)
    large = large + rng.normal(0.0, 2.0, large.shape)
    fits_lite.writeto(p_large, [fits_lite.PrimaryHDU(large.astype(np.float32), h_large)], overwrite=True)
    return p_large, p_small, spec
