"""Synthetic HRIEUV-like / FSI-like image pair with a known pointing error.

A positive log-normal "sky" is defined on a master TAN grid; the small-FOV and large-FOV images are
rendered by sampling it at their true sky positions; the small image's header is then written with a
pointing error so that the search must recover `true_shift` (arcsec). Everything is seeded.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
from scipy.ndimage import gaussian_filter, map_coordinates

from .._compat import fits_lite
from .._compat.wcs import TanWcs


@dataclass
class PairSpec:
    small_n: int = 2048
    large_n: int = 3072
    small_cdelt: float = 0.492        # arcsec / px
    large_cdelt: float = 4.44
    master_n: int = 4096
    master_cdelt: float = 0.35
    true_crval: tuple = (-100.0, 50.0)  # arcsec, true pointing of the small image
    true_shift: tuple = (24.0, 6.0)     # arcsec, correction the search must find
    crota: float = 3.0                  # deg, both images
    seed: int = 174
    noise_seed: int = 175
    wavelnth: int = 174
    date: str = "2022-03-17T09:50:45.000"
    dsun_obs: float = 5.7e10
    crln_obs: float = 250.0
    crlt_obs: float = -2.0
    jitter: tuple = (0.0, 0.0)          # extra true pointing offset (config 5 frames)


def _tan_header(n1, n2, cdelt, crval, crota, spec: PairSpec, telescop, detector):
    rho = np.deg2rad(crota)
    h = fits_lite.Header()
    h["NAXIS1"], h["NAXIS2"] = int(n1), int(n2)
    h["CTYPE1"], h["CTYPE2"] = "HPLN-TAN", "HPLT-TAN"
    h["CUNIT1"], h["CUNIT2"] = "arcsec", "arcsec"
    h["CRPIX1"], h["CRPIX2"] = (n1 + 1) / 2.0, (n2 + 1) / 2.0
    h["CDELT1"], h["CDELT2"] = float(cdelt), float(cdelt)
    h["CRVAL1"], h["CRVAL2"] = float(crval[0]), float(crval[1])
    h["CROTA"] = float(crota)
    h["PC1_1"], h["PC1_2"] = float(np.cos(rho)), float(-np.sin(rho))
    h["PC2_1"], h["PC2_2"] = float(np.sin(rho)), float(np.cos(rho))
    h["LONPOLE"] = 180.0
    h["DSUN_OBS"] = spec.dsun_obs
    h["RSUN_REF"] = 6.957e8
    h["CRLN_OBS"], h["CRLT_OBS"] = spec.crln_obs, spec.crlt_obs
    h["DATE-OBS"] = spec.date
    h["DATE-AVG"] = spec.date
    h["WAVELNTH"] = spec.wavelnth
    h["TELESCOP"] = telescop
    h["DETECTOR"] = detector
    h["BUNIT"] = "DN/s"
    return h


def master_scene(spec: PairSpec):
    """Log-normal field exp(sum_k a_k G_sigma_k * white), scaled to mean 500 / std 300."""
    rng = np.random.default_rng(spec.seed)
    n = spec.master_n
    white = rng.standard_normal((n, n)).astype(np.float32)
    field = np.zeros((n, n), dtype=np.float32)
    for sig_arcsec, amp in ((1.5, 0.45), (6.0, 0.7), (25.0, 1.0)):
        g = gaussian_filter(white, sig_arcsec / spec.master_cdelt, mode="wrap")
        field += amp * g / g.std()
    sky = np.exp(0.55 * field / field.std()).astype(np.float64)
    sky = (sky - sky.mean()) / sky.std() * 300.0 + 500.0
    return np.maximum(sky, 5.0)


def _render(sky, w_master: TanWcs, w_img: TanWcs, n1, n2, floor):
    x, y = np.meshgrid(np.arange(n1, dtype=np.float64), np.arange(n2, dtype=np.float64))
    lon, lat = w_img.pixel_to_world(x, y)
    mx, my = w_master.world_to_pixel(lon, lat)
    return map_coordinates(sky, np.stack((my.ravel(), mx.ravel())), order=1, mode="constant",
                           cval=floor).reshape(n2, n1)


def make_pair(out_dir, spec: PairSpec | None = None, tag="cfg", sky=None, write_large=True):
    """Write `<out_dir>/<tag>_small.fits` and `<tag>_large.fits`; returns (path_large, path_small, spec)."""
    spec = spec or PairSpec()
    os.makedirs(out_dir, exist_ok=True)
    p_small = os.path.join(out_dir, f"{tag}_small.fits")
    p_large = os.path.join(out_dir, f"{tag}_large.fits")
    if sky is None:
        sky = master_scene(spec)
    h_master = _tan_header(spec.master_n, spec.master_n, spec.master_cdelt, spec.true_crval, 0.0, spec, "", "")
    w_master = TanWcs.from_header(h_master)
    true_crval = (spec.true_crval[0] + spec.jitter[0], spec.true_crval[1] + spec.jitter[1])
    # small image at its TRUE pointing; header written with the pointing error
    h_true = _tan_header(spec.small_n, spec.small_n, spec.small_cdelt, true_crval, spec.crota, spec,
                         "SOLO/EUI/HRI_EUV", "HRI_EUV")
    small = _render(sky, w_master, TanWcs.from_header(h_true), spec.small_n, spec.small_n, 50.0)
    rng = np.random.default_rng(spec.noise_seed + 1)
    small = small + rng.normal(0.0, 3.0, small.shape)
    h_small = h_true.copy()
    h_small["CRVAL1"] = true_crval[0] - spec.true_shift[0]
    h_small["CRVAL2"] = true_crval[1] - spec.true_shift[1]
    fits_lite.writeto(p_small, [fits_lite.PrimaryHDU(small.astype(np.float32), h_small)], overwrite=True)
    if write_large:
        h_large = _tan_header(spec.large_n, spec.large_n, spec.large_cdelt, (0.0, 0.0), spec.crota, spec,
                              "SOLO/EUI/FSI", "FSI")
        blur = gaussian_filter(sky, 0.5 * spec.large_cdelt / spec.master_cdelt / 1.2, mode="nearest")
        large = _render(blur, w_master, TanWcs.from_header(h_large), spec.large_n, spec.large_n, 50.0)
        rng = np.random.default_rng(spec.noise_seed)
        large = large + rng.normal(0.0, 2.0, large.shape)
        fits_lite.writeto(p_large, [fits_lite.PrimaryHDU(large.astype(np.float32), h_large)], overwrite=True)
    return p_large, p_small, spec


def make_config1(out_dir):
    """BASELINE.json configs[0]/[1] images: HRIEUV-like 2048^2 vs FSI-174-like 3072^2."""
    return make_pair(out_dir, PairSpec(), tag="config1")


def small_spec(small_n=96, large_n=160, **kw):
    """A toy pair with the same geometry class (rotated TAN, arcsec units) for fast parity tests:
    the large image covers the small FOV plus margin."""
    small_cdelt = kw.pop("small_cdelt", 0.492 * 2048 / small_n / 8)
    large_cdelt = kw.pop("large_cdelt", small_cdelt * small_n * 2.5 / large_n)
    master_n = kw.pop("master_n", 512)
    master_cdelt = kw.pop("master_cdelt", small_cdelt * small_n * 1.6 / master_n)
    return PairSpec(small_n=small_n, large_n=large_n, small_cdelt=small_cdelt, large_cdelt=large_cdelt,
                    master_n=master_n, master_cdelt=master_cdelt, **kw)
