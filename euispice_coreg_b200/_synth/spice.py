"""Synthetic SPICE-like L2 raster + FSI-304-like imager sequence (BASELINE.json configs[2]).

The raster is a 4-axis cube [1, n_lambda, n_y, n_x] whose x columns are exposed at different times
(PC4_1 couples time to the x pixel); each imager frame of the sequence looks at the same synthetic sky with
its own small pointing jitter, so the synthetic raster built from the sequence differs from any single frame.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .._compat import fits_lite, timeutil
from .._compat.wcs import TanWcs
from . import scene


@dataclass
class SpiceSpec:
    n_x: int = 192
    n_y: int = 832
    n_lambda: int = 40
    cdelt1: float = 4.0          # arcsec per raster step
    cdelt2: float = 1.098        # arcsec per slit pixel
    step_s: float = 15.0         # seconds per raster step (the scan runs towards -x)
    crota: float = 1.5
    true_crval: tuple = (-100.0, 50.0)
    true_shift: tuple = (-23.0, 36.0)    # arcsec: mirrors the reference's SPICE expectation (test_alignment_spice.py:39-40)
    n_frames: int = 12
    cadence_s: float = 150.0
    large_n: int = 3072
    large_cdelt: float = 4.44
    master_n: int = 4096
    master_cdelt: float = 0.35
    seed: int = 304
    date_beg: str = "2022-03-17T00:20:00.000"
    nbin2: int = 1
    pxbeg2: int = 97


def _spice_header(spec: SpiceSpec, crval):
    rho = np.deg2rad(spec.crota)
    lam = spec.cdelt2 / spec.cdelt1
    h = fits_lite.Header()
    h["EXTNAME"] = "C III 977 - Peak"
    h["WCSAXES"] = 4
    h["CTYPE1"], h["CTYPE2"], h["CTYPE3"], h["CTYPE4"] = "HPLN-TAN", "HPLT-TAN", "WAVE", "UTC"
    h["CUNIT1"], h["CUNIT2"], h["CUNIT3"], h["CUNIT4"] = "arcsec", "arcsec", "nm", "s"
    h["CRPIX1"], h["CRPIX2"], h["CRPIX3"], h["CRPIX4"] = (spec.n_x + 1) / 2.0, (spec.n_y + 1) / 2.0, \
        (spec.n_lambda + 1) / 2.0, 1.0
    h["CDELT1"], h["CDELT2"], h["CDELT3"], h["CDELT4"] = spec.cdelt1, spec.cdelt2, 0.009, 1.0
    duration = spec.step_s * spec.n_x
    h["CRVAL1"], h["CRVAL2"], h["CRVAL3"], h["CRVAL4"] = float(crval[0]), float(crval[1]), 97.7, duration / 2.0
    h["PC1_1"], h["PC1_2"] = float(np.cos(rho)), float(-lam * np.sin(rho))
    h["PC2_1"], h["PC2_2"] = float(np.sin(rho) / lam), float(np.cos(rho))
    h["PC3_3"], h["PC4_4"] = 1.0, 1.0
    h["PC4_1"] = -spec.step_s
    h["CROTA"] = spec.crota
    h["LONPOLE"] = 180.0
    t0 = timeutil.to_seconds(spec.date_beg)
    h["DATE-BEG"] = spec.date_beg
    h["DATEREF"] = spec.date_beg
    h["DATE-OBS"] = spec.date_beg
    h["DATE-AVG"] = timeutil.from_seconds(t0 + duration / 2.0)
    h["DATE-END"] = timeutil.from_seconds(t0 + duration)
    h["DSUN_OBS"], h["RSUN_REF"], h["SOLAR_B0"] = 5.7e10, 6.957e8, -2.0
    h["CRLN_OBS"], h["CRLT_OBS"] = 250.0, -2.0
    h["NBIN2"], h["PXBEG2"], h["DETECTOR"] = spec.nbin2, spec.pxbeg2, "SW"
    h["TELESCOP"], h["INSTRUME"], h["LEVEL"], h["BUNIT"] = "SOLO/SPICE", "SPICE", "L2", "W/m2/sr/nm"
    return h


def make_spice_case(out_dir, spec: SpiceSpec | None = None, tag="config3"):
    """Writes `<tag>_spice_L2.fits` and `<tag>_fsi304_<k>.fits`; returns (path_spice, [imager paths], spec)."""
    spec = spec or SpiceSpec()
    os.makedirs(out_dir, exist_ok=True)
    pspec = scene.PairSpec(master_n=spec.master_n, master_cdelt=spec.master_cdelt, true_crval=spec.true_crval,
                           seed=spec.seed, wavelnth=304, large_n=spec.large_n, large_cdelt=spec.large_cdelt)
    sky = scene.master_scene(pspec)
    h_master = scene._tan_header(spec.master_n, spec.master_n, spec.master_cdelt, spec.true_crval, 0.0, pspec, "", "")
    w_master = TanWcs.from_header(h_master)
    rng = np.random.default_rng(spec.seed + 7)
    t0 = timeutil.to_seconds(spec.date_beg)
    # imager sequence: same sky, per-frame pointing jitter, small global brightness drift
    paths = []
    from scipy.ndimage import gaussian_filter
    blur = gaussian_filter(sky, 0.5 * spec.large_cdelt / spec.master_cdelt / 1.2, mode="nearest")
    for k in range(spec.n_frames):
        jit = rng.normal(0.0, 3.0, 2)
        h = scene._tan_header(spec.large_n, spec.large_n, spec.large_cdelt, (jit[0], jit[1]), 3.0, pspec,
                              "SOLO/EUI/FSI", "FSI")
        t = t0 + k * spec.cadence_s
        h["DATE-OBS"] = h["DATE-AVG"] = h["DATE-BEG"] = timeutil.from_seconds(t)
        img = scene._render(blur, w_master, TanWcs.from_header(h), spec.large_n, spec.large_n, 50.0)
        img = img * (1.0 + 0.01 * k) + rng.normal(0.0, 2.0, img.shape)
        p = os.path.join(out_dir, f"{tag}_fsi304_{k:02d}.fits")
        fits_lite.writeto(p, [fits_lite.PrimaryHDU(img.astype(np.float32), h)], overwrite=True)
        paths.append(p)
    # SPICE raster at its true pointing; header written with the pointing error
    h_true = _spice_header(spec, spec.true_crval)
    from .._compat.wcs import SpiceWcs
    w_true = SpiceWcs(h_true).celestial()
    img = scene._render(gaussian_filter(sky, 1.2 / spec.master_cdelt, mode="nearest"), w_master, w_true,
                        spec.n_x, spec.n_y, 50.0)
    img = img + rng.normal(0.0, 4.0, img.shape)
    lam = np.arange(spec.n_lambda) - (spec.n_lambda - 1) / 2.0
    profile = np.exp(-0.5 * (lam / 3.0) ** 2)
    profile /= profile.sum()
    cube = (img[None, :, :] * profile[:, None, None] + 0.2).astype(np.float32)[None]
    h_spice = h_true.copy()
    h_spice["CRVAL1"] = spec.true_crval[0] - spec.true_shift[0]
    h_spice["CRVAL2"] = spec.true_crval[1] - spec.true_shift[1]
    p_spice = os.path.join(out_dir, f"solo_L2_spice-n-ras_{tag}.fits")
    fits_lite.writeto(p_spice, [fits_lite.PrimaryHDU(cube, h_spice)], overwrite=True)
    return p_spice, paths, spec


def small_spice_spec(**kw):
    """Toy version for tests: 24 x 80 raster, 10 wavelengths, 4 frames of 200^2."""
    base = dict(n_x=24, n_y=80, n_lambda=10, cdelt1=4.0, cdelt2=1.098, step_s=20.0, n_frames=4, cadence_s=150.0,
                large_n=200, large_cdelt=4.44, master_n=512, master_cdelt=0.9, true_crval=(-20.0, 10.0),
                true_shift=(-8.0, 12.0), nbin2=1, pxbeg2=472)
    base.update(kw)
    return SpiceSpec(**base)
