"""`AlignmentPixels` -- same public API as the reference's `pxlshift/alignment_pixels.py`, device lag loop underneath.

The reference brings the large image to the small image's pixel size, rotates the small image per rotation lag and
then runs a Python double loop over (dx, dy) that slices, masks, compacts and calls a numba Pearson per lag
(`pxlshift/alignment_pixels.py:35-84`). Here the one-shot resamplings go through the device `map_coordinates`
(bit-exact against scipy), both images stay resident in HBM, and one kernel launch (`coreg_pixel_shift_corr`: tiles of
the small image in registers, the reachable window of the large image staged in shared memory) evaluates every
(dx, dy, rotation) lag. There is no CPU fallback.
"""
from __future__ import annotations

import warnings

import numpy as np

from .. import _ext
from .._compat import timeutil, units
from ..utils import Util
from ..utils.matrix_transform import MatrixTransform


def _torch():
    import torch
    return torch


class AlignmentPixels:

    def __init__(self, large_fov_known_pointing: str, window_large: int, small_fov_to_correct: str,
                 window_small: int):
        """Same parameters as `pxlshift/alignment_pixels.py:16-31`."""
        fits = Util._fits()
        with fits.open(large_fov_known_pointing) as hdul_large:
            hdu_large = hdul_large[window_large]
            self.hdr_large = hdu_large.header.copy()
            self.data_large = np.array(hdu_large.data.copy(), dtype=np.float64)
        with fits.open(small_fov_to_correct) as hdul_small:
            hdu_small = hdul_small[window_small]
            self.hdr_small = hdu_small.header.copy()
            self.data_small = np.array(hdu_small.data.copy(), dtype=np.float64)
        self.slc_small_ref = None
        self.x_large = None
        self.y_large = None
        self.nvalid = None

    # ------------------------------------------------------------------------------------------------ device helpers
    @staticmethod
    def _device():
        torch = _torch()
        _ext.load()  # fail loudly when the CUDA library is missing
        if not torch.cuda.is_available():
            raise _ext.CoregLibraryError("no CUDA device: the pixel-shift search has no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    @staticmethod
    def _resample(img_dev, x, y, fill):
        """`interpol2d(img, x, y, order=1, fill)` with the fill value turned into NaN afterwards, on the device;
        x, y: host or device float64 planes."""
        torch = _torch()
        xd = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(img_dev.device)
        yd = y if torch.is_tensor(y) else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(img_dev.device)
        out = _ext.map_coordinates(img_dev, yd.contiguous(), xd.contiguous(), 1, float(fill), torch.float64)
        return torch.where(out == float(fill), torch.full_like(out, float("nan")), out)

    # ------------------------------------------------------------------------------------------------ the search
    def find_best_parameters(self, lag_dx: np.array, lag_dy: np.array, lag_drot: np.array, unit_rot="degree",
                             shift_solar_rotation_dx_large=False):
        """corr [len(lag_dx), len(lag_dy), len(lag_drot)] float64 (`pxlshift/alignment_pixels.py:60-84`)."""
        torch = _torch()
        dev = self._device()
        for name, lag in (("lag_dx", lag_dx), ("lag_dy", lag_dy)):
            arr = np.asarray(lag)
            if arr.size and not np.all(arr == np.round(arr)):
                raise TypeError(f"{name} must hold integer pixel shifts (they index slices, "
                                "pxlshift/alignment_pixels.py:42-46)")
        self.lag_dx = lag_dx
        self.lag_dy = lag_dy
        self.lag_drot = lag_drot
        self.unit_rot = unit_rot
        with torch.cuda.device(dev):
            d_large = torch.from_numpy(np.ascontiguousarray(self.data_large)).to(dev)
            if shift_solar_rotation_dx_large:
                d_large = self._shift_large_fov(d_large)
            d_large = self._sub_resolution_large_fov(d_large)
            self._initialise_slice_corresponding_to_small(tuple(d_large.shape))
            ly, lx = d_large.shape
            sy, sx = self.data_small.shape
            y0, x0 = self.slc_small_ref[0].start, self.slc_small_ref[1].start
            # `_check_boundaries` of every lag (the reference raises at the first offending one, :47)
            for dx in (int(v) for v in np.asarray(lag_dx)):
                for dy in (int(v) for v in np.asarray(lag_dy)):
                    self._check_boundaries((slice(y0 + dy, y0 + sy + dy), slice(x0 + dx, x0 + sx + dx)), (ly, lx))
            d_small = torch.from_numpy(np.ascontiguousarray(self.data_small)).to(dev)
            rotated = []
            for drot in lag_drot:
                if drot != 0:
                    xx, yy = np.meshgrid(np.arange(sx), np.arange(sy))
                    nx, ny = MatrixTransform.polar_transform(xx, yy, theta=drot, units=self.unit_rot)
                    rotated.append(self._resample(d_small, nx, ny, -32762))
                else:
                    rotated.append(d_small.clone())
            smalls = torch.stack(rotated).contiguous()
            pivots = torch.zeros(2, dtype=torch.float64, device=dev)
            _ext.finite_mean(d_large, pivots[0:1])
            _ext.finite_mean(smalls, pivots[1:2])
            corr, nvalid = _ext.pixel_shift_corr(d_large, smalls, x0, y0, np.asarray(lag_dx).astype(np.int64),
                                                 np.asarray(lag_dy).astype(np.int64), pivots, return_nvalid=True)
            self.data_large = d_large.cpu().numpy()
            self.data_small_rotated = smalls[-1].cpu().numpy() if len(rotated) else self.data_small.copy()
            self.nvalid = nvalid.cpu().numpy()
            return corr.cpu().numpy()

    # ------------------------------------------------------------------------------------------------ one-shot steps
    def _shift_large_fov(self, d_large):
        """`pxlshift/alignment_pixels.py:86-108`: the large image displaced along the rotated x axis by the solar
        rotation between the two exposures (two order-1 resamplings, like the reference)."""
        torch = _torch()
        ly, lx = d_large.shape
        yy, xx = torch.meshgrid(torch.arange(ly, dtype=torch.float64, device=d_large.device),
                                torch.arange(lx, dtype=torch.float64, device=d_large.device), indexing="ij")
        data_large = self._resample(d_large, xx, yy, -32762)
        dcrval = self._return_shift_large_fov_solar_rotation()    # arcsec
        h = self.hdr_large
        if "CROTA" in h:
            warnings.warn("CROTA must be in degree", Warning)
            theta = np.deg2rad(h["CROTA"])
            dx = (units.convert(dcrval, "arcsec", h["CUNIT1"]) / h["CDELT1"]) * np.cos(-theta)
            dy = (units.convert(dcrval, "arcsec", h["CUNIT2"]) / h["CDELT2"]) * np.sin(-theta)
        else:
            dx = units.convert(dcrval, "arcsec", h["CUNIT1"]) / h["CDELT1"]
            dy = 0
        # linear_transform with displacement_matrix(dx, dy): nx = xx + dx, ny = yy + dy
        data_large = self._resample(data_large, xx + float(dx), yy + float(dy), -32762)
        print(f"corrected solar rotation on FSI on CRVAL1: {dx=}, {dy=}")
        return data_large

    def _return_shift_large_fov_solar_rotation(self):
        """`pxlshift/alignment_pixels.py:110-124`, arcsec."""
        band = self.hdr_large['WAVELNTH']
        b0 = np.deg2rad(self.hdr_large['SOLAR_B0'])
        omega_car = np.deg2rad(360 / 25.38 / 86400)
        if band == 174:
            band = 171
        omega = omega_car + Util.diff_rot(b0, f'EIT {band}')
        rsun = self.hdr_large['RSUN_REF']
        dsun = self.hdr_large['DSUN_OBS']
        phi = omega * rsun / (dsun - rsun)
        phi = np.rad2deg(phi) * 3600
        dt = timeutil.to_seconds(self.hdr_small["DATE-AVG"]) - timeutil.to_seconds(self.hdr_large["DATE-AVG"])
        return dt * phi

    def _sub_resolution_large_fov(self, d_large):
        """`pxlshift/alignment_pixels.py:126-143`: the large image sampled (order 1) every `ratio` of its pixels."""
        torch = _torch()
        hs, hl = self.hdr_small, self.hdr_large
        cdelt1_conv = units.convert(hs["CDELT1"], hs["CUNIT1"], hl["CUNIT1"])
        cdelt2_conv = units.convert(hs["CDELT2"], hs["CUNIT2"], hl["CUNIT2"])
        self.ratio_res_1 = cdelt1_conv / hl["CDELT1"]
        self.ratio_res_2 = cdelt2_conv / hl["CDELT2"]
        ly, lx = d_large.shape
        xv = torch.from_numpy(np.arange(0, lx, self.ratio_res_1, dtype=np.float64)).to(d_large.device)
        yv = torch.from_numpy(np.arange(0, ly, self.ratio_res_2, dtype=np.float64)).to(d_large.device)
        x = xv[None, :].expand(yv.numel(), xv.numel())
        y = yv[:, None].expand(yv.numel(), xv.numel())
        return self._resample(d_large, x, y, -32768)

    def _initialise_slice_corresponding_to_small(self, shape_large=None):
        shape_large = self._large_shape_after_sub_resolution() if shape_large is None else shape_large
        lo = [int((shape_large[n] - self.data_small.shape[n] - 1) / 2) for n in range(2)]
        self.slc_small_ref = (slice(lo[0], lo[0] + self.data_small.shape[0]),
                              slice(lo[1], lo[1] + self.data_small.shape[1]))

    def _large_shape_after_sub_resolution(self):
        return (len(np.arange(0, self.data_large.shape[0], self.ratio_res_2)),
                len(np.arange(0, self.data_large.shape[1], self.ratio_res_1)))

    @staticmethod
    def _check_boundaries(slc, shape):
        for n in range(2):
            if slc[n].start < 0:
                raise ValueError("too large shift : outside FSI")
            if slc[n].stop > shape[n]:
                raise ValueError("too large shift : outside FSI")
