"""Synthetic-raster builder: same API as the reference's `synras/map_builder.py`, device gather underneath.

The reference loops over the raster's slit positions; per position it opens the imager file closest in time,
builds its WCS, converts the slit's sky coordinates to that frame's pixels and interpolates
(`synras/map_builder.py:95-131`). Here the time bookkeeping stays on the host (it is a few hundred scalars),
the frames that are actually used are stacked once in HBM, and one kernel (`coreg_synras_build`, K6) produces
the whole raster: per output pixel it picks its column's frame, maps the sky point through that frame's TAN WCS
and takes the order-2 sample. The slit's sky coordinates come from the device pixel->world kernel (K3).
"""
from __future__ import annotations

import os
import random
import warnings

import numpy as np

from .. import _ext
from .._compat import timeutil, units
from .._compat.wcs import SpiceWcs, TanWcs
from ..utils.Util import _fits

_MAX_FRAMES_PER_CALL = 64


def _seconds(q):
    """`u.Quantity(..).to('s').value` for plain numbers (seconds) or quantity-like objects."""
    if hasattr(q, "to"):
        return float(q.to("s").value)
    if hasattr(q, "value") and hasattr(q, "unit"):
        scale = {"s": 1.0, "min": 60.0, "h": 3600.0, "d": 86400.0}[str(q.unit)]
        return float(q.value) * scale
    return float(q)


class MapBuilder:

    def __init__(self):
        pass

    def process(self, path_output: str):
        pass


class ComposedMapBuilder(MapBuilder):

    def __init__(self, path_to_spectro: str, list_imager_paths: list, threshold_time, window_imager=-1,
                 window_spectro=0):
        """Same parameters as `synras/map_builder.py:26-37`; `threshold_time` is seconds (number) or an
        astropy-like time quantity."""
        super().__init__()
        self.path_to_spectro = path_to_spectro
        self.list_imager_paths = np.array(list_imager_paths, dtype="str")
        self.window_imager = window_imager
        self.window_spectro = window_spectro
        self.threshold_time = threshold_time
        self.path_composed_map = None
        self._extract_imager_metadata()
        self.path_output = None
        self.use_sunpy = False
        self.order = 2
        self.use_windows = True     # read / upload only the part of each imager frame the raster can reach

    # ------------------------------------------------------------------------------------------ public
    def process(self, folder_path_output=None, basename_output=None, print_filename=True, level=2,
                keep_original_imager_pixel_size=False, return_synras_name=False):
        """`synras/map_builder.py:57-79`."""
        self.path_output = folder_path_output
        with _fits().open(self.path_to_spectro) as hdul_spice:
            hdr_spice = hdul_spice[self.window_spectro].header.copy()
        name = self._create_map_from_hdu(hdr_spice, basename_output, folder_path_output,
                                         print_filename=print_filename, level=level,
                                         keep_original_imager_pixel_size=keep_original_imager_pixel_size)
        if return_synras_name:
            return name

    def process_from_header(self, hdr_spice, path_output=None, basename_output=None, print_filename=False, level=2,
                            keep_original_imager_pixel_size=False):
        self.path_output = path_output
        self._create_map_from_hdu(hdr_spice, basename_output, path_output, print_filename=print_filename,
                                  level=level, keep_original_imager_pixel_size=keep_original_imager_pixel_size)

    def get_path_to_composed_map(self):
        return self.path_composed_map

    # ------------------------------------------------------------------------------------------ host bookkeeping
    def _extract_imager_metadata(self):
        """DATE-AVG (seconds) and header of every imager file (`synras/map_builder.py:218-225`)."""
        n = len(self.list_imager_paths)
        self.dates = np.empty(n, dtype=np.float64)
        self.headers = np.empty(n, dtype="object")
        for ii, path in enumerate(self.list_imager_paths):
            with _fits().open(path) as hdul:
                hdr = hdul[self.window_imager].header.copy()
            self.dates[ii] = timeutil.to_seconds(hdr["DATE-AVG"])
            self.headers[ii] = hdr

    def _find_closest_imager_time(self, utc_ref):
        delta = np.abs(utc_ref - self.dates)
        return int(delta.argmin()), float(delta.min())

    @staticmethod
    def _return_mean_time(utc_list):
        """`synras/map_builder.py:231-237`: utc_ref - mean(utc_ref - utc_i), evaluated in that order."""
        utc_list = np.asarray(utc_list, dtype=np.float64)
        utc_ref = utc_list[0]
        delta = utc_ref - utc_list
        return utc_ref - delta.mean(), delta

    def _prepare_spectro_data(self, hdr_spice, keep_original_imager_pixel_size, level):
        raise NotImplementedError

    # ------------------------------------------------------------------------------------------ the build
    def _create_map_from_hdu(self, hdr_spice, basename_output=None, path_output=None, print_filename=True, level=2,
                             keep_original_imager_pixel_size=False):
        import torch
        hdr_im, lng_dev, lat_dev, naxis1, naxis2, naxis_long, utc_cols, spice_wcs = \
            self._prepare_spectro_data(hdr_spice, keep_original_imager_pixel_size, level)
        threshold = _seconds(self.threshold_time)
        frame_of_col = np.empty(naxis_long, dtype=np.int64)
        self.dates_selected = np.empty(naxis_long, dtype=np.float64)
        for ii in range(naxis_long):
            utc_slit, _ = self._return_mean_time(utc_cols[:, ii])
            index_closest, dt = self._find_closest_imager_time(utc_slit)
            self.dates_selected[ii] = self.dates[index_closest]
            if dt > threshold:
                raise ValueError(f"dt={dt} s: Could not find imager sufficiently close in time")
            frame_of_col[ii] = index_closest
        used = np.unique(frame_of_col)
        if len(used) > _MAX_FRAMES_PER_CALL:
            raise NotImplementedError("more than 64 distinct imager frames in one raster")
        # Only the part of each imager frame the raster can reach is read, converted and uploaded (a full-disc FSI frame
        # is 38 MB, a SPICE raster covers ~200 x 250 of its pixels): every frame contributes a window of one common
        # shape at its own origin, found from the raster grid's four corners (`LagSearchEngine.large_window`: a
        # projective map has its extremes there; 4 pixels of margin for the spline support). The kernel computes
        # coordinates in the full image and subtracts the integer origin exactly: the raster has the bits of the
        # whole-frame build (tests/test_gpu_spice.py).
        from ..hdrshift.engine import LagSearchEngine
        hdus, wcs_list, wins = [], [], []
        for k in used:
            if print_filename:
                print(f"\nUse imager {os.path.basename(self.list_imager_paths[k])}")
            hdu = _fits().open(self.list_imager_paths[k])[self.window_imager]
            hdus.append(hdu)
            wcs_list.append(TanWcs.from_header(hdu.header))
            shape = tuple(hdu.shape) if getattr(hdu, "shape", None) is not None else np.asarray(hdu.data).shape
            wins.append((shape, LagSearchEngine.large_window(wcs_list[-1], self._w_grid, shape)
                         if self.use_windows and len(shape) == 2 else None))
        if len({w[0] for w in wins}) != 1:
            raise NotImplementedError("imager frames of different shapes in one raster")
        origins = None
        if all(w[1] is not None for w in wins):
            (ny_f, nx_f) = wins[0][0]
            wx = min(nx_f, max(w[1][1] - w[1][0] for w in wins))
            wy = min(ny_f, max(w[1][3] - w[1][2] for w in wins))
            origins = [(min(w[1][0], nx_f - wx), min(w[1][2], ny_f - wy)) for w in wins]
            frames = [(h.read_window(y0, y0 + wy, x0, x0 + wx) if hasattr(h, "read_window")
                       else np.asarray(h.data)[y0:y0 + wy, x0:x0 + wx]) for h, (x0, y0) in zip(hdus, origins)]
        else:
            frames = [np.asarray(h.data) for h in hdus]
        dt = np.float32 if all(f.dtype == np.float32 for f in frames) else np.float64
        stack = torch.from_numpy(np.ascontiguousarray(np.stack([f.astype(dt, copy=False) for f in frames]))).cuda()
        remap = {int(k): i for i, k in enumerate(used)}
        cols = [remap[int(k)] for k in frame_of_col]
        out = _ext.synras_build(stack, wcs_list, cols, lng_dev, lat_dev, self.order, origins=origins)
        self.data_composed = out.cpu().numpy()
        list_hdr_imagers_used = [self.headers[k] for k in frame_of_col]

        keys = ["CRPIX1", "CRPIX2", "CRPIX3", "CRPIX4", "CRVAL1", "CRVAL2", "CRVAL3", "CRVAL4",
                "CDELT1", "CDELT2", "CDELT3", "CDELT4", "CUNIT1", "CUNIT2", "CUNIT3", "CUNIT4", "CROTA2", "CROTA"]
        keys += [f"PC{i}_{j}" for i in (1, 2, 3, 4) for j in (1, 2, 3, 4)]
        self.hdr_composed = list_hdr_imagers_used[len(list_hdr_imagers_used) // 2].copy()
        for k in keys:
            if k in self.hdr_spice_:
                self.hdr_composed[k] = self.hdr_spice_[k]
            else:
                warnings.warn(f"{k} no in original header. It is not added to the synthetic raster header")
        self.hdr_composed["DATE-AVG"] = hdr_spice["DATE-AVG"]
        self.hdr_composed["DATE-OBS"] = hdr_spice["DATE-OBS"]
        self.hdr_composed["DATE-BEG"] = hdr_spice["DATE-BEG"]
        self.hdr_composed["SPECPATH"] = os.path.basename(self.path_to_spectro)
        utc_composed, _ = self._return_mean_time(self.dates_selected)
        wave = self.hdr_composed["WAVELNTH"]
        if "DETECTOR" in self.hdr_composed:
            detector = self.hdr_composed["DETECTOR"]
        elif "INSTRUME" in self.hdr_composed:
            detector = self.hdr_composed["INSTRUME"]
        else:
            raise ValueError("No info on reference instrument")
        if keep_original_imager_pixel_size:
            # `synras/map_builder.py:165-192`: the composed header gets the imager's pixel size, a PCi_j rebuilt for the
            # new CDELT ratio, and its reference pixel at the centre of the new grid (KeyError, like the reference, when
            # the raster's PCi_j is the identity and `to_header()` left it out)
            w_xy = spice_wcs.celestial()
            x_mid = (naxis1 - 1) / 2
            y_mid = (naxis2 - 1) / 2
            lon_mid, lat_mid = w_xy.pixel_to_world(np.array([x_mid]), np.array([y_mid]))
            h = self.hdr_composed
            h["CDELT1"] = float(units.convert(hdr_im["CDELT1"], hdr_im["CUNIT1"], h["CUNIT1"]))
            h["CDELT2"] = float(units.convert(hdr_im["CDELT2"], hdr_im["CUNIT2"], h["CUNIT2"]))
            lam = h["CDELT2"] / h["CDELT1"]
            rho = np.arccos(h["PC1_1"])
            rho = rho * (-np.sign(h["PC1_2"]))
            h["PC1_2"] = float(-lam * np.sin(rho))
            h["PC2_1"] = float((1 / lam) * np.sin(rho))
            h["CRPIX1"] = (self.data_composed.shape[1] + 1) / 2
            h["CRPIX2"] = (self.data_composed.shape[0] + 1) / 2
            h["CRVAL1"] = float(units.convert(lon_mid[0], "deg", h["CUNIT1"]))
            h["CRVAL2"] = float(units.convert(lat_mid[0], "deg", h["CUNIT2"]))
        if basename_output is None:
            date = timeutil.from_seconds(utc_composed)[:19].replace(":", "_")
            basename_new = f"solo_L3_{detector}{wave}-image-composed-{date}_{random.randint(1, 99999):05d}.fits"
        else:
            basename_new = basename_output
        if path_output is not None:
            fits = _fits()
            hdul = fits.HDUList([fits.PrimaryHDU(self.data_composed, header=self.hdr_composed)])
            hdul.writeto(os.path.join(self.path_output, basename_new), overwrite=True)
            self.path_composed_map = os.path.join(self.path_output, basename_new)
            return self.path_composed_map
        if level == 2:
            self.hdr_composed["NAXIS1"] = self.data_composed.shape[1]
            self.hdr_composed["NAXIS2"] = self.data_composed.shape[0]
            return None
        raise NotImplementedError


class SPICEComposedMapBuilder(ComposedMapBuilder):

    def __init__(self, path_to_spectro: str, list_imager_paths: list, threshold_time, window_imager=-1,
                 window_spectro=0):
        super().__init__(path_to_spectro=path_to_spectro, list_imager_paths=list_imager_paths,
                         threshold_time=threshold_time, window_imager=window_imager, window_spectro=window_spectro)

    def _prepare_spectro_data(self, hdr_spice, keep_original_imager_pixel_size, level):
        """`synras/map_builder.py:249-344`: sky coordinates of the output grid at the first time index (device, K3) and
        the exposure time of every column (host). Level 2: the raster's (x, y) = FITS axes (1, 2); level 3 (fitted
        files, axes: parameter, x, y, time): axes (2, 3). `keep_original_imager_pixel_size`: the grid steps through
        the raster in units of the imager's pixel, `np.arange(0, NAXIS, CDELT_imager / CDELT_raster)` -- a plain ratio
        of header values, no unit conversion, as in the reference (`:262-263, 314-315`)."""
        if level == 2:
            ax1, ax2 = 1, 2
        elif level == 3:
            ax1, ax2 = 2, 3
        else:
            raise ValueError("level must be 2 or 3")
        sw = SpiceWcs(hdr_spice)
        naxis1, naxis2 = int(hdr_spice["NAXIS%d" % ax1]), int(hdr_spice["NAXIS%d" % ax2])
        with _fits().open(self.list_imager_paths[0]) as hdul_im:
            hdr_im = hdul_im[self.window_imager].header.copy()
        w = sw.celestial()
        if keep_original_imager_pixel_size:
            r1 = hdr_im["CDELT1"] / hdr_spice["CDELT%d" % ax1]
            r2 = hdr_im["CDELT2"] / hdr_spice["CDELT%d" % ax2]
            x = np.arange(0, naxis1, r1, dtype=np.float64)
            y = np.arange(0, naxis2, r2, dtype=np.float64)
            # pixel k * r of the raster == pixel k of a grid with PCi_j scaled column-wise and CRPIX moved
            w_grid = w.replace(crpix1=1.0 + (w.crpix1 - 1.0) / r1, crpix2=1.0 + (w.crpix2 - 1.0) / r2,
                               pc11=w.pc11 * r1, pc21=w.pc21 * r1, pc12=w.pc12 * r2, pc22=w.pc22 * r2)
        else:
            x = np.arange(naxis1, dtype=np.float64)
            y = np.arange(naxis2, dtype=np.float64)
            w_grid = w
        lng, lat = _ext.tan_pix2world(w_grid, len(x), len(y), wrap_pipi=False)
        self._w_grid = w_grid.replace(naxis1=len(x), naxis2=len(y))     # the raster's grid: bounds the imager windows
        t_ref = timeutil.to_seconds(hdr_spice.get("DATEREF", hdr_spice.get("DATE-BEG", hdr_spice["DATE-OBS"])))
        utc_cols = t_ref + sw.time_seconds(x[None, :], 0.0, y[:, None])      # [len(y), len(x)]
        self.hdr_spice_ = sw.xy_header()
        return hdr_im, lng, lat, naxis1, naxis2, len(x), utc_cols, sw
