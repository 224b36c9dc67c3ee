"""`AlignmentSpicePixel` -- same public API as the reference's `pxlshift/alignment_spice_pixel.py`: the pixel-shift
search with a SPICE raster (L2 cube summed over its spectral window, or one fitted parameter of an L3 file) as the
small image. Host preparation only; the search is `AlignmentPixels.find_best_parameters` on the device."""
from __future__ import annotations

import numpy as np

from .._compat import units
from .._compat.wcs import SpiceWcs
from ..utils import Util
from .alignment_pixels import AlignmentPixels


class AlignmentSpicePixel(AlignmentPixels):
    def __init__(self, fsi_path: str, fsi_window: int, spice_path: str, spice_window: int, index_amplitude=None):
        """Same parameters as `pxlshift/alignment_spice_pixel.py:10-23`."""
        super().__init__(fsi_path, fsi_window, spice_path, spice_window)
        self.fsi_path = fsi_path
        self.spice_path = spice_path
        self.fsi_window = fsi_window
        self.spice_window = spice_window
        level = None
        if "L2" in self.spice_path:
            level = 2
        elif "L3" in self.spice_path:
            level = 3
        self._extract_spice_data_header(level=level, index_amplitude=index_amplitude)

    def find_best_parameters(self, lag_dx: np.array, lag_dy: np.array, lag_drot: np.array, unit_rot="degree",
                             shift_solar_rotation_dx_large=False):
        return super().find_best_parameters(lag_dx, lag_dy, lag_drot, unit_rot, shift_solar_rotation_dx_large)

    def _extract_spice_data_header(self, level: int, index_amplitude=None):
        """`pxlshift/alignment_spice_pixel.py:30-46`."""
        with Util._fits().open(self.spice_path) as hdul_small:
            hdu = hdul_small[self.spice_window]
            dt = hdu.header.copy()["PC4_1"]
            if level == 2:
                self._prepare_spice_from_l2(hdu)
            elif level == 3:
                self._prepare_spice_from_l3(hdu, index_amplitude)
            # any other level leaves data_small / hdr_small as `AlignmentPixels.__init__` read them (like the reference)
            for k in ("SOLAR_B0", "RSUN_REF", "DSUN_OBS"):
                self.hdr_small[k] = hdu.header[k]
            self._correct_solar_rotation(dt)

    def _correct_solar_rotation(self, dt):
        """`pxlshift/alignment_spice_pixel.py:48-63`: effective CDELT1 of a raster that steps against solar rotation."""
        b0 = np.deg2rad(self.hdr_small['SOLAR_B0'])
        band = self.hdr_large['WAVELNTH']
        omega_car = np.deg2rad(360 / 25.38 / 86400)
        if band == 174:
            band = 171
        omega = omega_car + Util.diff_rot(b0, f'EIT {band}')
        rsun = self.hdr_small['RSUN_REF']
        dsun = self.hdr_small['DSUN_OBS']
        phi = omega * rsun / (dsun - rsun)
        phi = np.rad2deg(phi) * 3600                     # arcsec / s
        unit = self.hdr_small['CUNIT1']
        dtx_old = float(self.hdr_small['CDELT1'])
        # Quantity(CDELT1, CUNIT1) - dt * phi * u.arcsec: astropy converts the right operand to the left one's unit
        dtx_new = dtx_old - float(units.convert(dt * phi, "arcsec", unit))
        self.hdr_small['CDELT1'] = dtx_new
        print(f'Corrected solar rotation : changed SPICE CDELT1 from {dtx_old} {unit} to {dtx_new} {unit}')

    def _prepare_spice_from_l2(self, hdu):
        """`pxlshift/alignment_spice_pixel.py:65-87`: spectral sum of the rows inside the slit, 2-D celestial header.
        (The reference also calls `AlignSpiceUtil.recenter_crpix_in_header_L2` at :68 and
        `AlignEUIUtil.recenter_crpix_in_header` at :77; both are `pass` in v0.4.0 -- `utils/Util.py:348-349, 565-566`,
        bodies commented out -- so there is nothing to restate.)"""
        data_small = np.array(hdu.data.copy(), dtype=np.float64)
        header_spice = hdu.header.copy()
        ymin, ymax = Util.AlignSpiceUtil.vertical_edges_limits(header_spice)
        self.hdr_small = SpiceWcs(header_spice).xy_header().copy()
        ylen = data_small.shape[2]
        ylim = np.array([ymin, ylen - ymax - 1]).max()
        self.data_small = np.nansum(data_small[0, :, ylim:(ylen - ylim), :], axis=0)
        self.hdr_small["CRPIX1"] = (self.data_small.shape[1] + 1) / 2
        self.hdr_small["CRPIX2"] = (self.data_small.shape[0] + 1) / 2
        self.hdr_small["NAXIS1"] = self.data_small.shape[1]
        self.hdr_small["NAXIS2"] = self.data_small.shape[0]

    def _prepare_spice_from_l3(self, hdu, index_amplitude):
        """`pxlshift/alignment_spice_pixel.py:89-101`: one fitted parameter of an L3 file, ANA_MISS -> NaN."""
        data_small = np.array(hdu.data.copy(), dtype=np.float64)
        self.data_small = data_small[:, :, index_amplitude]
        self.data_small[self.data_small == hdu.header["ANA_MISS"]] = np.nan
        self.hdr_small = SpiceWcs(hdu.header.copy()).xy_header().copy()
        self.hdr_small["NAXIS1"] = self.data_small.shape[1]
        self.hdr_small["NAXIS2"] = self.data_small.shape[0]
