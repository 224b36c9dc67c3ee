"""FITS header -> gnomonic (TAN) WCS constants, and a small host-side pixel<->world evaluator.

Stands in for `astropy.wcs.WCS(hdr)` on the pointing-search path. Reference call sites:
`hdrshift/alignment.py:1041-1065` (`WCS(hdr)`, `world_to_pixel`), `utils/Util.py:283-290`
(`WCS(hdr)`, `pixel_to_world`), `synras/map_builder.py:119-127`.

Only what that path needs is modelled: two celestial axes with a `-TAN` projection,
CUNIT in {deg, arcsec, arcmin, rad}, PCi_j (or CROTA/CROTA2 in the AIPS convention, or CDi_j),
LONPOLE. wcslib rescales CRVAL/CDELT to degrees when the WCS is set up; so does `TanWcs`.

The *per-pixel* work is done on the device (csrc/coreg_wcs.cu, `coreg_tan_pix2world`,
`coreg_tan_world2pix`). The numpy evaluator at the bottom of this file is for a handful of points
(synras bookkeeping, FOV limits) and is written in closed form; the wcslib-structured restatement
used as the checker lives in `oracle/`, not here.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import ClassVar

import numpy as np

from . import units

R2D = 180.0 / math.pi
D2R = math.pi / 180.0


def _axis_unit(hdr, i):
    unit = hdr.get("CUNIT%d" % i, "deg") if hasattr(hdr, "get") else hdr["CUNIT%d" % i]
    if unit is None or str(unit).strip() == "":
        unit = "deg"
    return units.canon(unit)


def celestial_axes(hdr, lon_prefixes=("HPLN", "RA--", "CRLN", "GLON", "ELON", "SOLX"),
                   naxis=None):
    """Return the 1-based indices (ilon, ilat) of the celestial pair in `hdr`."""
    n = int(naxis if naxis is not None else hdr.get("WCSAXES", hdr.get("NAXIS", 2)))
    ilon = ilat = None
    for i in range(1, max(n, 2) + 1):
        key = "CTYPE%d" % i
        if key not in hdr:
            continue
        ct = str(hdr[key]).upper()
        if ct[:4] in lon_prefixes or ct[1:4] == "LON":
            ilon = i
        elif ct[:4] in ("HPLT", "DEC-", "CRLT", "GLAT", "ELAT", "SOLY") or ct[1:4] == "LAT":
            ilat = i
    if ilon is None or ilat is None:
        raise ValueError("no celestial axis pair found in header")
    return ilon, ilat


@dataclass(frozen=True)
class TanWcs:
    """Constants of a 2-D TAN WCS, all angles in degrees (as wcslib holds them after wcsset)."""
    crpix1: float
    crpix2: float
    cdelt1: float
    cdelt2: float
    pc11: float
    pc12: float
    pc21: float
    pc22: float
    crval1: float
    crval2: float
    lonpole: float
    naxis1: int
    naxis2: int
    unit_scale1: float = 1.0  # CUNIT1 -> deg factor that was applied
    unit_scale2: float = 1.0

    _PROJ: ClassVar[str] = "TAN"

    @staticmethod
    def _default_lonpole(crval2):
        """wcslib celset: LONPOLE defaults to 0 when CRVAL2 >= theta0 (90 deg for zenithal projections), else 180."""
        return 0.0 if crval2 >= 90.0 else 180.0

    # -- construction ----------------------------------------------------
    @classmethod
    def from_header(cls, hdr, ilon: int | None = None, ilat: int | None = None):
        if ilon is None or ilat is None:
            ilon, ilat = celestial_axes(hdr)
        for i in (ilon, ilat):
            ct = str(hdr["CTYPE%d" % i]).upper()
            if not ct.endswith(cls._PROJ):
                raise NotImplementedError(
                    f"CTYPE{i}={ct!r}: a -{cls._PROJ} projection is expected here (the device path handles -TAN and, "
                    "for align_using_initial_carrington, -CAR)")
        s1 = units.factor(_axis_unit(hdr, ilon), "deg")
        s2 = units.factor(_axis_unit(hdr, ilat), "deg")
        cdelt1 = float(hdr.get("CDELT%d" % ilon, 1.0))
        cdelt2 = float(hdr.get("CDELT%d" % ilat, 1.0))
        k = lambda a, b: "PC%d_%d" % (a, b)  # noqa: E731
        c = lambda a, b: "CD%d_%d" % (a, b)  # noqa: E731
        if any(k(a, b) in hdr for a in (ilon, ilat) for b in (ilon, ilat)):
            pc11 = float(hdr.get(k(ilon, ilon), 1.0))
            pc12 = float(hdr.get(k(ilon, ilat), 0.0))
            pc21 = float(hdr.get(k(ilat, ilon), 0.0))
            pc22 = float(hdr.get(k(ilat, ilat), 1.0))
        elif any(c(a, b) in hdr for a in (ilon, ilat) for b in (ilon, ilat)):
            # CDi_j form: wcslib sets PC=CD and CDELT=1
            pc11 = float(hdr.get(c(ilon, ilon), 0.0))
            pc12 = float(hdr.get(c(ilon, ilat), 0.0))
            pc21 = float(hdr.get(c(ilat, ilon), 0.0))
            pc22 = float(hdr.get(c(ilat, ilat), 0.0))
            cdelt1 = cdelt2 = 1.0
        else:
            rot = None
            for key in ("CROTA%d" % ilat, "CROTA"):
                if key in hdr:
                    rot = float(hdr[key])
                    break
            if rot is None or rot == 0.0:
                pc11, pc12, pc21, pc22 = 1.0, 0.0, 0.0, 1.0
            else:
                # AIPS convention as translated by wcslib
                cr, sr = math.cos(rot * D2R), math.sin(rot * D2R)
                pc11, pc22 = cr, cr
                pc12 = -sr * cdelt2 / cdelt1
                pc21 = sr * cdelt1 / cdelt2
        crval1 = float(hdr.get("CRVAL%d" % ilon, 0.0)) * s1
        crval2 = float(hdr.get("CRVAL%d" % ilat, 0.0)) * s2
        if "LONPOLE" in hdr:
            lonpole = float(hdr["LONPOLE"])
        else:
            lonpole = cls._default_lonpole(crval2)
        nax1 = hdr.get("ZNAXIS%d" % ilon, hdr.get("NAXIS%d" % ilon, 0))
        nax2 = hdr.get("ZNAXIS%d" % ilat, hdr.get("NAXIS%d" % ilat, 0))
        return cls(crpix1=float(hdr.get("CRPIX%d" % ilon, 0.0)), crpix2=float(hdr.get("CRPIX%d" % ilat, 0.0)),
                   cdelt1=cdelt1 * s1, cdelt2=cdelt2 * s2,
                   pc11=pc11, pc12=pc12, pc21=pc21, pc22=pc22,
                   crval1=crval1, crval2=crval2, lonpole=lonpole,
                   naxis1=int(nax1), naxis2=int(nax2), unit_scale1=s1, unit_scale2=s2)

    def replace(self, **kw):
        return replace(self, **kw)

    # -- derived constants --------------------------------------------------
    def forward_matrix(self):
        """2x2 matrix taking pixel offsets (p - CRPIX) to projection-plane degrees."""
        return np.array([[self.cdelt1 * self.pc11, self.cdelt1 * self.pc12],
                         [self.cdelt2 * self.pc21, self.cdelt2 * self.pc22]], dtype=np.float64)

    def inverse_matrix(self):
        """2x2 matrix taking projection-plane degrees to pixel offsets."""
        m = self.forward_matrix()
        det = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
        if det == 0.0 or not np.isfinite(det):
            raise ValueError("singular PCi_j/CDELT matrix")
        return np.array([[m[1, 1], -m[0, 1]], [-m[1, 0], m[0, 0]]], dtype=np.float64) / det

    def as_array(self):
        """Packed `CoregTanWcs` (include/coreg_b200.h): 11 doubles."""
        return np.array([self.crpix1, self.crpix2, self.cdelt1, self.cdelt2,
                         self.pc11, self.pc12, self.pc21, self.pc22,
                         self.crval1, self.crval2, self.lonpole], dtype=np.float64)

    # -- small-N host evaluation (closed form) ------------------------------------
    def pixel_to_world(self, x, y):
        """0-based pixel -> (lon, lat) in degrees; lon normalised like wcslib
        ([0,360) when CRVAL1 >= 0, (-360,0] otherwise)."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        u1 = x + 1.0 - self.crpix1
        u2 = y + 1.0 - self.crpix2
        m = self.forward_matrix()
        px = (m[0, 0] * u1 + m[0, 1] * u2) * D2R
        py = (m[1, 0] * u1 + m[1, 1] * u2) * D2R
        # unit vector in the native frame (pole at the reference point), rotated by LONPOLE
        r = np.hypot(px, py)
        phi = np.arctan2(px, -py) - self.lonpole * D2R
        ct = r / np.sqrt(1.0 + r * r)
        st = 1.0 / np.sqrt(1.0 + r * r)
        s0, c0 = math.sin(self.crval2 * D2R), math.cos(self.crval2 * D2R)
        xx = st * c0 - ct * s0 * np.cos(phi)
        yy = -ct * np.sin(phi)
        zz = st * s0 + ct * c0 * np.cos(phi)
        lon = self.crval1 + np.arctan2(yy, xx) * R2D
        if self.crval1 >= 0.0:
            lon = np.where(lon < 0.0, lon + 360.0, lon)
        else:
            lon = np.where(lon > 0.0, lon - 360.0, lon)
        lat = np.arctan2(zz, np.hypot(xx, yy)) * R2D
        return lon, lat

    def world_to_pixel(self, lon, lat):
        """(lon, lat) in degrees -> 0-based pixel; NaN behind the tangent hemisphere."""
        lon = np.asarray(lon, dtype=np.float64) * D2R
        lat = np.asarray(lat, dtype=np.float64) * D2R
        a0, d0 = self.crval1 * D2R, self.crval2 * D2R
        da = lon - a0
        sl, cl = np.sin(lat), np.cos(lat)
        s0, c0 = math.sin(d0), math.cos(d0)
        den = sl * s0 + cl * c0 * np.cos(da)
        xs = sl * c0 - cl * s0 * np.cos(da)   # native x before LONPOLE
        ys = -cl * np.sin(da)
        phi = self.lonpole * D2R + np.arctan2(ys, xs)
        rr = np.hypot(xs, ys) / den * R2D
        px = rr * np.sin(phi)
        py = -rr * np.cos(phi)
        mi = self.inverse_matrix()
        x = mi[0, 0] * px + mi[0, 1] * py + (self.crpix1 - 1.0)
        y = mi[1, 0] * px + mi[1, 1] * py + (self.crpix2 - 1.0)
        bad = den <= 0.0
        return np.where(bad, np.nan, x), np.where(bad, np.nan, y)


# ------------------------------------------------------------------------------------------------------
# plate-carree (-CAR) headers: Carrington maps as inputs (`Alignment.align_using_initial_carrington`)
# ------------------------------------------------------------------------------------------------------
def _sind(a):
    a = np.asarray(a, dtype=np.float64)
    q = a / 90.0
    exact = np.array([0.0, 1.0, 0.0, -1.0])[np.mod(np.round(q), 4).astype(np.int64)]
    return np.where(q == np.round(q), exact, np.sin(a * D2R))


def _cosd(a):
    return _sind(np.asarray(a, dtype=np.float64) + 90.0)


def car_celestial_pole(lng0, lat0, lonpole, latpole=90.0, tol=1.0e-10):
    """Celestial coordinates (lng_p, lat_p) of the native pole and the native longitude phi_p of the celestial pole
    for a cylindrical projection (fiducial point (phi0, theta0) = (0, 0)) whose reference point is (lng0, lat0), as
    wcslib's `celset` derives them (FITS-WCS Paper II, eqs. 8-10). Degrees, vectorised over candidate headers.
    `lonpole` NaN = keyword absent: 0 when lat0 >= 0, else 180. Returns (lng_p, lat_p, phi_p, valid); `valid` is False
    where wcslib rejects the header (no native pole within [-90, 90])."""
    lng0, lat0, lonpole, latpole = np.broadcast_arrays(*(np.asarray(v, dtype=np.float64)
                                                         for v in (lng0, lat0, lonpole, latpole)))
    phip = np.where(np.isnan(lonpole), np.where(lat0 < 0.0, 180.0, 0.0), lonpole)
    slat0, clat0 = _sind(lat0), _cosd(lat0)
    sphip, cphip = _sind(phip), _cosd(phip)
    # theta0 = 0: x = cos(theta0) cos(phi_p - phi0) = cphip, y = sin(theta0) = 0
    z = np.abs(cphip)
    with np.errstate(divide="ignore", invalid="ignore"):
        slz = slat0 / z
    valid = np.ones(lat0.shape, dtype=bool)
    over = np.abs(slz) > 1.0
    valid &= ~(over & (np.abs(slz) - 1.0 >= tol))
    slz = np.clip(slz, -1.0, 1.0)
    u = np.where(cphip < 0.0, 180.0, 0.0)                        # atan2d(0, cphip)
    v = np.arccos(slz) * R2D
    norm = lambda a: np.where(a > 180.0, a - 360.0, np.where(a < -180.0, a + 360.0, a))   # noqa: E731
    latp1, latp2 = norm(u + v), norm(u - v)
    ok1, ok2 = np.abs(latp1) < 90.0 + tol, np.abs(latp2) < 90.0 + tol
    closer1 = np.abs(latpole - latp1) < np.abs(latpole - latp2)
    latp = np.where(ok1 & ok2, np.where(closer1, latp1, latp2), np.where(ok1, latp1, latp2))
    valid &= ok1 | ok2
    flat = z == 0.0                                              # phi_p = +-90: any pole latitude serves
    latp = np.where(flat, latpole, latp)
    valid = np.where(flat, slat0 == 0.0, valid)
    latp = np.where(np.abs(latp) > 90.0, np.sign(latp) * 90.0, latp)
    zz = _cosd(latp) * clat0
    with np.errstate(divide="ignore", invalid="ignore"):
        x = (0.0 - _sind(latp) * slat0) / zz
        y = sphip / clat0
        general = lng0 - np.arctan2(y, x) * R2D
    lngp = np.where(np.abs(zz) < tol,
                    np.where(np.abs(clat0) < tol, lng0, np.where(latp > 0.0, lng0 + phip - 180.0, lng0 - phip)),
                    general)
    lngp = np.where(lng0 >= 0.0,
                    np.where(lngp < 0.0, lngp + 360.0, np.where(lngp > 360.0, lngp - 360.0, lngp)),
                    np.where(lngp > 0.0, lngp - 360.0, np.where(lngp < -360.0, lngp + 360.0, lngp)))
    return lngp, latp, phip, valid


def car_rotation(lngp, latp, phip):
    """[n, 3, 3] rotations taking celestial unit vectors to native ones: R = Rz(phi_p) . A(lat_p) . Rz(-lng_p) with
    A = [[-sin lat_p, 0, cos lat_p], [0, -1, 0], [cos lat_p, 0, sin lat_p]] (wcslib's sphs2x written as a matrix)."""
    lngp, latp, phip = (np.atleast_1d(np.asarray(v, dtype=np.float64)) for v in (lngp, latp, phip))
    n = lngp.size

    def rz(a):
        m = np.zeros((n, 3, 3))
        m[:, 0, 0], m[:, 0, 1] = _cosd(a), -_sind(a)
        m[:, 1, 0], m[:, 1, 1] = _sind(a), _cosd(a)
        m[:, 2, 2] = 1.0
        return m

    a = np.zeros((n, 3, 3))
    a[:, 0, 0], a[:, 0, 2] = -_sind(latp), _cosd(latp)
    a[:, 1, 1] = -1.0
    a[:, 2, 0], a[:, 2, 2] = _cosd(latp), _sind(latp)
    return rz(phip) @ a @ rz(-lngp)


@dataclass(frozen=True)
class CarWcs(TanWcs):
    """Constants of a 2-D plate-carree WCS (CRLN-CAR / CRLT-CAR), degrees. `lonpole` is NaN when the header has no
    LONPOLE keyword (wcslib then picks 0 or 180 from the sign of CRVAL2, i.e. per candidate header)."""
    latpole: float = 90.0

    _PROJ: ClassVar[str] = "CAR"

    @staticmethod
    def _default_lonpole(crval2):
        return float("nan")

    @classmethod
    def from_header(cls, hdr, ilon: int | None = None, ilat: int | None = None):
        w = super().from_header(hdr, ilon, ilat)
        if "LATPOLE" in hdr:
            w = replace(w, latpole=float(hdr["LATPOLE"]))
        return w

    @staticmethod
    def lag_rows(crval1, crval2, cdelt1, cdelt2, pc11, pc12, pc21, pc22, crpix1, crpix2, lonpole, latpole=90.0):
        """[n, 16] float64 `CoregLagCar` rows (include/coreg_b200.h) of candidate headers given in degrees, and the
        mask of headers wcslib would reject (their rows are NaN)."""
        crval1 = np.atleast_1d(np.asarray(crval1, dtype=np.float64))
        n = crval1.size
        b = lambda v: np.broadcast_to(np.asarray(v, dtype=np.float64), (n,))   # noqa: E731
        lngp, latp, phip, valid = car_celestial_pole(crval1, b(crval2), b(lonpole), b(latpole))
        rot = car_rotation(lngp, latp, phip)
        f11, f12 = b(cdelt1) * b(pc11), b(cdelt1) * b(pc12)
        f21, f22 = b(cdelt2) * b(pc21), b(cdelt2) * b(pc22)
        det = f11 * f22 - f12 * f21
        tab = np.empty((n, 16), dtype=np.float64)
        tab[:, :9] = rot.reshape(n, 9)
        tab[:, 9], tab[:, 10], tab[:, 11], tab[:, 12] = f22 / det, -f12 / det, -f21 / det, f11 / det
        tab[:, 13], tab[:, 14] = b(crpix1) - 1.0, b(crpix2) - 1.0
        tab[:, 15] = lngp
        tab[~valid] = np.nan
        return tab, ~valid

    # -- small-N host evaluation (numpy; the per-pixel work is on the device) ------------------------------------
    def pixel_to_world(self, x, y):
        """0-based pixel -> (lon, lat) degrees, lon in wcslib's range (sign of the native pole's longitude)."""
        row = self.lag_row()
        r = row[:9].reshape(3, 3)
        f = np.linalg.inv(np.array([[row[9], row[10]], [row[11], row[12]]]))
        u1 = np.asarray(x, dtype=np.float64) - row[13]
        u2 = np.asarray(y, dtype=np.float64) - row[14]
        phi, theta = (f[0, 0] * u1 + f[0, 1] * u2) * D2R, (f[1, 0] * u1 + f[1, 1] * u2) * D2R
        nat = np.stack([np.cos(theta) * np.cos(phi), np.cos(theta) * np.sin(phi), np.sin(theta)])
        c = np.tensordot(r.T, nat, 1)
        lon = np.arctan2(c[1], c[0]) * R2D
        lon = np.where(lon < 0.0, lon + 360.0, lon) if row[15] >= 0.0 else np.where(lon > 0.0, lon - 360.0, lon)
        return lon, np.arctan2(c[2], np.hypot(c[0], c[1])) * R2D

    def world_to_pixel(self, lon, lat):
        row = self.lag_row()
        r = row[:9].reshape(3, 3)
        lon = np.asarray(lon, dtype=np.float64) * D2R
        lat = np.asarray(lat, dtype=np.float64) * D2R
        c = np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)])
        v = np.tensordot(r, c, 1)
        phi = np.arctan2(v[1], v[0]) * R2D
        theta = np.arctan2(v[2], np.hypot(v[0], v[1])) * R2D
        return row[9] * phi + row[10] * theta + row[13], row[11] * phi + row[12] * theta + row[14]

    def lag_row(self):
        """This header's own `CoregLagCar` row."""
        tab, bad = self.lag_rows(self.crval1, self.crval2, self.cdelt1, self.cdelt2, self.pc11, self.pc12, self.pc21,
                                 self.pc22, self.crpix1, self.crpix2, self.lonpole, self.latpole)
        if bad[0]:
            raise ValueError("invalid -CAR header: no native pole within [-90, 90] deg (CRVAL2 / LONPOLE / LATPOLE)")
        return tab[0]


# ------------------------------------------------------------------------------------------------------
# SPICE-style 4-axis headers (x = HPLN-TAN, y = HPLT-TAN, dispersion, time)
# ------------------------------------------------------------------------------------------------------
class SpiceWcs:
    """The pieces of a 4-axis SPICE L2 WCS that the pointing search needs, replacing the astropy calls
    `WCS(hdr).dropaxis(2)`, `.pixel_to_world(x, y, t)`, `.sub(['spectral'])` and `w_xy.to_header()` in
    `hdrshift/alignment_spice.py:250-261` and `synras/map_builder.py:249-294`.

    Axis roles are found from CTYPEi. The celestial pair may only couple to itself through PCi_j (SPICE
    files satisfy this); the time axis may depend on the x pixel through PC<t>_<x> (the raster scan).
    """

    def __init__(self, hdr):
        n = int(hdr.get("WCSAXES", hdr.get("NAXIS", 4)))
        self.n = n
        self.ilon, self.ilat = celestial_axes(hdr, naxis=n)
        self.iwave = self.itime = None
        for i in range(1, n + 1):
            ct = str(hdr.get("CTYPE%d" % i, "")).upper()
            if ct.startswith("WAVE") or ct.startswith("AWAV") or ct.startswith("FREQ"):
                self.iwave = i
            elif ct in ("UTC", "TIME", "TAI", "TT"):
                self.itime = i
        self.hdr = hdr
        self.pc = np.eye(n)
        for i in range(1, n + 1):
            for j in range(1, n + 1):
                k = "PC%d_%d" % (i, j)
                if k in hdr:
                    self.pc[i - 1, j - 1] = float(hdr[k])
        self.crpix = np.array([float(hdr.get("CRPIX%d" % i, 0.0)) for i in range(1, n + 1)])
        self.cdelt = np.array([float(hdr.get("CDELT%d" % i, 1.0)) for i in range(1, n + 1)])
        self.crval = np.array([float(hdr.get("CRVAL%d" % i, 0.0)) for i in range(1, n + 1)])
        for i in (self.ilon, self.ilat):
            for j in range(1, n + 1):
                if j not in (self.ilon, self.ilat) and self.pc[i - 1, j - 1] != 0.0:
                    raise NotImplementedError(f"PC{i}_{j} couples a celestial axis to a non-celestial one")

    def celestial(self) -> TanWcs:
        return TanWcs.from_header(self.hdr, self.ilon, self.ilat)

    def time_seconds(self, x, t=0.0, y=0.0):
        """Seconds relative to DATE-REF of pixel (x, y, t), 0-based, with the dispersion axis dropped
        (`w_spice.dropaxis(2)` removes its row and column of PCi_j)."""
        if self.itime is None:
            raise ValueError("no time axis in header")
        it = self.itime - 1
        off = (self.pc[it, self.ilon - 1] * (np.asarray(x, dtype=np.float64) + 1.0 - self.crpix[self.ilon - 1])
               + self.pc[it, self.ilat - 1] * (np.asarray(y, dtype=np.float64) + 1.0 - self.crpix[self.ilat - 1])
               + self.pc[it, it] * (np.asarray(t, dtype=np.float64) + 1.0 - self.crpix[it]))
        unit = str(self.hdr.get("CUNIT%d" % self.itime, "s")).strip()
        scale = {"s": 1.0, "min": 60.0, "h": 3600.0, "d": 86400.0}.get(unit, 1.0)
        return (self.crval[it] + self.cdelt[it] * off) * scale

    def wavelength(self, z):
        """World value of dispersion pixel z (0-based) in the header's CUNIT (`w_spice.sub(['spectral'])`)."""
        if self.iwave is None:
            raise ValueError("no spectral axis in header")
        iw = self.iwave - 1
        return self.crval[iw] + self.cdelt[iw] * self.pc[iw, iw] * (np.asarray(z, dtype=np.float64) + 1.0
                                                                    - self.crpix[iw])

    def xy_header(self):
        """What `w_xy.to_header()` yields for the celestial pair: degrees, PCi_j only where they differ from
        the identity, axes renumbered 1, 2."""
        from .fits_lite import Header
        w = self.celestial()
        h = Header()
        h["WCSAXES"] = 2
        h["CRPIX1"], h["CRPIX2"] = w.crpix1, w.crpix2
        for key, val, default in (("PC1_1", w.pc11, 1.0), ("PC1_2", w.pc12, 0.0), ("PC2_1", w.pc21, 0.0),
                                  ("PC2_2", w.pc22, 1.0)):
            if val != default:
                h[key] = val
        h["CDELT1"], h["CDELT2"] = w.cdelt1, w.cdelt2
        h["CUNIT1"], h["CUNIT2"] = "deg", "deg"
        h["CTYPE1"] = str(self.hdr["CTYPE%d" % self.ilon])
        h["CTYPE2"] = str(self.hdr["CTYPE%d" % self.ilat])
        h["CRVAL1"], h["CRVAL2"] = w.crval1, w.crval2
        h["LONPOLE"] = w.lonpole
        h["LATPOLE"] = float(self.hdr.get("LATPOLE", w.crval2))
        for k in ("DATE-OBS", "DATE-BEG", "DATE-AVG", "DATE-END", "RSUN_REF", "DSUN_OBS", "HGLN_OBS", "HGLT_OBS"):
            if k in self.hdr:
                h[k] = self.hdr[k]
        return h
