"""wcslib-structured restatement of the TAN pixel<->world chain (TEST INFRASTRUCTURE).

The reference gets this arithmetic from `astropy.wcs.WCS` (wcslib inside astropy 7.2.0,
`poetry.lock:190-191`); call sites: `WCS(hdr)` `hdrshift/alignment.py:1041`, `utils/Util.py:284`;
`pixel_to_world` `utils/Util.py:290`; `world_to_pixel` `hdrshift/alignment.py:1065`.
wcslib is not available in this image => this file is restated from wcslib's published algorithm
(linp2x -> tanx2s -> sphx2s for pixel->world, sphs2x -> tans2x -> linx2p for world->pixel;
Calabretta & Greisen 2002, FITS-WCS Paper II) and is "parity unpinned" against wcslib itself.
Each step keeps wcslib's operation order (degree-based trig helpers, eul[] Euler-angle vector,
the |z| > 0.99 branch for the latitude, longitude normalisation by the sign of CRVAL1).
"""
from __future__ import annotations

import numpy as np

D2R = np.pi / 180.0
R2D = 180.0 / np.pi

_UNIT_TO_DEG = {"deg": 1.0, "arcsec": 1.0 / 3600.0, "arcmin": 1.0 / 60.0, "rad": 180.0 / np.pi,
                "mas": 1.0 / 3600000.0}


def _cosd(a):
    a = np.asarray(a, dtype=np.float64)
    out = np.cos(a * D2R)
    m = np.fmod(a, 90.0) == 0.0
    if np.any(m):
        i = np.abs(np.floor(a / 90.0 + 0.5)).astype(np.int64) % 4
        exact = np.array([1.0, 0.0, -1.0, 0.0])[i]
        out = np.where(m, exact, out)
    return out


def _sind(a):
    a = np.asarray(a, dtype=np.float64)
    out = np.sin(a * D2R)
    m = np.fmod(a - 90.0, 90.0) == 0.0
    if np.any(m):
        i = np.abs(np.floor((a - 90.0) / 90.0 + 0.5)).astype(np.int64) % 4
        exact = np.array([1.0, 0.0, -1.0, 0.0])[i]
        out = np.where(m, exact, out)
    return out


def _sincosd(a):
    return _sind(a), _cosd(a)


def _atan2d(y, x):
    return np.arctan2(y, x) * R2D


class WcsTan:
    """What `astropy.wcs.WCS(header)` holds for a 2-axis TAN header after `wcsset`."""

    def __init__(self, hdr):
        ct1, ct2 = str(hdr["CTYPE1"]), str(hdr["CTYPE2"])
        if not (ct1.endswith("TAN") and ct2.endswith("TAN")):
            raise NotImplementedError("oracle handles -TAN only")
        s1 = _UNIT_TO_DEG[str(hdr["CUNIT1"]).strip()] if "CUNIT1" in hdr else 1.0
        s2 = _UNIT_TO_DEG[str(hdr["CUNIT2"]).strip()] if "CUNIT2" in hdr else 1.0
        self.crpix = (float(hdr["CRPIX1"]), float(hdr["CRPIX2"]))
        cdelt = [float(hdr["CDELT1"]), float(hdr["CDELT2"])]
        if "PC1_1" in hdr or "PC1_2" in hdr or "PC2_1" in hdr or "PC2_2" in hdr:
            pc = [[float(hdr["PC1_1"]) if "PC1_1" in hdr else 1.0, float(hdr["PC1_2"]) if "PC1_2" in hdr else 0.0],
                  [float(hdr["PC2_1"]) if "PC2_1" in hdr else 0.0, float(hdr["PC2_2"]) if "PC2_2" in hdr else 1.0]]
        else:
            rho = None
            if "CROTA2" in hdr:
                rho = float(hdr["CROTA2"])
            elif "CROTA" in hdr:
                rho = float(hdr["CROTA"])
            if rho is None or rho == 0.0:
                pc = [[1.0, 0.0], [0.0, 1.0]]
            else:
                c, s = np.cos(rho * D2R), np.sin(rho * D2R)
                pc = [[c, -s * cdelt[1] / cdelt[0]], [s * cdelt[0] / cdelt[1], c]]
        # wcsset: unit rescale to degrees is applied to CRVAL and CDELT
        self.cdelt = (cdelt[0] * s1, cdelt[1] * s2)
        self.crval = (float(hdr["CRVAL1"]) * s1, float(hdr["CRVAL2"]) * s2)
        self.pc = pc
        if "LONPOLE" in hdr:
            self.lonpole = float(hdr["LONPOLE"])
        else:
            self.lonpole = 0.0 if self.crval[1] >= 90.0 else 180.0
        # linset: piximg[i][j] = cdelt[i]*pc[i][j]; imgpix = inverse
        self.piximg = np.array([[self.cdelt[0] * pc[0][0], self.cdelt[0] * pc[0][1]],
                                [self.cdelt[1] * pc[1][0], self.cdelt[1] * pc[1][1]]], dtype=np.float64)
        self.imgpix = np.linalg.inv(self.piximg)
        self.unity = (pc[0][0] == 1.0 and pc[1][1] == 1.0 and pc[0][1] == 0.0 and pc[1][0] == 0.0)
        # celset for a zenithal projection (theta0 = 90): the native pole is the reference point
        self.eul = np.empty(5, dtype=np.float64)
        self.eul[0] = self.crval[0]
        self.eul[1] = 90.0 - self.crval[1]
        self.eul[2] = self.lonpole
        self.eul[3] = float(_cosd(self.eul[1]))
        self.eul[4] = float(_sind(self.eul[1]))
        self.pixel_shape = (int(hdr["ZNAXIS1"] if "ZNAXIS1" in hdr else hdr["NAXIS1"]),
                            int(hdr["ZNAXIS2"] if "ZNAXIS2" in hdr else hdr["NAXIS2"]))

    # ------------------------------------------------------------------
    def pixel_to_world(self, x, y):
        """0-based pixel coordinates -> (lng, lat) degrees, like `w.pixel_to_world(x, y)` on a plain WCS."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        # linp2x
        t1 = (x + 1.0) - self.crpix[0]
        t2 = (y + 1.0) - self.crpix[1]
        if self.unity:
            xi = self.cdelt[0] * t1
            eta = self.cdelt[1] * t2
        else:
            xi = self.piximg[0, 0] * t1 + self.piximg[0, 1] * t2
            eta = self.piximg[1, 0] * t1 + self.piximg[1, 1] * t2
        # tanx2s
        r = np.sqrt(xi * xi + eta * eta)
        phi = np.where(r == 0.0, 0.0, _atan2d(xi, -eta))
        theta = _atan2d(R2D, r)
        # sphx2s
        eul = self.eul
        dphi = phi - eul[2]
        sinthe, costhe = _sincosd(theta)
        costhe3 = costhe * eul[3]
        costhe4 = costhe * eul[4]
        sinthe3 = sinthe * eul[3]
        sinthe4 = sinthe * eul[4]
        sinphi, cosphi = _sincosd(dphi)
        xx = sinthe4 - costhe3 * cosphi
        small = np.abs(xx) < 1.0e-5
        if np.any(small):
            xx = np.where(small, -_cosd(theta + eul[1]) + costhe3 * (1.0 - cosphi), xx)
        yy = -costhe * sinphi
        both0 = (xx == 0.0) & (yy == 0.0)
        dlng = np.where(both0, dphi + 180.0, _atan2d(yy, xx))
        lng = eul[0] + dlng
        if eul[0] >= 0.0:
            lng = np.where(lng < 0.0, lng + 360.0, lng)
        else:
            lng = np.where(lng > 0.0, lng - 360.0, lng)
        lng = np.where(lng > 360.0, lng - 360.0, lng)
        lng = np.where(lng < -360.0, lng + 360.0, lng)
        z = sinthe3 + costhe4 * cosphi
        with np.errstate(invalid="ignore"):
            lat = np.where(np.abs(z) > 0.99,
                           np.copysign(np.arccos(np.minimum(np.sqrt(xx * xx + yy * yy), 1.0)) * R2D, z),
                           np.arcsin(np.clip(z, -1.0, 1.0)) * R2D)
        m180 = np.fmod(dphi, 180.0) == 0.0
        if np.any(m180):
            lat180 = theta + cosphi * eul[1]
            lat180 = np.where(lat180 > 90.0, 180.0 - lat180, lat180)
            lat180 = np.where(lat180 < -90.0, -180.0 - lat180, lat180)
            lat = np.where(m180, lat180, lat)
        return lng, lat

    # ------------------------------------------------------------------
    def world_to_pixel(self, lng, lat):
        """(lng, lat) degrees -> 0-based pixel coordinates, like `w.world_to_pixel(lng, lat)`.
        Points behind the tangent hemisphere (theta < 0) come back NaN (astropy's invalid marker)."""
        lng = np.asarray(lng, dtype=np.float64)
        lat = np.asarray(lat, dtype=np.float64)
        eul = self.eul
        # sphs2x
        dlng = lng - eul[0]
        sinlat, coslat = _sincosd(lat)
        coslat3 = coslat * eul[3]
        coslat4 = coslat * eul[4]
        sinlat3 = sinlat * eul[3]
        sinlat4 = sinlat * eul[4]
        sinlng, coslng = _sincosd(dlng)
        xx = sinlat4 - coslat3 * coslng
        small = np.abs(xx) < 1.0e-5
        if np.any(small):
            xx = np.where(small, -_cosd(lat + eul[1]) + coslat3 * (1.0 - coslng), xx)
        yy = -coslat * sinlng
        both0 = (xx == 0.0) & (yy == 0.0)
        dphi = np.where(both0, dlng - 180.0, _atan2d(yy, xx))
        phi = np.fmod(eul[2] + dphi, 360.0)
        phi = np.where(phi > 180.0, phi - 360.0, phi)
        phi = np.where(phi < -180.0, phi + 360.0, phi)
        z = sinlat3 + coslat4 * coslng
        with np.errstate(invalid="ignore"):
            theta = np.where(np.abs(z) > 0.99,
                             np.copysign(np.arccos(np.minimum(np.sqrt(xx * xx + yy * yy), 1.0)) * R2D, z),
                             np.arcsin(np.clip(z, -1.0, 1.0)) * R2D)
        m180 = np.fmod(dlng, 180.0) == 0.0
        if np.any(m180):
            th180 = lat + coslng * eul[1]
            th180 = np.where(th180 > 90.0, 180.0 - th180, th180)
            th180 = np.where(th180 < -90.0, -180.0 - th180, th180)
            theta = np.where(m180, th180, theta)
        # tans2x
        sinphi, cosphi = _sincosd(phi)
        s = _sind(theta)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = R2D * _cosd(theta) / s
        xi = r * sinphi
        eta = -r * cosphi
        bad = (theta < 0.0) | (s == 0.0)
        # linx2p
        if self.unity:
            p1 = xi / self.cdelt[0] + self.crpix[0]
            p2 = eta / self.cdelt[1] + self.crpix[1]
        else:
            p1 = (self.imgpix[0, 0] * xi + self.imgpix[0, 1] * eta) + self.crpix[0]
            p2 = (self.imgpix[1, 0] * xi + self.imgpix[1, 1] * eta) + self.crpix[1]
        p1 = np.where(bad, np.nan, p1 - 1.0)
        p2 = np.where(bad, np.nan, p2 - 1.0)
        return p1, p2


def ang2pipi_deg(ang):
    """`AlignCommonUtil.ang2pipi` on degree values (`utils/Util.py:76-80`)."""
    return -((-np.asarray(ang, dtype=np.float64) + 180.0) % 360.0 - 180.0)


def extract_coordinates(hdr):
    """`AlignEUIUtil.extract_EUI_coordinates(hdr, dsun=False)` on the plain-WCS branch
    (`utils/Util.py:283-312`): lon/lat [deg] of every pixel, wrapped to (-180, 180]."""
    w = WcsTan(hdr)
    x, y = np.meshgrid(np.arange(w.pixel_shape[0]), np.arange(w.pixel_shape[1]))
    lng, lat = w.pixel_to_world(x, y)
    return ang2pipi_deg(lng), ang2pipi_deg(lat)


def extract_coordinates_pixels(hdr_initial, hdr_target, world=None):
    """`Alignment._extract_coordinates_pixels` (`hdrshift/alignment.py:1038-1069`, non-sunpy branch):
    pixel coordinates in `hdr_target`'s image of every pixel centre of `hdr_initial`'s grid.
    `world` lets a caller reuse the (lag-independent) world grid of `hdr_initial`."""
    if world is None:
        world = extract_coordinates(hdr_initial)
    return WcsTan(hdr_target).world_to_pixel(world[0], world[1])
