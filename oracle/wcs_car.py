"""wcslib-structured restatement of the plate-carree (-CAR) pixel<->world chain (TEST INFRASTRUCTURE).

The reference gets this arithmetic from `astropy.wcs.WCS` (wcslib inside astropy 7.2.0, `poetry.lock:190-191`)
when both inputs already are Carrington maps: `Alignment.align_using_initial_carrington`
(`hdrshift/alignment.py:344-399`) sets `lon_ctype = "CRLN-CAR"`, `lat_ctype = "CRLT-CAR"` and runs the same
`_extract_coordinates_pixels` (`:1038-1069`) -> `extract_EUI_coordinates(..., dsun=False)` (`utils/Util.py:283-305`,
no longitude wrapping for CAR) -> `WCS(hdr_shifted).world_to_pixel` chain as the helioprojective search.
wcslib is not available in this image => restated from its published algorithm (Calabretta & Greisen 2002,
FITS-WCS Paper II): `celset` (native pole from CRVAL, LONPOLE, LATPOLE; eqs. 8-10), `linp2x -> carx2s -> sphx2s`
for pixel->world, `sphs2x -> cars2x -> linx2p` for world->pixel; "parity unpinned" against wcslib itself.
The product evaluates the same maps as a 3x3 sphere rotation and two atan2 (`CoregLagCar`), an independent
formulation, so agreement of the two is a real check of both.
"""
from __future__ import annotations

import numpy as np

from .wcs_tan import D2R, R2D, _UNIT_TO_DEG, _atan2d, _cosd, _sincosd, _sind

_TOL = 1.0e-10


def _acosd(v):
    return np.arccos(np.clip(v, -1.0, 1.0)) * R2D


def celset_cylindrical(lng0, lat0, lonpole=None, latpole=None):
    """wcslib `celset` for a projection with fiducial point (phi0, theta0) = (0, 0): returns eul[5] =
    (lng_p, 90 - lat_p, phi_p, cos, sin of eul[1]). Raises ValueError where wcslib reports
    "Invalid coordinate transformation parameters"."""
    phi0, theta0 = 0.0, 0.0
    if lonpole is None:
        lonpole = (180.0 if lat0 < theta0 else 0.0) + phi0
    if latpole is None:
        latpole = 90.0
    phip = lonpole
    slat0, clat0 = float(_sind(lat0)), float(_cosd(lat0))
    sphip, cphip = float(_sind(phip - phi0)), float(_cosd(phip - phi0))
    sthe0, cthe0 = float(_sind(theta0)), float(_cosd(theta0))
    x = cthe0 * cphip
    y = sthe0
    z = np.sqrt(x * x + y * y)
    if z == 0.0:
        if slat0 != 0.0:
            raise ValueError("Invalid coordinate transformation parameters")
        latp = latpole
    else:
        slz = slat0 / z
        if abs(slz) > 1.0:
            if abs(slz) - 1.0 < _TOL:
                slz = 1.0 if slz > 0.0 else -1.0
            else:
                raise ValueError("Invalid coordinate transformation parameters")
        u = float(_atan2d(y, x))
        v = float(_acosd(slz))
        latp1 = u + v
        if latp1 > 180.0:
            latp1 -= 360.0
        elif latp1 < -180.0:
            latp1 += 360.0
        latp2 = u - v
        if latp2 > 180.0:
            latp2 -= 360.0
        elif latp2 < -180.0:
            latp2 += 360.0
        if abs(latpole - latp1) < abs(latpole - latp2):
            latp = latp1 if abs(latp1) < 90.0 + _TOL else latp2
        else:
            latp = latp2 if abs(latp2) < 90.0 + _TOL else latp1
        if abs(latp) < 90.0 + _TOL:
            if latp > 90.0:
                latp = 90.0
            elif latp < -90.0:
                latp = -90.0
        else:
            raise ValueError("Invalid coordinate transformation parameters")
    eul = np.empty(5, dtype=np.float64)
    eul[1] = 90.0 - latp
    z = float(_cosd(latp)) * clat0
    if abs(z) < _TOL:
        if abs(clat0) < _TOL:
            lngp = lng0
        elif latp > 0.0:
            lngp = lng0 + phip - phi0 - 180.0
        else:
            lngp = lng0 - phip + phi0
    else:
        x = (sthe0 - float(_sind(latp)) * slat0) / z
        y = sphip * cthe0 / clat0
        if x == 0.0 and y == 0.0:
            raise ValueError("Invalid coordinate transformation parameters")
        lngp = lng0 - float(_atan2d(y, x))
    if lng0 >= 0.0:
        if lngp < 0.0:
            lngp += 360.0
        elif lngp > 360.0:
            lngp -= 360.0
    else:
        if lngp > 0.0:
            lngp -= 360.0
        elif lngp < -360.0:
            lngp += 360.0
    eul[0] = lngp
    eul[2] = phip
    eul[3] = float(_cosd(eul[1]))
    eul[4] = float(_sind(eul[1]))
    return eul


def sphx2s(eul, phi, theta):
    """wcslib `sphx2s`: native (phi, theta) -> celestial (lng, lat), degrees."""
    phi = np.asarray(phi, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    if eul[4] == 0.0:
        if eul[1] == 0.0:
            dlng = np.fmod(eul[0] + 180.0 - eul[2], 360.0)
            lng = phi + dlng
            lat = theta + 0.0
        else:
            dlng = np.fmod(eul[0] + eul[2], 360.0)
            lng = dlng - phi
            lat = -theta
        if eul[0] >= 0.0:
            lng = np.where(lng < 0.0, lng + 360.0, lng)
        else:
            lng = np.where(lng > 0.0, lng - 360.0, lng)
        lng = np.where(lng > 360.0, lng - 360.0, lng)
        lng = np.where(lng < -360.0, lng + 360.0, lng)
        return lng, lat
    dphi = phi - eul[2]
    sinthe, costhe = _sincosd(theta)
    costhe3, costhe4 = costhe * eul[3], costhe * eul[4]
    sinthe3, sinthe4 = sinthe * eul[3], sinthe * eul[4]
    sinphi, cosphi = _sincosd(dphi)
    x = sinthe4 - costhe3 * cosphi
    small = np.abs(x) < 1.0e-5
    if np.any(small):
        x = np.where(small, -_cosd(theta + eul[1]) + costhe3 * (1.0 - cosphi), x)
    y = -costhe * sinphi
    both0 = (x == 0.0) & (y == 0.0)
    dlng = np.where(both0, np.where(eul[1] < 90.0, dphi + 180.0, -dphi), _atan2d(y, x))
    lng = eul[0] + dlng
    if eul[0] >= 0.0:
        lng = np.where(lng < 0.0, lng + 360.0, lng)
    else:
        lng = np.where(lng > 0.0, lng - 360.0, lng)
    lng = np.where(lng > 360.0, lng - 360.0, lng)
    lng = np.where(lng < -360.0, lng + 360.0, lng)
    z = sinthe3 + costhe4 * cosphi
    with np.errstate(invalid="ignore"):
        lat = np.where(np.abs(z) > 0.99, np.copysign(_acosd(np.sqrt(x * x + y * y)), z),
                       np.arcsin(np.clip(z, -1.0, 1.0)) * R2D)
    m180 = np.fmod(dphi, 180.0) == 0.0
    if np.any(m180):
        l180 = theta + cosphi * eul[1]
        l180 = np.where(l180 > 90.0, 180.0 - l180, l180)
        l180 = np.where(l180 < -90.0, -180.0 - l180, l180)
        lat = np.where(m180, l180, lat)
    return lng, lat


def sphs2x(eul, lng, lat):
    """wcslib `sphs2x`: celestial (lng, lat) -> native (phi, theta), degrees, phi in [-180, 180]."""
    lng = np.asarray(lng, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    if eul[4] == 0.0:
        if eul[1] == 0.0:
            dphi = np.fmod(eul[2] - 180.0 - eul[0], 360.0)
            phi = np.fmod(lng + dphi, 360.0)
            theta = lat + 0.0
        else:
            dphi = np.fmod(eul[2] + eul[0], 360.0)
            phi = np.fmod(dphi - lng, 360.0)
            theta = -lat
        phi = np.where(phi > 180.0, phi - 360.0, phi)
        phi = np.where(phi < -180.0, phi + 360.0, phi)
        return phi, theta
    dlng = lng - eul[0]
    sinlat, coslat = _sincosd(lat)
    coslat3, coslat4 = coslat * eul[3], coslat * eul[4]
    sinlat3, sinlat4 = sinlat * eul[3], sinlat * eul[4]
    sinlng, coslng = _sincosd(dlng)
    x = sinlat4 - coslat3 * coslng
    small = np.abs(x) < 1.0e-5
    if np.any(small):
        x = np.where(small, -_cosd(lat + eul[1]) + coslat3 * (1.0 - coslng), x)
    y = -coslat * sinlng
    both0 = (x == 0.0) & (y == 0.0)
    dphi = np.where(both0, np.where(eul[1] < 90.0, dlng - 180.0, -dlng), _atan2d(y, x))
    phi = np.fmod(eul[2] + dphi, 360.0)
    phi = np.where(phi > 180.0, phi - 360.0, phi)
    phi = np.where(phi < -180.0, phi + 360.0, phi)
    z = sinlat3 + coslat4 * coslng
    with np.errstate(invalid="ignore"):
        theta = np.where(np.abs(z) > 0.99, np.copysign(_acosd(np.sqrt(x * x + y * y)), z),
                         np.arcsin(np.clip(z, -1.0, 1.0)) * R2D)
    m180 = np.fmod(dlng, 180.0) == 0.0
    if np.any(m180):
        t180 = lat + coslng * eul[1]
        t180 = np.where(t180 > 90.0, 180.0 - t180, t180)
        t180 = np.where(t180 < -90.0, -180.0 - t180, t180)
        theta = np.where(m180, t180, theta)
    return phi, theta


class WcsCar:
    """What `astropy.wcs.WCS(header)` holds for a 2-axis -CAR header after `wcsset`."""

    def __init__(self, hdr):
        ct1, ct2 = str(hdr["CTYPE1"]), str(hdr["CTYPE2"])
        if not (ct1.endswith("CAR") and ct2.endswith("CAR")):
            raise NotImplementedError("this oracle class handles -CAR only")
        s1 = _UNIT_TO_DEG[str(hdr["CUNIT1"]).strip()] if "CUNIT1" in hdr else 1.0
        s2 = _UNIT_TO_DEG[str(hdr["CUNIT2"]).strip()] if "CUNIT2" in hdr else 1.0
        self.crpix = (float(hdr["CRPIX1"]), float(hdr["CRPIX2"]))
        cdelt = [float(hdr["CDELT1"]), float(hdr["CDELT2"])]
        if any(k in hdr for k in ("PC1_1", "PC1_2", "PC2_1", "PC2_2")):
            pc = [[float(hdr["PC1_1"]) if "PC1_1" in hdr else 1.0, float(hdr["PC1_2"]) if "PC1_2" in hdr else 0.0],
                  [float(hdr["PC2_1"]) if "PC2_1" in hdr else 0.0, float(hdr["PC2_2"]) if "PC2_2" in hdr else 1.0]]
        else:
            rho = float(hdr["CROTA2"]) if "CROTA2" in hdr else (float(hdr["CROTA"]) if "CROTA" in hdr else None)
            if rho is None or rho == 0.0:
                pc = [[1.0, 0.0], [0.0, 1.0]]
            else:
                c, s = np.cos(rho * D2R), np.sin(rho * D2R)
                pc = [[c, -s * cdelt[1] / cdelt[0]], [s * cdelt[0] / cdelt[1], c]]
        self.cdelt = (cdelt[0] * s1, cdelt[1] * s2)
        self.crval = (float(hdr["CRVAL1"]) * s1, float(hdr["CRVAL2"]) * s2)
        self.piximg = np.array([[self.cdelt[0] * pc[0][0], self.cdelt[0] * pc[0][1]],
                                [self.cdelt[1] * pc[1][0], self.cdelt[1] * pc[1][1]]], dtype=np.float64)
        self.imgpix = np.linalg.inv(self.piximg)
        self.unity = (pc[0][0] == 1.0 and pc[1][1] == 1.0 and pc[0][1] == 0.0 and pc[1][0] == 0.0)
        self.eul = celset_cylindrical(self.crval[0], self.crval[1],
                                      float(hdr["LONPOLE"]) if "LONPOLE" in hdr else None,
                                      float(hdr["LATPOLE"]) if "LATPOLE" in hdr else None)
        self.pc = pc
        self.lonpole = float(self.eul[2])
        self.pixel_shape = (int(hdr["ZNAXIS1"] if "ZNAXIS1" in hdr else hdr["NAXIS1"]),
                            int(hdr["ZNAXIS2"] if "ZNAXIS2" in hdr else hdr["NAXIS2"]))

    def pixel_to_world(self, x, y):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        t1 = (x + 1.0) - self.crpix[0]
        t2 = (y + 1.0) - self.crpix[1]
        if self.unity:
            xi, eta = self.cdelt[0] * t1, self.cdelt[1] * t2
        else:
            xi = self.piximg[0, 0] * t1 + self.piximg[0, 1] * t2
            eta = self.piximg[1, 0] * t1 + self.piximg[1, 1] * t2
        # carx2s with r0 = 180/pi: phi = x, theta = y
        return sphx2s(self.eul, xi, eta)

    def world_to_pixel(self, lng, lat):
        phi, theta = sphs2x(self.eul, lng, lat)
        # cars2x: x = phi, y = theta
        if self.unity:
            p1 = phi / self.cdelt[0] + self.crpix[0]
            p2 = theta / self.cdelt[1] + self.crpix[1]
        else:
            p1 = (self.imgpix[0, 0] * phi + self.imgpix[0, 1] * theta) + self.crpix[0]
            p2 = (self.imgpix[1, 0] * phi + self.imgpix[1, 1] * theta) + self.crpix[1]
        return p1 - 1.0, p2 - 1.0


def extract_coordinates(hdr):
    """`AlignEUIUtil.extract_EUI_coordinates(hdr, dsun=False, lon_ctype="CRLN-CAR", ...)` (`utils/Util.py:283-312`):
    lon / lat [deg] of every pixel, NOT wrapped (`:302-305`)."""
    w = WcsCar(hdr)
    x, y = np.meshgrid(np.arange(w.pixel_shape[0]), np.arange(w.pixel_shape[1]))
    return w.pixel_to_world(x, y)


def extract_coordinates_pixels(hdr_initial, hdr_target, world=None):
    """`Alignment._extract_coordinates_pixels` (`hdrshift/alignment.py:1038-1069`, non-sunpy branch) for -CAR headers."""
    if world is None:
        world = extract_coordinates(hdr_initial)
    return WcsCar(hdr_target).world_to_pixel(world[0], world[1])
