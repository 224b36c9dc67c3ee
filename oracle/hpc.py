"""Helioprojective lag-grid search, CPU restatement (TEST INFRASTRUCTURE).

Follows the `parallelism=True` branch of `Alignment._find_best_header_parameters`
(`hdrshift/alignment.py:613-756`) as driven by `align_using_helioprojective` (`:263-342`):

* `check_and_create_pcij`        <- `_check_ant_create_pcij_matrix`            `:580-611`
* `threshold_to_nan`             <- `_set_threshold_minmax_to_nan`             `:876-887`
* `initial_header_values`        <- `_set_initial_header_values`               `:799-842`
* `shift_header`                 <- `_shift_header`                            `:401-468`
* `create_submap_of_large_data`  <- `_create_submap_of_large_data`             `:987-1016`
* `interpolate_on_large_data_grid` <- `_interpolate_on_large_data_grid`        `:1018-1029`
* `step`                         <- `_step`                                    `:509-542`
* `hpc_cube`                     <- lag enumeration / scatter                  `:635-756`

Bug-compatible choices: the common grid is the UNSHIFTED small grid (B3, parallel
branch); a non-zero CDELT1 lag only triggers the PC rebuild (B1); a non-zero CDELT2 lag kills the
worker in the reference so its cube entries stay 0.0 (B1, B11) -- `cdelt_mode="intended"` switches to
the evident intent (CDELTi = ref + lag, then PC rebuild), labelled "restated, not reference-verified".
Headers are plain mappings (dict); images are numpy arrays.
"""
from __future__ import annotations

import multiprocessing as mp
import warnings

import numpy as np

from . import wcs_tan
from .pearson import masked_pearson
from .resample import interpol2d

_ARCSEC_PER = {"arcsec": 1.0, "deg": 3600.0, "arcmin": 60.0, "rad": 3600.0 * 180.0 / np.pi}


def _convert(values, src, dst):
    """`u.Quantity(values, src).to(dst).value` for angular units."""
    values = np.asarray(values, dtype=np.float64)
    if src == dst:
        return values
    if (src, dst) == ("arcsec", "deg"):
        return values * (1.0 / 3600.0)
    if (src, dst) == ("deg", "arcsec"):
        return values * 3600.0
    return values * (_ARCSEC_PER[src] / _ARCSEC_PER[dst])


def _ang2pipi(values, unit):
    """`AlignCommonUtil.ang2pipi(u.Quantity(values, unit))` -- the arithmetic is done in `unit`."""
    pi = float(_convert(180.0, "deg", unit))
    values = np.asarray(values, dtype=np.float64)
    return -((-values + pi) % (2 * pi) - pi)


def check_and_create_pcij(hdr, force_crota_0=False):
    if "PC1_1" not in hdr:
        if "CROTA" in hdr:
            crot = hdr["CROTA"]
        elif "CROTA2" in hdr:
            crot = hdr["CROTA2"]
        elif force_crota_0:
            crot = 0.0
            hdr["CROTA"] = 0.0
        else:
            raise ValueError("No, CROTA, CROTA2 or PCi_j matrix in your FITS file.")
        rho = np.deg2rad(crot)
        lam = hdr["CDELT2"] / hdr["CDELT1"]
        hdr["PC1_1"] = np.cos(rho)
        hdr["PC2_2"] = np.cos(rho)
        hdr["PC1_2"] = -lam * np.sin(rho)
        hdr["PC2_1"] = (1 / lam) * np.sin(rho)
    if hdr["PC1_1"] >= 1.0:
        hdr["PC1_1"] = 1.0
        hdr["PC2_2"] = 1.0
        hdr["PC1_2"] = 0.0
        hdr["PC2_1"] = 0.0
        hdr["CROTA"] = 0.0
    if "CROTA" not in hdr:
        s = -np.sign(hdr["PC1_2"]) + (hdr["PC1_2"] == 0)
        hdr["CROTA"] = s * np.rad2deg(np.arccos(hdr["PC1_1"]))


def threshold_to_nan(data_small, value_min=None, value_max=None):
    c1 = np.ones(data_small.shape, dtype=bool)
    c2 = np.ones(data_small.shape, dtype=bool)
    with np.errstate(invalid="ignore"):
        if value_min is not None:
            c1[np.abs(data_small) < value_min] = False
        if value_max is not None:
            c2[np.abs(data_small) > value_max] = False
    data_small[np.logical_not(np.logical_and(c1, c2))] = np.nan


class Refs:
    """Reference header values + lag arrays in header units (`_set_initial_header_values`)."""

    def __init__(self, hdr_small, lag_crval1, lag_crval2, lag_cdelt1, lag_cdelt2, lag_crota, lag_solar_r,
                 unit_lag="arcsec", ang2pipi=True):
        none0 = lambda v: np.array([0.0]) if v is None else np.asarray(v, dtype=np.float64)  # noqa: E731
        self.lag_crval1, self.lag_crval2 = none0(lag_crval1), none0(lag_crval2)
        self.lag_cdelt1, self.lag_cdelt2 = none0(lag_cdelt1), none0(lag_cdelt2)
        self.lag_crota = none0(lag_crota)
        self.unit_lag = unit_lag
        self.crval1_ref = hdr_small["CRVAL1"]
        self.crval2_ref = hdr_small["CRVAL2"]
        if "CROTA" in hdr_small:
            self.crota_ref = hdr_small["CROTA"]
        elif "CROTA2" in hdr_small:
            self.crota_ref = hdr_small["CROTA2"]
        else:
            s = -np.sign(hdr_small["PC1_2"]) + (hdr_small["PC1_2"] == 0)
            self.crota_ref = np.rad2deg(np.arccos(hdr_small["PC1_1"])) * s
            hdr_small["CROTA"] = np.rad2deg(np.arccos(hdr_small["PC1_1"]))
        self.cdelt1_ref = hdr_small["CDELT1"]
        self.cdelt2_ref = hdr_small["CDELT2"]
        unit1, unit2 = hdr_small["CUNIT1"], hdr_small["CUNIT2"]
        if self.unit_lag in unit1:
            pass
        else:
            f = _ang2pipi if ang2pipi else (lambda v, u: np.asarray(v, dtype=np.float64))
            self.lag_crval1 = _convert(f(self.lag_crval1, unit_lag), unit_lag, unit1)
            self.lag_crval2 = _convert(f(self.lag_crval2, unit_lag), unit_lag, unit2)
            self.lag_cdelt1 = _convert(f(self.lag_cdelt1, unit_lag), unit_lag, unit1)
            self.lag_cdelt2 = _convert(f(self.lag_cdelt2, unit_lag), unit_lag, unit2)
            self.unit_lag = unit1
        if unit1 != unit2:
            raise ValueError("CUNIT1 and CUNIT2 must be equal")
        self.lag_solar_r = np.array([1.004]) if lag_solar_r is None else np.asarray(lag_solar_r, dtype=np.float64)


class LagKillsWorker(Exception):
    """The reference's worker dies on this lag (non-zero CDELT2 lag, `alignment.py:440`)."""


def shift_header(hdr, refs: Refs, d_crval1, d_crval2, d_cdelt1, d_cdelt2, d_crota, cdelt_mode="reference"):
    if refs.unit_lag != hdr["CUNIT1"] or refs.unit_lag != hdr["CUNIT2"]:
        raise ValueError("lag.unit and cUNIT are not the same")
    hdr["CRVAL1"] = refs.crval1_ref + d_crval1
    hdr["CRVAL2"] = refs.crval2_ref + d_crval2
    change_pcij = False
    crot = refs.crota_ref
    if d_cdelt1 != 0.0:
        change_pcij = True
        if cdelt_mode == "intended":
            hdr["CDELT1"] = refs.cdelt1_ref + d_cdelt1
    if d_cdelt2 != 0.0:
        change_pcij = True
        if cdelt_mode == "intended":
            hdr["CDELT2"] = refs.cdelt2_ref + d_cdelt2
        else:
            raise LagKillsWorker("float has no attribute 'to' (alignment.py:440)")
    if d_crota != 0.0:
        change_pcij = True
        if "CROTA" in hdr:
            hdr["CROTA"] = refs.crota_ref + d_crota
        elif "CROTA2" in hdr:
            hdr["CROTA2"] = refs.crota_ref + d_crota
        crot = refs.crota_ref + d_crota
    if change_pcij:
        rho = np.deg2rad(crot)
        lam = hdr["CDELT2"] / hdr["CDELT1"]
        hdr["PC1_1"] = np.cos(rho)
        hdr["PC2_2"] = np.cos(rho)
        hdr["PC1_2"] = -lam * np.sin(rho)
        hdr["PC2_1"] = (1 / lam) * np.sin(rho)


def build_regular_grid(longitude, latitude, lonlims=None, latlims=None):
    """`Util.PlotFits.build_regular_grid` (`utils/Util.py:874-904`) on degree arrays."""
    x = np.abs(longitude[0, 1] - longitude[0, 0])
    y = np.abs(latitude[0, 1] - latitude[0, 0])
    dlon = np.sqrt(x ** 2 + y ** 2)
    x = np.abs(longitude[1, 0] - longitude[0, 0])
    y = np.abs(latitude[1, 0] - latitude[0, 0])
    dlat = np.sqrt(x ** 2 + y ** 2)
    lon1d = np.arange(np.min(longitude), np.max(longitude), dlon)
    lat1d = np.arange(np.min(latitude), np.max(latitude), dlat)
    if (lonlims is not None) or (latlims is not None):
        lon1d = lon1d[(lon1d > lonlims[0]) & (lon1d < lonlims[1])]
        lat1d = lat1d[(lat1d > latlims[0]) & (lat1d < latlims[1])]
    lon_g, lat_g = np.meshgrid(lon1d, lat1d)
    return lon_g, lat_g, dlon, dlat


def select_fov_in_small_data(data_small, hdr_small, fov_limits_arcsec, order):
    """`Alignment._select_fov_in_small_data` (`hdrshift/alignment.py:1082-1127`): small image on a regular,
    unrotated lon / lat grid inside the limits; returns (data float64, new header). NAXIS1 / CRPIX1 come from the
    grid's row count, as in the reference."""
    lonlims = _convert(fov_limits_arcsec[0], "arcsec", "deg")
    latlims = _convert(fov_limits_arcsec[1], "arcsec", "deg")
    lon, lat = wcs_tan.extract_coordinates(hdr_small)
    long, latg, dlon, dlat = build_regular_grid(lon, lat, lonlims, latlims)
    mid = [long.shape[0] // 2, long.shape[1] // 2]
    h = dict(hdr_small)
    h["CRVAL1"] = float(_convert(long[mid[0], mid[1]], "deg", h["CUNIT1"]))
    h["CRVAL2"] = float(_convert(latg[mid[0], mid[1]], "deg", h["CUNIT2"]))
    h["CRPIX1"], h["CRPIX2"] = mid[0] + 1, mid[1] + 1
    h["CDELT1"] = float(_convert(dlon, "deg", h["CUNIT1"]))
    h["CDELT2"] = float(_convert(dlat, "deg", h["CUNIT2"]))
    h["PC1_1"], h["PC2_2"], h["PC1_2"], h["PC2_1"] = 1.0, 1.0, 0.0, 0.0
    h["CROTA"], h["CROTA2"] = 0.0, 0.0
    h["NAXIS1"], h["NAXIS2"] = long.shape[0], long.shape[1]
    xg, yg = wcs_tan.extract_coordinates_pixels(h, hdr_small)
    out = np.zeros_like(xg)
    interpol2d(np.array(data_small, dtype=np.float64), x=xg, y=yg, order=order, fill=np.nan, dst=out)
    return out, h


def create_submap_of_large_data(data_large, hdr_large, hdr_small, order, coords=wcs_tan):
    x_cut, y_cut = coords.extract_coordinates_pixels(hdr_small, hdr_large)
    cut = np.zeros_like(x_cut, dtype="float32")
    interpol2d(data_large.copy(), x=x_cut, y=y_cut, dst=cut, order=order, fill=np.nan)
    return np.array(cut)


def interpolate_on_large_data_grid(data_small, hdr_grid, hdr_shifted, order, world=None, coords=wcs_tan):
    x, y = coords.extract_coordinates_pixels(hdr_grid, hdr_shifted, world=world)
    out = np.zeros_like(x, dtype="float32")
    interpol2d(data_small.copy(), x=x, y=y, order=order, fill=np.nan, dst=out)
    return out


class HpcSearch:
    """State of one helioprojective search after the one-time preparation. frame="car" is the same search on two
    Carrington maps (`align_using_initial_carrington`, `hdrshift/alignment.py:344-399`): -CAR headers, no longitude
    wrapping of lags or world coordinates, both images rounded to float32 when read (`:373, 387`)."""

    def __init__(self, data_large, hdr_large, data_small, hdr_small,
                 lag_crval1, lag_crval2, lag_cdelt1, lag_cdelt2, lag_crota, lag_solar_r=None,
                 small_fov_value_min=None, small_fov_value_max=None, order=2, unit_lag="arcsec",
                 force_crota_0=False, cdelt_mode="reference", fov_limits=None, frame="hpc"):
        if frame not in ("hpc", "car"):
            raise ValueError("frame must be 'hpc' or 'car'")
        self.frame = frame
        if frame == "car":
            from . import wcs_car
            self.coords = wcs_car
            data_large = np.array(data_large, dtype="float32")
            data_small = np.array(data_small, dtype="float32")
            if fov_limits is not None:
                raise NotImplementedError("fov_limits on -CAR inputs (the reference's helper is TAN-only)")
        else:
            self.coords = wcs_tan
        self.order = order
        self.cdelt_mode = cdelt_mode
        self.hdr_small = dict(hdr_small)
        hdr_large = dict(hdr_large)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            check_and_create_pcij(self.hdr_small, force_crota_0)
            check_and_create_pcij(hdr_large, force_crota_0)
        self.data_small = np.array(data_small, dtype=np.float64)
        threshold_to_nan(self.data_small, small_fov_value_min, small_fov_value_max)
        if fov_limits is not None:   # alignment.py:859-860, before _set_initial_header_values (:623)
            self.data_small, self.hdr_small = select_fov_in_small_data(self.data_small, self.hdr_small, fov_limits,
                                                                       order)
        self.refs = Refs(self.hdr_small, lag_crval1, lag_crval2, lag_cdelt1, lag_cdelt2, lag_crota,
                         lag_solar_r, unit_lag=unit_lag, ang2pipi=(frame == "hpc"))
        if np.isnan(self.data_small).all():
            raise ValueError("minimum or maximum value have set all small FOV to nan")
        # one-time: large image onto the unshifted small grid, float32 (alignment.py:649-651, 987-1016)
        self.data_large = create_submap_of_large_data(np.array(data_large, dtype=np.float64), hdr_large,
                                                      self.hdr_small, order, coords=self.coords)
        self.hdr_grid = dict(self.hdr_small)  # self.hdr_large = hdr_cut.copy()  (:1000)
        self._world = None

    @property
    def shape(self):
        r = self.refs
        return (len(r.lag_crval1), len(r.lag_crval2), len(r.lag_cdelt1), len(r.lag_cdelt2),
                len(r.lag_crota), len(r.lag_solar_r))

    def world(self):
        """Lag-independent world grid of the common grid. The reference recomputes it for every lag
        (`alignment.py:1061`); the values are identical, so the oracle may cache them."""
        if self._world is None:
            self._world = self.coords.extract_coordinates(self.hdr_grid)
        return self._world

    def reprojected(self, d_crval1, d_crval2, d_cdelt1, d_cdelt2, d_crota, reuse_world=True):
        hdr = dict(self.hdr_small)
        shift_header(hdr, self.refs, d_crval1, d_crval2, d_cdelt1, d_cdelt2, d_crota, self.cdelt_mode)
        world = self.world() if reuse_world else None
        try:
            return interpolate_on_large_data_grid(self.data_small, self.hdr_grid, hdr, self.order, world=world,
                                                  coords=self.coords)
        except ValueError as exc:    # WCS(hdr_shifted) rejected by wcslib's celset: the reference's worker dies
            if "Invalid coordinate transformation" in str(exc):
                raise LagKillsWorker(str(exc))
            raise

    def step(self, d_crval1, d_crval2, d_cdelt1, d_cdelt2, d_crota, reuse_world=True):
        """One lag -> Pearson r (`_step`, `alignment.py:509-542`)."""
        try:
            interp = self.reprojected(d_crval1, d_crval2, d_cdelt1, d_cdelt2, d_crota, reuse_world)
        except LagKillsWorker:
            return 0.0  # cube entry never written by the dead worker (np.zeros, alignment.py:635)
        return masked_pearson(self.data_large, interp)

    def flat_lags(self):
        r = self.refs
        g = np.meshgrid(r.lag_crval1, r.lag_crval2, r.lag_cdelt1, r.lag_cdelt2, r.lag_crota, indexing="ij")
        return [a.ravel() for a in g]

    def cube(self, reuse_world=True, lag_slice=None):
        l1, l2, l3, l4, l5 = self.flat_lags()
        n = l1.size
        out = np.zeros(n, dtype=np.float64)
        idx = range(n) if lag_slice is None else range(*lag_slice.indices(n))
        for i in idx:
            out[i] = self.step(l1[i], l2[i], l3[i], l4[i], l5[i], reuse_world)
        return out.reshape(self.shape[:5] + (1,)) if len(self.refs.lag_solar_r) == 1 else out


def hpc_cube(data_large, hdr_large, data_small, hdr_small, **kw):
    """Correlation cube [n_crval1, n_crval2, n_cdelt1, n_cdelt2, n_crota, n_solar_r] (float64)."""
    return HpcSearch(data_large, hdr_large, data_small, hdr_small, **kw).cube()


# ----------------------------------------------------------------------------------------------
# reference-structured multiprocessing run, for the CPU baseline timing (bench.py only)
# ----------------------------------------------------------------------------------------------
_G = {}


def _worker(args):
    lo, hi = args
    s = _G["search"]
    l1, l2, l3, l4, l5 = _G["lags"]
    out = np.empty(hi - lo, dtype=np.float64)
    for k, i in enumerate(range(lo, hi)):
        # reuse_world=False: like the reference, redo pixel->world of the common grid for every lag
        out[k] = s.step(l1[i], l2[i], l3[i], l4[i], l5[i], reuse_world=False)
    return lo, out


def cube_multiprocess(search: HpcSearch, n_workers: int, lag_indices=None):
    """Evaluate `lag_indices` (default: all) with `n_workers` forked processes, the way
    `alignment.py:667-744` fans chunks out (np.array_split of the flattened C-order lag list)."""
    lags = search.flat_lags()
    n = lags[0].size
    sel = np.arange(n) if lag_indices is None else np.asarray(lag_indices)
    sub = [a[sel] for a in lags]
    _G["search"], _G["lags"] = search, sub
    bounds = np.linspace(0, sel.size, n_workers + 1).astype(int)
    tasks = [(int(bounds[i]), int(bounds[i + 1])) for i in range(n_workers) if bounds[i + 1] > bounds[i]]
    out = np.zeros(sel.size, dtype=np.float64)
    ctx = mp.get_context("fork")
    with ctx.Pool(len(tasks)) as pool:
        for lo, part in pool.map(_worker, tasks):
            out[lo:lo + part.size] = part
    _G.clear()
    return out
