"""CPU tests of the host layer: FITS/WCS stand-ins, lag tables vs the oracle's shifted headers, results object
against the reference's golden cube, header writing, C-ABI symbol export, lag sharding with gloo."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_pair

GOLD = os.path.join(ROOT, "tests", "golden")


# ----------------------------------------------------------------------------------------------- fits / units
def test_fits_lite_round_trip(tmp_path):
    from euispice_coreg_b200._compat import fits_lite
    rng = np.random.default_rng(0)
    h = fits_lite.Header()
    h["CRVAL1"], h["CUNIT1"], h["FLAG"], h["N"], h["DATE-OBS"] = -100.25, "arcsec", True, 7, "2022-03-17T09:50:45.000"
    h["TINY"] = 1.23456789012345e-17
    prim = fits_lite.PrimaryHDU(rng.normal(size=(5, 7)).astype(np.float32), h)
    ext = fits_lite.ImageHDU(rng.integers(0, 100, (2, 3, 4)).astype(np.int16), h, name="WIN B")
    p = str(tmp_path / "x.fits")
    fits_lite.writeto(p, [prim, ext], overwrite=True)
    assert os.path.getsize(p) % 2880 == 0
    with fits_lite.open(p) as hdul:
        assert len(hdul) == 2
        assert np.array_equal(hdul[0].data, prim.data) and hdul[0].data.dtype == np.float32
        assert np.array_equal(hdul["WIN B"].data, ext.data) and np.array_equal(hdul[-1].data, ext.data)
        for k in ("CRVAL1", "CUNIT1", "FLAG", "N", "DATE-OBS", "TINY"):
            assert hdul[0].header[k] == h[k]
    with pytest.raises(OSError):
        fits_lite.writeto(p, [prim])


def test_units_match_astropy_conventions():
    from euispice_coreg_b200._compat import units
    assert units.convert(3600.0, "arcsec", "deg") == 1.0
    assert units.convert(1.0, "deg", "arcsec") == 3600.0
    assert units.factor("arcsec", "arcsec") == 1.0
    assert np.allclose(units.ang2pipi(np.array([190.0, -190.0, 180.0]), "deg"), [-170.0, 170.0, 180.0])
    assert np.allclose(units.ang2pipi(np.array([30.0, -30.0]), "arcsec"), [30.0, -30.0])
    with pytest.raises(ValueError):
        units.canon("parsec")


def test_timeutil():
    from euispice_coreg_b200._compat import timeutil
    assert timeutil.diff_seconds("2022-03-17T09:50:45.281", "2022-03-17T09:50:45.000") == pytest.approx(0.281)
    assert timeutil.diff_days("2022-03-18T09:50:45", "2022-03-17T09:50:45") == 1.0
    assert timeutil.from_seconds(timeutil.to_seconds("2022-03-17T09:50:45.281")) == "2022-03-17T09:50:45.281"


# ----------------------------------------------------------------------------------------------- lag tables
def _kernel_map(table_row, lng, lat, alpha_ref_deg):
    """numpy emulation of TanCoord::map (csrc/coreg_common.cuh) from the lag-table row."""
    s_da, c_da, s_d0, c_d0, m11, m12, m21, m22, x0, y0 = table_row
    p0 = np.sin(np.deg2rad(lat))
    a = np.deg2rad(lng) - np.deg2rad(alpha_ref_deg)
    p1, p2 = np.cos(np.deg2rad(lat)) * np.sin(a), np.cos(np.deg2rad(lat)) * np.cos(a)
    qs = p1 * c_da - p2 * s_da
    pc = p2 * c_da + p1 * s_da
    den = pc * c_d0 + p0 * s_d0
    en = p0 * c_d0 - pc * s_d0
    return m11 * qs / den + m12 * en / den + x0, m21 * qs / den + m22 * en / den + y0


@pytest.mark.parametrize("sem", ["reference", "intended"])
def test_lag_table_reproduces_oracle_shifted_headers(toy_pair, sem):
    """Host `tan_lag_table` + the kernel's coordinate formula == oracle `_shift_header` + wcslib-structured
    world_to_pixel, for CRVAL / CROTA / CDELT lags (SURVEY App. B1 semantics), within 1e-9 px."""
    from euispice_coreg_b200.hdrshift import engine
    from euispice_coreg_b200._compat.wcs import TanWcs
    from oracle import hpc, wcs_tan
    _, _, _, hs = load_pair(*toy_pair[:2])
    hs = dict(hs)
    hpc.check_and_create_pcij(hs)
    refs = hpc.Refs(hs, [0.0], [0.0], [0.0], [0.0], [0.0], None)
    lags = [(24.0, 6.0, 0.0, 0.0, 0.0), (-30.0, 12.5, 0.0, 0.0, 0.75), (3.0, -4.0, 0.004, 0.0, 0.0),
            (5.0, 5.0, 0.002, -0.003, -0.5)]
    d = [np.array(c, dtype=np.float64) for c in zip(*lags)]
    alpha = TanWcs.from_header(hs).crval1
    table, dead = engine.tan_lag_table(hs, refs, *d, alpha, sem)
    lng, lat = wcs_tan.extract_coordinates(hs)
    for k, lag in enumerate(lags):
        h = dict(hs)
        try:
            hpc.shift_header(h, refs, *lag, cdelt_mode=sem)
        except hpc.LagKillsWorker:
            assert dead[k] and sem == "reference"
            continue
        assert not dead[k]
        xo, yo = wcs_tan.WcsTan(h).world_to_pixel(lng, lat)
        x, y = _kernel_map(table[k], lng, lat, alpha)
        assert np.max(np.abs(x - xo)) < 1e-9 and np.max(np.abs(y - yo)) < 1e-9


def _homography_map(grid_row, lag_row, i, j):
    """numpy restatement of csrc `make_hom_grid` + `tan_homography_kernel` + `TanHom::map_half` (division form):
    common-grid pixel (i, j) -> 0-based pixel of the candidate header."""
    d2r = np.pi / 180.0

    def euler(dec, lonpole):
        sd, cd, sl, cl = np.sin(dec * d2r), np.cos(dec * d2r), np.sin(lonpole * d2r), np.cos(lonpole * d2r)
        return np.array([[-sd * cl, -sd * sl, cd], [sl, -cl, 0.0], [cd * cl, cd * sl, sd]])

    def fmat(r):
        return np.array([[r[2] * r[4], r[2] * r[5]], [r[3] * r[6], r[3] * r[7]]]) * d2r

    f = fmat(grid_row)
    c = f @ (1.0 - grid_row[0:2])
    cmat = np.array([[-f[1, 0], -f[1, 1], -c[1]], [f[0, 0], f[0, 1], c[0]], [0.0, 0.0, 1.0]])
    a = (grid_row[8] - lag_row[8]) * d2r
    rz = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
    r = euler(lag_row[9], lag_row[10]).T @ rz @ euler(grid_row[9], grid_row[10]) @ cmat
    inv = np.linalg.inv(fmat(lag_row))
    hx = inv[0, 0] * r[1] - inv[0, 1] * r[0]
    hy = inv[1, 0] * r[1] - inv[1, 1] * r[0]
    v = np.stack([i, j, np.ones_like(i)])
    den = np.tensordot(r[2], v, 1)
    return (np.tensordot(hx, v, 1) / den + lag_row[0] - 1.0, np.tensordot(hy, v, 1) / den + lag_row[1] - 1.0, den)


@pytest.mark.parametrize("sem", ["reference", "intended"])
def test_homography_of_candidate_headers_reproduces_oracle_world_to_pixel(toy_pair, sem):
    """Two gnomonic projections of one sphere are related by a plane homography: the per-lag 3x3 matrix the fast
    kernel derives from `tan_wcs_table` rows maps common-grid pixels where the oracle's pixel->world->pixel
    round trip (wcslib-structured, per lag) puts them, within 1e-9 px, for CRVAL / CROTA / CDELT lags."""
    from euispice_coreg_b200.hdrshift import engine
    from oracle import hpc, wcs_tan
    _, _, _, hs = load_pair(*toy_pair[:2])
    hs = dict(hs)
    hpc.check_and_create_pcij(hs)
    refs = hpc.Refs(hs, [0.0], [0.0], [0.0], [0.0], [0.0], None)
    lags = [(24.0, 6.0, 0.0, 0.0, 0.0), (-30.0, 12.5, 0.0, 0.0, 0.75), (3.0, -4.0, 0.004, 0.0, 0.0),
            (5.0, 5.0, 0.002, -0.003, -0.5), (3000.0, -2000.0, 0.0, 0.0, 10.0)]
    d = [np.array(c, dtype=np.float64) for c in zip(*lags)]
    table, dead = engine.tan_wcs_table(hs, refs, *d, sem)
    grid, _ = engine.tan_wcs_table(hs, refs, *[np.zeros(1)] * 5, sem)
    assert table.shape == (len(lags), 11)
    lng, lat = wcs_tan.extract_coordinates(hs)
    jj, ii = np.mgrid[0:lng.shape[0], 0:lng.shape[1]].astype(np.float64)
    for k, lag in enumerate(lags):
        h = dict(hs)
        try:
            hpc.shift_header(h, refs, *lag, cdelt_mode=sem)
        except hpc.LagKillsWorker:
            assert dead[k] and sem == "reference"
            continue
        xo, yo = wcs_tan.WcsTan(h).world_to_pixel(lng, lat)
        x, y, den = _homography_map(grid[0], table[k], ii, jj)
        assert np.max(np.abs(x - xo)) < 1e-9 and np.max(np.abs(y - yo)) < 1e-9
        assert np.all(den > 0.9)    # e = 1 - den stays tiny: the small-angle reciprocal series applies


def test_flat_lag_grid_and_shard_bounds_match_array_split():
    from euispice_coreg_b200.hdrshift import engine
    d = engine.flat_lag_grid([1, 2, 3], [10, 20], [0], [0], [0.0, 0.5])
    assert d[0].tolist() == [1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3]
    assert d[4].tolist() == [0.0, 0.5] * 6
    for n, w in ((3600, 8), (25, 2), (7, 4), (1, 8), (10, 3)):
        chunk, b = engine.shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(hi - lo <= chunk for lo, hi in b)


def test_carrington_offsets_and_vectors_match_oracle(toy_pair):
    from euispice_coreg_b200.hdrshift.engine import LagSearchEngine
    from oracle.carrington import CarringtonTransform
    _, _, _, hs = load_pair(*toy_pair[:2])
    hs = dict(hs)
    d1, d2 = np.array([24.0, -3.0]), np.array([6.0, 2.0])
    x0, y0 = LagSearchEngine.carrington_offset(hs, hs["CRVAL1"] + d1, hs["CRVAL2"] + d2, hs["CROTA"])
    for k in range(2):
        t = CarringtonTransform(dict(hs, CRVAL1=hs["CRVAL1"] + d1[k], CRVAL2=hs["CRVAL2"] + d2[k]), 1.004)
        assert t.x == x0[k] and t.y == y0[k]
    sl, cl, sb, cb = LagSearchEngine.carrington_vectors((248.0, 252.0), (-4.0, 0.0), (12, 10), hs["CRLN_OBS"])
    lon = np.linspace(248.0, 252.0, 12, dtype=np.float32)
    lat = np.linspace(-4.0, 0.0, 10, dtype=np.float32)
    assert np.array_equal(sl, np.sin(np.radians(lon) - np.radians(hs["CRLN_OBS"]))) and sl.dtype == np.float64
    assert np.array_equal(sb, np.sin(np.radians(lat)).astype(np.float64)) and np.radians(lat).dtype == np.float32


# ----------------------------------------------------------------------------------------------- results
def test_alignment_results_on_reference_golden_cube():
    """hdrshift/test/test_AlignmentResults.py:35-126: 11x6 cube, max 0.9700199 at (9, 1) i.e. lag (24, 6); the
    reference expects the fitted peak (9.33682107, 1.42187891) +-1e-2. Our restatement of the reference's current
    fit gives (9.34903, 1.41708) on the printed 8-digit cube (insensitive to the print truncation, 2e-6), so the gate
    is 2e-2 on the fitted peak and exact on the arg-max."""
    from euispice_coreg_b200.hdrshift import AlignmentResults
    z = np.load(os.path.join(GOLD, "results_golden.npz"))
    r = AlignmentResults(corr=z["cube_a"], lag_crval1=z["a_lag_crval1"], lag_crval2=z["a_lag_crval2"], lag_cdelt1=None,
                         lag_cdelt2=[0], lag_crota=[0.75], unit_lag="arcsec")
    assert tuple(int(v) for v in r.max_index) == (9, 1, 0, 0, 0, 0)
    assert (z["a_lag_crval1"][r.max_index[0]], z["a_lag_crval2"][r.max_index[1]]) == (24, 6)
    assert abs(r.shift_pixels[0] - z["a_peak"][0]) < 2e-2 and abs(r.shift_pixels[1] - z["a_peak"][1]) < 2e-2
    assert r.shift_arcsec[0] == pytest.approx(15 + r.shift_pixels[0]) and r.shift_arcsec[4] == 0.75
    rb = AlignmentResults(corr=z["cube_b"], lag_crval1=z["b_lag_crval1"], lag_crval2=z["b_lag_crval2"],
                          lag_cdelt1=None, lag_cdelt2=[0], lag_crota=[0.75], unit_lag="arcsec")
    assert tuple(int(v) for v in rb.max_index[:2]) == (2, 1)      # Carrington golden: lag (24, 7)
    assert 22.0 < rb.shift_arcsec[0] < 26.0 and 5.0 < rb.shift_arcsec[1] < 9.0


def test_alignment_results_quirks_kept():
    """Reference quirks (AlignmentResults.py:224-239): offsets of -2 wrap around below index 0, so a lag axis of
    length 1 raises IndexError exactly like the reference; an all-NaN cube raises ValueError (nanargmax)."""
    from euispice_coreg_b200.hdrshift import AlignmentResults
    c = np.zeros((1, 2, 1, 1, 1, 1))
    c[0, 1] = 0.5
    with pytest.raises(IndexError):
        AlignmentResults(corr=c, lag_crval1=[3.0], lag_crval2=[1.0, 2.0], lag_cdelt1=None, lag_cdelt2=None,
                         lag_crota=None, unit_lag="arcsec")
    x, y = np.meshgrid(np.arange(4.0), np.arange(5.0), indexing="ij")
    c = (0.9 * np.exp(-((x - 1.3) ** 2 + (y - 2.2) ** 2) / 8.0)).reshape(4, 5, 1, 1, 1, 1)
    r = AlignmentResults(corr=c, lag_crval1=np.arange(4.0) * 2, lag_crval2=np.arange(5.0), lag_cdelt1=None,
                         lag_cdelt2=None, lag_crota=None, unit_lag="arcsec")
    assert tuple(int(v) for v in r.max_index[:2]) == (1, 2) and len(r.shift_arcsec) == 5
    with pytest.raises(ValueError):
        r.write_corrected_fits([0], "/tmp/nowhere.fits")
    c = np.full((3, 3, 1, 1, 1, 1), np.nan)
    with pytest.raises(ValueError):
        AlignmentResults(corr=c, lag_crval1=[1, 2, 3], lag_crval2=[1, 2, 3], lag_cdelt1=None, lag_cdelt2=None,
                         lag_crota=None, unit_lag="arcsec")


def test_correct_pointing_header_and_write(tmp_path, toy_pair):
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.utils.Util import AlignCommonUtil
    h0 = fits_lite.open(toy_pair[1])[0].header
    h = h0.copy()
    AlignCommonUtil.correct_pointing_header(h, lag_cdelt1=0.01, lag_cdelt2=None, lag_crota=0.5, lag_crval1=24.0,
                                            lag_crval2=-6.0)
    assert h["CRVAL1"] == h0["CRVAL1"] + 24.0 and h["CRVAL2"] == h0["CRVAL2"] - 6.0
    assert h["CDELT1"] == h0["CDELT1"] + 0.01 and h["CROTA"] == h0["CROTA"] + 0.5
    th, lam = np.deg2rad(h0["CROTA"] + 0.5), h0["CDELT2"] / (h0["CDELT1"] + 0.01)
    assert h["PC1_1"] == np.cos(th) and h["PC1_2"] == -lam * np.sin(th) and h["PC2_1"] == np.sin(th) / lam
    hd = h0.copy()
    for k in ("CRVAL1", "CRVAL2", "CDELT1", "CDELT2"):
        hd[k] = hd[k] / 3600.0
    hd["CUNIT1"] = hd["CUNIT2"] = "deg"
    AlignCommonUtil.correct_pointing_header(hd, None, None, None, 36.0, 0.0)
    assert hd["CRVAL1"] == h0["CRVAL1"] / 3600.0 + 36.0 * (1.0 / 3600.0)
    out = str(tmp_path / "o.fits")
    AlignCommonUtil.write_corrected_fits(toy_pair[1], [-1], out, corr=None, shift_arcsec=[1.0, 2.0, None, None, None])
    o = fits_lite.open(out)[0]
    assert o.header["CRVAL1"] == h0["CRVAL1"] + 1.0 and o.data.dtype == np.float32
    with pytest.raises(ValueError):
        AlignCommonUtil.write_corrected_fits(toy_pair[1], ["NOPE"], out, corr=None, shift_arcsec=[0, 0, 0, 0, 0])


def test_alignment_host_preparation(toy_pair):
    """Constructor defaults, PC creation, unit handling and error behaviour mirror alignment.py:47-140, 580-611."""
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.hdrshift import Alignment
    a = Alignment(toy_pair[0], toy_pair[1], lag_crval1=[1.0], lag_crval2=None, lag_cdelt1=None, lag_cdelt2=None,
                  lag_crota=None)
    assert a.lag_crval2.tolist() == [0.0] and a.use_pcij is False and a.order == 2 and a.counts == 40
    h = fits_lite.Header({"CDELT1": 2.0, "CDELT2": 4.0, "CROTA2": 30.0})
    a._check_ant_create_pcij_matrix(h)
    assert h["PC1_2"] == pytest.approx(-2.0 * np.sin(np.pi / 6)) and h["PC2_1"] == pytest.approx(0.5 * np.sin(np.pi / 6))
    assert h["CROTA"] == pytest.approx(30.0)
    with pytest.raises(ValueError):
        a._check_ant_create_pcij_matrix(fits_lite.Header({"CDELT1": 1.0, "CDELT2": 1.0}))
    a.force_crota_0 = True
    h = fits_lite.Header({"CDELT1": 1.0, "CDELT2": 1.0})
    a._check_ant_create_pcij_matrix(h)
    assert (h["PC1_1"], h["PC1_2"], h["CROTA"]) == (1.0, 0.0, 0.0)
    # HPLN-TAN files are not Carrington maps: no CPU fallback without a device, a -CAR projection demanded with one
    from euispice_coreg_b200._ext import CoregLibraryError
    with pytest.raises((NotImplementedError, CoregLibraryError)):
        a.align_using_initial_carrington()


def test_product_does_not_import_oracle_and_fails_without_library(tmp_path):
    """No CPU fallback: nothing under the package imports oracle/, and compute calls raise when the .so is missing."""
    pkg = os.path.join(ROOT, "euispice_coreg_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
    from euispice_coreg_b200 import _ext
    with pytest.raises(_ext.CoregLibraryError):
        _ext.load(str(tmp_path / "missing.so"))


# ----------------------------------------------------------------------------------------------- C ABI
def test_c_abi_library_exports_every_declared_symbol():
    from euispice_coreg_b200 import _ext
    hdr = open(os.path.join(ROOT, "include", "coreg_b200.h")).read()
    declared = set(re.findall(r"\b(coreg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    if not os.path.exists(_ext.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_ext.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_ext.EXPORTED_SYMBOLS)
    assert ctypes.sizeof(_ext.CoregTanWcs) == 88 and ctypes.sizeof(_ext.CoregCarrington) == 48
    lib.coreg_version.restype = ctypes.c_int
    assert lib.coreg_version() >= 100
    lib.coreg_lag_corr_workspace_bytes.restype = ctypes.c_size_t
    lib.coreg_lag_corr_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64]
    # rolling kernel: [32 x 43 tiles of 64x48][8 warps] record rows x (3 + 3 doubles per lag) + 3 constants per row
    # + a 4-byte mask per (tile, lag), then one 96-byte homography row per lag
    rows = 32 * 43 * 8
    assert lib.coreg_lag_corr_workspace_bytes(2048, 2048, 3600) == (rows * 3600 * 48 + rows * 24 + 32 * 43 * 3600 * 4
                                                                    + 3600 * 96)
    assert lib.coreg_lag_corr_workspace_bytes(0, 5, 5) == 0


def test_c_abi_argument_validation_needs_no_device():
    """Error behaviour at the boundary: bad arguments come back as negative COREG_E* codes with a message, before any
    CUDA call (these calls touch no device memory, so they run on the CPU box)."""
    from euispice_coreg_b200 import _ext
    lib = _ext.load()
    null, fake = None, ctypes.c_void_p(0x1000)          # never dereferenced: validation fails first
    EINVAL, ENOMEM = -1, -3
    assert ctypes.sizeof(ctypes.c_double) * _ext.LAG_CAR_DOUBLES == 128

    def msg():
        return lib.coreg_last_error().decode()

    rc = lib.coreg_hpc_lag_corr(null, null, 0, 8, 8, 8, 8, null, null, 4, 2, null, null, 0, null, null, 0, null)
    assert rc == EINVAL and "null pointer" in msg()
    rc = lib.coreg_car_lag_corr(fake, fake, 7, 8, 8, 8, 8, fake, fake, 4, 2, fake, fake, 1 << 30, fake, null, 0, null)
    assert rc == EINVAL and "small_dtype" in msg()
    rc = lib.coreg_hpc_lag_corr(fake, fake, _ext.F64, 8, 8, 8, 8, fake, fake, 4, 2, fake, fake, 16, fake, null, 0, null)
    assert rc == ENOMEM and "workspace" in msg()
    rc = lib.coreg_car_pix2world(null, 8, 8, fake, fake, null)
    assert rc == EINVAL and "CoregLagCar" in msg()
    singular = np.zeros(_ext.LAG_CAR_DOUBLES)
    rc = lib.coreg_car_pix2world(singular.ctypes.data_as(ctypes.c_void_p), 8, 8, fake, fake, null)
    assert rc == EINVAL and "singular" in msg()
    tan = _ext.CoregTanWcs(1, 1, 0.0, 1.0, 1, 0, 0, 1, 0, 0, 180)     # CDELT1 = 0
    rc = lib.coreg_tan_pix2world(ctypes.byref(tan), 8, 8, 0, fake, fake, null)
    assert rc == EINVAL and "singular" in msg()
    # pixel shift: `_check_boundaries` (pxlshift/alignment_pixels.py:150-156) on the host lag arrays
    dx, dy = (ctypes.c_int * 2)(0, 9), (ctypes.c_int * 1)(0)
    need = lib.coreg_pixel_shift_workspace_bytes(10, 10, 2, 1, 1)
    assert need == 1 * 2 * 64 + (2 + 1 + 2) * 4
    rc = lib.coreg_pixel_shift_corr(fake, 20, 20, fake, 1, 10, 10, 5, 5, dx, 2, dy, 1, fake, fake, need, fake, null, null)
    assert rc == EINVAL and msg() == "too large shift : outside FSI"
    rc = lib.coreg_pixel_shift_corr(fake, 20, 20, fake, 1, 10, 10, 5, 5, dx, 2, dy, 1, fake, fake, need - 1, fake, null,
                                    null)
    assert rc == ENOMEM
    rc = lib.coreg_map_coordinates(fake, 5, 4, 4, fake, fake, 10, 2, 0.0, fake, _ext.F64, null)
    assert rc == EINVAL and "dtype" in msg()
    good = _ext.CoregTanWcs(1, 1, 1.0, 1.0, 1, 0, 0, 1, 0, 0, 180)
    rc = lib.coreg_hpc_cut(ctypes.byref(good), 8, 8, ctypes.byref(tan), fake, _ext.F32, 8, 8, 0, 0, 2, fake, null)
    assert rc == EINVAL and "singular" in msg()
    rc = lib.coreg_hpc_cut(ctypes.byref(good), 8, 8, ctypes.byref(good), null, _ext.F32, 8, 8, 0, 0, 2, fake, null)
    assert rc == EINVAL and "null" in msg()
    rc = lib.coreg_hpc_cut(ctypes.byref(good), 8, 8, ctypes.byref(good), fake, 5, 8, 8, 0, 0, 2, fake, null)
    assert rc == EINVAL and "dtype" in msg()


# ----------------------------------------------------------------------------------------------- multi-rank
_GLOO = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from euispice_coreg_b200.hdrshift import engine
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
for n in (25, 7, 2, 1, 3600):
    chunk, bounds = engine.shard_bounds(n, world)
    lo, hi = bounds[rank]
    local = torch.full((chunk,), float("nan"), dtype=torch.float64)
    local[:hi - lo] = torch.arange(lo, hi, dtype=torch.float64) * 0.5      # stands in for the rank's r values
    full = engine.gather_slices(local, n, chunk)
    assert full.shape == (n,) and torch.equal(full, torch.arange(n, dtype=torch.float64) * 0.5), (n, full)
# frame sharding of a sequence (hdrshift/sequence.py): round robin, one all-gather of the cubes -- also with more
# ranks than frames (a rank without frames must still contribute a block of the common shape)
from euispice_coreg_b200.hdrshift import sequence
for n_frames, n_lags in ((5, 12), (1, 7), (world + 3, 4), (world - 1, 6), (0, 3)):
    mine = sequence.frames_of_rank(n_frames, rank, world)
    cubes = {{k: np.full(n_lags, 100.0 * k) + np.arange(n_lags) for k in mine}}
    allc = sequence.gather_frame_cubes(cubes, n_frames, n_lags, torch.device("cpu"))
    assert sorted(allc) == list(range(n_frames)), (n_frames, sorted(allc))
    for k in range(n_frames):
        assert np.array_equal(allc[k], np.full(n_lags, 100.0 * k) + np.arange(n_lags)), (n_frames, k)
dist.barrier()
if rank == 0:
    print("GLOO_OK")
"""


@pytest.mark.parametrize("world", [2, 3])
def test_lag_and_frame_sharding_gather_gloo(tmp_path, world):
    script = tmp_path / "gloo_worker.py"
    script.write_text(_GLOO.format(root=ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29531 + world), str(script)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_spice_solar_rotation_cdelt1_correction():
    """`extend_pixel_size=True` (alignment_spice.py:223-248): CDELT1 shrinks by dt * helioprojective rotation rate *
    cos(phi); checked against the formula evaluated by hand for a disc-centre and a near-limb pointing."""
    from euispice_coreg_b200.hdrshift.alignment_spice import AlignmentSpice
    from euispice_coreg_b200.utils import Util
    assert abs(Util.diff_rot(0.0, "EIT 171") - np.deg2rad((14.56 - 360 / 25.38) / 86400)) < 1e-20
    assert Util.AlignEUIUtil.diff_rot(0.3) == Util.diff_rot(0.3, "EIT 195")
    for crval1, unit in ((0.0, "arcsec"), (800.0, "arcsec"), (0.2, "deg")):
        a = AlignmentSpice.__new__(AlignmentSpice)
        a.hdr_large = {"WAVELNTH": 174}
        cd = 4.0 if unit == "arcsec" else 4.0 / 3600
        a.hdr_small = {"SOLAR_B0": -2.0, "RSUN_REF": 6.957e8, "DSUN_OBS": 5.7e10, "CRVAL1": crval1, "CUNIT1": unit,
                       "CDELT1": cd}
        dt = 21.3
        a._correct_solar_rotation(dt)
        omega = np.deg2rad(360 / 25.38 / 86400) + np.deg2rad(
            (14.56 - 360 / 25.38 - 2.65 * np.sin(np.deg2rad(-2.0)) ** 2 + 0.96 * np.sin(np.deg2rad(-2.0)) ** 4) / 86400)
        rate = np.rad2deg(1.004 * omega * 6.957e8 / (5.7e10 - 1.004 * 6.957e8)) * 3600          # arcsec / s
        alpha = np.deg2rad(crval1 / 3600 if unit == "arcsec" else crval1)
        phi = np.arcsin((5.7e10 - 1.004 * 6.957e8) / (1.004 * 6.957e8) * np.sin(alpha))
        want = 4.0 - dt * rate * np.cos(phi)
        got = a.hdr_small["CDELT1"] * (1.0 if unit == "arcsec" else 3600.0)
        assert abs(got - want) < 1e-12 and 3.8 < got < 4.0, (got, want)


def test_fits_lite_lazy_payload_window_and_raw_view(tmp_path):
    """Image payloads stay in the memory-mapped file until used: `read_window` converts a sub-window only,
    `raw_big_endian` hands a float32 payload over as stored (for the device-side byte swap); both agree with `.data`,
    also for scaled integer images; overwriting a file that is still mapped is safe."""
    from euispice_coreg_b200._compat import fits_lite
    rng = np.random.default_rng(3)
    img = rng.normal(500.0, 300.0, (37, 53)).astype(np.float32)
    img[4, 7] = np.nan
    p = str(tmp_path / "a.fits")
    fits_lite.writeto(p, [fits_lite.PrimaryHDU(img, fits_lite.Header())])
    h = fits_lite.open(p)[0]
    assert h.shape == (37, 53)
    raw = h.raw_big_endian()
    assert raw.dtype == np.dtype(">f4") and not raw.flags.writeable
    assert np.array_equal(raw.astype(np.float32), img, equal_nan=True)
    assert np.array_equal(h.read_window(3, 20, 5, 41), img[3:20, 5:41], equal_nan=True)
    assert np.array_equal(h.data, img, equal_nan=True) and h.data.dtype == np.float32 and h.data.dtype.isnative
    assert h.raw_big_endian() is None                     # materialised: the raw view is gone
    assert np.array_equal(h.read_window(0, 2, 0, 3), img[:2, :3], equal_nan=True)
    # 16-bit integers with BSCALE / BZERO: no raw float view, window == slice of the scaled data
    ints = rng.integers(-3000, 3000, (20, 30)).astype(np.int16)
    hdr = fits_lite.Header()
    hdr["BSCALE"], hdr["BZERO"] = 0.5, 100.0
    q = str(tmp_path / "b.fits")
    fits_lite.writeto(q, [fits_lite.PrimaryHDU(ints, hdr)])
    hq = fits_lite.open(q)[0]
    assert hq.raw_big_endian() is None
    win = hq.read_window(2, 9, 4, 11)
    assert win.dtype == hq.data.dtype and np.array_equal(win, hq.data[2:9, 4:11])
    stored = ints.astype(">i2")
    conv = fits_lite.ImageHDU._convert
    assert conv(stored, 0.5, 100.0).dtype == np.float32        # astropy's promotion rule for <= 16-bit integers
    assert np.array_equal(conv(stored[2:9, 4:11], 0.5, 100.0), conv(stored, 0.5, 100.0)[2:9, 4:11])
    assert conv(stored, 1, 32768).dtype == np.uint16           # unsigned-integer convention
    # BLANK pixels of an integer image that is scaled to float become NaN (astropy's behaviour); unscaled stay integers
    blank = int(stored[5, 6])
    out = conv(stored, 0.5, 100.0, blank)
    assert np.isnan(out[5, 6]) and np.array_equal(np.isnan(out), stored == blank)
    assert conv(stored, 1, 0, blank).dtype == np.int16
    # overwrite while mapped
    h2 = fits_lite.open(p)
    fits_lite.writeto(p, [fits_lite.PrimaryHDU(img * 2, h2[0].header)], overwrite=True)
    assert np.array_equal(h2[0].data, img, equal_nan=True)
    assert np.array_equal(fits_lite.open(p)[0].data, img * 2, equal_nan=True)


def test_large_image_window_covers_every_cut_coordinate(toy_pair):
    """`LagSearchEngine.large_window`: the one-time cut of the large image onto the small grid touches only a window of
    it; the window found from the grid's four edges contains every coordinate of the full map (oracle) plus the
    spline support."""
    from euispice_coreg_b200._compat.wcs import TanWcs
    from euispice_coreg_b200.hdrshift.engine import LagSearchEngine
    from oracle import wcs_tan
    dl, hl, ds, hs = load_pair(*toy_pair[:2])
    from oracle.hpc import check_and_create_pcij
    check_and_create_pcij(hl)
    check_and_create_pcij(hs)
    win = LagSearchEngine.large_window(TanWcs.from_header(hl), TanWcs.from_header(hs), dl.shape)
    x, y = wcs_tan.extract_coordinates_pixels(hs, hl)
    x0, x1, y0, y1 = win
    inside = (x >= 0) & (x <= dl.shape[1] - 1) & (y >= 0) & (y <= dl.shape[0] - 1)
    assert inside.any()
    # order-3 support reaches 2 pixels beyond floor(x); the window keeps 4
    assert x0 <= max(0, np.floor(x[inside].min()) - 2) and x1 >= min(dl.shape[1], np.ceil(x[inside].max()) + 3)
    assert y0 <= max(0, np.floor(y[inside].min()) - 2) and y1 >= min(dl.shape[0], np.ceil(y[inside].max()) + 3)
    assert (x1 - x0) * (y1 - y0) < 0.5 * dl.size
    # the window comes from the grid's four CORNERS (a projective map is monotonic along straight edges and has no interior
    # extrema): rotated, rescaled, strongly offset small grids -- every coordinate of the full map stays inside
    rng = np.random.default_rng(4)
    checked = 0
    for k in range(16):
        h2 = dict(hs)
        rot = np.radians(rng.uniform(-180, 180))
        h2["PC1_1"], h2["PC1_2"], h2["PC2_1"], h2["PC2_2"] = np.cos(rot), -np.sin(rot), np.sin(rot), np.cos(rot)
        h2["CDELT1"] = hs["CDELT1"] * rng.uniform(0.3, 1.5)
        h2["CDELT2"] = hs["CDELT2"] * rng.uniform(0.3, 1.5)
        h2["CRVAL1"] = hs["CRVAL1"] + rng.uniform(-400, 400) * (k % 3)
        h2["CRVAL2"] = hs["CRVAL2"] + rng.uniform(-400, 400) * (k % 3)
        win = LagSearchEngine.large_window(TanWcs.from_header(hl), TanWcs.from_header(h2), dl.shape)
        x, y = wcs_tan.extract_coordinates_pixels(h2, hl)
        if win is None:      # no overlap worth cropping to
            continue
        checked += 1
        x0, x1, y0, y1 = win
        assert x0 <= max(0, np.floor(x.min()) - 2) and x1 >= min(dl.shape[1], np.ceil(x.max()) + 3), k
        assert y0 <= max(0, np.floor(y.min()) - 2) and y1 >= min(dl.shape[0], np.ceil(y.max()) + 3), k
    assert checked >= 5


def test_offset_patch_order_properties():
    """Lag order of the Carrington-frame kernel (`engine.offset_patch_order`): every lag gets its own slot; a block of
    256 consecutive slots holds one 16 x 16 patch of the CRVAL grid; a warp's 32 slots hold 16 x 2 lags; patches
    start at the list's own first indices (a rank's slice has no ragged leading patch); lags of different groups never
    share a patch; a list with repeated (i1, i2) pairs is left in its own order."""
    from euispice_coreg_b200.hdrshift.engine import OFFSET_CHUNK, offset_patch_order
    i1, i2 = (a.ravel() for a in np.meshgrid(np.arange(120), np.arange(120), indexing="ij"))
    slot, n_slots = offset_patch_order(i1, i2)
    assert n_slots % OFFSET_CHUNK == 0 and n_slots == 64 * 256 and np.unique(slot).size == slot.size
    inv = np.full(n_slots, -1)
    inv[slot] = np.arange(slot.size)
    for blk in range(0, n_slots, OFFSET_CHUNK):
        k = inv[blk:blk + OFFSET_CHUNK]
        k = k[k >= 0]
        assert np.ptp(i1[k]) < 16 and np.ptp(i2[k]) < 16
        for w in range(blk, blk + OFFSET_CHUNK, 32):
            kw = inv[w:w + 32]
            kw = kw[kw >= 0]
            if kw.size:
                assert np.ptp(i1[kw]) < 16 and np.ptp(i2[kw]) < 2
    # the second half of the grid (what rank 1 of 2 gets): 60 rows -> 4 patch columns, not 5
    half = i1 >= 60
    _, n_half = offset_patch_order(i1[half], i2[half])
    assert n_half == 4 * 8 * 256
    # groups (e.g. CDELT indices) keep apart
    g = np.repeat([0, 1], 8)
    s, n = offset_patch_order(np.tile(np.arange(8), 2), np.zeros(16, int), g)
    assert n == 512 and set(s[:8] // 256) == {0} and set(s[8:] // 256) == {1}
    # repeated pairs: identity
    s, n = offset_patch_order(np.zeros(5, int), np.zeros(5, int))
    assert n == 5 and s.tolist() == [0, 1, 2, 3, 4]
