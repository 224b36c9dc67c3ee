"""Carrington maps as inputs (CRLN-CAR / CRLT-CAR): `Alignment.align_using_initial_carrington`
(`hdrshift/alignment.py:344-399`), SURVEY.md section 8f rank 4.

CPU part: the wcslib-structured oracle (`oracle/wcs_car.py`: celset, sphx2s / sphs2x with Euler angles) against the
product's host formulation (`_compat/wcs.py`: one sphere rotation per header) -- two independent restatements of the
same published algorithm -- and the oracle search recovering the pointing error planted in the synthetic pair.
GPU part (`-m gpu`): the CUDA path through the C ABI against the oracle, r within 1e-6, arg-max identical.
"""
import numpy as np
import pytest

from conftest import load_pair

R_TOL = 1e-6  # tolerance stated by BASELINE.json north_star


def _hdr(crval=(250.0, 0.0), crota=0.0, n=(120, 80), cdelt=(0.05, 0.04), lonpole=None):
    c, s = np.cos(np.deg2rad(crota)), np.sin(np.deg2rad(crota))
    lam = cdelt[1] / cdelt[0]
    h = {"NAXIS1": n[0], "NAXIS2": n[1], "CTYPE1": "CRLN-CAR", "CTYPE2": "CRLT-CAR", "CUNIT1": "deg", "CUNIT2": "deg",
         "CRPIX1": (n[0] + 1) / 2, "CRPIX2": (n[1] + 1) / 2, "CDELT1": cdelt[0], "CDELT2": cdelt[1],
         "CRVAL1": crval[0], "CRVAL2": crval[1], "PC1_1": c, "PC1_2": -lam * s, "PC2_1": s / lam, "PC2_2": c,
         "CROTA": crota}
    if lonpole is not None:
        h["LONPOLE"] = lonpole
    return h


def _apply_row(row, lng, lat):
    """numpy evaluation of one CoregLagCar row (what `car_map_unit` does on the device)."""
    r = row[:9].reshape(3, 3)
    lng, lat = np.deg2rad(lng), np.deg2rad(lat)
    c = np.stack([np.cos(lat) * np.cos(lng), np.cos(lat) * np.sin(lng), np.sin(lat)])
    v = np.tensordot(r, c, 1)
    phi = np.rad2deg(np.arctan2(v[1], v[0]))
    theta = np.rad2deg(np.arctan2(v[2], np.hypot(v[0], v[1])))
    return row[9] * phi + row[10] * theta + row[13], row[11] * phi + row[12] * theta + row[14]


CASES = [((250.0, 0.0), 0.0, None), ((250.0, 0.013), 0.0, None), ((250.0, -0.013), 0.0, None),
         ((10.0, 35.0), 2.0, None), ((-30.0, -20.0), -1.0, None), ((250.0, 1.0), 3.0, 0.0),
         ((250.0, -1.0), 0.0, 180.0), ((0.0, 0.0), 0.5, None), ((359.9, 60.0), 0.0, None)]


@pytest.mark.parametrize("crval,crota,lonpole", CASES)
def test_car_rotation_form_matches_wcslib_structured_oracle(crval, crota, lonpole):
    from euispice_coreg_b200._compat.wcs import CarWcs
    from oracle.wcs_car import WcsCar
    h = _hdr(crval, crota, lonpole=lonpole)
    wo, w = WcsCar(h), CarWcs.from_header(h)
    x, y = np.meshgrid(np.arange(120.0), np.arange(80.0))
    lng, lat = wo.pixel_to_world(x, y)
    px, py = wo.world_to_pixel(lng, lat)
    assert np.max(np.abs(px - x)) < 1e-10 and np.max(np.abs(py - y)) < 1e-10          # oracle round trip
    qx, qy = _apply_row(w.lag_row(), lng, lat)
    assert np.max(np.abs(qx - px)) < 1e-10 and np.max(np.abs(qy - py)) < 1e-10        # product form == oracle
    hl, hb = w.pixel_to_world(x, y)
    assert np.max(np.abs((hl - lng + 180.0) % 360.0 - 180.0)) < 1e-11 and np.max(np.abs(hb - lat)) < 1e-11
    # a candidate header of the search: CRVAL shifted, other keywords untouched
    h2 = dict(h, CRVAL1=crval[0] + 0.02, CRVAL2=crval[1] - 0.03)
    rows, bad = CarWcs.lag_rows(w.crval1 + 0.02, w.crval2 - 0.03, w.cdelt1, w.cdelt2, w.pc11, w.pc12, w.pc21, w.pc22,
                                w.crpix1, w.crpix2, w.lonpole, w.latpole)
    assert not bad[0]
    ax, ay = WcsCar(h2).world_to_pixel(lng, lat)
    bx, by = _apply_row(rows[0], lng, lat)
    assert np.max(np.abs(ax - bx)) < 1e-10 and np.max(np.abs(ay - by)) < 1e-10


def test_car_equatorial_reference_is_the_linear_map():
    """CRVAL2 = 0: the native pole is the celestial pole and the projection is lng = CRVAL1 + CDELT1 (p - CRPIX1)."""
    from oracle.wcs_car import WcsCar
    h = _hdr((250.0, 0.0))
    x, y = np.meshgrid(np.arange(120.0), np.arange(80.0))
    lng, lat = WcsCar(h).pixel_to_world(x, y)
    assert np.max(np.abs(lng - (250.0 + 0.05 * (x + 1 - 60.5)))) < 1e-12
    assert np.max(np.abs(lat - 0.04 * (y + 1 - 40.5))) < 1e-13


def test_car_invalid_candidate_is_flagged_like_wcslib():
    """An explicit LONPOLE = 0 admits no native pole once CRVAL2 < 0: wcslib's celset fails, the reference's worker
    dies; the host table flags the lag (its cube entry is 0.0 like every never-written entry)."""
    from euispice_coreg_b200._compat.wcs import CarWcs
    from oracle.wcs_car import WcsCar
    h = _hdr((250.0, -0.5), lonpole=0.0)
    with pytest.raises(ValueError):
        WcsCar(h)
    w = CarWcs.from_header(_hdr((250.0, 0.5), lonpole=0.0))
    rows, bad = CarWcs.lag_rows([250.0, 250.0], [0.5, -0.5], w.cdelt1, w.cdelt2, w.pc11, w.pc12, w.pc21, w.pc22,
                                w.crpix1, w.crpix2, w.lonpole, w.latpole)
    assert list(bad) == [False, True] and np.isnan(rows[1]).all() and np.isfinite(rows[0]).all()


@pytest.fixture(scope="module")
def car_pair(tmp_path_factory):
    from euispice_coreg_b200._synth.carmaps import make_car_pair
    d = tmp_path_factory.mktemp("car")
    return make_car_pair(str(d))


# arcsec (the default unit_lag); headers are in degrees. LAG2 never hits 0, so no candidate header equals the grid's
# own header: at that single lag the map is the identity to ~1e-12 px and whole border rows / columns sit on
# map_coordinates' closed bound [0, n-1] -- membership is rounding noise of the pixel -> world -> pixel chain in wcslib
# as much as here; it is checked loosely in its own test.
LAG = np.arange(-0.2, 0.21, 0.04) * 3600.0
LAG2 = np.arange(-0.22, 0.2, 0.04) * 3600.0


def test_oracle_car_search_recovers_planted_shift(car_pair):
    from oracle.hpc import HpcSearch
    p_large, p_small, spec = car_pair
    dl, hl, ds, hs = load_pair(p_large, p_small)
    lag1 = np.array([0.08, 0.12, 0.16]) * 3600.0
    lag2 = np.array([-0.10, -0.06, -0.02]) * 3600.0
    cube = HpcSearch(dl, hl, ds, hs, lag1, lag2, None, None, None, frame="car").cube()
    i = np.unravel_index(np.nanargmax(cube), cube.shape)
    assert (i[0], i[1]) == (1, 1) and cube.max() > 0.95


# --------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from euispice_coreg_b200 import _ext
    _ext.load()
    return torch


@pytest.mark.gpu
@pytest.mark.parametrize("crval,crota,lonpole", CASES)
def test_gpu_car_pix2world_world2pix_match_oracle(torch_cuda, crval, crota, lonpole):
    from euispice_coreg_b200 import _ext
    from euispice_coreg_b200._compat.wcs import CarWcs
    from oracle.wcs_car import WcsCar
    h = _hdr(crval, crota, lonpole=lonpole)
    w = CarWcs.from_header(h)
    lng, lat = _ext.car_pix2world(w.lag_row(), 120, 80)
    lng_o, lat_o = WcsCar(h).pixel_to_world(*np.meshgrid(np.arange(120.0), np.arange(80.0)))
    assert np.max(np.abs((lng.cpu().numpy() - lng_o + 180.0) % 360.0 - 180.0)) < 1e-11
    assert np.max(np.abs(lat.cpu().numpy() - lat_o)) < 1e-11
    h2 = dict(h, CRVAL1=crval[0] - 0.07, CRVAL2=crval[1] + 0.05)
    x, y = _ext.car_world2pix(CarWcs.from_header(h2).lag_row(), torch_cuda.from_numpy(lng_o).cuda(),
                              torch_cuda.from_numpy(lat_o).cuda())
    x_o, y_o = WcsCar(h2).world_to_pixel(lng_o, lat_o)
    assert np.max(np.abs(x.cpu().numpy() - x_o)) < 1e-9 and np.max(np.abs(y.cpu().numpy() - y_o)) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("order,strict", [(2, False), (2, True), (1, False), (3, False)])
def test_gpu_initial_carrington_cube_matches_oracle(torch_cuda, car_pair, order, strict):
    from euispice_coreg_b200.hdrshift import Alignment
    from oracle.hpc import HpcSearch
    p_large, p_small, spec = car_pair
    dl, hl, ds, hs = load_pair(p_large, p_small)
    a = Alignment(p_large, p_small, LAG, LAG2, None, None, None, reprojection_order=order, strict_arithmetic=strict)
    cube = a.align_using_initial_carrington(return_type="corr")
    ref = HpcSearch(dl, hl, ds, hs, LAG, LAG2, None, None, None, order=order, frame="car").cube()
    assert cube.shape == ref.shape == (11, 11, 1, 1, 1, 1)
    assert np.array_equal(np.isnan(cube), np.isnan(ref))
    assert np.nanmax(np.abs(cube - ref)) < R_TOL
    assert np.nanargmax(cube) == np.nanargmax(ref)
    i = np.unravel_index(np.nanargmax(cube), cube.shape)
    assert abs(LAG[i[0]] / 3600.0 - spec.true_shift[0]) < 1e-9 and abs(LAG2[i[1]] / 3600.0 - spec.true_shift[1]) < 1e-9


@pytest.mark.gpu
def test_gpu_initial_carrington_identity_lag_is_the_closed_bound_knife_edge(torch_cuda, car_pair):
    """Lag (0, 0): candidate header == grid header. Only border rows / columns may differ (see LAG2 above)."""
    from euispice_coreg_b200.hdrshift import Alignment
    from oracle.hpc import HpcSearch
    p_large, p_small, spec = car_pair
    dl, hl, ds, hs = load_pair(p_large, p_small)
    z = np.array([0.0])
    a = Alignment(p_large, p_small, z, z, None, None, None)
    cube = a.align_using_initial_carrington(return_type="corr")
    ref = HpcSearch(dl, hl, ds, hs, z, z, None, None, None, frame="car").cube()
    ny, nx = ds.shape
    assert abs(cube.ravel()[0] - ref.ravel()[0]) < 5e-3
    assert nx * ny - 2 * (nx + ny) <= int(a.nvalid.ravel()[0]) <= nx * ny


@pytest.mark.gpu
def test_gpu_initial_carrington_rotation_lags_results_and_dead_lags(torch_cuda, car_pair, tmp_path):
    """CROTA lags (PCi_j rebuilt per lag), AlignmentResults with the lag arrays in header units (degrees), and a header
    with an explicit LONPOLE whose negative-CRVAL2 candidates wcslib rejects (cube entries 0.0 on both sides)."""
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.hdrshift import Alignment
    from oracle.hpc import HpcSearch
    p_large, p_small, spec = car_pair
    dl, hl, ds, hs = load_pair(p_large, p_small)
    lag1 = np.array([0.08, 0.12, 0.16]) * 3600.0
    lag2 = np.array([-0.10, -0.06, -0.02]) * 3600.0
    rot = np.array([-0.5, 0.0, 0.5])
    a = Alignment(p_large, p_small, lag1, lag2, None, None, rot)
    res = a.align_using_initial_carrington()
    ref = HpcSearch(dl, hl, ds, hs, lag1, lag2, None, None, rot, frame="car").cube()
    assert np.nanmax(np.abs(res.corr - ref)) < R_TOL
    assert tuple(res.max_index[:2]) == (1, 1) and res.max_index[4] == 1
    assert res.unit_lag == "deg"
    vals, unit = res.parameters_alignment["lag_crval1"]
    assert unit == "deg" and np.allclose(vals, lag1 / 3600.0, rtol=0, atol=1e-15)
    assert np.allclose(res.parameters_alignment_arcsec["lag_crval1"], lag1, rtol=0, atol=1e-9)
    # explicit LONPOLE = 0 on a small map whose header latitude is +0.02 deg: candidates below the equator are invalid
    hd = fits_lite.open(p_small)[0]
    h = hd.header.copy()
    h["CRVAL2"] = 0.02
    h["LONPOLE"] = 0.0
    p2 = str(tmp_path / "small_lonpole.fits")
    fits_lite.writeto(p2, [fits_lite.PrimaryHDU(hd.data, h)], overwrite=True)
    lagb = np.array([-0.04, 0.0, 0.04]) * 3600.0
    cube = Alignment(p_large, p2, lag1, lagb, None, None, None).align_using_initial_carrington(return_type="corr")
    ref = HpcSearch(dl, hl, hd.data, dict(h.items()), lag1, lagb, None, None, None, frame="car").cube()
    assert np.all(cube[:, 0] == 0.0) and np.all(ref[:, 0] == 0.0)
    assert np.nanmax(np.abs(cube - ref)) < R_TOL
