"""The "sunpy" Carrington search (`hdrshift/alignment.py:939-985`): solar-surface reprojection + bilinear lag search.

sunpy / reproject / astropy are absent from the image: `oracle/surface_reproject.py` restates their published
algorithm (PARITY UNPINNED, see its header). CPU tests check the restatement against what the algorithm must do
(identity, rotation rates, limb, edge rule); the `-m gpu` tests compare the device path with it."""
import ctypes

import numpy as np
import pytest

from conftest import load_pair

LAGS = dict(lag_crval1=np.arange(20, 29, 2.0), lag_crval2=np.arange(2, 11, 2.0), lag_cdelt1=[0], lag_cdelt2=[0],
            lag_crota=[0.0, 0.5])


def _with_observers(pair, tmp_path, small=(10.0, -3.0, None, None), large=(10.0, -3.0, None, None), tag="surf"):
    """Copies of a synthetic pair whose headers carry Stonyhurst observer keywords: (HGLN_OBS, HGLT_OBS, DSUN_OBS or None,
    DATE-AVG or None) per image."""
    from euispice_coreg_b200._compat import fits_lite
    out = []
    for path, (lon, lat, dsun, date), name in ((pair[0], large, "large"), (pair[1], small, "small")):
        hd = fits_lite.open(path)[0]
        h = hd.header.copy()
        h["HGLN_OBS"], h["HGLT_OBS"] = float(lon), float(lat)
        if dsun is not None:
            h["DSUN_OBS"] = float(dsun)
        if date is not None:
            h["DATE-AVG"] = date
        p = str(tmp_path / f"{tag}_{name}.fits")
        fits_lite.writeto(p, [fits_lite.PrimaryHDU(np.array(hd.data), h)], overwrite=True)
        out.append(p)
    return out[0], out[1]


# ------------------------------------------------------------------------------------------------ oracle, CPU
def test_change_of_observer_on_the_surface():
    from oracle import surface_reproject as sr
    au = 1.495978707e11
    f0 = dict(lon=0.0, lat=0.0, dsun=au, t=0.0, rsun=sr.R_SUN_M)
    # same observer, same time: identity (sunpy's hpc_to_hpc short-circuit)
    tx, ty = np.array([0.0, 0.1, -0.15]), np.array([0.0, -0.05, 0.12])
    a, b, vis = sr.hpc_to_hpc_on_surface(tx, ty, f0, dict(f0))
    assert np.array_equal(a, tx) and np.array_equal(b, ty) and vis.all()
    # one day later, observer at the same Stonyhurst position: the disc-centre point has moved by the SYNODIC rate
    # (Howard 2.894 urad/s = 14.326 deg/day sidereal, minus 0.9856), seen under asin-ish R sin(dlon) / (D - R cos(dlon))
    f1 = dict(f0, t=86400.0)
    a, b, vis = sr.hpc_to_hpc_on_surface(np.array([0.0]), np.array([0.0]), f0, f1)
    dlon = np.deg2rad(2.894e-6 * 86400 * 180 / np.pi - 0.9856)
    want = np.rad2deg(np.arctan2(sr.R_SUN_M * np.sin(dlon), au - sr.R_SUN_M * np.cos(dlon)))
    assert abs(a[0] - want) < 1e-12 and abs(b[0]) < 1e-15 and vis[0]
    # differential: a point at 60 deg latitude rotates slower (2.894 - 0.428 * 0.75 - 0.370 * 0.5625 urad/s)
    assert abs(sr.differential_rotation_deg(1.0, np.deg2rad(60.0))
               - ((2.894 - 0.428 * 0.75 - 0.370 * 0.5625) * 1e-6 * 86400 * 180 / np.pi - 0.9856)) < 1e-12
    # off-disc directions have no surface point
    a, b, vis = sr.hpc_to_hpc_on_surface(np.array([0.5]), np.array([0.0]), f0, f1)
    assert np.isnan(a[0]) and not vis[0]
    # a second observer 120 deg away cannot see the first one's disc centre; one 30 deg away can, displaced eastwards
    a, b, vis = sr.hpc_to_hpc_on_surface(np.array([0.0]), np.array([0.0]), f0, dict(f0, lon=np.deg2rad(120.0)))
    assert not vis[0]
    a, b, vis = sr.hpc_to_hpc_on_surface(np.array([0.0]), np.array([0.0]), f0, dict(f0, lon=np.deg2rad(30.0)))
    assert vis[0] and a[0] < 0 and abs(a[0] + np.rad2deg(np.arctan2(sr.R_SUN_M * 0.5, au - sr.R_SUN_M * np.cos(np.pi / 6)))) < 1e-12
    # there and back again
    f2 = dict(lon=np.deg2rad(12.0), lat=np.deg2rad(-4.0), dsun=0.6 * au, t=5400.0, rsun=sr.R_SUN_M)
    a, b, _ = sr.hpc_to_hpc_on_surface(tx, ty, f0, f2)
    a2, b2, _ = sr.hpc_to_hpc_on_surface(a, b, f2, f0)
    assert np.max(np.abs(a2 - tx)) < 1e-12 and np.max(np.abs(b2 - ty)) < 1e-12


def test_bilinear_with_reprojects_edge_rule():
    from oracle import surface_reproject as sr
    img = np.arange(12, dtype=np.float64).reshape(3, 4) ** 1.5
    x = np.array([0.0, 1.5, 3.0, -0.5, -0.25, 3.5, 3.50001, -0.50001, 2.25, np.nan])
    y = np.array([0.0, 0.5, 2.0, 0.0, 2.5, 2.5, 1.0, 1.0, 1.75, 1.0])
    got = sr.bilinear_edge(img, x, y)
    assert got[0] == img[0, 0] and got[2] == img[2, 3]
    assert abs(got[1] - 0.25 * (img[0, 1] + img[0, 2] + img[1, 1] + img[1, 2])) < 1e-14
    assert got[3] == img[0, 0] and abs(got[4] - img[2, 0]) < 1e-14 and abs(got[5] - img[2, 3]) < 1e-14   # edge pixels extend half a pixel
    assert np.isnan(got[6]) and np.isnan(got[7]) and np.isnan(got[9])
    fx, fy = 0.25, 0.75
    want = (img[1, 2] * (1 - fx) + img[1, 3] * fx) * (1 - fy) + (img[2, 2] * (1 - fx) + img[2, 3] * fx) * fy
    assert abs(got[8] - want) < 1e-13
    img[1, 2] = np.nan                      # a NaN tap poisons every sample that touches it, weight 0 included
    assert np.isnan(sr.bilinear_edge(img, np.array([2.0, 1.0]), np.array([1.0, 1.0]))).all()


def test_oracle_search_recovers_the_synthetic_shift(toy_pair, tmp_path):
    from oracle import surface_reproject as sr
    pl, ps = _with_observers(toy_pair, tmp_path)
    dl, hl, ds, hs = load_pair(pl, ps)
    s = sr.SurfaceSearch(dl, hl, ds, hs, **LAGS)
    # same observer, same time: the one-time reprojection is the plain TAN -> TAN bilinear cut
    from oracle import wcs_tan
    x, y = wcs_tan.extract_coordinates_pixels(hs, hl)
    assert np.allclose(s.ref, sr.bilinear_edge(dl, x, y), rtol=0, atol=1e-6, equal_nan=True)
    cube = s.cube()
    assert cube.shape == (5, 5, 1, 1, 2, 1)
    am = np.unravel_index(np.nanargmax(cube), cube.shape)
    assert (LAGS["lag_crval1"][am[0]], LAGS["lag_crval2"][am[1]], LAGS["lag_crota"][am[4]]) == (24.0, 6.0, 0.0)
    assert cube.max() > 0.99


def test_frames_are_read_from_the_headers_like_sunpy_does():
    """Host side of the sunpy path: observer = HGLN_OBS / HGLT_OBS / DSUN_OBS, time = DATE-AVG (else DATE-OBS); the product
    and the oracle read the same frame; a header without the observer keywords is refused."""
    from euispice_coreg_b200.hdrshift import Alignment
    from oracle import surface_reproject as sr
    h = {"HGLN_OBS": 10.25, "HGLT_OBS": -3.5, "DSUN_OBS": 5.7e10, "DATE-OBS": "2022-03-17T09:50:45.000",
         "DATE-AVG": "2022-03-17T09:50:50.500"}
    lon, lat, dsun, t = Alignment._surface_frame(h)
    f = sr.frame_of(h, 1.004 * sr.R_SUN_M)
    assert (lon, lat, dsun) == (f["lon"], f["lat"], f["dsun"]) == (np.radians(10.25), np.radians(-3.5), 5.7e10)
    assert t == f["t"]
    h2 = dict(h)
    del h2["DATE-AVG"]
    assert Alignment._surface_frame(h2)[3] == t - 5.5 == sr.frame_of(h2, 1.0)["t"]
    for k in ("HGLN_OBS", "HGLT_OBS", "DSUN_OBS"):
        bad = {kk: v for kk, v in h.items() if kk != k}
        with pytest.raises(ValueError, match=k):
            Alignment._surface_frame(bad)
        with pytest.raises(ValueError, match=k):
            sr.frame_of(bad, 1.0)


def test_c_abi_rejects_bad_arguments_without_a_device():
    from euispice_coreg_b200 import _ext
    lib = _ext.load()
    fake, null = ctypes.c_void_p(16), ctypes.c_void_p(0)
    good = _ext.CoregTanWcs(1, 1, 1.0, 1.0, 1, 0, 0, 1, 0, 0, 180)
    fr = _ext.CoregSurfaceFrames(0.0, 0.0, 1.5e11, 0.0, 0.0, 1.5e11, 0.0, 6.957e8)

    class E:
        EINVAL, ENOMEM = -1, -3

    def msg():
        return lib.coreg_last_error().decode()
    assert lib.coreg_pad_edge(null, _ext.F32, 4, 4, fake, null) == E.EINVAL and "null" in msg()
    assert lib.coreg_pad_edge(fake, 7, 4, 4, fake, null) == E.EINVAL and "dtype" in msg()
    assert lib.coreg_surface_cut(ctypes.byref(good), 4, 4, ctypes.byref(good), null, 4, 4, ctypes.byref(fr), fake,
                                 null) == E.EINVAL and "null" in msg()
    inside = _ext.CoregSurfaceFrames(0.0, 0.0, 1.0e8, 0.0, 0.0, 1.5e11, 0.0, 6.957e8)     # observer inside the sphere
    assert lib.coreg_surface_cut(ctypes.byref(good), 4, 4, ctypes.byref(good), fake, 4, 4, ctypes.byref(inside), fake,
                                 null) == E.EINVAL and "outside" in msg()
    assert lib.coreg_hpc_lag_corr_edge(null, fake, 4, 4, 4, 4, fake, fake, 3, fake, fake, 1 << 30, fake, null, 0,
                                       null) == E.EINVAL and "null" in msg()
    assert lib.coreg_hpc_lag_corr_edge(fake, fake, 4, 4, 4, 4, fake, fake, 3, fake, fake, 8, fake, null, 0,
                                       null) == E.ENOMEM and "workspace" in msg()
    assert ctypes.sizeof(_ext.CoregSurfaceFrames) == 64
    assert lib.coreg_spice_wave_sum(null, 1, 4, 4, 4, fake, 0, 4, fake, null) == E.EINVAL and "null" in msg()
    assert lib.coreg_spice_wave_sum(fake, 1, 0, 4, 4, fake, 0, 4, fake, null) == E.EINVAL and "empty" in msg()
    assert lib.coreg_synras_build_windows(null, _ext.F32, 1, 4, 4, ctypes.byref(good), None, None, fake, fake, 2, 2, 2,
                                          fake, null) == E.EINVAL and "null" in msg()
    assert lib.coreg_surface_search_host(null, _ext.F32, 4, 4, ctypes.byref(good), fake, _ext.F32, 4, 4,
                                         ctypes.byref(good), ctypes.byref(fr), fake, 3, 0, fake,
                                         null) == E.EINVAL and "null" in msg()
    assert lib.coreg_surface_search_host(fake, 9, 4, 4, ctypes.byref(good), fake, _ext.F32, 4, 4, ctypes.byref(good),
                                         ctypes.byref(fr), fake, 3, 0, fake, null) == E.EINVAL and "dtype" in msg()


# ------------------------------------------------------------------------------------------------ device
@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


CASES = {
    "same_frame": dict(small=(10.0, -3.0, None, None), large=(10.0, -3.0, None, None)),
    # a second spacecraft 0.3 deg away in longitude, closer to the Sun, half an hour later
    "two_observers": dict(small=(10.0, -3.0, None, None), large=(10.3, -2.9, 5.5e10, "2022-03-17T10:20:45.000")),
}


@pytest.mark.gpu
def test_pad_edge_and_surface_cut_match_the_oracle(torch_cuda, toy_pair, tmp_path):
    torch = torch_cuda
    from euispice_coreg_b200 import _ext
    from euispice_coreg_b200._compat.wcs import TanWcs
    from euispice_coreg_b200.hdrshift.alignment import Alignment
    from oracle import surface_reproject as sr
    from oracle.hpc import check_and_create_pcij
    img = np.random.default_rng(3).normal(size=(7, 9)).astype(np.float32)
    assert np.array_equal(_ext.pad_edge(torch.from_numpy(img).cuda()).cpu().numpy(), np.pad(img.astype(np.float64), 1, mode="edge"))
    for name, case in CASES.items():
        pl, ps = _with_observers(toy_pair, tmp_path, tag=name, **case)
        dl, hl, ds, hs = load_pair(pl, ps)
        check_and_create_pcij(hl)
        check_and_create_pcij(hs)
        rsun = 1.004 * sr.R_SUN_M
        want = sr.reproject_to(np.asarray(dl, dtype=np.float64), hl, hs, rsun)
        g, i = Alignment._surface_frame(hs), Alignment._surface_frame(hl)
        fr = _ext.CoregSurfaceFrames(g[0], g[1], g[2], i[0], i[1], i[2], (i[3] - g[3]) / 86400.0, rsun)
        got = _ext.surface_cut(TanWcs.from_header(hs), TanWcs.from_header(hl),
                               _ext.pad_edge(torch.from_numpy(np.ascontiguousarray(dl)).cuda()), fr).cpu().numpy()
        assert got.shape == want.shape and np.isfinite(want).mean() > 0.9
        assert (np.isfinite(got) != np.isfinite(want)).sum() <= 4            # a coordinate 1e-10 px from the edge rule
        both = np.isfinite(got) & np.isfinite(want)
        assert np.max(np.abs(got[both] - want[both])) < 1e-6 * np.abs(want[both]).max(), name   # coordinates to ~1e-10 px
    # the two frames really differ: the reprojected image moves by a good fraction of a pixel
    assert np.nanmax(np.abs(got - sr.reproject_to(np.asarray(dl, dtype=np.float64), hl, hs, rsun * 1.2))) > 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_sunpy_carrington_search_matches_the_oracle(torch_cuda, toy_pair, tmp_path, case):
    from euispice_coreg_b200.hdrshift import Alignment
    from oracle import surface_reproject as sr
    pl, ps = _with_observers(toy_pair, tmp_path, tag=case, **CASES[case])
    a = Alignment(pl, ps, parallelism=True, **LAGS)
    gpu = a.align_using_carrington(method="correlation", method_carrington_reprojection="sunpy", return_type="corr")
    dl, hl, ds, hs = load_pair(pl, ps)
    ref = sr.SurfaceSearch(dl, hl, ds, hs, **LAGS).cube()
    assert gpu.shape == ref.shape == (5, 5, 1, 1, 2, 1)
    err = np.nanmax(np.abs(gpu - ref))
    assert err <= 1e-6, err                     # the contract; observed ~1e-12
    assert np.unravel_index(np.nanargmax(gpu), gpu.shape) == np.unravel_index(np.nanargmax(ref), ref.shape)
    assert a.nvalid.max() > 0.8 * 96 * 96
    if case == "same_frame":
        am = np.unravel_index(np.nanargmax(gpu), gpu.shape)
        assert (LAGS["lag_crval1"][am[0]], LAGS["lag_crval2"][am[1]], LAGS["lag_crota"][am[4]]) == (24.0, 6.0, 0.0)
        res = Alignment(pl, ps, parallelism=True, **LAGS).align_using_carrington(method_carrington_reprojection="sunpy")
        assert res.max_index[:2] == (2, 2)


@pytest.mark.gpu
def test_surface_search_host_entry_equals_the_public_api(torch_cuda, toy_pair, tmp_path):
    """coreg_surface_search_host (host buffers in, cube out: the non-Python caller's entry at the seam `alignment.py:237`
    with `_carrington_transform_sunpy`) against `align_using_carrington(method_carrington_reprojection="sunpy")`: the same
    kernels; the per-lag constants are derived on the device instead of by numpy (last-bit differences of sin / cos)."""
    from euispice_coreg_b200 import _ext
    from euispice_coreg_b200._compat.wcs import TanWcs
    from euispice_coreg_b200.hdrshift import Alignment, engine
    from euispice_coreg_b200.hdrshift.engine import R_SUN_M
    pl, ps = _with_observers(toy_pair, tmp_path, tag="host", **CASES["two_observers"])
    a = Alignment(pl, ps, parallelism=True, **LAGS)
    gpu = a.align_using_carrington(method="correlation", method_carrington_reprojection="sunpy", return_type="corr")
    dl, hl, ds, hs = load_pair(pl, ps)
    from oracle.hpc import check_and_create_pcij
    check_and_create_pcij(hl)
    check_and_create_pcij(hs)
    d = engine.flat_lag_grid(LAGS["lag_crval1"], LAGS["lag_crval2"], [0.0], [0.0], LAGS["lag_crota"])
    table, dead = engine.tan_wcs_table(a.hdr_small, a, *d)
    assert not dead.any()
    g, i = Alignment._surface_frame(hs), Alignment._surface_frame(hl)
    fr = _ext.CoregSurfaceFrames(g[0], g[1], g[2], i[0], i[1], i[2], (i[3] - g[3]) / 86400.0, 1.004 * R_SUN_M)
    corr, nvalid = _ext.surface_search_host(dl, TanWcs.from_header(hl), ds, TanWcs.from_header(hs), fr, table)
    assert np.max(np.abs(corr.reshape(gpu.shape) - gpu)) < 1e-11
    assert np.array_equal(nvalid.reshape(a.nvalid.shape), a.nvalid)


@pytest.mark.gpu
def test_sunpy_path_needs_the_observer_keywords(torch_cuda, toy_pair):
    from euispice_coreg_b200.hdrshift import Alignment
    with pytest.raises(ValueError, match="HGLN_OBS"):
        Alignment(toy_pair[0], toy_pair[1], parallelism=True, **LAGS).align_using_carrington(
            method_carrington_reprojection="sunpy", return_type="corr")
    with pytest.raises(ValueError, match="fa"):
        Alignment(toy_pair[0], toy_pair[1], parallelism=True, **LAGS).align_using_carrington(
            method_carrington_reprojection="nearest")
