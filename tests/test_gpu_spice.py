"""GPU parity of the synthetic-raster builder (K6) and the SPICE helioprojective search (config 3 path)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
R_TOL = 1e-6


@pytest.fixture(scope="module")
def spice_case(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from euispice_coreg_b200._synth.spice import make_spice_case, small_spice_spec
    d = tmp_path_factory.mktemp("spice")
    return make_spice_case(str(d), small_spice_spec(nbin2=8, pxbeg2=192), tag="toy") + (str(d),)


def _load(path):
    from euispice_coreg_b200._compat import fits_lite
    h = fits_lite.open(path)[0]
    return h.data, dict(h.header.items())


def test_synras_matches_oracle(spice_case):
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.synras import SPICEComposedMapBuilder
    from oracle.synras import build_synras
    p_spice, imagers, spec, d = spice_case
    b = SPICEComposedMapBuilder(p_spice, imagers, threshold_time=100.0)
    name = b.process(folder_path_output=d, basename_output="synras.fits", print_filename=False,
                     return_synras_name=True)
    assert name == os.path.join(d, "synras.fits") == b.get_path_to_composed_map()
    out = fits_lite.open(name)[0]
    _, h4 = _load(p_spice)
    frames, hdrs = zip(*[_load(p) for p in imagers])
    ref, chosen = build_synras(h4, frames, hdrs, 100.0)
    assert len(set(chosen.tolist())) >= 3            # the raster really spans several imager frames
    assert out.data.shape == ref.shape == (spec.n_y, spec.n_x) and out.data.dtype == np.float64
    assert np.array_equal(np.isnan(out.data), np.isnan(ref))
    # float32 frames -> float32-rounded samples on both sides; device trig differs by <= 1 ulp in the coordinates
    assert np.nanmax(np.abs(out.data - ref) / np.abs(ref)) < 2e-7
    assert np.mean(out.data == ref) > 0.99
    # header: SPICE WCS keys in degrees on top of the middle imager header
    assert out.header["CUNIT1"] == "deg" and out.header["CRVAL1"] == h4["CRVAL1"] * (1.0 / 3600.0)
    assert out.header["CDELT2"] == h4["CDELT2"] * (1.0 / 3600.0) and out.header["PC1_2"] == h4["PC1_2"]
    assert out.header["DATE-AVG"] == h4["DATE-AVG"] and out.header["SPECPATH"] == os.path.basename(p_spice)
    with pytest.raises(ValueError):
        SPICEComposedMapBuilder(p_spice, imagers, threshold_time=1.0).process(print_filename=False)


def test_synras_from_frame_windows_equals_whole_frames(spice_case, tmp_path):
    """`SPICEComposedMapBuilder` reads and uploads only the window of each imager frame the raster can reach
    (`coreg_synras_build_windows`: coordinates in the full image, integer origin subtracted exactly): the raster has the
    bits of the whole-frame build, and the windows really are small."""
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.hdrshift.engine import LagSearchEngine
    from euispice_coreg_b200.synras import SPICEComposedMapBuilder
    p_spice, imagers, spec, d = spice_case
    outs = {}
    for windows in (True, False):
        b = SPICEComposedMapBuilder(p_spice, imagers, threshold_time=100.0)
        b.use_windows = windows
        name = b.process(folder_path_output=str(tmp_path), basename_output=f"synras_w{int(windows)}.fits",
                         print_filename=False, return_synras_name=True)
        outs[windows] = fits_lite.open(name)[0].data
        grid = b._w_grid
    assert outs[True].dtype == np.float64 and np.isfinite(outs[True]).mean() > 0.9
    assert np.array_equal(outs[True].view(np.uint64), outs[False].view(np.uint64))
    from euispice_coreg_b200._compat.wcs import TanWcs
    hd = fits_lite.open(imagers[0])[0]
    win = LagSearchEngine.large_window(TanWcs.from_header(hd.header), grid, hd.data.shape)
    assert win is not None and (win[1] - win[0]) * (win[3] - win[2]) < 0.5 * hd.data.size


def test_synras_keep_original_imager_pixel_size(spice_case):
    """`keep_original_imager_pixel_size=True` (`map_builder.py:259-275, 165-192`): the raster is stepped in units of
    the imager's pixel along both axes (fractional raster pixels), and the composed header is rebuilt around the centre
    of the new grid with the imager's CDELT."""
    from euispice_coreg_b200._compat.wcs import TanWcs
    from euispice_coreg_b200.synras import SPICEComposedMapBuilder
    from oracle.synras import build_synras, xy_header
    p_spice, imagers, spec, d = spice_case
    _, h4 = _load(p_spice)
    frames, hdrs = zip(*[_load(p) for p in imagers])
    b = SPICEComposedMapBuilder(p_spice, imagers, threshold_time=100.0)
    b.process(print_filename=False, keep_original_imager_pixel_size=True)
    ref, chosen = build_synras(h4, frames, hdrs, 100.0, keep_original_imager_pixel_size=True)
    r1, r2 = hdrs[0]["CDELT1"] / h4["CDELT1"], hdrs[0]["CDELT2"] / h4["CDELT2"]
    assert b.data_composed.shape == ref.shape == (len(np.arange(0, spec.n_y, r2)), len(np.arange(0, spec.n_x, r1)))
    assert ref.shape != (spec.n_y, spec.n_x)
    assert np.array_equal(np.isnan(b.data_composed), np.isnan(ref))
    assert np.nanmax(np.abs(b.data_composed - ref) / np.abs(ref)) < 2e-7
    h = b.hdr_composed
    s_im = 1.0 / 3600.0 if hdrs[0]["CUNIT1"].strip() == "arcsec" else 1.0
    assert h["CDELT1"] == hdrs[0]["CDELT1"] * s_im and h["CDELT2"] == hdrs[0]["CDELT2"] * s_im
    assert h["CRPIX1"] == (ref.shape[1] + 1) / 2 and h["CRPIX2"] == (ref.shape[0] + 1) / 2
    assert h["NAXIS1"] == ref.shape[1] and h["NAXIS2"] == ref.shape[0]
    from oracle import wcs_tan
    lon_mid, lat_mid = wcs_tan.WcsTan(dict(xy_header(h4), NAXIS1=spec.n_x, NAXIS2=spec.n_y)).pixel_to_world(
        np.array([(spec.n_x - 1) / 2]), np.array([(spec.n_y - 1) / 2]))
    assert abs((h["CRVAL1"] - lon_mid[0] + 180.0) % 360.0 - 180.0) < 1e-11 and abs(h["CRVAL2"] - lat_mid[0]) < 1e-11
    lam = h["CDELT2"] / h["CDELT1"]
    rho = np.arccos(h["PC1_1"]) * (-np.sign(h4["PC1_2"]))
    assert abs(h["PC1_2"] + lam * np.sin(rho)) < 1e-15 and abs(h["PC2_1"] - np.sin(rho) / lam) < 1e-15
    assert TanWcs.from_header(h).cdelt1 == h["CDELT1"]


def test_synras_level3_header_gives_the_level2_raster(spice_case, tmp_path):
    """Level-3 (fitted) SPICE files carry the axes (parameter, x, y, time): `map_builder.py:297-344` builds the same
    raster from axes (2, 3) and the time dependence PC4_2. The same geometry written both ways must give the same
    synthetic raster bit for bit; without an output folder the reference raises NotImplementedError for level 3."""
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.synras import SPICEComposedMapBuilder
    p_spice, imagers, spec, d = spice_case
    hdu = fits_lite.open(p_spice)[0]
    h2 = hdu.header
    h3 = fits_lite.Header()
    for k, v in h2.items():
        if k in ("SIMPLE", "BITPIX", "EXTEND") or k.startswith("NAXIS"):
            continue
        if k[:5] in ("CTYPE", "CUNIT", "CRPIX", "CRVAL", "CDELT") or k[:2] == "PC" and "_" in k:
            continue
        h3[k] = v
    old_to_new = {1: 2, 2: 3, 4: 4}          # x, y, time; the dispersion axis is gone, axis 1 = fit parameter
    h3["WCSAXES"] = 4
    h3["CTYPE1"], h3["CUNIT1"], h3["CRPIX1"], h3["CRVAL1"], h3["CDELT1"] = "PARAMETER", "", 1.0, 1.0, 1.0
    for o, n in old_to_new.items():
        for key in ("CTYPE", "CUNIT", "CRPIX", "CRVAL", "CDELT"):
            h3[f"{key}{n}"] = h2[f"{key}{o}"]
        for o2, n2 in old_to_new.items():
            if f"PC{o}_{o2}" in h2:
                h3[f"PC{n}_{n2}"] = h2[f"PC{o}_{o2}"]
    data3 = np.zeros((spec.n_y, spec.n_x, 3), dtype=np.float32)
    p3 = str(tmp_path / "solo_L3_spice_toy.fits")
    fits_lite.writeto(p3, [fits_lite.PrimaryHDU(data3, h3)], overwrite=True)
    b2 = SPICEComposedMapBuilder(p_spice, imagers, threshold_time=100.0)
    b2.process(print_filename=False)
    b3 = SPICEComposedMapBuilder(p3, imagers, threshold_time=100.0)
    b3.process(folder_path_output=str(tmp_path), basename_output="synras_l3.fits", print_filename=False, level=3)
    assert np.array_equal(b2.data_composed, b3.data_composed, equal_nan=True)
    for k in ("CRVAL1", "CRVAL2", "CDELT1", "CDELT2", "PC1_2", "PC2_1", "CRPIX1", "CRPIX2"):
        assert b2.hdr_composed[k] == b3.hdr_composed[k], k
    with pytest.raises(NotImplementedError):
        SPICEComposedMapBuilder(p3, imagers, threshold_time=100.0).process(print_filename=False, level=3)


def test_alignment_spice_parity_and_recovers_shift(spice_case, tmp_path):
    """SPICE-style inputs: 4-axis L2 cube, degrees after the 2-D header extraction, anisotropic pixels, NaN slit
    edges. A synthetic raster has the SPICE header by construction, so the one-time cut of the large image maps
    pixel (i, j) onto (i, j) +- 1e-11 and whole border rows/columns sit exactly on map_coordinates' closed
    [0, n-1] bound: whether they are "inside" is decided by the last bit of the WCS round trip (in the reference
    too). The strict parity check therefore uses a synras whose CRPIX is moved by a fraction of a pixel (no
    coordinate within 1e-9 of the bound); the unmodified case is checked for arg-max and a loose bound."""
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.hdrshift import AlignmentSpice
    from euispice_coreg_b200.synras import SPICEComposedMapBuilder
    from oracle.hpc import HpcSearch
    from oracle.synras import spice_l2_image
    p_spice, imagers, spec, d = spice_case
    synras = SPICEComposedMapBuilder(p_spice, imagers, threshold_time=100.0).process(
        folder_path_output=d, basename_output="synras2.fits", print_filename=False, return_synras_name=True)
    hd = fits_lite.open(synras)[0]
    h_off = hd.header.copy()
    h_off["CRPIX1"] = h_off["CRPIX1"] + 0.37
    h_off["CRPIX2"] = h_off["CRPIX2"] - 0.41
    synras_off = str(tmp_path / "synras_off.fits")
    fits_lite.writeto(synras_off, [fits_lite.PrimaryHDU(hd.data, h_off)], overwrite=True)
    lags = dict(lag_crval1=np.arange(-14, -1, 2.0), lag_crval2=np.arange(6, 19, 2.0), lag_cdelt1=[0], lag_cdelt2=[0],
                lag_crota=[0])
    d4, h4 = _load(p_spice)
    img, hdr = spice_l2_image(d4, h4)
    for large, tol in ((synras_off, R_TOL), (synras, 2e-2)):
        a = AlignmentSpice(large, p_spice, parallelism=True, small_fov_window=0, large_fov_window=-1, **lags)
        gpu = a.align_using_helioprojective(return_type="corr")
        assert np.array_equal(np.isnan(img), np.isnan(a.data_small)) and np.nanmax(np.abs(img - a.data_small)) == 0.0
        dl, hl = _load(large)
        # the reference's AlignmentSpice skips the PCi_j check; both headers carry PCi_j so the oracle's check is a no-op
        ref = HpcSearch(dl, hl, img, hdr, **lags).cube()
        assert np.array_equal(np.isnan(gpu), np.isnan(ref))
        assert np.nanmax(np.abs(gpu - ref)) <= tol
        am = np.unravel_index(np.nanargmax(gpu), gpu.shape)
        assert am == np.unravel_index(np.nanargmax(ref), ref.shape)
    assert (lags["lag_crval1"][am[0]], lags["lag_crval2"][am[1]]) == spec.true_shift
    # wavelength interval + results object
    a2 = AlignmentSpice(synras, p_spice, small_fov_window=0, wavelength_interval_to_sum=[97.68, 97.72], **lags)
    res = a2.align_using_helioprojective()
    img2, _ = spice_l2_image(d4, h4, (97.68, 97.72))
    assert np.nanmax(np.abs(img2 - a2.data_small)) == 0.0 and not np.array_equal(img2, img, equal_nan=True)
    assert abs(res.shift_arcsec[0] - spec.true_shift[0]) < 2.0 and abs(res.shift_arcsec[1] - spec.true_shift[1]) < 2.0


def test_synras_host_buffer_entry_point_matches_device_entry():
    """coreg_synras_build_host (frame stack, slit coordinates and raster in host memory: the body of
    `SPICEComposedMapBuilder.process`, synras/map_builder.py:57-79, 95-131) == coreg_synras_build on device tensors."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from euispice_coreg_b200 import _ext
    from euispice_coreg_b200._compat.wcs import TanWcs
    rng = np.random.default_rng(11)
    frames = rng.lognormal(5, 1, (3, 64, 80)).astype(np.float32)
    frames[1, 10, 12] = np.nan

    def hdr(crval, crota):
        rho = np.deg2rad(crota)
        return {"NAXIS1": 80, "NAXIS2": 64, "CTYPE1": "HPLN-TAN", "CTYPE2": "HPLT-TAN", "CUNIT1": "arcsec",
                "CUNIT2": "arcsec", "CRPIX1": 40.5, "CRPIX2": 32.5, "CDELT1": 4.4, "CDELT2": 4.4, "CRVAL1": crval[0],
                "CRVAL2": crval[1], "PC1_1": np.cos(rho), "PC1_2": -np.sin(rho), "PC2_1": np.sin(rho),
                "PC2_2": np.cos(rho), "LONPOLE": 180.0}

    wcs = [TanWcs.from_header(hdr((10.0 * k, -5.0 * k), 2.0 + k)) for k in range(3)]
    yy, xx = np.meshgrid(np.linspace(-3, 66, 45), np.linspace(-4, 83, 12), indexing="ij")
    lng, lat = wcs[0].pixel_to_world(xx, yy)
    cols = [0, 1, 2, -1, 2, 1, 0, 0, 1, 2, 2, 1]
    for order in (1, 2):
        host = _ext.synras_build_host(frames, wcs, cols, lng, lat, order)
        dev = _ext.synras_build(torch.from_numpy(frames).cuda(), wcs, cols, torch.from_numpy(np.ascontiguousarray(lng)).cuda(),
                                torch.from_numpy(np.ascontiguousarray(lat)).cuda(), order).cpu().numpy()
        assert host.shape == (45, 12) and np.array_equal(host, dev, equal_nan=True)
        assert np.isnan(host[:, 3]).all() and np.isfinite(host).any()
