"""CPU tests of the oracle itself: it must agree with the reference's golden material before it is trusted."""
import os
import sys

import numpy as np
import pytest
from scipy.ndimage import map_coordinates

from oracle import wcs_tan
from oracle.pearson import masked_pearson, pearson
from oracle.resample import interpol2d, map_coordinates_restated

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_pearson_matches_reference_numba_golden():
    """r from the reference's own `c_correlate` (hdrshift/c_correlate.py:39-72), bit for bit."""
    sys.path.insert(0, GOLD)
    import make_pearson_golden as g
    z = np.load(os.path.join(GOLD, "pearson_golden.npz"))
    for k, (a, b) in enumerate(g.regenerate_inputs()):
        if f"a{k}" in z.files:
            assert np.array_equal(z[f"a{k}"], a) and np.array_equal(z[f"b{k}"], b)
        assert pearson(a, b) == z[f"r{k}"][0]


def test_pearson_edge_cases():
    assert np.isnan(pearson(np.array([]), np.array([])))
    assert np.isnan(pearson(np.array([1.0]), np.array([2.0])))
    a = np.array([1.0, np.nan, 3.0, 4.0, np.inf])
    b = np.array([2.0, 5.0, np.nan, 9.0, 1.0], dtype=np.float32)
    assert masked_pearson(a, b) == pearson(np.array([1.0, 4.0]), np.array([2.0, 9.0]))


@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_restated_map_coordinates_is_bit_exact_vs_scipy(order, dtype):
    rng = np.random.default_rng(order * 7 + 1)
    img = rng.lognormal(5, 1, (37, 53)).astype(dtype)
    img[5, 7] = np.nan
    n = 20000
    y = rng.uniform(-3, 40, n)
    x = rng.uniform(-3, 56, n)
    # exact integers, half-integers, the closed borders and NaN coordinates
    y[:8] = [0.0, 36.0, 36.0000001, -1e-12, 0.5, 35.5, np.nan, 18.0]
    x[:8] = [0.0, 52.0, 3.0, 4.0, 51.5, 0.5, 2.0, np.nan]
    for out_dtype in (np.float32, np.float64):
        ref = np.empty(n, dtype=out_dtype)
        map_coordinates(img, np.stack((y, x)), order=order, mode="constant", cval=-7.5, output=ref, prefilter=False)
        got = map_coordinates_restated(img, y, x, order, -7.5, out_dtype)
        assert np.array_equal(ref, got, equal_nan=True)


def test_interpol2d_writes_dst_like_reference():
    rng = np.random.default_rng(3)
    img = rng.normal(size=(20, 30))
    x = rng.uniform(-2, 32, (6, 9))
    y = rng.uniform(-2, 22, (6, 9))
    dst = np.zeros_like(x, dtype="float32")
    assert interpol2d(img, x=x, y=y, fill=np.nan, order=2, dst=dst) is None
    ref = map_coordinates(img, np.stack((y.ravel(), x.ravel())), order=2, mode="constant", cval=np.nan,
                          prefilter=False).astype(np.float32).reshape(x.shape)
    assert np.array_equal(dst, ref, equal_nan=True)
    out = interpol2d(img, x=x, y=y, fill=0.0, order=1)
    assert out.dtype == img.dtype and out.shape == x.shape


def _hdr(crval=(-100.0, 50.0), crota=3.0, unit="arcsec", n=(64, 48), cdelt=0.492):
    rho = np.deg2rad(crota)
    return {"NAXIS1": n[0], "NAXIS2": n[1], "CTYPE1": "HPLN-TAN", "CTYPE2": "HPLT-TAN", "CUNIT1": unit,
            "CUNIT2": unit, "CRPIX1": (n[0] + 1) / 2, "CRPIX2": (n[1] + 1) / 2, "CDELT1": cdelt, "CDELT2": cdelt,
            "CRVAL1": crval[0], "CRVAL2": crval[1], "PC1_1": np.cos(rho), "PC1_2": -np.sin(rho),
            "PC2_1": np.sin(rho), "PC2_2": np.cos(rho), "LONPOLE": 180.0, "CROTA": crota}


def test_wcs_tan_round_trip_and_reference_pixel():
    w = wcs_tan.WcsTan(_hdr())
    x, y = np.meshgrid(np.arange(64.0), np.arange(48.0))
    lng, lat = w.pixel_to_world(x, y)
    x2, y2 = w.world_to_pixel(lng, lat)
    assert np.max(np.abs(x2 - x)) < 1e-9 and np.max(np.abs(y2 - y)) < 1e-9
    # CRPIX maps onto CRVAL (wcslib normalises lng to (-360, 0] when CRVAL1 < 0)
    l0, b0 = w.pixel_to_world(np.array([31.5]), np.array([23.5]))
    assert abs(wcs_tan.ang2pipi_deg(l0)[0] - (-100.0 / 3600)) < 1e-13 and abs(b0[0] - 50.0 / 3600) < 1e-13
    assert np.all(lng <= 0.0)


def test_wcs_tan_agrees_with_independent_closed_form():
    """Second derivation (closed-form gnomonic in the product's host layer) within 1e-9 px / 1e-12 deg."""
    from euispice_coreg_b200._compat.wcs import TanWcs
    for unit, crval, cdelt in (("arcsec", (-100.0, 50.0), 0.492), ("deg", (0.3, -0.2), 4.44 / 3600),
                               ("arcsec", (2000.0, -1500.0), 4.44)):
        h = _hdr(crval=crval, unit=unit, cdelt=cdelt, crota=-7.0)
        w = wcs_tan.WcsTan(h)
        t = TanWcs.from_header(h)
        x, y = np.meshgrid(np.arange(0, 64.0, 3), np.arange(0, 48.0, 5))
        lng, lat = w.pixel_to_world(x, y)
        lng2, lat2 = t.pixel_to_world(x, y)
        assert np.max(np.abs(wcs_tan.ang2pipi_deg(lng) - wcs_tan.ang2pipi_deg(lng2))) < 1e-12
        assert np.max(np.abs(lat - lat2)) < 1e-12
        h2 = dict(h, CRVAL1=crval[0] + 30 * (1 if unit == "arcsec" else 1 / 3600.0))
        xa, ya = wcs_tan.WcsTan(h2).world_to_pixel(lng, lat)
        xb, yb = TanWcs.from_header(h2).world_to_pixel(lng, lat)
        assert np.max(np.abs(xa - xb)) < 1e-9 and np.max(np.abs(ya - yb)) < 1e-9


def test_ang2pipi():
    a = np.array([-190.0, -180.0, 180.0, 181.0, 359.0, -359.99, 0.0])
    out = wcs_tan.ang2pipi_deg(a)
    assert np.allclose(out, [170.0, 180.0, 180.0, -179.0, -1.0, 0.01, 0.0])


def test_hpc_oracle_recovers_injected_pointing_error(toy_pair):
    from conftest import load_pair
    from oracle.hpc import HpcSearch
    dl, hl, ds, hs = load_pair(*toy_pair[:2])
    s = HpcSearch(dl, hl, ds, hs, lag_crval1=np.arange(20, 29, 2.0), lag_crval2=np.arange(2, 11, 2.0),
                  lag_cdelt1=[0], lag_cdelt2=[0], lag_crota=[0])
    c = s.cube()
    assert c.shape == (5, 5, 1, 1, 1, 1)
    i, j = np.unravel_index(np.nanargmax(c), c.shape)[:2]
    assert (20 + 2 * i, 2 + 2 * j) == (24, 6)
    assert c.max() > 0.95
    # the reference recomputes the world grid per lag; identical result
    assert s.step(24.0, 6.0, 0.0, 0.0, 0.0, reuse_world=False) == c[2, 2, 0, 0, 0, 0]


def test_hpc_oracle_reference_cdelt_quirks(toy_pair):
    """App. B1: a CDELT2 lag kills the reference's worker (entry stays 0.0); a CDELT1 lag is ignored."""
    from conftest import load_pair
    from oracle.hpc import HpcSearch
    dl, hl, ds, hs = load_pair(*toy_pair[:2])
    s = HpcSearch(dl, hl, ds, hs, lag_crval1=[24.0], lag_crval2=[6.0], lag_cdelt1=[0.0, 0.01], lag_cdelt2=[0.0, 0.01],
                  lag_crota=[0])
    c = s.cube()[0, 0, :, :, 0, 0]
    assert c[0, 1] == 0.0 and c[1, 1] == 0.0
    assert abs(c[1, 0] - c[0, 0]) < 1e-9 and c[0, 0] > 0.9
    s2 = HpcSearch(dl, hl, ds, hs, lag_crval1=[24.0], lag_crval2=[6.0], lag_cdelt1=[0.0, 0.01], lag_cdelt2=[0.0, 0.01],
                   lag_crota=[0], cdelt_mode="intended")
    c2 = s2.cube()[0, 0, :, :, 0, 0]
    assert np.all(c2 > 0.5) and c2[0, 0] == c[0, 0] and c2[1, 1] != c2[0, 0]


# ----------------------------------------------------------------------------------------------- RICE tiles
def test_rice_dither_sequence_check_value():
    """cfitsio documents that the 10 000th seed of `fits_init_randoms` must be 1043618065: the one known-answer
    vector for the tiled-image codec that is available offline."""
    from oracle import rice
    from euispice_coreg_b200 import _ext
    vals, seed = rice.fits_rand_values()
    assert seed == 1043618065
    assert vals.dtype == np.float32 and 0.0 < vals.min() and vals.max() < 1.0
    assert np.array_equal(_ext.fits_rand_values(), vals)


@pytest.mark.parametrize("bytepix", [1, 2, 4])
def test_rice_codec_round_trip_edge_cases(bytepix):
    from oracle import rice
    bits = 8 * bytepix
    lo, hi = -(1 << (bits - 1)), (1 << (bits - 1)) - 1
    rng = np.random.default_rng(bytepix)
    cases = [np.array([5]), np.full(64, 7), np.arange(33), rng.integers(lo, hi, 100),        # single, constant, ramp, noise
             np.clip(np.cumsum(rng.integers(-3, 4, 1000)), lo, hi),                            # low entropy
             np.array([lo, hi, lo, hi, 0, -1, 1] * 9),                                         # wrap-around differences
             np.concatenate([np.zeros(32), rng.integers(lo, hi, 32), np.zeros(31)])]           # zero / verbatim / short block
    for a in cases:
        a = a.astype(np.int64)
        buf = rice.rice_encode(a, 32, bytepix)
        assert np.array_equal(rice.rice_decode(buf, a.size, 32, bytepix), a)
    # the constant tile costs the first pixel + one FS code (3 / 4 / 5 bits) per block
    fsbits = {1: 3, 2: 4, 4: 5}[bytepix]
    assert len(rice.rice_encode(np.full(64, 7), 32, bytepix)) == bytepix + (2 * fsbits + 7) // 8


def test_rice_quantisation_round_trip_and_compressed_header(tmp_path):
    from euispice_coreg_b200._compat import fits_lite
    from oracle import rice
    rng = np.random.default_rng(3)
    tile = 500 + 80 * rng.standard_normal(300)
    for method in (1, 2):
        tile[7] = 0.0
        q = rice.quantize_tile(tile, 0.5, tile.min(), row=17, zdither0=123, method=method)
        back = rice.unquantize_tile(q, 0.5, tile.min(), row=17, zdither0=123, method=method, out_dtype=np.float64)
        assert np.max(np.abs(back - tile)) <= 0.25 + 1e-9
        assert (back[7] == 0.0) == (method == 2)
    img = (500 + 100 * rng.standard_normal((12, 40))).astype(np.float32)
    p = str(tmp_path / "c.fits")
    rice.write_compressed_image(p, img, extra_cards=[("CRVAL1", 12.5), ("CTYPE1", "HPLN-TAN"), ("EXTNAME", "IMG")],
                                quantize_scale=0.25, zdither0=7)
    hdus = fits_lite.open(p)
    assert len(hdus) == 2 and type(hdus[-1]).__name__ == "CompImageHDU" and hdus["IMG"] is hdus[1]
    h = hdus[1].header
    assert (h["BITPIX"], h["NAXIS"], h["NAXIS1"], h["NAXIS2"], h["CRVAL1"], h["CTYPE1"]) == (-32, 2, 40, 12, 12.5, "HPLN-TAN")
    assert not any(k.startswith("Z") or k.startswith("TFORM") for k in h.keys())
