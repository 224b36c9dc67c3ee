"""Pixel-shift co-alignment (`pxlshift.AlignmentPixels`).

Golden vectors: `tests/golden/pxlshift_golden.npz` is the output of the REFERENCE's own
`AlignmentPixels.find_best_parameters` (its scipy resampling, polar transform and numba Pearson with a float32
numerator) on the seeded inputs of `tests/golden/make_pxlshift_golden.py`, generated in the build container.
CPU part: the oracle restatement reproduces them bit for bit. GPU part (`-m gpu`): the CUDA path through the C ABI
reproduces the one-shot resamplings bit for bit and every r within 1e-6 (observed <= 1e-7: the float32 rounding of the
numerator can flip by one unit when the float64 sums differ in their last bits), same arg-max.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_pxlshift_golden as G  # noqa: E402

R_TOL = 1e-6
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pxlshift_golden.npz"))


@pytest.mark.parametrize("case", [0, 1, 2])
def test_oracle_reproduces_reference_golden_bit_exact(case):
    from oracle.pxlshift import PixelShiftSearch
    large, hl, small, hs = G.make_inputs(case)
    dx, dy, drot = G.LAGS[case]
    s = PixelShiftSearch(large, hl, small, hs)
    corr = s.find_best_parameters(dx, dy, drot, shift_solar_rotation_dx_large=(case == 2))
    assert np.array_equal(corr, GOLD[f"corr{case}"])
    assert np.array_equal(s.data_large, GOLD[f"large_sub{case}"], equal_nan=True)
    assert np.array_equal(s.data_small_rotated, GOLD[f"rot_last{case}"], equal_nan=True)


def test_oracle_pearson_f32_numerator():
    """pxlshift/c_correlate.py stores the lag sums in a float32 array: r differs from the float64 Pearson by up to
    one float32 rounding of the numerator."""
    from oracle.pearson import pearson
    from oracle.pxlshift import pearson_f32num
    rng = np.random.default_rng(3)
    a = rng.lognormal(5, 1, 5000)
    b = 0.7 * a + rng.normal(0, 50, 5000)
    r64, r32 = pearson(a, b), pearson_f32num(a, b)
    assert r64 != r32 and abs(r64 - r32) < 1.2e-7 * abs(r64)


def test_polar_transform_restatement_matches_oracle_and_rounds_half_to_even():
    from euispice_coreg_b200.utils.matrix_transform import MatrixTransform
    from oracle.pxlshift import polar_transform
    for shape in ((5, 7), (6, 8), (37, 51)):
        xx, yy = np.meshgrid(np.arange(shape[1]), np.arange(shape[0]))
        nx, ny = MatrixTransform.polar_transform(xx, yy, theta=1.5, units="degree")
        ox, oy = polar_transform(xx, yy, theta=1.5, units="degree")
        assert np.array_equal(nx, ox) and np.array_equal(ny, oy)
        c = (round(shape[0] / 2), round(shape[1] / 2))      # Python round: 2.5 -> 2, 3.5 -> 4
        assert nx[c] == xx[c] and ny[c] == yy[c]


def test_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from euispice_coreg_b200._ext import CoregLibraryError
    p_large, p_small = _write_case(tmp_path, 0)
    from euispice_coreg_b200.pxlshift.alignment_pixels import AlignmentPixels
    a = AlignmentPixels(p_large, 0, p_small, 0)
    with pytest.raises(CoregLibraryError):
        a.find_best_parameters([0], [0], [0.0])


def _write_case(tmp_path, case):
    from euispice_coreg_b200._compat import fits_lite
    large, hl, small, hs = G.make_inputs(case)
    p_large, p_small = str(tmp_path / f"large{case}.fits"), str(tmp_path / f"small{case}.fits")
    for p, d, h in ((p_large, large, hl), (p_small, small, hs)):
        hdr = fits_lite.Header()
        for k, v in h.items():
            hdr[k] = v
        fits_lite.writeto(p, [fits_lite.PrimaryHDU(d, hdr)], overwrite=True)
    return p_large, p_small


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
@pytest.mark.parametrize("case", [0, 1, 2])
def test_gpu_pixel_shift_matches_reference_golden(torch_cuda, tmp_path, case):
    from euispice_coreg_b200.pxlshift.alignment_pixels import AlignmentPixels
    p_large, p_small = _write_case(tmp_path, case)
    dx, dy, drot = G.LAGS[case]
    a = AlignmentPixels(p_large, 0, p_small, 0)
    corr = a.find_best_parameters(dx, dy, drot, unit_rot="degree", shift_solar_rotation_dx_large=(case == 2))
    gold = GOLD[f"corr{case}"]
    assert corr.shape == gold.shape and corr.dtype == np.float64
    # the one-shot resamplings are bit-exact against the reference's scipy calls
    assert np.array_equal(a.data_large, GOLD[f"large_sub{case}"], equal_nan=True)
    assert np.array_equal(a.data_small_rotated, GOLD[f"rot_last{case}"], equal_nan=True)
    assert np.max(np.abs(corr - gold)) < R_TOL
    assert np.argmax(corr) == np.argmax(gold)


@pytest.mark.gpu
def test_gpu_pixel_shift_larger_pair_sparse_lags_and_boundaries(torch_cuda, tmp_path):
    """Several tiles in both directions (ragged), NaN holes, a sparse lag array (window chunks fall back to one lag
    each when the staged window would not fit), nvalid, and the reference's boundary error."""
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.pxlshift.alignment_pixels import AlignmentPixels
    from oracle.pxlshift import PixelShiftSearch
    rng = np.random.default_rng(77)
    from scipy.ndimage import gaussian_filter
    large = np.exp(gaussian_filter(rng.standard_normal((400, 520)), 4.0) * 8.0) * 100.0
    small = large[120:120 + 150, 170:170 + 201] * 0.9 + rng.normal(0, 1.0, (150, 201))
    small[40:44, 100:108] = np.nan
    large[200, :] = np.nan
    hl = {"CDELT1": 4.0, "CDELT2": 4.0, "CUNIT1": "arcsec", "CUNIT2": "arcsec"}
    hs = {"CDELT1": 4.0, "CDELT2": 4.0, "CUNIT1": "arcsec", "CUNIT2": "arcsec"}
    p_large, p_small = str(tmp_path / "L.fits"), str(tmp_path / "S.fits")
    for p, d, h in ((p_large, large, hl), (p_small, small, hs)):
        hdr = fits_lite.Header()
        for k, v in h.items():
            hdr[k] = v
        fits_lite.writeto(p, [fits_lite.PrimaryHDU(d, hdr)], overwrite=True)
    # centred slice starts at ((400-150-1)//2, (520-201-1)//2) = (124, 159): the planted offset is dx=+11, dy=-4
    dx = np.arange(-2, 21)
    dy = np.array([-120, -40, -8, -6, -4, -2, 0, 5, 60, 120])
    drot = np.array([0.0, 0.4])
    a = AlignmentPixels(p_large, 0, p_small, 0)
    corr = a.find_best_parameters(dx, dy, drot)
    ref = PixelShiftSearch(large, hl, small, hs).find_best_parameters(dx, dy, drot)
    assert np.max(np.abs(corr - ref)) < R_TOL
    i = np.unravel_index(np.argmax(corr), corr.shape)
    assert (dx[i[0]], dy[i[1]], drot[i[2]]) == (11, -4, 0.0) and corr.max() > 0.99
    assert a.nvalid.shape == corr.shape and a.nvalid.max() <= 150 * 201 - 32
    with pytest.raises(ValueError, match="too large shift"):
        AlignmentPixels(p_large, 0, p_small, 0).find_best_parameters(np.array([0, 200]), np.array([0]), np.array([0.0]))
    with pytest.raises(TypeError):
        AlignmentPixels(p_large, 0, p_small, 0).find_best_parameters(np.array([0.5]), np.array([0]), np.array([0.0]))
