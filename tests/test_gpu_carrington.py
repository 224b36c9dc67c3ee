"""GPU parity of the Carrington ("fa") search (K4/K5) against oracle/carrington.py."""
import numpy as np
import pytest

from conftest import load_pair

pytestmark = pytest.mark.gpu
R_TOL = 1e-6

GRID = dict(lonlims=(248.0, 252.0), latlims=(-4.0, 0.0), shape=(120, 100))
LAGS = dict(lag_crval1=np.arange(20, 29, 2.0), lag_crval2=np.arange(2, 11, 2.0), lag_cdelt1=[0], lag_cdelt2=[0],
            lag_crota=[0])


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _both(pair, lags, grid, **kw):
    from euispice_coreg_b200.hdrshift import Alignment
    from oracle.carrington import CarringtonSearch
    a = Alignment(pair[0], pair[1], parallelism=True, **lags, **kw)
    gpu = a.align_using_carrington(method="correlation", return_type="corr", **grid)
    dl, hl, ds, hs = load_pair(*pair[:2])
    okw = {("order" if k == "reprojection_order" else k): v for k, v in kw.items()}
    s = CarringtonSearch(dl, hl, ds, hs, **lags, **grid, **okw)
    return gpu, s.cube(), a, s


def _assert_parity(gpu, ref):
    assert gpu.shape == ref.shape
    assert np.array_equal(np.isnan(gpu), np.isnan(ref))
    err = np.nanmax(np.abs(gpu - ref))
    assert err <= R_TOL, err
    assert np.unravel_index(np.nanargmax(gpu), gpu.shape) == np.unravel_index(np.nanargmax(ref), ref.shape)
    return err


def test_carrington_large_image_projection_matches_oracle(torch_cuda, toy_pair):
    gpu, ref, a, s = _both(toy_pair, dict(LAGS, lag_crval1=[24.0], lag_crval2=[6.0]), GRID)
    got = a.engine.ref.cpu().numpy()
    assert got.shape == s.data_large.shape == (100, 120) and got.dtype == np.float64
    assert np.array_equal(np.isnan(got), np.isnan(s.data_large))
    # device atan vs libm atan differ by <= 1 ulp in the coordinates -> ~1e-12 relative in the samples
    assert np.nanmax(np.abs(got - s.data_large) / np.abs(s.data_large)) < 1e-10


def test_carrington_cube_parity(torch_cuda, toy_pair):
    gpu, ref, a, _ = _both(toy_pair, LAGS, GRID)
    err = _assert_parity(gpu, ref)
    assert err < 1e-9
    i, j = np.unravel_index(np.nanargmax(gpu), gpu.shape)[:2]
    assert (LAGS["lag_crval1"][i], LAGS["lag_crval2"][j]) == (24.0, 6.0)


def test_carrington_cube_parity_rotation_lags_thresholds_grid_beyond_fov(torch_cuda, toy_pair):
    """CROTA lags (one plane pair per value), masked small pixels, and a grid much larger than the small FOV
    (most Carrington pixels fall outside the small image)."""
    lags = dict(LAGS, lag_crota=[-0.4, 0.0, 0.6])
    grid = dict(lonlims=(244.0, 256.0), latlims=(-8.0, 4.0), shape=(150, 130))
    gpu, ref, _, _ = _both(toy_pair, lags, grid, small_fov_value_min=80.0, small_fov_value_max=1500.0)
    _assert_parity(gpu, ref)


@pytest.mark.parametrize("order", [1, 3])
def test_carrington_orders(torch_cuda, toy_pair, order):
    gpu, ref, _, _ = _both(toy_pair, dict(LAGS, lag_crval1=[22.0, 24.0], lag_crval2=[6.0, 8.0]), GRID,
                           reprojection_order=order)
    _assert_parity(gpu, ref)


def test_carrington_size_deg_grid_and_results_object(torch_cuda, toy_pair):
    from euispice_coreg_b200.hdrshift import Alignment
    a = Alignment(toy_pair[0], toy_pair[1], parallelism=True, **LAGS)
    res = a.align_using_carrington(size_deg_carrington=(4.0, 4.0))
    assert a.shape == [96, 96] and res.corr.shape == (5, 5, 1, 1, 1, 1)
    assert abs(res.shift_arcsec[0] - 24.0) < 1.0 and abs(res.shift_arcsec[1] - 6.0) < 1.0
    with pytest.raises(ValueError):
        Alignment(toy_pair[0], toy_pair[1], **LAGS).align_using_carrington(lonlims=(1, 2))
