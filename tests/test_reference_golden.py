"""Golden cubes produced by the REFERENCE's own `hdrshift.Alignment` (parallel branch) on the seeded synthetic pairs:
`tests/golden/alignment_golden.npz`, generated in the build container by `tests/golden/make_alignment_golden.py`
(astropy / matplotlib replaced by import stand-ins; the WCS numbers come from the oracle's wcslib restatement, so these
goldens pin everything the reference does around the WCS calls and the whole Carrington "fa" chain, not wcslib).

CPU part: the oracle reproduces every golden cube bit for bit. GPU part (`-m gpu`): the CUDA path through the public
API reproduces them within 1e-6 (north_star), same arg-max, same NaN / never-evaluated pattern.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_alignment_golden as G  # noqa: E402

from conftest import load_pair  # noqa: E402

R_TOL = 1e-6
GOLD = np.load(os.path.join(HERE, "golden", "alignment_golden.npz"))


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("refgold"))
    f = {"toy": G.toy_files(d), "car": G.car_files(d)}
    for pair, (p_large, p_small) in f.items():
        dl, _, ds, _ = load_pair(p_large, p_small)
        if G.digest(dl, ds) != str(GOLD[f"sha_{pair}"]):
            pytest.skip(f"synthetic pair '{pair}' differs from the one the goldens were made with (library versions)")
    return f


def _oracle_cube(name, files):
    from oracle.carrington import CarringtonSearch
    from oracle.hpc import HpcSearch
    pair, entry, ctor, call = G.CASES[name]
    dl, hl, ds, hs = load_pair(*files[pair])
    kw = dict(ctor)
    lags = [kw.pop(k) for k in ("lag_crval1", "lag_crval2", "lag_cdelt1", "lag_cdelt2", "lag_crota")]
    if "reprojection_order" in kw:
        kw["order"] = kw.pop("reprojection_order")
    if entry == "align_using_carrington":
        return CarringtonSearch(dl, hl, ds, hs, *lags, call["lonlims"], call["latlims"], call["shape"], **kw).cube()
    frame = "car" if entry == "align_using_initial_carrington" else "hpc"
    return HpcSearch(dl, hl, ds, hs, *lags, frame=frame, fov_limits=call.get("fov_limits"), **kw).cube()


@pytest.mark.parametrize("name", list(G.CASES))
def test_oracle_reproduces_reference_cube_bit_exact(name, files):
    cube = _oracle_cube(name, files)
    gold = GOLD[name]
    assert cube.shape == gold.shape
    assert np.array_equal(cube, gold, equal_nan=True)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from euispice_coreg_b200 import _ext
    _ext.load()
    return torch


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_gpu_public_api_reproduces_reference_cube(torch_cuda, name, files):
    from euispice_coreg_b200.hdrshift import Alignment
    pair, entry, ctor, call = G.CASES[name]
    p_large, p_small = files[pair]
    a = Alignment(large_fov_known_pointing=p_large, small_fov_to_correct=p_small, parallelism=True, counts_cpu_max=4,
                  display_progress_bar=False, **ctor)
    cube = getattr(a, entry)(method="correlation", return_type="corr", **call)
    gold = GOLD[name]
    assert cube.shape == gold.shape and cube.dtype == np.float64
    assert np.array_equal(np.isnan(cube), np.isnan(gold)) and np.array_equal(cube == 0.0, gold == 0.0)
    assert np.nanmax(np.abs(cube - gold)) < R_TOL
    assert np.nanargmax(cube) == np.nanargmax(gold)
