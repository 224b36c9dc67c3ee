"""Golden outputs of the REFERENCE's own `synras.SPICEComposedMapBuilder` and `hdrshift.AlignmentSpice` on the seeded
synthetic SPICE case (BASELINE configs[2] path): `tests/golden/spice_golden.npz`, generated in the build container by
`tests/golden/make_spice_golden.py` (import stand-ins; the 4-axis WCS is the stand-in's restatement, so the goldens pin
the reference's logic around it: frame selection per raster column, composed header, slit-edge rows, spectral sum,
wavelength / sub-FOV selections, the search).

CPU part: the oracle reproduces them bit for bit. GPU part (`-m gpu`): the public API reproduces the synthetic raster to
float32 rounding and every cube within 1e-6 with the same arg-max.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_spice_golden as G  # noqa: E402

R_TOL = 1e-6
GOLD = np.load(os.path.join(HERE, "golden", "spice_golden.npz"))
CASES = {"all": {}, "wave": dict(wave=G.WAVE_NM), "subfov": dict(sub=G.SUB_FOV)}


def _load(path):
    from euispice_coreg_b200._compat import fits_lite
    h = fits_lite.open(path)[0]
    return h.data, dict(h.header.items())


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("spicegold"))
    p_spice, imagers, spec = G.spice_files(d)
    return p_spice, imagers, spec, d


def test_oracle_reproduces_reference_synras_and_cubes_bit_exact(case, tmp_path):
    from euispice_coreg_b200._compat import fits_lite
    from oracle.hpc import HpcSearch
    from oracle.synras import build_synras, spice_l2_image, xy_header
    p_spice, imagers, spec, d = case
    d4, h4 = _load(p_spice)
    frames, hdrs = zip(*[_load(p) for p in imagers])
    syn, chosen = build_synras(h4, frames, hdrs, 100.0)
    assert np.array_equal(syn, GOLD["synras"], equal_nan=True) and len(set(chosen.tolist())) >= 3
    hxy = xy_header(h4)
    for k in ("CRVAL1", "CRVAL2", "CDELT1", "CDELT2", "CRPIX1", "CRPIX2", "PC1_1", "PC1_2", "PC2_1", "PC2_2"):
        assert hxy[k] == float(GOLD[f"synras_{k}"]), k
    # the header of the synthetic raster the reference wrote = middle imager header + the SPICE WCS keys
    mid = hdrs[int(chosen[len(chosen) // 2])]
    assert str(GOLD["synras_TELESCOP"]) == mid["TELESCOP"] and str(GOLD["synras_DATE-AVG"]) == h4["DATE-AVG"]
    h_syn = dict(mid)
    h_syn.update({k: v for k, v in hxy.items() if k != "WCSAXES"})
    h_syn["CRPIX1"] += G.CRPIX_OFFSET[0]
    h_syn["CRPIX2"] += G.CRPIX_OFFSET[1]
    h_syn["NAXIS1"], h_syn["NAXIS2"] = syn.shape[1], syn.shape[0]
    for name, kw in CASES.items():
        img, hdr = spice_l2_image(d4, h4, kw.get("wave", "all"), kw.get("sub"))
        cube = HpcSearch(syn, h_syn, img, hdr, G.LAG1, G.LAG2, [0], [0], [0]).cube()
        assert np.array_equal(cube, GOLD[f"cube_{name}"], equal_nan=True), name


@pytest.mark.gpu
def test_gpu_public_api_reproduces_reference_synras_and_cubes(case):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from euispice_coreg_b200.hdrshift import AlignmentSpice
    from euispice_coreg_b200.synras import SPICEComposedMapBuilder
    p_spice, imagers, spec, d = case
    b = SPICEComposedMapBuilder(p_spice, imagers, threshold_time=100.0, window_imager=0, window_spectro=0)
    name = b.process(folder_path_output=d, basename_output="synras.fits", print_filename=False, return_synras_name=True)
    syn, h_syn = _load(name)
    gold = GOLD["synras"]
    assert syn.shape == gold.shape and np.array_equal(np.isnan(syn), np.isnan(gold))
    assert np.nanmax(np.abs(syn - gold) / np.abs(gold)) < 2e-7 and np.mean(syn == gold) > 0.99
    for k in ("CRVAL1", "CRVAL2", "CDELT1", "CDELT2", "CRPIX1", "CRPIX2", "PC1_1", "PC1_2", "PC2_1", "PC2_2"):
        assert h_syn[k] == float(GOLD[f"synras_{k}"]), k
    assert h_syn["TELESCOP"] == str(GOLD["synras_TELESCOP"]) and h_syn["DATE-AVG"] == str(GOLD["synras_DATE-AVG"])
    p_off = os.path.join(d, "synras_offset.fits")
    G.offset_synras(name, p_off)
    kws = {"all": {}, "wave": dict(wavelength_interval_to_sum=list(G.WAVE_NM)), "subfov": dict(sub_fov_window=list(G.SUB_FOV))}
    for cname, kw in kws.items():
        a = AlignmentSpice(p_off, p_spice, lag_crval1=G.LAG1, lag_crval2=G.LAG2, lag_cdelt1=[0], lag_cdelt2=[0],
                           lag_crota=[0], parallelism=True, large_fov_window=0, small_fov_window=0, **kw)
        cube = a.align_using_helioprojective(return_type="corr")
        g = GOLD[f"cube_{cname}"]
        assert cube.shape == g.shape and np.array_equal(np.isnan(cube), np.isnan(g)), cname
        assert np.nanmax(np.abs(cube - g)) < R_TOL, cname
        assert np.nanargmax(cube) == np.nanargmax(g), cname
