"""Frame-sequence driver (BASELINE config 5 pattern): every frame's cube equals the single-pair `Alignment` cube."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LAGS = dict(lag_crval1=np.arange(18, 31, 2.0), lag_crval2=np.arange(0, 13, 2.0), lag_cdelt1=[0], lag_cdelt2=[0],
            lag_crota=[0.0, 0.5])


@pytest.fixture(scope="module")
def frames(tmp_path_factory):
    from euispice_coreg_b200._synth.scene import make_pair, master_scene, small_spec
    d = str(tmp_path_factory.mktemp("seq"))
    spec0 = small_spec(96, 160, true_crval=(-12.0, 8.0))
    sky = master_scene(spec0)
    p_large, p0, _ = make_pair(d, spec0, tag="f0", sky=sky)
    paths = [p0]
    for i, jit in enumerate([(1.3, -0.8), (-2.1, 0.4), (0.6, 2.2)], start=1):
        spec = small_spec(96, 160, true_crval=(-12.0, 8.0), jitter=jit, noise_seed=200 + i)
        paths.append(make_pair(d, spec, tag=f"f{i}", sky=sky, write_large=False)[1])
    return p_large, paths


def test_sequence_equals_per_pair_alignment(frames):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from euispice_coreg_b200.hdrshift import Alignment, SequenceAlignment
    p_large, paths = frames
    seq = SequenceAlignment(p_large, paths, **LAGS)
    cubes = seq.align_using_helioprojective(return_type="corr")
    assert len(cubes) == len(paths) and seq.frames_per_s > 0
    for p, cube in zip(paths, cubes):
        one = Alignment(p_large, p, parallelism=True, **LAGS).align_using_helioprojective(return_type="corr")
        assert np.array_equal(cube, one, equal_nan=True)
    res = SequenceAlignment(p_large, paths[:2], **LAGS).align_using_helioprojective()
    ref = Alignment(p_large, paths[1], parallelism=True, **LAGS).align_using_helioprojective()
    assert res[1].max_index == ref.max_index and res[1].shift_arcsec == ref.shift_arcsec
    # the jittered frames peak at different sub-lag shifts
    assert res[0].shift_arcsec != res[1].shift_arcsec


def test_jitter_correction_imagers_chain(tmp_path):
    """`jitter_correction_imagers` (reference: jitter_correction/jitter_correction.py:14-174): frames whose headers all
    claim the same pointing while the true pointing jitters; sublists [0, 1, 2], [2, 3] -- the second one is aligned to
    the CORRECTED frame 2 the first one wrote. Helioprojective and Carrington ("fa") co-alignment."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200._synth.scene import make_pair, master_scene, small_spec
    from euispice_coreg_b200.jitter_correction.jitter_correction import jitter_correction_imagers
    d = str(tmp_path)
    true = (-12.0, 8.0)
    jitters = [(0.0, 0.0), (1.3, -0.8), (-2.1, 0.4), (0.6, 2.2)]
    sky = master_scene(small_spec(96, 160, true_crval=true))
    paths = []
    for i, jit in enumerate(jitters):
        spec = small_spec(96, 160, true_crval=true, jitter=jit, true_shift=jit, noise_seed=300 + i,
                          date=f"2022-03-17T09:5{i}:45.000")
        paths.append(make_pair(d, spec, tag=f"j{i}", sky=sky, write_large=False)[1])
        assert abs(fits_lite.open(paths[-1])[0].header["CRVAL1"] - true[0]) < 1e-12
    lag = np.arange(-4.0, 4.1, 0.5)
    for method, kw in (("helioprojective", {}),
                       ("carrington", dict(lonlims=(249.3, 250.7), latlims=(-2.6, -1.4), shape=[120, 110]))):
        out = str(tmp_path / f"out_{method}")
        import os
        os.makedirs(out)
        res = jitter_correction_imagers(paths, out, lag_crval1=lag, lag_crval2=lag, sublist_length=2, overlap=1,
                                        alignement_method=method, **kw)
        assert len(res) == 3
        for i, jit in enumerate(jitters):
            h = fits_lite.open(os.path.join(out, os.path.basename(paths[i])))[0].header
            assert abs(h["CRVAL1"] - (true[0] + jit[0])) < 0.3 and abs(h["CRVAL2"] - (true[1] + jit[1])) < 0.3, (method, i)
