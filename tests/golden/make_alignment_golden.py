"""Generate tests/golden/alignment_golden.npz by running the REFERENCE's own `hdrshift.Alignment`
(`/root/reference/euispice_coreg/hdrshift/alignment.py`, `parallelism=True` branch: its multiprocessing fan-out,
shared memory, `_shift_header`, `_step`, scipy resampling, numba `c_correlate`, and -- for the Carrington frame --
all of `utils/rectify.py`) on the seeded synthetic pairs of `euispice_coreg_b200/_synth`.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    NUMBA_CACHE_DIR=$(mktemp -d) python tests/golden/make_alignment_golden.py
astropy / matplotlib / multiprocess are not installed here; `_ref_standins.py` replaces them (see its header). The
one arithmetic stand-in is `astropy.wcs.WCS`, whose pixel<->world numbers come from `oracle/wcs_tan.py` /
`oracle/wcs_car.py`: the goldens therefore pin everything the reference does around the WCS calls (lag enumeration and
cube axis order, header shifting and PCi_j rebuild, common grid, float32 rounding, masks, the Pearson function, the
Carrington "fa" chain with its NumPy dtype trail, which involves no WCS at all), not wcslib's own arithmetic.
"""
import hashlib
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def toy_files(d):
    from euispice_coreg_b200._synth.scene import make_pair, small_spec
    return make_pair(d, small_spec(96, 160, true_crval=(-12.0, 8.0)), tag="toy")[:2]


def car_files(d):
    from euispice_coreg_b200._synth.carmaps import make_car_pair
    return make_car_pair(d)[:2]


# every case: (pair, entry point, constructor kwargs, call kwargs)
CASES = {
    "hpc_crval_crota": ("toy", "align_using_helioprojective",
                        dict(lag_crval1=np.arange(18, 31, 3.0), lag_crval2=np.arange(0, 13, 3.0), lag_cdelt1=[0],
                             lag_cdelt2=[0], lag_crota=np.array([-0.5, 0.0, 0.5])), {}),
    "hpc_threshold_cdelt1_order1": ("toy", "align_using_helioprojective",
                                    dict(lag_crval1=np.array([21.0, 24.0, 27.0]), lag_crval2=np.array([3.0, 6.0]),
                                         lag_cdelt1=np.array([0.0, 0.01]), lag_cdelt2=[0], lag_crota=[0],
                                         small_fov_value_min=120.0, small_fov_value_max=1500.0, reprojection_order=1),
                                    {}),
    "hpc_deg_lags": ("toy", "align_using_helioprojective",
                     dict(lag_crval1=np.array([21.0, 24.0, 27.0]) / 3600.0, lag_crval2=np.array([3.0, 6.0]) / 3600.0,
                          lag_cdelt1=None, lag_cdelt2=None, lag_crota=None, unit_lag="deg"), {}),
    # fov_limits: [[lon_min, lon_max], [lat_min, lat_max]] in arcsec (astropy quantities for the reference)
    "hpc_fov_limits": ("toy", "align_using_helioprojective",
                       dict(lag_crval1=np.array([21.0, 24.0, 27.0]), lag_crval2=np.array([3.0, 6.0, 9.0]),
                            lag_cdelt1=[0], lag_cdelt2=[0], lag_crota=[0]),
                       dict(fov_limits=[[-80.0, 10.0], [-40.0, 45.0]])),
    "carrington_fa": ("toy", "align_using_carrington",
                      dict(lag_crval1=np.arange(18, 31, 3.0), lag_crval2=np.arange(0, 13, 6.0), lag_cdelt1=[0],
                           lag_cdelt2=[0], lag_crota=np.array([0.0, 0.75])),
                      dict(lonlims=(248.6, 251.4), latlims=(-3.2, -0.8), shape=[140, 120])),
    "initial_carrington": ("car", "align_using_initial_carrington",
                           dict(lag_crval1=np.array([0.08, 0.12, 0.16]) * 3600.0,
                                lag_crval2=np.array([-0.10, -0.06, -0.02]) * 3600.0, lag_cdelt1=[0], lag_cdelt2=[0],
                                lag_crota=np.array([0.0, 0.5])), {}),
}


def main():
    import _ref_standins
    _ref_standins.install(ROOT)
    warnings.simplefilter("ignore")
    from euispice_coreg.hdrshift.alignment import Alignment
    from euispice_coreg_b200._compat import fits_lite
    d = tempfile.mkdtemp()
    files = {"toy": toy_files(d), "car": car_files(d)}
    out = {}
    for pair, (p_large, p_small) in files.items():
        L, S = fits_lite.open(p_large)[0], fits_lite.open(p_small)[0]
        out[f"sha_{pair}"] = np.array(digest(L.data, S.data))
    for name, (pair, entry, ctor, call) in CASES.items():
        p_large, p_small = files[pair]
        a = Alignment(large_fov_known_pointing=p_large, small_fov_to_correct=p_small, parallelism=True,
                      counts_cpu_max=4, display_progress_bar=False, **ctor)
        call = dict(call)
        if "fov_limits" in call:
            call["fov_limits"] = [_ref_standins.Quantity(np.array(v), "arcsec") for v in call["fov_limits"]]
        cube = getattr(a, entry)(method="correlation", return_type="corr", **call)
        out[name] = np.asarray(cube, dtype=np.float64)
        i = np.unravel_index(np.nanargmax(cube), cube.shape)
        print(name, cube.shape, "max", np.nanmax(cube), "at", i, "zeros", int((cube == 0).sum()))
    dst = os.path.join(HERE, "alignment_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
