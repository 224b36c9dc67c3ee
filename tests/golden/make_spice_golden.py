"""Generate tests/golden/spice_golden.npz by running the REFERENCE's own `synras.SPICEComposedMapBuilder` and
`hdrshift.AlignmentSpice` on the seeded synthetic SPICE case of `euispice_coreg_b200/_synth/spice.py`.

Run in the build container only:   NUMBA_CACHE_DIR=$(mktemp -d) python tests/golden/make_spice_golden.py
Same stand-ins as `make_alignment_golden.py` (`_ref_standins.py`); here `astropy.wcs.WCS` also has to answer for the
4-axis SPICE header (`dropaxis`, `sub(['spectral'])`, `pixel_to_world(x, y, t)`, `to_header()`), all of it the
stand-in's restatement, so these goldens pin the reference's logic AROUND the WCS: which imager frame each raster column
takes (`_return_mean_time`, `_find_closest_imager_time`), the per-column sampling and the composed header
(`synras/map_builder.py:87-215`), slit-edge rows, spectral sum, wavelength / sub-FOV selections and the header the
search starts from (`hdrshift/alignment_spice.py:189-323`), then the search itself.
"""
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

LAG1 = np.arange(-14.0, -1.0, 2.0)      # arcsec; the planted correction of the toy case is (-8, 12)
LAG2 = np.arange(6.0, 19.0, 2.0)
WAVE_NM = (97.68, 97.72)                # wavelength_interval_to_sum
SUB_FOV = (-45.0, 15.0, -30.0, 25.0)    # arcsec: lon_min, lon_max, lat_min, lat_max (the header claims (-12, -2))
CRPIX_OFFSET = (0.37, -0.21)            # see tests/test_gpu_spice.py: keeps the one-time cut off the closed-bound knife edge


def spice_files(d):
    from euispice_coreg_b200._synth.spice import make_spice_case, small_spice_spec
    return make_spice_case(d, small_spice_spec(nbin2=8, pxbeg2=192), tag="toy")


def offset_synras(path_in, path_out):
    """Copy of the synthetic raster whose CRPIX is moved by a fraction of a pixel (CRVAL kept): a synras has the SPICE
    header by construction, so without this whole border rows of the one-time cut sit on map_coordinates' closed bound."""
    from euispice_coreg_b200._compat import fits_lite
    hdu = fits_lite.open(path_in)[0]
    h = hdu.header.copy()
    h["CRPIX1"] = h["CRPIX1"] + CRPIX_OFFSET[0]
    h["CRPIX2"] = h["CRPIX2"] + CRPIX_OFFSET[1]
    fits_lite.writeto(path_out, [fits_lite.PrimaryHDU(hdu.data, h)], overwrite=True)


def main():
    import _ref_standins
    from _ref_standins import Quantity
    _ref_standins.install(ROOT)
    warnings.simplefilter("ignore")
    from euispice_coreg.hdrshift.alignment_spice import AlignmentSpice
    from euispice_coreg.synras.map_builder import SPICEComposedMapBuilder
    from euispice_coreg_b200._compat import fits_lite
    d = tempfile.mkdtemp()
    p_spice, imagers, spec = spice_files(d)
    out = {}
    c = SPICEComposedMapBuilder(path_to_spectro=p_spice, list_imager_paths=imagers,
                                threshold_time=Quantity(100.0, "s"), window_imager=0, window_spectro=0)
    c.process(folder_path_output=d, basename_output="synras.fits", print_filename=False)
    syn = fits_lite.open(os.path.join(d, "synras.fits"))[0]
    out["synras"] = np.asarray(syn.data, dtype=np.float64)
    for k in ("CRVAL1", "CRVAL2", "CDELT1", "CDELT2", "CRPIX1", "CRPIX2", "PC1_1", "PC1_2", "PC2_1", "PC2_2"):
        out[f"synras_{k}"] = np.float64(syn.header[k])
    out["synras_DATE-AVG"] = np.array(str(syn.header["DATE-AVG"]))
    out["synras_TELESCOP"] = np.array(str(syn.header["TELESCOP"]))
    print("synras", out["synras"].shape, "NaN", int(np.isnan(out["synras"]).sum()), "CRVAL1", out["synras_CRVAL1"])
    p_syn = os.path.join(d, "synras_offset.fits")
    offset_synras(os.path.join(d, "synras.fits"), p_syn)
    cases = {"all": {}, "wave": dict(wavelength_interval_to_sum=[Quantity(WAVE_NM[0], "nm"), Quantity(WAVE_NM[1], "nm")]),
             "subfov": dict(sub_fov_window=[Quantity(v, "arcsec") for v in SUB_FOV])}
    for name, kw in cases.items():
        a = AlignmentSpice(large_fov_known_pointing=p_syn, small_fov_to_correct=p_spice, lag_crval1=LAG1, lag_crval2=LAG2,
                           lag_cdelt1=[0], lag_cdelt2=[0], lag_crota=[0], parallelism=True, counts_cpu_max=4,
                           large_fov_window=0, small_fov_window=0, **kw)
        cube = a.align_using_helioprojective(method="correlation", return_type="corr")
        out[f"cube_{name}"] = np.asarray(cube, dtype=np.float64)
        i = np.unravel_index(np.nanargmax(cube), cube.shape)
        print(name, cube.shape, "max", np.nanmax(cube), "at", (LAG1[i[0]], LAG2[i[1]]))
    dst = os.path.join(HERE, "spice_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
