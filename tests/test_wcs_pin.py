"""Pin of the WCS restatements against the real wcslib -- runs wherever astropy is importable, skips elsewhere.

`oracle/wcs_tan.py`, `oracle/wcs_car.py` and the product's `_compat/wcs.py` restate what the reference obtains from
`astropy.wcs.WCS` (wcslib): `utils/Util.py:283-312`, `hdrshift/alignment.py:1038-1069, 344-399`. astropy is absent from
the build image and from the GPU boxes seen so far, so this boundary is "parity unpinned" (DESIGN.md section 4); the
probe below closes it on the first machine that has astropy, with no change to the repository.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from _ref_standins import real_astropy  # noqa: E402

needs_astropy = pytest.mark.skipif(not real_astropy(), reason="astropy / wcslib not installed: WCS boundary stays unpinned")


def _tan_header(crval=(-100.0, 50.0), crota=3.0, unit="arcsec", n=(200, 120), cdelt=0.492, lonpole=180.0):
    rho = np.deg2rad(crota)
    return {"NAXIS": 2, "NAXIS1": n[0], "NAXIS2": n[1], "CTYPE1": "HPLN-TAN", "CTYPE2": "HPLT-TAN", "CUNIT1": unit,
            "CUNIT2": unit, "CRPIX1": (n[0] + 1) / 2, "CRPIX2": (n[1] + 1) / 2, "CDELT1": cdelt, "CDELT2": cdelt,
            "CRVAL1": crval[0], "CRVAL2": crval[1], "PC1_1": np.cos(rho), "PC1_2": -np.sin(rho), "PC2_1": np.sin(rho),
            "PC2_2": np.cos(rho), "LONPOLE": lonpole}


def _car_header(crval=(250.0, -2.0), n=(180, 90), cdelt=0.05):
    return {"NAXIS": 2, "NAXIS1": n[0], "NAXIS2": n[1], "CTYPE1": "CRLN-CAR", "CTYPE2": "CRLT-CAR", "CUNIT1": "deg",
            "CUNIT2": "deg", "CRPIX1": (n[0] + 1) / 2, "CRPIX2": (n[1] + 1) / 2, "CDELT1": cdelt, "CDELT2": cdelt,
            "CRVAL1": crval[0], "CRVAL2": crval[1], "PC1_1": 1.0, "PC1_2": 0.0, "PC2_1": 0.0, "PC2_2": 1.0}


def test_probe_reports_the_environment():
    """Always runs: records in the test log whether this box can pin the WCS boundary."""
    print("astropy importable:", real_astropy())


@needs_astropy
@pytest.mark.parametrize("hdr", [_tan_header(), _tan_header(unit="deg", crval=(0.3, -0.2), cdelt=4.44 / 3600),
                                 _tan_header(crval=(2000.0, -1500.0), cdelt=4.44, crota=-12.5)])
def test_tan_restatements_against_wcslib(hdr):
    from astropy.io import fits
    from astropy.wcs import WCS
    from euispice_coreg_b200._compat.wcs import TanWcs
    from oracle import wcs_tan
    w = WCS(fits.Header(hdr))
    x, y = np.meshgrid(np.arange(hdr["NAXIS1"], dtype=float), np.arange(hdr["NAXIS2"], dtype=float))
    lon, lat = w.wcs_pix2world(x, y, 0)
    o = wcs_tan.WcsTan(hdr)
    lon_o, lat_o = o.pixel_to_world(x, y)
    d = (lon - lon_o + 180.0) % 360.0 - 180.0
    assert np.max(np.abs(d)) < 1e-11 and np.max(np.abs(lat - lat_o)) < 1e-11
    xb, yb = w.wcs_world2pix(lon, lat, 0)
    xo, yo = o.world_to_pixel(lon, lat)
    assert np.max(np.abs(xb - xo)) < 1e-9 and np.max(np.abs(yb - yo)) < 1e-9
    p = TanWcs.from_header(hdr)
    lon_p, lat_p = p.pixel_to_world(x, y)
    assert np.max(np.abs((lon - lon_p + 180.0) % 360.0 - 180.0)) < 1e-11 and np.max(np.abs(lat - lat_p)) < 1e-11


@needs_astropy
@pytest.mark.parametrize("hdr", [_car_header(), _car_header(crval=(10.0, 0.0)), _car_header(crval=(300.0, 35.0))])
def test_car_restatements_against_wcslib(hdr):
    from astropy.io import fits
    from astropy.wcs import WCS
    from oracle import wcs_car
    w = WCS(fits.Header(hdr))
    x, y = np.meshgrid(np.arange(hdr["NAXIS1"], dtype=float), np.arange(hdr["NAXIS2"], dtype=float))
    lon, lat = w.wcs_pix2world(x, y, 0)
    o = wcs_car.WcsCar(hdr)
    lon_o, lat_o = o.pixel_to_world(x, y)
    assert np.max(np.abs((lon - lon_o + 180.0) % 360.0 - 180.0)) < 1e-10 and np.max(np.abs(lat - lat_o)) < 1e-10
    xb, yb = w.wcs_world2pix(lon, lat, 0)
    xo, yo = o.world_to_pixel(lon, lat)
    assert np.max(np.abs(xb - xo)) < 1e-8 and np.max(np.abs(yb - yo)) < 1e-8


def _sunpy_stack():
    try:
        import astropy  # noqa: F401
        import reproject  # noqa: F401
        import sunpy.map  # noqa: F401
        return True
    except Exception:
        return False


@pytest.mark.skipif(not _sunpy_stack(), reason="sunpy / reproject not installed: the solar-surface reprojection stays unpinned")
def test_surface_reprojection_against_sunpy(tmp_path):
    """`oracle/surface_reproject.reproject_to` against the real thing -- `Map.reproject_to` under
    `propagate_with_solar_surface()` (`hdrshift/alignment.py:939-985`) -- on a synthetic pair seen by two observers half an
    hour apart. Runs wherever sunpy + reproject are importable (never in the build image: sunpy 6.1.2 / reproject 0.14.1
    are the reference's locked versions)."""
    import copy

    import astropy.constants
    from astropy.io import fits
    from astropy.wcs import WCS
    from sunpy.coordinates import propagate_with_solar_surface
    from sunpy.map import Map
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200._synth.scene import make_pair, small_spec
    from oracle import surface_reproject as sr
    from oracle.hpc import check_and_create_pcij
    pl, ps, _ = make_pair(str(tmp_path), small_spec(96, 160, true_crval=(-12.0, 8.0)), tag="pin")
    L, S = fits_lite.open(pl)[0], fits_lite.open(ps)[0]
    hl, hs = dict(L.header.items()), dict(S.header.items())
    hl.update({"HGLN_OBS": 10.3, "HGLT_OBS": -2.9, "DSUN_OBS": 5.5e10, "DATE-AVG": "2022-03-17T10:20:45.000"})
    hs.update({"HGLN_OBS": 10.0, "HGLT_OBS": -3.0})
    for h in (hl, hs):
        check_and_create_pcij(h)
        h.pop("CRLN_OBS", None), h.pop("CRLT_OBS", None)       # one observer definition per header
    rsun = (1.004 * astropy.constants.R_sun).to("m").value
    assert abs(rsun - 1.004 * sr.R_SUN_M) < 1e-3
    want_oracle = sr.reproject_to(np.asarray(L.data, dtype=np.float64), hl, hs, rsun)
    map_ref = Map(np.asarray(L.data, dtype=np.float64), fits.Header(hl))
    map_ref.meta["rsun_ref"] = rsun
    hdr_out = copy.deepcopy(hs)
    hdr_out["RSUN_REF"] = rsun
    with propagate_with_solar_surface():
        got = map_ref.reproject_to(WCS(fits.Header(hdr_out))).data
    both = np.isfinite(got) & np.isfinite(want_oracle)
    assert both.mean() > 0.9 and (np.isfinite(got) != np.isfinite(want_oracle)).mean() < 0.01
    assert np.max(np.abs(got[both] - want_oracle[both])) < 1e-6 * np.abs(want_oracle[both]).max()
