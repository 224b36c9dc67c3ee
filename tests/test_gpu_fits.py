"""GPU decoder of RICE tile-compressed FITS images (`coreg_rice_decode`) against the CPU restatement in
oracle/rice.py, bit for bit, and the public API reading compressed inputs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _oracle_image(path):
    """Decode every tile of the file's compressed HDU with the oracle (slow, pure Python)."""
    from euispice_coreg_b200._compat import fits_lite
    from oracle import rice
    hdu = fits_lite.open(path)[1]
    th, body = hdu._thdr, hdu._body
    nx, ny, tw, tht = th["ZNAXIS1"], th["ZNAXIS2"], th["ZTILE1"], th["ZTILE2"]
    rowb, nrows = th["NAXIS1"], th["NAXIS2"]
    table = np.frombuffer(body, np.uint8, rowb * nrows).reshape(nrows, rowb)
    dsc = np.ascontiguousarray(table[:, :8]).view(">i4")
    heap = body[rowb * nrows:]
    bytepix = int(th["ZVAL2"])
    is_float = th["ZBITPIX"] < 0
    out = np.zeros((ny, nx), dtype=np.float64 if is_float else np.int64)
    rand = rice.fits_rand_values()[0]
    method = {"NO_DITHER": 0, "SUBTRACTIVE_DITHER_1": 1, "SUBTRACTIVE_DITHER_2": 2}.get(str(th.get("ZQUANTIZ", "")).strip())
    n = 0
    for y0 in range(0, ny, tht):
        for x0 in range(0, nx, tw):
            w, h = min(tw, nx - x0), min(tht, ny - y0)
            cnt, off = int(dsc[n, 0]), int(dsc[n, 1])
            q = rice.rice_decode(heap[off:off + cnt], w * h, int(th["ZVAL1"]), bytepix)
            if is_float:
                zs, zz = np.ascontiguousarray(table[n, 8:24]).view(">f8")
                q = rice.unquantize_tile(q, float(zs), float(zz), n + 1, int(th["ZDITHER0"]), rand, method,
                                         th.get("ZBLANK", None), np.float32 if th["ZBITPIX"] == -32 else np.float64)
            out[y0:y0 + h, x0:x0 + w] = np.asarray(q).reshape(h, w)
            n += 1
    return out


@pytest.mark.parametrize("dtype,bytepix,tile", [(np.int32, 4, None), (np.int16, 2, None), (np.int32, 4, (17, 5)),
                                                (np.int16, 2, (64, 3))])
def test_rice_integer_images_bit_exact(cuda, tmp_path, dtype, bytepix, tile):
    from euispice_coreg_b200._compat import fits_lite
    from oracle import rice
    rng = np.random.default_rng(bytepix + (0 if tile is None else 10))
    ii = np.iinfo(dtype)
    img = np.clip(1000 + np.cumsum(rng.integers(-30, 31, (37, 150)), axis=1), ii.min, ii.max).astype(dtype)
    img[5] = 42                                     # constant row: zero-difference blocks
    img[9, ::2], img[9, 1::2] = ii.min, ii.max      # verbatim blocks and wrap-around differences
    p = str(tmp_path / "i.fits")
    rice.write_compressed_image(p, img, tile=tile, bytepix=bytepix)
    got = fits_lite.open(p)[1].data
    assert got.dtype == dtype and np.array_equal(got, img) and np.array_equal(_oracle_image(p), img)


@pytest.mark.parametrize("method,tile", [(1, None), (2, None), (1, (50, 4))])
def test_rice_float_images_dither_and_blank(cuda, tmp_path, method, tile):
    from euispice_coreg_b200._compat import fits_lite
    from oracle import rice
    rng = np.random.default_rng(method)
    img = (500 + 100 * rng.standard_normal((41, 130))).astype(np.float32)
    img[3, 5] = img[20, 129] = np.nan
    img[7, 7] = 0.0
    p = str(tmp_path / "f.fits")
    rice.write_compressed_image(p, img, tile=tile, quantize_scale=0.25, zdither0=9876, method=method, blank=-2147483647)
    hdu = fits_lite.open(p)[1]
    dev = hdu.device_data()
    got = hdu.data
    assert got.dtype == np.float32 and dev.is_cuda and np.array_equal(dev.cpu().numpy(), got, equal_nan=True)
    assert np.array_equal(got, _oracle_image(p).astype(np.float32), equal_nan=True)      # same bits as the oracle
    assert np.isnan(got[3, 5]) and np.isnan(got[20, 129]) and np.nanmax(np.abs(got - img)) <= 0.126
    assert (got[7, 7] == 0.0) == (method == 2)


def test_alignment_reads_compressed_inputs(cuda, toy_pair, tmp_path):
    """Both inputs as RICE tile-compressed HDUs (the form Solar Orbiter L2 files have): same cube as from plain FITS
    files holding the de-quantised pixels."""
    from euispice_coreg_b200._compat import fits_lite
    from euispice_coreg_b200.hdrshift import Alignment
    from oracle import rice
    lags = dict(lag_crval1=np.arange(20, 29, 2.0), lag_crval2=np.arange(2, 11, 2.0), lag_cdelt1=[0], lag_cdelt2=[0],
                lag_crota=[0])
    comp, plain = [], []
    for k, src in enumerate(toy_pair[:2]):
        h = fits_lite.open(src)[0]
        cards = [(key, h.header[key]) for key in h.header.keys() if key not in ("SIMPLE", "BITPIX", "NAXIS", "NAXIS1",
                                                                                  "NAXIS2", "EXTEND")]
        pc = str(tmp_path / f"c{k}.fits")
        rice.write_compressed_image(pc, h.data.astype(np.float32), extra_cards=cards, quantize_scale=1.0 / 64, zdither0=k + 1)
        comp.append(pc)
        dq = fits_lite.open(pc)[-1]
        assert dq.header["CRVAL1"] == h.header["CRVAL1"] and dq.header["NAXIS1"] == h.header["NAXIS1"]
        pp = str(tmp_path / f"p{k}.fits")
        fits_lite.writeto(pp, [fits_lite.PrimaryHDU(dq.data, h.header)], overwrite=True)
        plain.append(pp)
    a = Alignment(comp[0], comp[1], parallelism=True, **lags).align_using_helioprojective(return_type="corr")
    b = Alignment(plain[0], plain[1], parallelism=True, **lags).align_using_helioprojective(return_type="corr")
    assert np.array_equal(a, b)
    i, j = np.unravel_index(np.nanargmax(a), a.shape)[:2]
    assert (lags["lag_crval1"][i], lags["lag_crval2"][j]) == (24.0, 6.0)


def test_lossless_fallback_tiles_are_patched_in(cuda, tmp_path):
    """A float tile cfitsio cannot quantise is stored losslessly (empty COMPRESSED_DATA descriptor, gzip of the
    big-endian pixels in GZIP_COMPRESSED_DATA): such tiles are decoded on the host and patched into the device image;
    the other tiles are the RICE decoder's, bit for bit."""
    from euispice_coreg_b200._compat import fits_lite
    from oracle import rice
    rng = np.random.default_rng(21)
    img = rng.lognormal(5, 1, (24, 50)).astype(np.float32)
    img[3, 4] = np.nan
    plain, mixed = str(tmp_path / "plain.fits"), str(tmp_path / "mixed.fits")
    kw = dict(tile=(25, 8), quantize_scale=0.25, zdither0=7, method=1, blank=-2147483647)
    rice.write_compressed_image(plain, img, **kw)
    rice.write_compressed_image(mixed, img, gzip_tiles=(1, 4), **kw)      # tiles (row 0, col 1) and (row 2, col 0)
    a = fits_lite.open(plain)[1].data
    b = fits_lite.open(mixed)[1].data
    assert a.dtype == b.dtype == np.float32
    lossless = np.zeros(img.shape, dtype=bool)
    lossless[0:8, 25:50] = True
    lossless[16:24, 0:25] = True
    assert np.array_equal(b[lossless], img[lossless], equal_nan=True)          # exact pixels where stored losslessly
    assert np.array_equal(b[~lossless], a[~lossless], equal_nan=True)          # RICE + de-quantisation elsewhere
    assert not np.array_equal(a[lossless], img[lossless], equal_nan=True)      # (the quantised version differs)
    dev = fits_lite.open(mixed)[1].device_data()
    assert np.array_equal(dev.cpu().numpy(), b, equal_nan=True)
