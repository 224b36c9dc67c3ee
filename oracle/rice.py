"""CPU restatement of the FITS tiled-image RICE_1 codec -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference reads its inputs with `astropy.io.fits` (`hdrshift/alignment.py:299-316`, `utils/Util.py:144-145`);
real Solar Orbiter L2 files are tile-compressed `CompImageHDU`s, which astropy decodes with its bundled cfitsio
(astropy 7.2.0, `poetry.lock:190-191`; cfitsio 4.x `ricecomp.c`, `quantize.c`, `imcompress.c`). Neither is under
`/root/reference` nor installed here, so this file restates the published algorithm:

  * Rice coding of pixel differences (White & Greenfield; cfitsio `fits_rcomp` / `fits_rdecomp`): the first pixel of a
    tile verbatim (BYTEPIX bytes, big-endian); differences to the previous pixel zig-zag mapped to unsigned; per block
    of BLOCKSIZE pixels an FS code of FSBITS bits holding fs + 1 (0: all differences zero; FSMAX + 1: differences
    verbatim in BBITS bits), else each difference as (diff >> fs) zero bits, a one bit and the low fs bits;
  * quantisation of floating-point tiles, value = (q - r + 0.5) * ZSCALE + ZZERO with the subtractive dither r taken
    from cfitsio's 10 000-number Park-Miller sequence (`fits_init_randoms`; its documented check value -- the 10 000th
    seed equals 1043618065 -- is asserted in tests/test_oracle.py) starting, for the tile in table row n (1-based),
    at index int(rand[(n + ZDITHER0 - 2) % 10000] * 500);
  * the FITS tiled-image convention (Pence et al. 2010): one tile per binary-table row, `COMPRESSED_DATA` as a
    variable-length byte array (`1PB` / `1QB` descriptors into the heap), `ZSCALE` / `ZZERO` / `ZBLANK` columns.

PARITY UNPINNED against cfitsio itself: the only known-answer vector available offline is the random-sequence check
value; everything else is encoder <-> decoder self-consistency plus the CUDA decoder == this decoder, bit for bit.
"""
from __future__ import annotations

import numpy as np

N_RANDOM = 10000
ZERO_VALUE = -2147483646      # SUBTRACTIVE_DITHER_2: this quantised value means exactly 0.0
_PARAMS = {1: (3, 6, 8), 2: (4, 14, 16), 4: (5, 25, 32)}    # BYTEPIX -> (FSBITS, FSMAX, BBITS)


def fits_rand_values():
    """cfitsio `fits_init_randoms`: (float32[10000], last seed)."""
    a, m = 16807.0, 2147483647.0
    seed = 1.0
    out = np.empty(N_RANDOM, dtype=np.float32)
    for i in range(N_RANDOM):
        temp = a * seed
        seed = temp - m * int(temp / m)
        out[i] = np.float32(seed / m)
    return out, int(seed)


class _BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, value, nbits):
        if nbits == 0:
            return
        self.acc = (self.acc << nbits) | (value & ((1 << nbits) - 1))
        self.n += nbits
        while self.n >= 8:
            self.n -= 8
            self.out.append((self.acc >> self.n) & 0xFF)
        self.acc &= (1 << self.n) - 1

    def finish(self):
        if self.n:
            self.out.append((self.acc << (8 - self.n)) & 0xFF)
            self.acc, self.n = 0, 0
        return bytes(self.out)


def rice_encode(values, blocksize=32, bytepix=4):
    """One tile of integers -> RICE_1 byte string (cfitsio `fits_rcomp`, same choice of fs per block)."""
    fsbits, fsmax, bbits = _PARAMS[bytepix]
    mask = (1 << bbits) - 1
    a = [int(v) for v in np.asarray(values).ravel()]
    w = _BitWriter()
    w.put(a[0] & mask, bbits)
    last = a[0]
    for i0 in range(0, len(a), blocksize):
        blk = a[i0:i0 + blocksize]
        diffs = []
        for v in blk:
            d = v - last
            d = ((d + (1 << (bbits - 1))) & mask) - (1 << (bbits - 1))      # wrap like the C integer arithmetic
            diffs.append(((d << 1) ^ (d >> (bbits - 1))) & mask)            # zig-zag: >= 0 -> 2d, < 0 -> ~(2d)
            last = v
        pixelsum = float(sum(diffs))
        dpsum = (pixelsum - (len(blk) // 2) - 1) / len(blk)
        if dpsum < 0:
            dpsum = 0.0
        psum = int(dpsum) >> 1
        fs = 0
        while psum > 0:
            psum >>= 1
            fs += 1
        if fs >= fsmax:
            w.put(fsmax + 1, fsbits)
            for d in diffs:
                w.put(d, bbits)
        elif fs == 0 and pixelsum == 0:
            w.put(0, fsbits)
        else:
            w.put(fs + 1, fsbits)
            for d in diffs:
                top = d >> fs
                w.put(0, top)            # `top` zero bits ...
                w.put(1, 1)              # ... closed by a one
                w.put(d, fs)
    return w.finish()


def rice_decode(buf, nx, blocksize=32, bytepix=4):
    """RICE_1 byte string -> nx integers (cfitsio `fits_rdecomp`). Returns int64 values wrapped to BBITS bits."""
    fsbits, fsmax, bbits = _PARAMS[bytepix]
    mask = (1 << bbits) - 1
    half = 1 << (bbits - 1)
    data = bytes(buf)
    pos = [0]
    state = [0, 0]   # bit accumulator, number of valid bits

    def take(n):
        while state[1] < n:
            state[0] = (state[0] << 8) | (data[pos[0]] if pos[0] < len(data) else 0)
            pos[0] += 1
            state[1] += 8
        state[1] -= n
        v = (state[0] >> state[1]) & ((1 << n) - 1)
        state[0] &= (1 << state[1]) - 1
        return v

    def signed(v):
        v &= mask
        return v - (1 << bbits) if v >= half else v

    last = signed(take(bbits))
    out = np.empty(nx, dtype=np.int64)
    i = 0
    while i < nx:
        fs = take(fsbits) - 1
        imax = min(nx, i + blocksize)
        while i < imax:
            if fs < 0:
                diff = 0
            elif fs == fsmax:
                diff = take(bbits)
            else:
                nzero = 0
                while take(1) == 0:
                    nzero += 1
                diff = (nzero << fs) | take(fs)
            d = (diff >> 1) if (diff & 1) == 0 else ~(diff >> 1)
            last = signed(d + last)
            out[i] = last
            i += 1
    return out


def quantize_tile(tile, scale, zero, row, zdither0, rand=None, method=1):
    """float tile -> int32 with subtractive dither (`fits_quantize_float`'s final loop): NINT((x - zero) / scale + r - 0.5).
    `row` is the tile's 1-based table row."""
    rand = fits_rand_values()[0] if rand is None else rand
    x = np.asarray(tile, dtype=np.float64).ravel()
    out = np.empty(x.size, dtype=np.int64)
    iseed = (row + zdither0 - 2) % N_RANDOM
    nextrand = int(rand[iseed] * 500)
    for i, v in enumerate(x):
        if method == 2 and v == 0.0:
            out[i] = ZERO_VALUE
        else:
            t = (v - zero) / scale + float(rand[nextrand]) - 0.5
            out[i] = int(t + 0.5) if t >= 0 else int(t - 0.5)      # cfitsio NINT
        nextrand += 1
        if nextrand == N_RANDOM:
            iseed = (iseed + 1) % N_RANDOM
            nextrand = int(rand[iseed] * 500)
    return out


def unquantize_tile(q, scale, zero, row, zdither0, rand=None, method=1, blank=None, out_dtype=np.float32):
    """int32 tile -> float (`unquantize_i4r4` / `_i4r8`): (q - r + 0.5) * scale + zero, NaN where q == blank."""
    rand = fits_rand_values()[0] if rand is None else rand
    q = np.asarray(q).ravel()
    out = np.empty(q.size, dtype=np.float64)
    iseed = (row + zdither0 - 2) % N_RANDOM
    nextrand = int(rand[iseed] * 500)
    for i, v in enumerate(q):
        if blank is not None and v == blank:
            out[i] = np.nan
        elif method == 2 and v == ZERO_VALUE:
            out[i] = 0.0
        elif method == 0:
            out[i] = float(v) * scale + zero
        else:
            out[i] = (float(v) - float(rand[nextrand]) + 0.5) * scale + zero
        nextrand += 1
        if nextrand == N_RANDOM:
            iseed = (iseed + 1) % N_RANDOM
            nextrand = int(rand[iseed] * 500)
    return out.astype(out_dtype)


# ---------------------------------------------------------------------------------------------------------
# a writer for tests: image -> tile-compressed FITS file (row tiles by default)
# ---------------------------------------------------------------------------------------------------------
def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20}"
    elif isinstance(value, (float, np.floating)):
        v = f"{float(value):>20.15G}"
    else:
        v = "'" + f"{str(value):<8}" + "'"
        v = f"{v:<20}"
    return f"{key:<8}= {v} / {comment}"[:80].ljust(80)


def write_compressed_image(path, data, extra_cards=(), tile=None, quantize_scale=None, zdither0=1, method=1,
                           blank=None, bytepix=4, blocksize=32, gzip_tiles=()):
    """Write `data` (2-D int16/int32, or float32/float64 with `quantize_scale`) as PRIMARY (no data) + one ZIMAGE
    binary-table extension with RICE_1 tiles. `extra_cards`: (key, value) pairs copied into the extension header.
    gzip_tiles: 0-based tile numbers of a float image stored losslessly instead -- empty COMPRESSED_DATA descriptor,
    gzip-compressed big-endian pixel values in a GZIP_COMPRESSED_DATA column: what cfitsio does with a tile it cannot
    quantise (imcompress.c, imcomp_compress_tile)."""
    import zlib
    data = np.asarray(data)
    ny, nx = data.shape
    tw, th = (nx, 1) if tile is None else tile
    is_float = data.dtype.kind == "f"
    rand = fits_rand_values()[0]
    rows = []
    for ty in range(0, ny, th):
        for tx in range(0, nx, tw):
            rows.append(data[ty:ty + th, tx:tx + tw])
    heap = bytearray()
    desc, scales, zeros, gz = [], [], [], []
    for n, t in enumerate(rows, start=1):
        if is_float and (n - 1) in gzip_tiles:
            b = zlib.compress(np.ascontiguousarray(t).astype(t.dtype.newbyteorder(">")).tobytes())
            desc.append((0, 0))
            gz.append((len(b), len(heap)))
            heap += b
            scales.append(0.0)
            zeros.append(0.0)
            continue
        gz.append((0, 0))
        if is_float:
            finite = np.isfinite(t)
            zero = float(np.min(t[finite])) if finite.any() else 0.0
            q = quantize_tile(np.where(finite, t, zero), quantize_scale, zero, n, zdither0, rand, method)
            if blank is not None:
                q[~finite.ravel()] = blank
            scales.append(quantize_scale)
            zeros.append(zero)
        else:
            q = t.ravel().astype(np.int64)
        b = rice_encode(q, blocksize, bytepix)
        desc.append((len(b), len(heap)))
        heap += b
    with_gz = bool(gzip_tiles) and is_float
    ncols_bytes = 8 + (16 if is_float else 0) + (8 if with_gz else 0)
    table = bytearray()
    for i, (cnt, off) in enumerate(desc):
        table += np.array([cnt, off], dtype=">i4").tobytes()
        if is_float:
            table += np.array([scales[i], zeros[i]], dtype=">f8").tobytes()
        if with_gz:
            table += np.array(gz[i], dtype=">i4").tobytes()
    cards = [_card("XTENSION", "BINTABLE"), _card("BITPIX", 8), _card("NAXIS", 2), _card("NAXIS1", ncols_bytes),
             _card("NAXIS2", len(rows)), _card("PCOUNT", len(heap)), _card("GCOUNT", 1),
             _card("TFIELDS", (3 if is_float else 1) + (1 if with_gz else 0)), _card("TTYPE1", "COMPRESSED_DATA"),
             _card("TFORM1", f"1PB({max(c for c, _ in desc)})")]
    if is_float:
        cards += [_card("TTYPE2", "ZSCALE"), _card("TFORM2", "1D"), _card("TTYPE3", "ZZERO"), _card("TFORM3", "1D")]
    if with_gz:
        cards += [_card("TTYPE4", "GZIP_COMPRESSED_DATA"), _card("TFORM4", f"1PB({max(c for c, _ in gz)})")]
    zbitpix = {"f4": -32, "f8": -64, "i2": 16, "i4": 32, "u1": 8}[data.dtype.str[1:]]
    cards += [_card("ZIMAGE", True), _card("ZCMPTYPE", "RICE_1"), _card("ZBITPIX", zbitpix), _card("ZNAXIS", 2),
              _card("ZNAXIS1", nx), _card("ZNAXIS2", ny), _card("ZTILE1", tw), _card("ZTILE2", th),
              _card("ZNAME1", "BLOCKSIZE"), _card("ZVAL1", blocksize), _card("ZNAME2", "BYTEPIX"),
              _card("ZVAL2", bytepix)]
    if is_float:
        cards += [_card("ZQUANTIZ", {0: "NO_DITHER", 1: "SUBTRACTIVE_DITHER_1", 2: "SUBTRACTIVE_DITHER_2"}[method]),
                  _card("ZDITHER0", zdither0)]
        if blank is not None:
            cards.append(_card("ZBLANK", blank))
    cards += [_card(k, v) for k, v in extra_cards]
    cards.append("END".ljust(80))
    prim = [_card("SIMPLE", True), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True), "END".ljust(80)]
    out = bytearray()
    for block in (prim, cards):
        raw = "".join(block).encode("ascii")
        out += raw + b" " * ((-len(raw)) % 2880)
        if block is cards:
            body = bytes(table) + bytes(heap)
            out += body + b"\0" * ((-len(body)) % 2880)
    with open(path, "wb") as f:
        f.write(out)
