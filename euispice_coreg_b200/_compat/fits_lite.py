"""Dependency-free FITS reader/writer for uncompressed image HDUs.

Stands in for the `astropy.io.fits` calls of the reference's pointing-search path
(`hdrshift/alignment.py:299-316`, `utils/Util.py:106-159`, `synras/map_builder.py:106-131,192-203`).
Scope: primary + IMAGE extension HDUs, BITPIX in {8, 16, 32, 64, -32, -64}, BSCALE/BZERO,
string/logical/int/float cards, CONTINUE-less headers. Tile-compressed (RICE) HDUs are the
"next" row of SURVEY.md section 8(f) and raise a clear error here.

If astropy is importable the caller may still hand astropy headers to the package: everything
downstream only needs the mapping protocol (`in`, `[]`, `.copy()`, `.keys()`).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

BLOCK = 2880
CARD = 80

_BITPIX_DTYPE = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
_DTYPE_BITPIX = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}


class Header:
    """Ordered FITS header with the subset of the `astropy.io.fits.Header` mapping API the
    reference uses: `hdr[k]`, `hdr[k] = v`, `k in hdr`, `hdr.copy()`, `hdr.keys()`, `hdr.get`."""

    def __init__(self, cards=None):
        self._cards: "OrderedDict[str, tuple]" = OrderedDict()
        self._commentary: list[tuple[str, str]] = []
        if cards is not None:
            if isinstance(cards, Header):
                self._cards = OrderedDict(cards._cards)
                self._commentary = list(cards._commentary)
            else:
                items = cards.items() if hasattr(cards, "items") else cards
                for k, v in items:
                    self[k] = v

    # -- mapping protocol -------------------------------------------------
    @staticmethod
    def _norm(key):
        return str(key).strip().upper()

    def __contains__(self, key):
        return self._norm(key) in self._cards

    def __getitem__(self, key):
        k = self._norm(key)
        if k not in self._cards:
            raise KeyError(f"Keyword {key!r} not found.")
        return self._cards[k][0]

    def __setitem__(self, key, value):
        k = self._norm(key)
        comment = ""
        if isinstance(value, tuple) and len(value) == 2 and isinstance(value[1], str):
            value, comment = value
        elif k in self._cards:
            comment = self._cards[k][1]
        if isinstance(value, np.generic):
            value = value.item()
        self._cards[k] = (value, comment)

    def __delitem__(self, key):
        del self._cards[self._norm(key)]

    def __iter__(self):
        return iter(self._cards)

    def __len__(self):
        return len(self._cards)

    def keys(self):
        return self._cards.keys()

    def items(self):
        return [(k, v[0]) for k, v in self._cards.items()]

    def values(self):
        return [v[0] for v in self._cards.values()]

    def get(self, key, default=None):
        k = self._norm(key)
        return self._cards[k][0] if k in self._cards else default

    def comment(self, key):
        return self._cards[self._norm(key)][1]

    def copy(self):
        return Header(self)

    def update(self, other):
        for k, v in (other.items() if hasattr(other, "items") else other):
            self[k] = v

    def __eq__(self, other):
        if not isinstance(other, Header):
            return NotImplemented
        return self.items() == other.items()

    def __repr__(self):
        return "\n".join(_format_card(k, v, c).rstrip() for k, (v, c) in self._cards.items())


# --------------------------------------------------------------------------
# card parsing / formatting
# --------------------------------------------------------------------------
def _parse_value(field: str):
    """Parse the value field (columns 11-80) of a card -> (value, comment)."""
    s = field.strip()
    if not s:
        return None, ""
    if s[0] == "'":
        # quoted string; '' is an escaped quote
        i = 1
        out = []
        while i < len(s):
            if s[i] == "'":
                if i + 1 < len(s) and s[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(s[i])
            i += 1
        val = "".join(out).rstrip()
        rest = s[i + 1:]
        comment = rest.split("/", 1)[1].strip() if "/" in rest else ""
        return val, comment
    if "/" in s:
        vs, comment = s.split("/", 1)
        vs, comment = vs.strip(), comment.strip()
    else:
        vs, comment = s, ""
    if vs == "T":
        return True, comment
    if vs == "F":
        return False, comment
    if not vs:
        return None, comment
    try:
        return int(vs), comment
    except ValueError:
        pass
    try:
        return float(vs.replace("D", "E").replace("d", "e")), comment
    except ValueError:
        return vs, comment


def _format_float(v: float) -> str:
    if v != v:
        return "'NaN'"
    s = repr(float(v)).upper()
    if "INF" in s:
        return "'%s'" % s
    if "E" in s:
        mant, exp = s.split("E")
        if "." not in mant:
            mant += ".0"
        s = f"{mant}E{int(exp):+03d}"
    elif "." not in s:
        s += ".0"
    return s


def _format_card(key: str, value, comment: str = "") -> str:
    if key in ("COMMENT", "HISTORY", ""):
        return f"{key:<8}{str(value)}"[:CARD].ljust(CARD)
    if isinstance(value, bool):
        vs = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        vs = f"{int(value):>20d}"
    elif isinstance(value, (float, np.floating)):
        vs = f"{_format_float(float(value)):>20}"
    elif value is None:
        vs = " " * 20
    else:
        txt = str(value).replace("'", "''")
        vs = f"'{txt:<8}'"
        vs = f"{vs:<20}"
    if len(key) > 8:
        card = f"HIERARCH {key} = {vs.strip()}"
    else:
        card = f"{key:<8}= {vs}"
    if comment:
        card += f" / {comment}"
    return card[:CARD].ljust(CARD)


def _parse_header_block(raw: bytes):
    hdr = Header()
    for i in range(0, len(raw), CARD):
        card = raw[i:i + CARD].decode("ascii", errors="replace")
        key = card[:8].strip()
        if key == "END":
            return hdr, True
        if key in ("COMMENT", "HISTORY", ""):
            hdr._commentary.append((key, card[8:].rstrip()))
            continue
        if key == "HIERARCH" and "=" in card:
            k, rest = card[8:].split("=", 1)
            v, c = _parse_value(rest)
            hdr._cards[k.strip().upper()] = (v, c)
            continue
        if card[8:10] != "= ":
            continue
        v, c = _parse_value(card[10:])
        hdr._cards[key] = (v, c)
    return hdr, False


# --------------------------------------------------------------------------
# HDUs
# --------------------------------------------------------------------------
class ImageHDU:
    is_primary = False

    def __init__(self, data=None, header=None, name=None):
        self.header = Header(header) if header is not None else Header()
        self.data = None if data is None else np.asarray(data)
        if name is not None:
            self.header["EXTNAME"] = name

    @property
    def name(self):
        return self.header.get("EXTNAME", "PRIMARY" if self.is_primary else "")

    def copy(self):
        return type(self)(None if self.data is None else self.data.copy(), self.header.copy())


class PrimaryHDU(ImageHDU):
    is_primary = True


class HDUList(list):
    def __getitem__(self, key):
        if isinstance(key, str):
            k = key.strip().upper()
            for h in self:
                if str(h.name).strip().upper() == k:
                    return h
            raise KeyError(f"Extension {key!r} not found.")
        return list.__getitem__(self, key)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def writeto(self, path, overwrite=False):
        writeto(path, self, overwrite=overwrite)


def _data_nbytes(hdr: Header) -> int:
    naxis = int(hdr.get("NAXIS", 0))
    if naxis == 0:
        return 0
    n = 1
    for i in range(1, naxis + 1):
        n *= int(hdr[f"NAXIS{i}"])
    n *= abs(int(hdr["BITPIX"])) // 8
    n *= int(hdr.get("GCOUNT", 1))
    n += int(hdr.get("PCOUNT", 0))
    return n


def open(path, mode="readonly", **_):  # noqa: A001 - mirrors astropy.io.fits.open
    """Read every HDU of a FITS file into memory. Returns an `HDUList`."""
    with builtins_open(os.fspath(path), "rb") as f:
        buf = f.read()
    hdus = HDUList()
    pos = 0
    first = True
    while pos < len(buf):
        raw = b""
        done = False
        hdr = Header()
        while not done:
            if pos >= len(buf):
                if first:
                    raise OSError(f"{path}: truncated FITS header")
                return hdus
            blk = buf[pos:pos + BLOCK]
            pos += BLOCK
            part, done = _parse_header_block(blk)
            hdr._cards.update(part._cards)
            hdr._commentary.extend(part._commentary)
        if first and "SIMPLE" not in hdr:
            raise OSError(f"{path}: not a FITS file (no SIMPLE card)")
        nbytes = _data_nbytes(hdr)
        data = None
        xt = str(hdr.get("XTENSION", "IMAGE")).strip().upper()
        if nbytes and (first or xt == "IMAGE"):
            naxis = int(hdr["NAXIS"])
            shape = tuple(int(hdr[f"NAXIS{i}"]) for i in range(naxis, 0, -1))
            dt = np.dtype(_BITPIX_DTYPE[int(hdr["BITPIX"])])
            arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape)), offset=pos).reshape(shape)
            bscale = hdr.get("BSCALE", 1)
            bzero = hdr.get("BZERO", 0)
            scaled = float(bscale) != 1.0 or float(bzero) != 0.0
            if scaled and dt.kind == "i" and float(bscale) == 1.0 \
                    and float(bzero) == float(2 ** (8 * dt.itemsize - 1)):
                # unsigned-integer convention (BZERO = 2**(bits-1))
                arr = (arr.astype(np.int64) + int(bzero)).astype(f"u{dt.itemsize}")
            elif scaled:
                # same promotion rule as astropy: <=16-bit ints -> float32, everything else float64
                out_dt = np.float32 if (dt.kind in "iu" and dt.itemsize <= 2) or dt.itemsize == 4 and dt.kind == "f" \
                    else np.float64
                arr = arr.astype(out_dt) * out_dt(bscale) + out_dt(bzero)
            else:
                arr = arr.astype(dt.newbyteorder("="))
            data = arr
        elif nbytes and xt == "BINTABLE" and "ZIMAGE" in hdr:
            raise NotImplementedError(
                f"{path}: tile-compressed image HDU (ZIMAGE) is not supported by fits_lite; "
                "decompress it first (e.g. funpack) or install astropy")
        pos += ((nbytes + BLOCK - 1) // BLOCK) * BLOCK
        hdu = PrimaryHDU(data, hdr) if first else ImageHDU(data, hdr)
        hdus.append(hdu)
        first = False
    return hdus


import builtins  # noqa: E402

builtins_open = builtins.open


def _structural_cards(hdu, primary: bool):
    data = hdu.data
    cards = []
    if primary:
        cards.append(("SIMPLE", True, "conforms to FITS standard"))
    else:
        cards.append(("XTENSION", "IMAGE", "Image extension"))
    if data is None:
        cards += [("BITPIX", 8, ""), ("NAXIS", 0, "")]
    else:
        key = data.dtype.newbyteorder("=").str[1:]
        if key not in _DTYPE_BITPIX:
            raise TypeError(f"unsupported FITS dtype {data.dtype}")
        cards += [("BITPIX", _DTYPE_BITPIX[key], "array data type"), ("NAXIS", data.ndim, "")]
        for i, n in enumerate(reversed(data.shape), 1):
            cards.append((f"NAXIS{i}", int(n), ""))
    if primary:
        cards.append(("EXTEND", True, ""))
    else:
        cards += [("PCOUNT", 0, ""), ("GCOUNT", 1, "")]
    return cards


_STRUCTURAL = {"SIMPLE", "XTENSION", "BITPIX", "NAXIS", "EXTEND", "PCOUNT", "GCOUNT", "BSCALE", "BZERO"}


def writeto(path, hdus, overwrite=False):
    """Write an `HDUList` (or a single HDU) as an uncompressed FITS file."""
    path = os.fspath(path)
    if os.path.exists(path) and not overwrite:
        raise OSError(f"File {path!r} already exists.")
    if isinstance(hdus, ImageHDU):
        hdus = [hdus]
    out = bytearray()
    for idx, hdu in enumerate(hdus):
        primary = idx == 0
        cards = [_format_card(k, v, c) for k, v, c in _structural_cards(hdu, primary)]
        for k, (v, c) in hdu.header._cards.items():
            if k in _STRUCTURAL or (k.startswith("NAXIS") and k[5:].isdigit()):
                continue
            cards.append(_format_card(k, v, c))
        for k, txt in hdu.header._commentary:
            cards.append(f"{k:<8}{txt}"[:CARD].ljust(CARD))
        cards.append("END".ljust(CARD))
        raw = "".join(cards).encode("ascii", errors="replace")
        raw += b" " * ((-len(raw)) % BLOCK)
        out += raw
        if hdu.data is not None:
            d = np.ascontiguousarray(hdu.data)
            d = d.astype(d.dtype.newbyteorder(">"))
            b = d.tobytes()
            out += b + b"\0" * ((-len(b)) % BLOCK)
    tmp = path + ".tmp~"
    with builtins_open(tmp, "wb") as f:
        f.write(out)
    os.replace(tmp, path)


def getheader(path, ext=0):
    return open(path)[ext].header


def getdata(path, ext=0):
    return open(path)[ext].data
