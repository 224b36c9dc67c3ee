"""ISO-8601 FITS date strings -> seconds / days. Replaces the `astropy.time.Time` differences
on the pointing-search path (`utils/rectify.py:416-418`, `synras/map_builder.py:96-105, 218-229`).

UTC leap seconds are ignored (none were inserted after 2016-12-31; Solar Orbiter launched in 2020).
"""
from __future__ import annotations

import datetime as _dt
import re

import numpy as np

_EPOCH = _dt.datetime(2000, 1, 1, 12, 0, 0)
_ISO = re.compile(r"^\s*(\d{4})-(\d{2})-(\d{2})(?:[T ](\d{2}):(\d{2})(?::(\d{2})(\.\d*)?)?)?\s*Z?\s*$")


def to_seconds(date) -> float:
    """Seconds since J2000.0 (UTC, no leap seconds) of an ISO `YYYY-MM-DDThh:mm:ss.sss` string."""
    if isinstance(date, (int, float, np.floating, np.integer)):
        return float(date)
    if isinstance(date, _dt.datetime):
        d = date.replace(tzinfo=None)
        return (d - _EPOCH).total_seconds()
    if hasattr(date, "isot"):  # astropy Time
        date = date.isot
    m = _ISO.match(str(date))
    if not m:
        raise ValueError(f"unparsable FITS date {date!r}")
    y, mo, d, hh, mm, ss, frac = m.groups()
    base = _dt.datetime(int(y), int(mo), int(d), int(hh or 0), int(mm or 0), int(ss or 0))
    whole = (base - _EPOCH).days * 86400.0 + (base - _EPOCH).seconds
    return whole + (float(frac) if frac and frac != "." else 0.0)


def diff_days(a, b) -> float:
    """`(Time(a) - Time(b)).value` (a TimeDelta in days)."""
    return (to_seconds(a) - to_seconds(b)) / 86400.0


def diff_seconds(a, b) -> float:
    return to_seconds(a) - to_seconds(b)


def from_seconds(t: float) -> str:
    """Inverse of `to_seconds`, millisecond resolution (FITS `isot`)."""
    whole = int(np.floor(t))
    ms = int(round((t - whole) * 1000.0))
    if ms == 1000:
        whole, ms = whole + 1, 0
    d = _EPOCH + _dt.timedelta(seconds=whole)
    return d.strftime("%Y-%m-%dT%H:%M:%S") + ".%03d" % ms
