"""Spline resampling as the reference performs it (TEST INFRASTRUCTURE).

`interpol2d` follows `utils/Util.py:82-104` (and its duplicate `utils/rectify.py:22-56`): it hands the
coordinates to `scipy.ndimage.map_coordinates(order=k, mode='constant', cval=fill, prefilter=False)`.
scipy is the reference's own third-party dependency for this step (scipy 1.17.1 in `poetry.lock:2183-2184`)
and IS installed in this image, so the oracle calls the real function.

`map_coordinates_restated` is a numpy restatement of the same algorithm (scipy's NI_GeometricTransform +
spline weights for orders 0..3, 'constant' mode without prefilter). tests/test_oracle_resample.py checks
it bit-for-bit against scipy; the CUDA kernels follow this restatement operation by operation.
"""
from __future__ import annotations

import numpy as np
from scipy.ndimage import map_coordinates


def interpol2d(image, x, y, fill, order, dst=None):
    """`AlignCommonUtil.interpol2d` (`utils/Util.py:82-104`). The `x == np.nan` guard of the reference
    is always False, so NaN coordinates reach scipy and produce `fill`."""
    coords = np.stack((np.asarray(y).ravel(), np.asarray(x).ravel()), axis=0)
    ret = dst is None
    if dst is None:
        dst = np.empty(np.shape(x), dtype=image.dtype)
    flat = dst.reshape(-1)  # same memory as dst for a contiguous array, like dst.ravel() in the reference
    map_coordinates(image, coords, order=order, mode="constant", cval=fill, output=flat, prefilter=False)
    if ret:
        return dst
    return None


def spline_weights(t, order):
    """Per-axis start index and weights; `t` float64 array. Returns (start int64, weights [order+1, ...])."""
    t = np.asarray(t, dtype=np.float64)
    if order & 1:
        base = np.floor(t)
    else:
        base = np.floor(t + 0.5)
    d = t - base
    start = base.astype(np.int64) - order // 2
    if order == 0:
        w = [np.ones_like(d)]
    elif order == 1:
        w0 = 1.0 - d
        w = [w0, 1.0 - w0]
    elif order == 2:
        w1 = 0.75 - d * d
        u = 0.5 - d
        w0 = 0.5 * u * u
        w2 = 1.0 - w0 - w1
        w = [w0, w1, w2]
    elif order == 3:
        z = 1.0 - d
        w1 = (d * d * (d - 2.0) * 3.0 + 4.0) / 6.0
        w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
        w0 = z * z * z / 6.0
        w3 = 1.0 - w0 - w1 - w2
        w = [w0, w1, w2, w3]
    else:
        raise NotImplementedError("orders 0..3")
    return start, w


def _mirror(i, n):
    """Reflect an out-of-range tap index about the edge pixel centre (scipy 'constant' mode keeps the
    spline support by mirroring: i<0 -> -i, i>n-1 -> 2(n-1)-i)."""
    if n == 1:
        return np.zeros_like(i)
    i = np.where(i < 0, -i, i)
    i = np.where(i > n - 1, 2 * (n - 1) - i, i)
    # one reflection is enough for |offset| <= order//2+1 < n
    return np.clip(i, 0, n - 1)


def map_coordinates_restated(image, y, x, order, cval, out_dtype=np.float64):
    """Numpy restatement of `map_coordinates(image, [y, x], order, mode='constant', cval, prefilter=False)`."""
    image = np.asarray(image)
    img = image.astype(np.float64, copy=False)
    H, W = img.shape
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    shape = x.shape
    y = y.ravel()
    x = x.ravel()
    with np.errstate(invalid="ignore"):
        inside = (y >= 0.0) & (y <= H - 1) & (x >= 0.0) & (x <= W - 1)
    ys = np.where(inside, y, 0.0)
    xs = np.where(inside, x, 0.0)
    sy, wy = spline_weights(ys, order)
    sx, wx = spline_weights(xs, order)
    acc = np.zeros(ys.shape, dtype=np.float64)
    for a in range(order + 1):
        iy = _mirror(sy + a, H)
        for b in range(order + 1):
            ix = _mirror(sx + b, W)
            acc = acc + (img[iy, ix] * wy[a]) * wx[b]
    out = np.where(inside, acc, np.float64(cval))
    return out.astype(out_dtype).reshape(shape)
