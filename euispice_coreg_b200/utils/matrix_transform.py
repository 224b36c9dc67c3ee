"""Pixel-plane transforms of the pixel-shift search (the role of the reference's `utils/matrix_transform.py`).

The pixel-shift search (`pxlshift/alignment_pixels.py:35-84`) needs one thing from this module: the small image's
pixel grid rotated by each rotation lag about the grid's centre pixel, which the reference obtains by going to polar
coordinates, adding the angle and coming back (`utils/matrix_transform.py:77-106`). `rotate_about` below does exactly
that arithmetic (radius by sqrt of the summed squares, angle by arctan2, NaN angles -> 0, r cos / r sin back), so the
rotated planes -- and the images resampled on them, pinned by `tests/golden/pxlshift_golden.npz` -- carry the
reference's bits. The planes are a handful per search (one per rotation lag); the resampling and the lag loop run on
the device. `MatrixTransform` keeps the reference's class and method names for callers written against it.
"""
import numpy as np


def centre_pixel(xx, yy):
    """Coordinates of the pixel the reference rotates about: index (round(H / 2), round(W / 2)) of the planes."""
    j, i = round(xx.shape[0] / 2), round(xx.shape[1] / 2)
    return xx[j, i], yy[j, i]


def _angle(theta, units):
    if units == "degree":
        return np.radians(theta)
    if units != "radian":
        raise ValueError("units must be 'radian' or 'degree'")
    return theta


def rotate_about(xx, yy, theta, centre=None, units="radian"):
    """(xx, yy) rotated by `theta` about `centre` (default: the planes' centre pixel), through polar coordinates."""
    if xx.shape != yy.shape:
        raise ValueError("coordinate planes must have the same shape")
    xc, yc = centre_pixel(xx, yy) if centre is None else centre
    dx, dy = xx - xc, yy - yc
    radius = np.sqrt(np.power(dx, 2) + np.power(dy, 2))
    phase = np.arctan2(dy, dx)
    phase[np.isnan(phase)] = 0
    phase = phase + _angle(theta, units)
    return radius * np.cos(phase) + xc, radius * np.sin(phase) + yc


def homogeneous(matrix, xx, yy):
    """Apply a 3 x 3 homogeneous matrix to the planes (translation / rotation helpers below build such matrices)."""
    m = np.asarray(matrix)
    if m.shape != (3, 3) or xx.shape != yy.shape:
        raise ValueError("a 3 x 3 matrix and two planes of one shape expected")
    pts = np.vstack([xx.ravel(), yy.ravel(), np.ones(xx.size)])
    out = m @ pts
    return out[0].reshape(xx.shape), out[1].reshape(yy.shape)


class MatrixTransform:
    """The reference's entry points (`utils/matrix_transform.py:4-106`), as thin wrappers."""

    @staticmethod
    def displacement_matrix(ndim=2, dx=0, dy=0):
        if ndim != 2:
            raise NotImplementedError
        m = np.eye(3, dtype=np.result_type(dx, dy, 1))
        m[0, 2], m[1, 2] = dx, dy
        return m

    @staticmethod
    def rotation_matrix(ndim=2, theta=0, units="radian"):
        if ndim != 2:
            raise NotImplementedError
        t = _angle(theta, units)
        c, s = np.cos(t), np.sin(t)
        return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])

    @staticmethod
    def linear_transform(*planes, matrix):
        if len(planes) != 2:
            raise NotImplementedError
        return homogeneous(matrix, *planes)

    @staticmethod
    def to_polar_coordinates(*args, direction="forward"):
        if len(args) not in (2, 4):
            raise NotImplementedError
        a, b = args[:2]
        if direction == "forward":
            xc, yc = args[2:] if len(args) == 4 else centre_pixel(a, b)
            radius = np.sqrt(np.power(a - xc, 2) + np.power(b - yc, 2))
            phase = np.arctan2(b - yc, a - xc)
            phase[np.isnan(phase)] = 0
            return radius, phase
        if direction == "backward":       # a = radius, b = angle
            xc, yc = args[2:] if len(args) == 4 else (0, 0)
            return a * np.cos(b) + xc, a * np.sin(b) + yc
        raise ValueError("direction must be 'forward' or 'backward'")

    @staticmethod
    def polar_transform(*args, theta=0, units="radian"):
        if len(args) == 2:
            return rotate_about(args[0], args[1], theta, None, units)
        if len(args) == 4:
            return rotate_about(args[0], args[1], theta, (args[2], args[3]), units)
        raise NotImplementedError
