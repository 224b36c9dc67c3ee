"""Homogeneous 2-D pixel transforms of the pixel-shift search (`utils/matrix_transform.py:4-106` of the reference),
host numpy: they produce a handful of coordinate planes per search (one per rotation lag); the sampling and the lag
loop run on the device."""
import numpy as np


class MatrixTransform:
    @staticmethod
    def displacement_matrix(ndim=2, dx=0, dy=0):
        if ndim != 2:
            raise NotImplementedError
        return np.array([[1, 0, dx], [0, 1, dy], [0, 0, 1]])

    @staticmethod
    def rotation_matrix(ndim=2, theta=0, units='radian'):
        if ndim != 2:
            raise NotImplementedError
        if units == 'degree':
            theta = np.radians(theta)
        return np.array([[np.cos(theta), -np.sin(theta), 0], [np.sin(theta), np.cos(theta), 0], [0, 0, 1]])

    @staticmethod
    def linear_transform(*args, matrix):
        if len(args) != 2:
            raise NotImplementedError
        assert matrix.ndim == 2
        xx, yy = args
        assert xx.shape == yy.shape
        xyz = np.stack((xx.ravel(), yy.ravel(), np.ones(xx.shape).ravel()))
        nx, ny, _ = np.matmul(matrix, xyz)
        return nx.reshape(xx.shape), ny.reshape(yy.shape)

    @staticmethod
    def to_polar_coordinates(*args, direction='forward'):
        if len(args) == 2:
            xx, yy = args
        elif len(args) == 4:
            xx, yy, xc, yc = args
        else:
            raise NotImplementedError
        assert xx.shape == yy.shape
        if direction == 'forward':
            if len(args) == 2:
                xc = xx[round(xx.shape[0] / 2), round(xx.shape[1] / 2)]
                yc = yy[round(xx.shape[0] / 2), round(xx.shape[1] / 2)]
            nr = np.sqrt(np.power(xx - xc, 2) + np.power(yy - yc, 2))
            ntheta = np.arctan2(yy - yc, xx - xc)
            ntheta[np.isnan(ntheta)] = 0
            return nr, ntheta
        if direction == 'backward':
            if len(args) == 2:
                xc = 0
                yc = 0
            # here xx = r and yy = theta
            return np.multiply(xx, np.cos(yy)) + xc, np.multiply(xx, np.sin(yy)) + yc
        raise ValueError("direction must be 'forward' or 'backward'")

    @staticmethod
    def polar_transform(*args, theta=0, units='radian'):
        """Rotation by theta about (xc, yc); with two arguments the centre is the pixel (round(H / 2), round(W / 2))
        of the coordinate planes (`utils/matrix_transform.py:77-106`)."""
        if units == 'degree':
            theta = np.radians(theta)
        if len(args) == 2:
            xx, yy = args
            assert xx.shape == yy.shape
            xc = xx[round(xx.shape[0] / 2), round(xx.shape[1] / 2)]
            yc = yy[round(xx.shape[0] / 2), round(xx.shape[1] / 2)]
        elif len(args) == 4:
            xx, yy, xc, yc = args
            assert xx.shape == yy.shape
        else:
            raise NotImplementedError
        nr, ntheta = MatrixTransform.to_polar_coordinates(xx, yy, xc, yc, direction='forward')
        ntheta = ntheta + theta
        return MatrixTransform.to_polar_coordinates(nr, ntheta, xc, yc, direction='backward')
