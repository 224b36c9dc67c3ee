"""Thin stand-in for the reference's matplotlib figures (`plot/plot.py:56-178, 608-924`).

Visualisation is outside the device hot path and matplotlib is not in the
image, so only the correlation map is provided, imported lazily; the API names are kept so user scripts
that call `results.plot_correlation(...)` keep working where matplotlib exists.
"""
from __future__ import annotations

import numpy as np


def _plt():
    try:
        import matplotlib
        matplotlib.use("Agg", force=False)
        from matplotlib import pyplot as plt
        return plt
    except Exception as exc:  # pragma: no cover
        raise ImportError("matplotlib is required for the plot_* helpers") from exc


class PlotFunctions:

    @staticmethod
    def plot_correlation(corr, lag_crval1, lag_crval2, lag_cdelt1=None, lag_cdelt2=None, lag_crota=None,
                         show=False, path_save_figure=None, fig=None, ax=None, shift=None, unit_to_plot="arcsec",
                         lag_dx_label="CRVAL1", lag_dy_label="CRVAL2", **_):
        plt = _plt()
        mi = np.unravel_index(np.nanargmax(corr), corr.shape)
        plane = corr[:, :, mi[2], mi[3], mi[4], 0]
        if fig is None:
            fig = plt.figure(figsize=(6, 5))
        if ax is None:
            ax = fig.add_subplot()
        dx = (lag_crval1[1] - lag_crval1[0]) if len(lag_crval1) > 1 else 1.0
        dy = (lag_crval2[1] - lag_crval2[0]) if len(lag_crval2) > 1 else 1.0
        im = ax.imshow(plane.T, origin="lower", interpolation="none", aspect="auto",
                       extent=(lag_crval1[0] - 0.5 * dx, lag_crval1[-1] + 0.5 * dx,
                               lag_crval2[0] - 0.5 * dy, lag_crval2[-1] + 0.5 * dy))
        if shift is not None:
            ax.plot(shift[0], shift[1], "r+")
        ax.set_xlabel(lag_dx_label)
        ax.set_ylabel(lag_dy_label)
        fig.colorbar(im, ax=ax, label="correlation")
        if path_save_figure is not None:
            fig.savefig(path_save_figure)
        if show:
            fig.show()
        return fig, ax

    @staticmethod
    def plot_co_alignment(*args, **kwargs):
        _plt()
        raise NotImplementedError("plot_co_alignment is outside the scope of the device hot path")
