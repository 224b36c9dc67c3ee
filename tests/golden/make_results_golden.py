"""Extract the reference's printed golden correlation cubes into tests/golden/results_golden.npz.

Run in the build container only (reads /root/reference):
    python tests/golden/make_results_golden.py
Sources: cube A = the `corr` fixture of hdrshift/test/test_AlignmentResults.py:33-126 (HPC, 11x6, real EUI
data, expected Gaussian-fit peak (9.33682107, 1.42187891) +-1e-2 at :172-173); cube B = the `corr` fixture
of plot/test/test_plot.py:31-57 (Carrington, 5x5). The fixtures are number literals; they are evaluated
with numpy only (the test modules themselves import astropy and cannot be imported here).
"""
import os
import re

import numpy as np


def _literal_after(src, marker):
    i = src.index(marker)
    j = src.index("np.array(", i) + len("np.array(")
    depth, k = 1, j
    while depth:
        c = src[k]
        depth += (c == "(") - (c == ")")
        k += 1
    return np.array(eval(src[j:k - 1].strip(), {"__builtins__": {}}), dtype=np.float64)


def main():
    a = open("/root/reference/euispice_coreg/hdrshift/test/test_AlignmentResults.py").read()
    b = open("/root/reference/euispice_coreg/plot/test/test_plot.py").read()
    cube_a = _literal_after(a, "def corr():")
    cube_b = _literal_after(b, "def corr():")
    assert cube_a.shape == (11, 6, 1, 1, 1, 1) and cube_b.shape == (5, 5, 1, 1, 1, 1)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "results_golden.npz")
    np.savez(dst, cube_a=cube_a, a_lag_crval1=np.arange(15, 26, 1), a_lag_crval2=np.arange(5, 11, 1),
             a_lag_cdelt2=np.array([0]), a_lag_crota=np.array([0.75]), a_peak=np.array([9.33682107, 1.42187891]),
             cube_b=cube_b, b_lag_crval1=np.arange(20, 30, 2), b_lag_crval2=np.arange(5, 15, 2),
             b_lag_cdelt2=np.array([0]), b_lag_crota=np.array([0.75]))
    print("wrote", dst, cube_a.max(), np.unravel_index(cube_a.argmax(), cube_a.shape)[:2],
          cube_b.max(), np.unravel_index(cube_b.argmax(), cube_b.shape)[:2])


if __name__ == "__main__":
    main()
