"""Generate tests/golden/pearson_golden.npz from the REFERENCE's own numba `c_correlate`.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_pearson_golden.py
The reference module `hdrshift/c_correlate.py` depends on numba+numpy only, so it is loaded by file
path without importing the (astropy-dependent) package around it.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np

REF = "/root/reference/euispice_coreg/hdrshift/c_correlate.py"


def main():
    spec = importlib.util.spec_from_file_location("ref_c_correlate", REF)
    mod = importlib.util.module_from_spec(spec)
    warnings.simplefilter("ignore")
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(20221017)
    cases = {}
    sizes = [2, 3, 17, 1000, 65537, 1 << 20]
    for k, n in enumerate(sizes):
        a = rng.lognormal(6.0, 0.6, n)
        noise = rng.normal(0.0, 40.0, n)
        b = (0.8 * a + noise + 100.0 * np.sin(np.arange(n) * 0.01)).astype(np.float32).astype(np.float64)
        r = mod.c_correlate(a, b, [0])
        cases[f"a{k}"] = a
        cases[f"b{k}"] = b
        cases[f"r{k}"] = np.asarray(r, dtype=np.float64)
    # keep the fixture small: store seeds for the big ones instead of data
    out = {}
    for k, n in enumerate(sizes):
        if n <= 1000:
            out[f"a{k}"] = cases[f"a{k}"]
            out[f"b{k}"] = cases[f"b{k}"]
        out[f"r{k}"] = cases[f"r{k}"]
    out["sizes"] = np.asarray(sizes)
    out["seed"] = np.asarray([20221017])
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pearson_golden.npz")
    np.savez(dst, **out)
    print("wrote", dst, {k: v.shape for k, v in out.items()})


def regenerate_inputs():
    """Inputs of every case, regenerated from the seed (used by the test for the large cases)."""
    rng = np.random.default_rng(20221017)
    sizes = [2, 3, 17, 1000, 65537, 1 << 20]
    res = []
    for n in sizes:
        a = rng.lognormal(6.0, 0.6, n)
        noise = rng.normal(0.0, 40.0, n)
        b = (0.8 * a + noise + 100.0 * np.sin(np.arange(n) * 0.01)).astype(np.float32).astype(np.float64)
        res.append((a, b))
    return res


if __name__ == "__main__":
    sys.exit(main())
