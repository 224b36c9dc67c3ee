"""Oracle cubes at BASELINE.json's full sizes, generated ONCE on CPU and committed (TEST INFRASTRUCTURE).

    python tests/golden/make_fullsize_golden.py [config1] [config2] [config4] [--workers 8]

* config1_full_cube.npz   every one of the 60 x 60 = 3600 helioprojective lags of configs[0]
                          (2048^2 vs 3072^2 synthetic pair of `bench.ensure_config1`), `oracle.hpc.HpcSearch.step`
* config2_sample.npz      ~290 of the 120 x 120 Carrington-grid lags of configs[1] (2048^2 grid, lon 200-300, lat +-20):
                          a 15 x 15 sub-lattice + 64 seeded random lags, `oracle.carrington.CarringtonSearch.step`
* config4_sample.npz      288 of the 20 x 20 x 16 x 16 x 10 = 1 024 000 lags of configs[3] (intended CDELT semantics):
                          the arg-max neighbourhood + seeded random lags

The oracle needs 1.5 - 4 s per lag per core, which is why the `-m gpu` tests compare against these files instead of
running it. Each file stores the flat C-order lag indices, the oracle's r, the lag arrays and the SHA-256 of the two
synthetic FITS payloads, so a test can tell a changed synthetic scene from a changed kernel. Partial results are
checkpointed next to the output (`*.part.npz`) and picked up again.
"""
from __future__ import annotations

import argparse
import hashlib
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIG2_LAGS = dict(lag_crval1=np.arange(-60, 60, 1.0), lag_crval2=np.arange(-60, 60, 1.0), lag_cdelt1=[0.0],
                    lag_cdelt2=[0.0], lag_crota=[0.0])
CONFIG2_GRID = dict(lonlims=(200.0, 300.0), latlims=(-20.0, 20.0), shape=(2048, 2048))
CONFIG4_LAGS = dict(lag_crval1=np.arange(14, 34, 1.0), lag_crval2=np.arange(-4, 16, 1.0),
                    lag_cdelt1=(np.arange(16) - 8) * 0.001, lag_cdelt2=(np.arange(16) - 8) * 0.001,
                    lag_crota=(np.arange(10) - 5) * 0.1)

_G = {}


def payload_sha(pl, ps):
    from conftest import load_pair
    dl, _, ds, _ = load_pair(pl, ps)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(dl).tobytes())
    h.update(np.ascontiguousarray(ds).tobytes())
    return h.hexdigest()


def _one(i):
    s, flat = _G["search"], _G["flat"]
    return i, float(s.step(*(f[i] for f in flat)))


def run(search, flat, sel, out, workers, extra):
    part = out.replace(".npz", ".part.npz")
    done = {}
    if os.path.exists(part):
        z = np.load(part)
        done = dict(zip(z["index"].tolist(), z["r"].tolist()))
    todo = [int(i) for i in sel if int(i) not in done]
    _G["search"], _G["flat"] = search, flat
    t0 = time.time()
    if todo:
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            for k, (i, r) in enumerate(pool.imap_unordered(_one, todo, chunksize=1)):
                done[i] = r
                if (k + 1) % 64 == 0 or k + 1 == len(todo):
                    idx = np.array(sorted(done), dtype=np.int64)
                    np.savez(part, index=idx, r=np.array([done[j] for j in idx]))
                    print(f"{os.path.basename(out)}: {k + 1}/{len(todo)} lags, {time.time() - t0:.0f} s", flush=True)
    sel = np.asarray(sel, dtype=np.int64)
    np.savez_compressed(out, index=sel, r=np.array([done[int(j)] for j in sel]), **extra)
    if os.path.exists(part):
        os.remove(part)


def main():
    import bench
    from conftest import load_pair
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["config1", "config2", "config4"])
    ap.add_argument("--workers", type=int, default=len(os.sched_getaffinity(0)))
    args = ap.parse_args()
    pl, ps = bench.ensure_config1()
    sha = payload_sha(pl, ps)
    dl, hl, ds, hs = load_pair(pl, ps)

    if "config1" in args.which:
        from oracle.hpc import HpcSearch
        s = HpcSearch(dl, hl, ds, hs, **bench.LAGS)
        s.world()                      # lag-independent; computed once before the fork
        flat = s.flat_lags()
        run(s, flat, np.arange(flat[0].size), os.path.join(HERE, "config1_full_cube.npz"), args.workers,
            dict(shape=np.array(s.shape), payload_sha256=sha,
                 **{k: np.asarray(v, dtype=np.float64) for k, v in bench.LAGS.items()}))

    if "config2" in args.which:
        from oracle.carrington import CarringtonSearch
        s = CarringtonSearch(dl, hl, ds, hs, **CONFIG2_LAGS, **CONFIG2_GRID)
        r = s.refs
        g = np.meshgrid(r.lag_crval1, r.lag_crval2, r.lag_cdelt1, r.lag_cdelt2, r.lag_crota, indexing="ij")
        flat = [a.ravel() for a in g]
        lat = (np.arange(15) * 8 + 4)
        sub = (lat[:, None] * 120 + lat[None, :]).ravel()
        rng = np.random.default_rng(22)
        sel = np.unique(np.concatenate([sub, rng.integers(0, 14400, 64), [84 * 120 + 66, 0, 14399]]))
        run(s, flat, sel, os.path.join(HERE, "config2_sample.npz"), args.workers,
            dict(shape=np.array(s.cube_shape), payload_sha256=sha,
                 **{k: np.asarray(v, dtype=np.float64) for k, v in CONFIG2_LAGS.items()}))

    if "config4" in args.which:
        from oracle.hpc import HpcSearch
        s = HpcSearch(dl, hl, ds, hs, cdelt_mode="intended", **CONFIG4_LAGS)
        s.world()
        flat = s.flat_lags()
        n = flat[0].size
        shape = s.shape[:5]
        peak = np.ravel_multi_index((10, 10, 8, 8, 5), shape)      # (24", 6", 0, 0, 0)
        near = [np.ravel_multi_index((10 + a, 10 + b, 8 + c, 8 + d, 5 + e), shape)
                for a in (-1, 0, 1) for b in (-1, 0, 1) for c in (-1, 1) for d in (-1, 1) for e in (-1, 1)][:31]
        rng = np.random.default_rng(44)
        sel = np.unique(np.concatenate([[peak, 0, n - 1], near, rng.integers(0, n, 256)]))
        run(s, flat, sel, os.path.join(HERE, "config4_sample.npz"), args.workers,
            dict(shape=np.array(s.shape), payload_sha256=sha,
                 **{k: np.asarray(v, dtype=np.float64) for k, v in CONFIG4_LAGS.items()}))


if __name__ == "__main__":
    main()
