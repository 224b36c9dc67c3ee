"""BASELINE.json configs[1], [3], [4] at full image size on one B200 (GPU box only): wall / device times through the
public API and, for a bounded sample of lags, the difference to the CPU oracle.

    python tools/config_runs.py [--out gpurun_out/config_runs.json] [--frames 8] [--skip carrington,grid5d,sequence]

configs[1]  Carrington grid 2048^2 (lon 200-300 deg, lat +-20 deg), 120 x 120 CRVAL lags
configs[3]  5-D lag grid with the intended CDELT semantics (default 24 x 24 x 6 x 5 x 5 = 86 400 lags; the 1e6-lag
            grid of the baseline is the same code path in more launches)
configs[4]  sequence of jittered HRIEUV-like frames against one FSI-like image (default 8 frames; 60 x 60 lags each)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bench
    from conftest import load_pair
    from euispice_coreg_b200.hdrshift import Alignment, SequenceAlignment
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/config_runs.json")
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--skip", default="")
    ap.add_argument("--grid5d-full", action="store_true",
                    help="configs[3] at BASELINE size: 20 x 20 x 10 x 16 x 16 = 1.024e6 lags (crota step 0.1 deg, "
                         "cdelt step 0.001 arcsec; SURVEY 8d)")
    args = ap.parse_args()
    skip = set(args.skip.split(","))
    pl, ps = bench.ensure_config1()
    dl, hl, ds, hs = load_pair(pl, ps)
    cores = len(os.sched_getaffinity(0))
    results = []

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, time.perf_counter() - t0

    if "c2f" not in skip:
        dense = Alignment(pl, ps, parallelism=True, **bench.LAGS)
        rd = dense.align_using_helioprojective()
        run = lambda: Alignment(pl, ps, parallelism=True, lag_search="coarse_to_fine", **bench.LAGS)\
            .align_using_helioprojective()  # noqa: E731
        run()
        rc, dt = timed(run)
        a = Alignment(pl, ps, parallelism=True, lag_search="coarse_to_fine", **bench.LAGS)
        rc = a.align_using_helioprojective()
        ev = ~np.isnan(rc.corr)
        _, dtd = timed(lambda: Alignment(pl, ps, parallelism=True, **bench.LAGS).align_using_helioprojective())
        results.append({"config": "configs[0] with lag_search='coarse_to_fine' (opt-in, SURVEY 8f-3)",
                        "lags_evaluated": int(a.lags_evaluated), "lags_total": int(rd.corr.size),
                        "wall_s_public_api": dt, "wall_s_public_api_dense": dtd,
                        "same_argmax_as_dense": bool(rc.max_index == rd.max_index),
                        "same_fitted_shift_as_dense": bool(rc.shift_arcsec == rd.shift_arcsec),
                        "evaluated_entries_bit_identical": bool(np.array_equal(rc.corr[ev], rd.corr[ev]))})
        print(json.dumps(results[-1]), flush=True)

    if "carrington" not in skip:
        from oracle.carrington import CarringtonSearch
        lags = dict(lag_crval1=np.arange(-60, 60, 1.0), lag_crval2=np.arange(-60, 60, 1.0), lag_cdelt1=[0],
                    lag_cdelt2=[0], lag_crota=[0])
        grid = dict(lonlims=(200.0, 300.0), latlims=(-20.0, 20.0), shape=(2048, 2048))
        run = lambda: Alignment(pl, ps, parallelism=True, **lags).align_using_carrington(  # noqa: E731
            method="correlation", return_type="corr", **grid)
        run()
        cube, dt = timed(run)
        a = Alignment(pl, ps, parallelism=True, **lags)
        cube = a.align_using_carrington(method="correlation", return_type="corr", **grid)
        nv = a.nvalid.ravel()
        am = np.unravel_index(np.nanargmax(cube), cube.shape)[:2]
        sel = [(84, 66), (0, 0), (119, 119), (60, 60), (30, 100)]
        s = CarringtonSearch(dl, hl, ds, hs, **lags, **grid)
        t0 = time.perf_counter()
        ref = [s.step(lags["lag_crval1"][i], lags["lag_crval2"][j], 0.0, 0.0, 0.0) for i, j in sel]
        t_or = (time.perf_counter() - t0) / len(sel)
        err = [abs(float(cube[i, j, 0, 0, 0, 0]) - r) for (i, j), r in zip(sel, ref)]
        results.append({"config": "configs[1] Carrington 2048^2 grid, 120x120 lags", "lags": 14400,
                        "wall_s_public_api": dt, "lag_evals_per_s": 14400 / dt,
                        "nominal_pixel_samples_per_s": 14400 * 2048 * 2048 / dt,
                        "effective_pixel_samples_per_s": float(nv.sum()) / dt,
                        "effective_fraction_of_grid": float(nv.mean()) / (2048 * 2048),
                        "argmax_lag_arcsec": [float(lags["lag_crval1"][am[0]]), float(lags["lag_crval2"][am[1]])],
                        "oracle_sample_lags": [[float(lags["lag_crval1"][i]), float(lags["lag_crval2"][j])] for i, j in sel],
                        "oracle_abs_err": err, "oracle_s_per_lag_one_core": t_or})
        print(json.dumps(results[-1]), flush=True)

    if "grid5d" not in skip:
        from oracle.hpc import HpcSearch, cube_multiprocess
        lags = dict(lag_crval1=np.arange(12, 36, 1.0), lag_crval2=np.arange(-6, 18, 1.0),
                    lag_cdelt1=np.arange(-0.002, 0.0021, 0.001), lag_cdelt2=np.arange(-0.002, 0.0021, 0.001),
                    lag_crota=np.arange(-0.25, 0.3, 0.1))
        if args.grid5d_full:
            lags = dict(lag_crval1=np.arange(14, 34, 1.0), lag_crval2=np.arange(-4, 16, 1.0),
                        lag_cdelt1=(np.arange(16) - 8) * 0.001, lag_cdelt2=(np.arange(16) - 8) * 0.001,
                        lag_crota=(np.arange(10) - 5) * 0.1)
        n = int(np.prod([len(v) for v in lags.values()]))
        run = lambda: Alignment(pl, ps, parallelism=True, cdelt_semantics="intended", **lags)\
            .align_using_helioprojective(return_type="corr")  # noqa: E731
        cube, dt = timed(run)
        am = np.unravel_index(np.nanargmax(cube), cube.shape)[:5]
        search = HpcSearch(dl, hl, ds, hs, cdelt_mode="intended", **lags)
        flat = cube.ravel()
        rng = np.random.default_rng(4)
        sel = np.concatenate([[int(np.nanargmax(flat))], rng.integers(0, n, max(1, min(cores, 16) - 1))])
        ref = cube_multiprocess(search, min(cores, len(sel)), sel)
        results.append({"config": "configs[3] 5-D lag grid (crval1 x crval2 x cdelt1 x cdelt2 x crota), intended CDELT "
                                  "semantics", "lags": n, "shape": list(cube.shape[:5]), "wall_s_public_api": dt,
                        "lag_evals_per_s": n / dt, "pixel_samples_per_s": n * 2048 * 2048 / dt,
                        "argmax_index": [int(v) for v in am],
                        "argmax_lag": [float(lags[k][i]) for k, i in zip(
                            ("lag_crval1", "lag_crval2", "lag_cdelt1", "lag_cdelt2", "lag_crota"), am)],
                        "oracle_sample": int(len(sel)), "oracle_max_abs_err": float(np.max(np.abs(flat[sel] - ref)))})
        print(json.dumps(results[-1]), flush=True)

    if "sequence" not in skip:
        from euispice_coreg_b200._synth.scene import PairSpec, make_pair, master_scene
        d = os.path.join(bench.synth_dir(), "sequence")
        os.makedirs(d, exist_ok=True)
        paths = [os.path.join(d, f"frame{i:03d}_small.fits") for i in range(args.frames)]
        if not all(os.path.exists(p) for p in paths):
            sky = master_scene(PairSpec())
            rng = np.random.default_rng(1000)
            for i in range(args.frames):
                jit = tuple(float(v) for v in rng.normal(0.0, 1.5, 2))
                make_pair(d, PairSpec(jitter=jit, noise_seed=1000 + i), tag=f"frame{i:03d}", sky=sky, write_large=False)
        seq = SequenceAlignment(pl, paths, **bench.LAGS)
        seq.align_using_helioprojective(return_type="corr")          # warm-up (allocations, page cache)
        seq = SequenceAlignment(pl, paths, **bench.LAGS)
        cubes, dt = timed(lambda: seq.align_using_helioprojective(return_type="corr"))
        one = Alignment(pl, paths[-1], parallelism=True, **bench.LAGS).align_using_helioprojective(return_type="corr")
        peaks = [[float(bench.LAGS["lag_crval1"][i]), float(bench.LAGS["lag_crval2"][j])]
                 for i, j in (np.unravel_index(np.nanargmax(c), c.shape)[:2] for c in cubes)]
        # steady state: the same sequence three times over (24 frames) in one call
        seq3 = SequenceAlignment(pl, paths * 3, **bench.LAGS)
        _, dt3 = timed(lambda: seq3.align_using_helioprojective(return_type="corr"))
        results.append({"config": "configs[4] frame sequence vs one reference, 60x60 lags per frame",
                        "frames_per_s_24_frames": 3 * args.frames / dt3,
                        "frames": args.frames, "wall_s_public_api": dt, "frames_per_s": args.frames / dt,
                        "lag_evals_per_s": args.frames * 3600 / dt, "device_loop_frames_per_s": seq.frames_per_s,
                        "last_frame_equals_single_pair_alignment": bool(np.array_equal(cubes[-1], one, equal_nan=True)),
                        "argmax_lags_arcsec": peaks})
        print(json.dumps(results[-1]), flush=True)

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
