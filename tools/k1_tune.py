"""Time the fused helioprojective lag kernel on config 1 for every tuning variant (GPU box only).
    python tools/k1_tune.py [--variants 0,1,2,3] [--out gpurun_out/k1_tune.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from euispice_coreg_b200 import _ext
    from euispice_coreg_b200._compat.wcs import TanWcs
    from euispice_coreg_b200.hdrshift import engine as E
    from euispice_coreg_b200.hdrshift.alignment import Alignment
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,1,2,3")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/k1_tune.json")
    ap.add_argument("--only-fast", action="store_true")
    args = ap.parse_args()
    pl, ps = bench.ensure_config1()
    a = Alignment(pl, ps, parallelism=True, **bench.LAGS)
    a.method, a.coordinate_frame = "correlation", "final_helioprojective"
    a._load_pair()
    a._set_initial_header_values(True)
    w_small, w_large = TanWcs.from_header(a.hdr_small), TanWcs.from_header(a.hdr_large)
    d = E.flat_lag_grid(a.lag_crval1, a.lag_crval2, a.lag_cdelt1, a.lag_cdelt2, a.lag_crota)
    results = []
    base = None
    combos = (("f64", False, False), ("f64", False, True), ("f64", True, True))
    if args.only_fast:
        combos = (("f64", False, False),)
    for storage, strict, no_fast in combos:
        if True:
            for v in [int(x) for x in args.variants.split(",")]:
                eng = E.LagSearchEngine(order=2, strict=strict, variant=v, small_storage=storage, no_fast=no_fast)
                eng.set_small(a.data_small)
                eng.prepare_hpc(a.data_large, w_large, w_small)
                table, _ = eng.hpc_lag_table(a.hdr_small, a, *d)
                fast = table.shape[1] == _ext.TAN_WCS_DOUBLES
                tab = eng._upload(table)
                out = torch.empty(table.shape[0], dtype=torch.float64, device=eng.device)
                eng.evaluate(tab, out)
                torch.cuda.synchronize()
                _ext.profile_begin()
                for _ in range(args.steps):
                    eng.evaluate(tab, out)
                ms, n = _ext.profile_end()
                c = out.cpu().numpy()
                if base is None:
                    base = c
                rec = {"storage": storage, "strict": strict, "fast_kernel": bool(fast), "variant": v,
                       "k1_ms_per_search": ms / args.steps, "launches_per_search": n // args.steps,
                       "max_abs_diff_vs_first": float(np.nanmax(np.abs(c - base))),
                       "argmax": int(np.nanargmax(c))}
                print(json.dumps(rec), flush=True)
                results.append(rec)
                del eng, tab, out
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
