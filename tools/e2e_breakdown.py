"""Stage-by-stage wall times of the host-buffer (e2e) path on config 1 (GPU box only)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from euispice_coreg_b200._compat.wcs import TanWcs
    from euispice_coreg_b200.hdrshift import engine as E
    from euispice_coreg_b200.hdrshift.alignment import Alignment
    pl, ps = bench.ensure_config1()
    a = Alignment(pl, ps, parallelism=True, **bench.LAGS)
    a.method, a.coordinate_frame = "correlation", "final_helioprojective"
    a._load_pair()
    a._set_initial_header_values(True)
    w_small, w_large = TanWcs.from_header(a.hdr_small), TanWcs.from_header(a.hdr_large)
    d = E.flat_lag_grid(a.lag_crval1, a.lag_crval2, a.lag_cdelt1, a.lag_cdelt2, a.lag_crota)
    h_large = torch.from_numpy(np.ascontiguousarray(a.data_large)).pin_memory().numpy()
    h_small = torch.from_numpy(np.ascontiguousarray(a.data_small)).pin_memory().numpy()

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    for rep in range(3):
        t = [sync()]
        e = E.LagSearchEngine(order=2)
        t.append(sync())
        e.set_small(h_small)
        t.append(sync())
        e.prepare_hpc(h_large, w_large, w_small)
        t.append(sync())
        table, _ = e.hpc_lag_table(a.hdr_small, a, *d)
        t.append(sync())
        cube = e.search(table)
        t.append(sync())
        names = ["engine()", "set_small", "prepare_hpc", "lag table (host)", "search"]
        print(rep, " ".join(f"{n}={1e3 * (t[i + 1] - t[i]):.1f}ms" for i, n in enumerate(names)),
              f"total={1e3 * (t[-1] - t[0]):.1f}ms", flush=True)
    t0 = sync()
    res = Alignment(pl, ps, parallelism=True, **bench.LAGS).align_using_helioprojective()
    print("align() wall", 1e3 * (sync() - t0), "ms")
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    res = Alignment(pl, ps, parallelism=True, **bench.LAGS).align_using_helioprojective()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
