"""Device-timed `align_using_initial_carrington` search on a synthetic pair of Carrington maps.
Usage: python tools/car_bench.py [small_nx small_ny n_lag]  -> one JSON line (lag-evals/s, pixel-samples/s)."""
import json
import os
import sys
import tempfile
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(snx=2048, sny=1024, nl=60):
    import torch
    from euispice_coreg_b200._synth.carmaps import CarPairSpec, make_car_pair
    from euispice_coreg_b200.hdrshift import Alignment
    spec = CarPairSpec(small_n=(snx, sny), large_n=(snx * 3 // 4, sny * 3 // 4), small_cdelt=0.02 * 96 / snx * 8,
                       large_cdelt=0.02 * 96 / snx * 8 * 2.5, master_n=4096, master_cdelt=0.02 * 96 / snx * 8 * snx * 1.7 / 4096)
    d = tempfile.mkdtemp()
    p_large, p_small, spec = make_car_pair(d, spec)
    step = spec.small_cdelt * 3600.0
    lag = (np.arange(nl) - nl // 2) * step
    warnings.simplefilter("ignore")
    a = Alignment(p_large, p_small, lag, lag, None, None, None)
    t0 = time.perf_counter()
    cube = a.align_using_initial_carrington(return_type="corr")
    wall_first = time.perf_counter() - t0
    eng = a.engine
    from euispice_coreg_b200.hdrshift import engine as E
    refs = type("R", (), dict(crval1_ref=a.crval1_ref, crval2_ref=a.crval2_ref, crota_ref=a.crota_ref,
                              cdelt1_ref=a.cdelt1_ref, cdelt2_ref=a.cdelt2_ref))()
    d1, d2, d3, d4, d5 = E.flat_lag_grid(a.lag_crval1, a.lag_crval2, a.lag_cdelt1, a.lag_cdelt2, a.lag_crota)
    table, dead = E.car_lag_table(a.hdr_small, refs, d1, d2, d3, d4, d5)
    tab = torch.from_numpy(table).cuda()
    out = torch.empty(table.shape[0], dtype=torch.float64, device="cuda")
    for _ in range(2):
        eng.evaluate(tab, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.evaluate(tab, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    i = np.unravel_index(np.nanargmax(cube), cube.shape)
    return ({"workload": f"initial_carrington: CAR maps {snx}x{sny} vs {spec.large_n[0]}x{spec.large_n[1]}, "
                                  f"{nl}x{nl} CRVAL lags", "lags": int(table.shape[0]), "ms_per_search": ms,
                      "lag_evals_per_s": table.shape[0] / ms * 1e3,
                      "pixel_samples_per_s": table.shape[0] * snx * sny / ms * 1e3, "align_wall_first_s": wall_first,
                      "argmax_lag_deg": [float(a.lag_crval1[i[0]]), float(a.lag_crval2[i[1]])],
                      "true_shift_deg": list(spec.true_shift), "max_r": float(np.nanmax(cube))})


if __name__ == "__main__":
    print(json.dumps(run(*(int(v) for v in sys.argv[1:4]))))
