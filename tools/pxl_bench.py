"""Device-timed pixel-shift lag search (`coreg_pixel_shift_corr`) on synthetic images.
Usage: python tools/pxl_bench.py [large_n small_n n_lag n_rot] -> one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(ln=2048, sn=1024, nl=61, nr=3):
    import torch
    from euispice_coreg_b200 import _ext
    rng = np.random.default_rng(5)
    large = torch.from_numpy(rng.lognormal(5, 1, (ln, ln))).cuda()
    y0 = x0 = (ln - sn - 1) // 2
    small = large[y0 + 3:y0 + 3 + sn, x0 - 5:x0 - 5 + sn].clone()
    smalls = torch.stack([small + 0.01 * k for k in range(nr)]).contiguous()
    pivots = torch.zeros(2, dtype=torch.float64, device="cuda")
    _ext.finite_mean(large, pivots[0:1])
    _ext.finite_mean(smalls, pivots[1:2])
    lag = np.arange(nl) - nl // 2
    for _ in range(2):
        corr = _ext.pixel_shift_corr(large, smalls, x0, y0, lag, lag, pivots)
    _ext.profile_begin()
    for _ in range(5):
        corr = _ext.pixel_shift_corr(large, smalls, x0, y0, lag, lag, pivots)
    ms, n = _ext.profile_end()
    ms /= n
    c = corr.cpu().numpy()
    i = np.unravel_index(np.argmax(c), c.shape)
    samples = float(nl * nl * nr) * sn * sn
    return ({"workload": f"pixel shift: small {sn}x{sn} in large {ln}x{ln}, {nl}x{nl}x{nr} lags",
                      "lags": nl * nl * nr, "kernel_ms": ms, "lag_evals_per_s": nl * nl * nr / ms * 1e3,
                      "pixel_samples_per_s": samples / ms * 1e3,
                      "unamortised_GBps": samples * 16 / ms * 1e3 / 1e9,
                      "argmax": [int(lag[i[0]]), int(lag[i[1]]), int(i[2])], "max_r": float(c.max())})


if __name__ == "__main__":
    print(json.dumps(run(*(int(v) for v in sys.argv[1:5]))))
