"""Time the GPU decoder of RICE tile-compressed images on a 2048-pixel-wide float32 image (GPU box only).
The test writer (oracle/rice.py) is pure Python, so only `--rows` rows are encoded; one thread decodes one tile, so
the decode time of a row-tiled image depends on the tile length, not on the number of rows (up to ~19 000 rows)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    from euispice_coreg_b200._compat import fits_lite
    from oracle import rice
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=256)
    args = ap.parse_args()
    _, ps = bench.ensure_config1()
    img = fits_lite.open(ps)[0].data[:args.rows].astype(np.float32)
    p = "/tmp/coreg_rice_bench.fits"
    t0 = time.perf_counter()
    rice.write_compressed_image(p, img, quantize_scale=float(np.std(np.diff(img, axis=1))) / 16.0, zdither0=1)
    t_enc = time.perf_counter() - t0
    raw = os.path.getsize(p)
    hdu = fits_lite.open(p)[1]
    hdu.device_data()
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        hdu = fits_lite.open(p)[1]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev = hdu.device_data()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hdu = fits_lite.open(p)[1]
    # kernel alone: everything already on the device is not exposed by the wrapper, so time the wrapper's GPU part
    e0.record()
    dev = hdu.device_data()
    e1.record()
    torch.cuda.synchronize()
    err = float(np.nanmax(np.abs(dev.cpu().numpy() - img)))
    print(json.dumps({"rows": args.rows, "cols": img.shape[1], "file_bytes": raw, "pixel_bytes": img.nbytes,
                      "compression_ratio": img.nbytes / raw, "python_encode_s": t_enc,
                      "decode_wall_ms_host_to_device_tensor": 1e3 * min(times),
                      "decode_device_ms_incl_h2d": e0.elapsed_time(e1), "max_abs_err": err}))


if __name__ == "__main__":
    main()
