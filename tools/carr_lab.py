"""BASELINE configs[1] (Carrington grid 2048^2, 120 x 120 CRVAL lags) alone: device time of the search on one GPU.
    python tools/carr_lab.py [--steps 5]        (GPU box; also the command profiled with ncu)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--shards", type=int, default=0, help="time the lag slice of every rank of a W-rank run, one by one")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    pl, ps = bench.ensure_config1()
    if args.shards:
        for r in range(args.shards):
            print(json.dumps(bench.carrington_secondary(pl, ps, args.steps, 1, lambda: None, torch, None,
                                                        align_wall=False, shard=(r, args.shards))), flush=True)
        return
    r = bench.carrington_secondary(pl, ps, args.steps, 1, lambda: None, torch, None, align_wall=False)
    print(json.dumps(r))


if __name__ == "__main__":
    main()
