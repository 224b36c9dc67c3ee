"""Weak scaling of SequenceAlignment under torchrun (GPU box): 8 config-1-sized frames per GPU, 60x60 lags each.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/seq_scale.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from euispice_coreg_b200._synth.scene import PairSpec, make_pair, master_scene
    from euispice_coreg_b200.hdrshift import SequenceAlignment
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)
    pl, ps = bench.ensure_config1(rank, barrier)
    d = os.path.join(bench.synth_dir(), "sequence")
    per_gpu = 8
    paths = [os.path.join(d, f"frame{i:03d}_small.fits") for i in range(per_gpu)]
    if rank == 0 and not all(os.path.exists(p) for p in paths):
        os.makedirs(d, exist_ok=True)
        sky = master_scene(PairSpec())
        rng = np.random.default_rng(1000)
        for i in range(per_gpu):
            jit = tuple(float(v) for v in rng.normal(0.0, 1.5, 2))
            make_pair(d, PairSpec(jitter=jit, noise_seed=1000 + i), tag=f"frame{i:03d}", sky=sky, write_large=False)
    barrier()
    frames = paths * world                       # the same 8 files per GPU: weak scaling
    SequenceAlignment(pl, frames, **bench.LAGS).align_using_helioprojective(return_type="corr")    # warm-up
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    cubes = SequenceAlignment(pl, frames, **bench.LAGS).align_using_helioprojective(return_type="corr")
    torch.cuda.synchronize()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        same = all(np.array_equal(cubes[i], cubes[i % per_gpu], equal_nan=True) for i in range(len(frames)))
        print(json.dumps({"n_gpus": world, "frames": len(frames), "wall_s": dt, "frames_per_s": len(frames) / dt,
                          "lag_evals_per_s": len(frames) * 3600 / dt, "scaling": "weak (8 frames per GPU)",
                          "repeated_frames_identical_across_ranks": bool(same)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
