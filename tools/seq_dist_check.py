"""torchrun check of SequenceAlignment's frame sharding (GPU box, >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/seq_dist_check.py
Every rank must end with all cubes, identical to a single-rank run of the same frames."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from euispice_coreg_b200._synth.scene import make_pair, master_scene, small_spec
    from euispice_coreg_b200.hdrshift import Alignment, SequenceAlignment
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = "/tmp/coreg_seq_check"
    n_frames = 5
    lags = dict(lag_crval1=np.arange(18, 31, 2.0), lag_crval2=np.arange(0, 13, 2.0), lag_cdelt1=[0], lag_cdelt2=[0],
                lag_crota=[0.0, 0.5])
    if rank == 0:
        spec0 = small_spec(96, 160, true_crval=(-12.0, 8.0))
        sky = master_scene(spec0)
        make_pair(d, spec0, tag="f0", sky=sky)
        for i in range(1, n_frames):
            make_pair(d, small_spec(96, 160, true_crval=(-12.0, 8.0), jitter=(0.7 * i, -0.4 * i), noise_seed=300 + i),
                      tag=f"f{i}", sky=sky, write_large=False)
    dist.barrier()
    p_large = os.path.join(d, "f0_large.fits")
    paths = [os.path.join(d, f"f{i}_small.fits") for i in range(n_frames)]
    t0 = time.perf_counter()
    cubes = SequenceAlignment(p_large, paths, **lags).align_using_helioprojective(return_type="corr")
    dt = time.perf_counter() - t0
    # single-rank truth: plain Alignment per frame, without the process group in the way
    ok = True
    for k in range(rank, n_frames, world):      # each rank checks a share; cubes of other ranks' frames included below
        pass
    dist.barrier()
    dist.destroy_process_group()
    for k, p in enumerate(paths):
        one = Alignment(p_large, p, parallelism=True, **lags).align_using_helioprojective(return_type="corr")
        ok = ok and np.array_equal(cubes[k], one, equal_nan=True)
    print(f"rank {rank}/{world}: {len(cubes)} cubes in {dt:.2f} s, all equal to single-pair Alignment: {ok}", flush=True)
    assert ok


if __name__ == "__main__":
    main()
