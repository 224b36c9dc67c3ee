#!/usr/bin/env python
"""bench.py -- lag-evaluations/s of the helioprojective pointing search on BASELINE.json configs[0]
(HRIEUV-like 2048^2 vs FSI-174-like 3072^2, 60x60 CRVAL lags at 1 arcsec, synthetic FITS).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over the whole 3600-lag grid (fused lag kernel + finalize, plus the
all-gather of the cube when N > 1; the lag list is sharded over the N ranks => strong scaling).
`value` is device-timed with inputs resident in HBM; `e2e` is the same search through the host-facing engine call
with HOST buffers (H2D of both images and the lag table, the one-time resampling, the search, D2H of the cube).
`--impl reference` times the reference's CPU path: the reference itself cannot be installed in this image
(astropy/sunpy/poetry-core absent, no network) so the reference-structured oracle port is run with one
process per host core on a bounded sample of the same lag grid.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "config1: HRIEUV-like 2048x2048 vs FSI174-like 3072x3072, helioprojective, 60x60 CRVAL lags @1arcsec"
FP64_INSTR_PER_SAMPLE = 69.0   # algorithmic FP64 instructions per pixel-sample (HPC), counted
#                                on the reference's formulation (pixel -> world -> pixel per lag)
K1_DRAM_BYTES_PER_LAUNCH = 1708.3e6  # dram__bytes_read.sum + dram__bytes_write.sum of one config-1 launch of the rolling
#                                      kernel (ncu --set full, profiles/r1_ncu_roll_v7_raw.csv): 733.6 MB + 974.6 MB (the
#                                      per-warp records of the barrier-free variant are written and read back once)
FP64_EXECUTED_PER_SAMPLE = 31.7  # FP64 thread-instructions the column-rolling kernel executes per pixel-sample = the
#                                  algorithmic count of ITS formulation (ncu: DADD + DMUL + DFMA of
#                                  one launch / pixel-samples, profiles/r1_roll_kernel.md)
# mixed-arithmetic rolling kernel (FP64 projection, FP32 spline; the default when the small image is float32), from one
# ncu --set full capture of a config-1 launch (profiles/r1_ncu_roll_mixed_v2.txt): 2.2539e10 warp-instructions and
# 9.42e10 FP64 thread-instructions over 1.51e10 pixel-samples = 4.72e8 warp-samples
MIXED_INSTR_PER_WARP_SAMPLE = 47.8   # all warp-instructions per warp-sample (32 pixel-samples)
MIXED_FP64_PER_SAMPLE = 6.2          # of which FP64 (4 per pixel for the two coordinate quadratics + per-lag set-up)
MIXED_DRAM_BYTES_PER_LAUNCH = 849.1e6   # dram read 144.0 MB + write 705.0 MB of that launch
BYTES_PER_SAMPLE = 8.0         # un-amortised: one f32 sample of each image per pixel-sample
LAGS = dict(lag_crval1=np.arange(-30, 30, 1.0), lag_crval2=np.arange(-30, 30, 1.0), lag_cdelt1=np.array([0.0]),
            lag_cdelt2=np.array([0.0]), lag_crota=np.array([0.0]))


def synth_dir():
    d = os.environ.get("COREG_BENCH_DATA", "/tmp/coreg_bench_data")
    os.makedirs(d, exist_ok=True)
    return d


def ensure_config1(rank=0, barrier=None):
    from euispice_coreg_b200._synth.scene import make_config1
    d = synth_dir()
    pl, ps = os.path.join(d, "config1_large.fits"), os.path.join(d, "config1_small.fits")
    if rank == 0 and not (os.path.exists(pl) and os.path.exists(ps)):
        make_config1(d)
    if barrier is not None:
        barrier()
    return pl, ps


def load_pair(pl, ps):
    from euispice_coreg_b200._compat import fits_lite
    L, S = fits_lite.open(pl)[0], fits_lite.open(ps)[0]
    return L.data, L.header, S.data, S.header


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm (reference-structured oracle port)
# ---------------------------------------------------------------------------------------------------------
_CPU_SEARCH = None


def cpu_search():
    """The CPU arm's one-time preparation (FITS read, PCi_j checks, the one-time cut of the large image), done once
    per process -- like the reference does it once per `align_using_helioprojective` call, outside the lag loop."""
    global _CPU_SEARCH
    if _CPU_SEARCH is None:
        from oracle.hpc import HpcSearch
        pl, ps = ensure_config1()
        dl, hl, ds, hs = load_pair(pl, ps)
        _CPU_SEARCH = HpcSearch(dl, dict(hl.items()), ds, dict(hs.items()), **LAGS)
    return _CPU_SEARCH


def cpu_sample(n_lags_per_core=1, cores=None):
    """Evaluate a bounded sample of the config-1 lag grid the way the reference does (per lag: pixel->world and
    world->pixel of the full grid, scipy map_coordinates, NaN compaction, two-pass Pearson) with one forked
    process per host core. Returns (lag_evals_per_s, seconds, n_lags, cores)."""
    from oracle.hpc import cube_multiprocess
    cores = cores or len(os.sched_getaffinity(0))
    search = cpu_search()
    n_total = 3600
    n = cores * n_lags_per_core
    sel = np.linspace(0, n_total - 1, n).astype(int)   # spread over the grid: in- and out-of-overlap lags alike
    t0 = time.perf_counter()
    cube_multiprocess(search, cores, sel)
    dt = time.perf_counter() - t0
    return n / dt, dt, n, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_core = 1
    secs = []
    cores = len(os.sched_getaffinity(0))
    cpu_search()
    for i in range(args.warmup + args.steps):
        v, dt, n, cores = cpu_sample(per_core, cores)
        if i >= args.warmup:
            secs.append(dt)
    total_lags = cores * per_core * args.steps
    value = total_lags / sum(secs)
    sample = (f"{cores * per_core} of 3600 lags per step (1 per core, spread over the grid), {args.steps} steps, "
              f"{sum(secs):.1f} s; one-time cut of the large image outside the timed region")
    line = {"impl": "reference", "metric": "lag_evals_per_s", "value": value, "unit": "lag-evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(secs) / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "lags": 3600, "grid": [2048, 2048], "spline_order": 2},
            "cpu_baseline": {"value": value, "unit": "lag-evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "lag-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference not installable here (astropy, sunpy, poetry-core absent); oracle port of "
                    "hdrshift/alignment.py:509-549 with multiprocessing over host cores"}
    print(json.dumps(line), flush=True)
    return 0


CARRINGTON_LAGS = dict(lag_crval1=np.arange(-60, 60, 1.0), lag_crval2=np.arange(-60, 60, 1.0), lag_cdelt1=np.array([0.0]),
                       lag_cdelt2=np.array([0.0]), lag_crota=np.array([0.0]))
CARRINGTON_GRID = dict(lonlims=(200.0, 300.0), latlims=(-20.0, 20.0), shape=(2048, 2048))


def carrington_secondary(pl, ps, steps, world, barrier, torch, dist, variant=0, align_wall=True):
    """BASELINE.json configs[1] (same image pair on a user Carrington grid 2048^2, 120 x 120 CRVAL lags), measured
    beside the headline: device-timed search with everything resident (lags sharded like the headline) and the wall
    time of the public call. Returns a dict for the JSON line."""
    from euispice_coreg_b200.hdrshift import engine as E
    from euispice_coreg_b200.hdrshift.alignment import Alignment
    a = Alignment(pl, ps, parallelism=True, **CARRINGTON_LAGS)
    a.method, a.coordinate_frame, a.method_carrington_reprojection = "correlation", "final_carrington", "fa"
    a._load_pair()
    a._set_initial_header_values(True)
    a.lonlims, a.latlims, a.shape = (CARRINGTON_GRID[k] for k in ("lonlims", "latlims", "shape"))
    d1, d2, _, _, _ = E.flat_lag_grid(a.lag_crval1, a.lag_crval2, a.lag_cdelt1, a.lag_cdelt2, a.lag_crota)
    eng = E.LagSearchEngine(order=2, small_storage="auto", variant=variant)   # float32 storage, as Alignment does here
    eng.set_small(a.data_small)
    r_sun = float(a.lag_solar_r[0])
    eng.prepare_carrington_large(a.data_large, a.hdr_large, r_sun, a.lonlims, a.latlims, a.shape)
    planes = eng.carrington_planes(a.hdr_small, r_sun, a.lonlims, a.latlims, a.shape)
    roll = a.hdr_small["CROTA"] if "CROTA" in a.hdr_small else a.hdr_small["CROTA2"]
    x0, y0 = eng.carrington_offset(a.hdr_small, a.crval1_ref + d1, a.crval2_ref + d2, roll)
    table = np.stack([x0, y0], axis=1).astype(np.float64)
    n = table.shape[0]
    rank = dist.get_rank() if world > 1 else 0
    chunk, bounds = E.shard_bounds(n, world)
    lo, hi = bounds[rank]
    # the kernel takes its lags in detector-plane patches (engine.offset_patch_order), padded with dummy lags
    i1, i2 = np.unravel_index(np.arange(lo, hi), (len(a.lag_crval1), len(a.lag_crval2)))
    slot, n_slots = E.offset_patch_order(i1, i2)
    padded = np.full((n_slots, 2), np.nan)
    padded[slot] = table[lo:hi]
    tab_dev = eng._upload(padded)
    slot_dev = torch.from_numpy(slot).to(eng.device)
    out_p = torch.empty(n_slots, dtype=torch.float64, device=eng.device)
    nv_p = torch.zeros(n_slots, dtype=torch.int64, device=eng.device)
    out = torch.full((chunk,), float("nan"), dtype=torch.float64, device=eng.device)
    full = torch.empty(chunk * world, dtype=torch.float64, device=eng.device)

    def step():
        eng.evaluate(tab_dev, out_p, nv_p, planes)
        out[:hi - lo] = out_p.index_select(0, slot_dev)
        if world > 1:
            dist.all_gather_into_tensor(full, out)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=eng.device)
    eff = nv_p.index_select(0, slot_dev).sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(eff, op=dist.ReduceOp.SUM)
    ms = float(t.item()) / steps
    barrier()
    wall, am = None, None
    if align_wall:
        t0 = time.perf_counter()
        res = Alignment(pl, ps, parallelism=True, **CARRINGTON_LAGS).align_using_carrington(method="correlation",
                                                                                          **CARRINGTON_GRID)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        am = tuple(int(v) for v in res.max_index[:2])
    else:
        flat = int(torch.argmax(torch.nan_to_num(out[:hi - lo], nan=-2.0)).item()) + lo
        am = (flat // len(a.lag_crval2), flat % len(a.lag_crval2))
    return {"workload": "configs[1]: same pair on a Carrington grid 2048x2048 (lon 200-300 deg, lat +-20 deg), 120x120 "
                        "CRVAL lags @1arcsec", "lags": n, "ms_per_search": ms, "lag_evals_per_s": n / (ms * 1e-3),
            "effective_pixel_samples_per_s": float(eff.item()) / (ms * 1e-3),
            "effective_fraction_of_grid": float(eff.item()) / (n * 2048.0 * 2048.0),
            "align_wall_s": wall,
            "argmax_lag_arcsec": [float(CARRINGTON_LAGS["lag_crval1"][am[0]]), float(CARRINGTON_LAGS["lag_crval2"][am[1]])]}


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from euispice_coreg_b200 import _ext
    from euispice_coreg_b200._compat.wcs import TanWcs
    from euispice_coreg_b200.hdrshift import engine as E
    from euispice_coreg_b200.hdrshift.alignment import Alignment

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    _ext.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local),
                                timeout=datetime.timedelta(seconds=180))
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    pl, ps = ensure_config1(rank, barrier)
    # host-side preparation exactly as Alignment does it (header checks, lag units, PC matrix)
    a = Alignment(pl, ps, parallelism=True, **LAGS)
    a.method, a.coordinate_frame = "correlation", "final_helioprojective"
    a._load_pair()
    a._set_threshold_minmax_to_nan()
    a._set_initial_header_values(True)
    w_small, w_large = TanWcs.from_header(a.hdr_small), TanWcs.from_header(a.hdr_large)
    d = E.flat_lag_grid(a.lag_crval1, a.lag_crval2, a.lag_cdelt1, a.lag_cdelt2, a.lag_crota)
    gny, gnx = a.data_small.shape
    n_pix = gnx * gny

    eng = E.LagSearchEngine(order=2, strict=args.strict, variant=args.variant, small_storage=args.small_storage,
                            no_fast=args.no_fast, arithmetic=args.arithmetic)
    eng.set_small(a.data_small)
    eng.prepare_hpc(a.data_large, w_large, w_small)
    table, _ = eng.hpc_lag_table(a.hdr_small, a, *d)
    n_lags = table.shape[0]
    fast = table.shape[1] == _ext.TAN_WCS_DOUBLES
    chunk, bounds = E.shard_bounds(n_lags, world)
    lo, hi = bounds[rank]
    tab_dev = eng._upload(table[lo:hi])
    local_out = torch.full((chunk,), float("nan"), dtype=torch.float64, device=eng.device)
    full = torch.empty(chunk * world, dtype=torch.float64, device=eng.device)

    # L2 hygiene: a 256 MiB buffer (> the 126 MB L2) is rewritten before every step, inside the timed region
    # (~0.05 ms per step), so no step starts with the previous step's images or partials in L2
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)

    def step():
        l2_flush.zero_()
        eng.evaluate(tab_dev, local_out[:hi - lo])
        if world > 1:
            dist.all_gather_into_tensor(full, local_out)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _ext.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    k1_ms, k1_launches = _ext.profile_end()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = n_lags * args.steps / (ms_total * 1e-3)

    # ---- end to end: host buffers -> cube on the host, every step -------------------------------------------
    # host inputs of the end-to-end arm: the images as the FITS files hold them (float32), in pinned memory
    h_large = torch.from_numpy(np.ascontiguousarray(a.data_large)).pin_memory().numpy()
    h_small = torch.from_numpy(np.ascontiguousarray(a.data_small)).pin_memory().numpy()
    e2e_steps = max(1, min(args.steps, 5))

    def e2e_step():
        e = E.LagSearchEngine(order=2, strict=args.strict, variant=args.variant, small_storage=args.small_storage,
                              no_fast=args.no_fast, arithmetic=args.arithmetic)
        e.pure_shift_hint = eng.pure_shift_hint     # what hpc_lag_table derived from the lag grid
        e.set_small(h_small)
        e.prepare_hpc(h_large, w_large, w_small)
        return e.search(table)

    e2e_step()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        cube = e2e_step()
    torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    small_bytes = eng.small.numel() * eng.small.element_size()
    h2d = h_large.nbytes + h_small.nbytes + table[lo:hi].nbytes
    d2h = n_lags * 8

    # full public API once on every rank (FITS read + host prep + sharded search + all-gather + Gaussian fit):
    # the "align() wall time" metric. It contains a collective, so all ranks take part.
    barrier()
    t0 = time.perf_counter()
    res = Alignment(pl, ps, parallelism=True, arithmetic=args.arithmetic, **LAGS).align_using_helioprojective()
    torch.cuda.synchronize()
    align_wall = time.perf_counter() - t0
    carr = None if args.no_carrington else carrington_secondary(pl, ps, args.steps, world, barrier, torch, dist,
                                                                 args.carrington_variant)
    widened = None
    if world == 1 and not args.no_carrington:
        # the SURVEY 8f-4 paths, timed for the record (single GPU only: not part of `value`)
        try:
            from tools import car_bench, pxl_bench
            widened = {"initial_carrington": car_bench.run(2048, 1024, 60), "pixel_shift": pxl_bench.run()}
        except Exception as exc:   # never lose the headline line over a secondary measurement
            widened = {"error": repr(exc)}
    if rank == 0:
        am = tuple(int(v) for v in res.max_index[:2])
        best = (float(LAGS["lag_crval1"][am[0]]), float(LAGS["lag_crval2"][am[1]]))
        assert np.array_equal(np.nan_to_num(cube), np.nan_to_num(res.corr.ravel())), "e2e cube != public API cube"

        fp64_peak = _ext.fp64_peak(40000)          # FP64 FMA lane-instructions / s, measured now on this GPU
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        k1_avg_ms = k1_ms / max(1, k1_launches)
        samples_per_launch = n_pix * (hi - lo) / max(1, k1_launches // args.steps)
        ach_instr = FP64_INSTR_PER_SAMPLE * samples_per_launch / (k1_avg_ms * 1e-3)
        per_sample = FP64_EXECUTED_PER_SAMPLE if fast else FP64_INSTR_PER_SAMPLE
        ach_exec = per_sample * samples_per_launch / (k1_avg_ms * 1e-3)
        ach_gbs = BYTES_PER_SAMPLE * samples_per_launch / (k1_avg_ms * 1e-3) / 1e9
        mixed = fast and eng.arithmetic == "mixed" and eng.small32 is not None
        roof_fp64 = {"bound": "fp64", "achieved": ach_exec / 1e12, "peak": fp64_peak / 1e12,
                     "unit": "T FP64-instr/s", "frac": ach_exec / fp64_peak,
                     "traffic": K1_DRAM_BYTES_PER_LAUNCH * (hi - lo) / 3600.0 if (fast and n_lags == 3600) else None,
                     "traffic_unit": "B per launch (ncu DRAM read + write of one full 3600-lag launch, scaled by "
                                     "this rank's share of the lags)",
                     "kernel": "lag_corr_roll_kernel" if fast else "lag_corr_kernel<TanCoord>",
                     "kernel_ms": k1_avg_ms,
                     "algorithmic": f"{per_sample:.1f} FP64 instr/pixel-sample x {samples_per_launch:.3e} "
                                    "pixel-samples/launch (homography + shared-floor formulation, DESIGN.md 5)"
                                    if fast else f"{per_sample:.0f} FP64 instr/pixel-sample (SURVEY 8d)",
                     "peak_source": "coreg_fp64_peak DFMA microbenchmark, this run",
                     "note": "the FP64 pipe is the binding unit but a DFMA blocks the warp scheduler's dispatch "
                             "port for 2 cycles and every other instruction for 1, so the reachable fraction "
                             "for this instruction mix is about 0.8 (profiles/r1_roll_kernel.md)"}
        extra = {}
        if mixed:
            # The mixed kernel took the spline off the FP64 pipe; what bounds it is the warp scheduler's dispatch
            # port (one instruction per cycle and SM sub-partition, an FP64 instruction holds it for two:
            # profiles/r1_fp64_issue_model.md). Needed dispatch cycles per warp-sample = all instructions + the
            # FP64 ones once more; peak = SMs x 4 sub-partitions x the SM clock sampled during the timed region.
            sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or 1965.0)
            sms = _ext.load().coreg_device_sm_count()
            sms = sms if sms > 0 else 148
            slots_peak = sms * 4 * sm_hz
            need = (MIXED_INSTR_PER_WARP_SAMPLE + MIXED_FP64_PER_SAMPLE) * samples_per_launch / 32.0
            ach_slots = need / (k1_avg_ms * 1e-3)
            ach_f = MIXED_FP64_PER_SAMPLE * samples_per_launch / (k1_avg_ms * 1e-3)
            roof = {"bound": "issue", "achieved": ach_slots / 1e9, "peak": slots_peak / 1e9,
                    "unit": "G warp-dispatch cycles/s", "frac": ach_slots / slots_peak,
                    "traffic": MIXED_DRAM_BYTES_PER_LAUNCH * (hi - lo) / 3600.0 if n_lags == 3600 else None,
                    "traffic_unit": roof_fp64["traffic_unit"],
                    "kernel": "lag_corr_roll_kernel<MIXED>", "kernel_ms": k1_avg_ms,
                    "algorithmic": f"({MIXED_INSTR_PER_WARP_SAMPLE} instr + {MIXED_FP64_PER_SAMPLE} FP64 counted "
                                   f"twice) dispatch cycles per warp-sample x {samples_per_launch / 32.0:.3e} "
                                   "warp-samples/launch (ncu instruction mix of this kernel, DESIGN.md 5)",
                    "peak_source": f"{sms} SMs x 4 sub-partitions x {sm_hz / 1e6:.0f} MHz (nvidia-smi median under "
                                   "load, this run)",
                    "note": "neither HBM nor tensor cores nor, after the FP32 spline, the FP64 pipe binds this "
                            "kernel: the warp scheduler does. roofline_fp64 gives the FP64 pipe's share, "
                            "roofline_survey the rate on SURVEY 8d's 69-instruction count"}
            extra["roofline_fp64"] = {"bound": "fp64", "achieved": ach_f / 1e12, "peak": fp64_peak / 1e12,
                                      "unit": "T FP64-instr/s", "frac": ach_f / fp64_peak,
                                      "algorithmic": f"{MIXED_FP64_PER_SAMPLE} FP64 instr/pixel-sample executed by "
                                                     "the mixed kernel (the all-FP64 kernel needs 31.7 and reaches "
                                                     "0.64 of this peak: --arithmetic fp64)",
                                      "peak_source": roof_fp64["peak_source"]}
        else:
            roof = roof_fp64
        arith_name = ("strict (scipy op order)" if args.strict else
                      ("mixed: fp64 projection, fp32 spline + segment sums" if mixed else "fp64 fma"))
        line = {
            "metric": "lag_evals_per_s", "value": value, "unit": "lag-evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64+f32" if mixed else "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "lags": n_lags, "grid": [gny, gnx], "spline_order": 2,
                       "arithmetic": arith_name, "variant": args.variant, "kernel": "generic" if (args.no_fast or args.strict) else "fast", "small_storage": str(eng.small.dtype).replace("torch.", ""),
                       "parallelism": f"lag-sharded x{world}",
                       "l2": "flushed: a 256 MiB device buffer is rewritten before every timed step (inside the timed "
                             f"region); every step also rewrites its {eng._work.numel() * 8 / 1e6:.0f} MB partials "
                             "workspace"},
            "pixel_samples_per_s": value * n_pix,
            "e2e": {"value": n_lags * e2e_steps / e2e_s, "unit": "lag-evals/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "what": f"LagSearchEngine from pinned host arrays ({h_large.dtype} large, {h_small.dtype} small, as "
                            "the FITS files hold them): H2D images + lag table, one-time resampling, search, "
                            "all-gather, D2H cube"},
            "align_wall_s": align_wall, "argmax_lag_arcsec": best,
            "carrington": carr,
            "widened": widened,
            # own kernels inside the timed region: per lag-kernel launch the homography table (fast path), the fused
            # lag kernel and the finalize kernel (the L2 flush memset and the NCCL all-gather are not ours)
            "gpu_launches": int(k1_launches * (3 if fast else 2)),
            "clocks": clocks,
            "roofline": roof,
            **extra,
            "roofline_survey": {"bound": "fp64", "achieved": ach_instr / 1e12, "peak": fp64_peak / 1e12,
                                "unit": "T FP64-instr/s", "frac": ach_instr / fp64_peak,
                                "algorithmic": f"{FP64_INSTR_PER_SAMPLE:.0f} FP64 instr/pixel-sample: SURVEY 8d's count "
                                               "for the reference's pixel->world->pixel formulation; > 1 means the "
                                               "kernel needs fewer instructions than that formulation's minimum"},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": ach_gbs / hbm_peak, "traffic": None, "peak_source": hbm_src,
                             "algorithmic": f"{BYTES_PER_SAMPLE:.0f} B/pixel-sample (un-amortised)"},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, dt, n, cores = cpu_sample(1)
            line["cpu_baseline"] = {"value": v, "unit": "lag-evals/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} of 3600 lags (1 per core, spread over the grid), {dt:.1f} s; "
                                              "oracle port of the reference's per-lag body with multiprocessing"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--strict", action="store_true", help="scipy operation order in the spline (no FMA)")
    ap.add_argument("--variant", type=int, default=0, help="kernel tuning variant (tile/occupancy)")
    ap.add_argument("--arithmetic", default=None, choices=["fp64", "mixed"],
                    help="homography kernel: everything in FP64, or FP64 projection + FP32 spline (engine default)")
    ap.add_argument("--no-fast", action="store_true", help="force the generic fused kernel")
    ap.add_argument("--small-storage", default="f64", choices=["auto", "f64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-carrington", action="store_true", help="skip the secondary configs[1] measurement")
    ap.add_argument("--carrington-variant", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
