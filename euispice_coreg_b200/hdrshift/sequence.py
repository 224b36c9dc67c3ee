"""`SequenceAlignment` -- many small-FOV frames of one sequence against ONE large-FOV reference image.

BASELINE.json configs[4] (a batch of HRIEUV frames aligned to one FSI 174 image); the batch pattern is the one of
the reference's `jitter_correction.jitter_correction_imagers` (`jitter_correction/jitter_correction.py:14-174`):
one `Alignment` per frame with the same lag grid, results written per frame. The reference re-opens, re-uploads
(shared memory) and re-prepares everything for every frame. Here the large image is uploaded once and stays
resident in HBM; per frame only the small image (17 MB as float32), its candidate-header table and the one-time
cut of the large image onto the frame's grid are new. Kernel launches are asynchronous, so the host reads and
decodes frame k+1 while the device searches frame k, and the cube of frame k is fetched after frame k+1 has been
enqueued. With `torch.distributed` initialised the FRAMES are sharded over the ranks (each GPU: resident
reference, its own frames, the full lag grid -- SURVEY.md section 8e) and one all-gather assembles the cubes.

Per frame the result is exactly what `Alignment(large, frame, ...).align_using_helioprojective()` returns.
"""
from __future__ import annotations

import numpy as np

from .._compat.wcs import TanWcs
from . import engine as _engine
from .alignment import Alignment, _Refs


_SIDE = {}


def _side_stream(device):
    """One upload / preparation stream per device for the life of the process (the caching allocator keeps a pool per
    stream: a fresh stream per call would re-allocate every per-frame buffer)."""
    torch = _engine._torch()
    key = str(device)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def frames_of_rank(n_frames, rank, world):
    """Round-robin frame sharding: rank r searches frames r, r + world, ..."""
    return list(range(rank, n_frames, world))


def gather_frame_cubes(cubes, n_frames, n_lags, device):
    """{frame index: flat cube} of this rank's frames -> the same for ALL frames, on every rank: one all-gather of a
    [frames per rank, n_lags] block (NCCL: `all_gather_into_tensor`; gloo, for the CPU tests of this logic: list
    `all_gather`). A rank without frames (more ranks than frames) contributes an all-NaN block of the same shape."""
    torch = _engine._torch()
    dist, rank, world = _engine._dist_info()
    if dist is None or world == 1 or n_frames == 0:
        return cubes
    per = (n_frames + world - 1) // world
    width = max(n_lags, 1)
    local = torch.full((per, width), float("nan"), dtype=torch.float64, device=device)
    for j, k in enumerate(frames_of_rank(n_frames, rank, world)):
        local[j, :n_lags] = torch.from_numpy(np.ascontiguousarray(cubes[k], dtype=np.float64).ravel()).to(device)
    if local.is_cuda:
        full = torch.empty((world * per, width), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(full, local)
    else:
        parts = [torch.empty((per, width), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, local)
        full = torch.cat(parts)
    full = full.cpu().numpy().reshape(world, per, -1)
    return {r + world * j: full[r, j, :n_lags] for r in range(world) for j in range(per) if r + world * j < n_frames}


class SequenceAlignment:

    def __init__(self, large_fov_known_pointing: str, list_small_fov_to_correct, lag_crval1, lag_crval2,
                 lag_cdelt1=None, lag_cdelt2=None, lag_crota=None, small_fov_value_min=None,
                 small_fov_value_max=None, large_fov_window=-1, small_fov_window=-1, reprojection_order=2,
                 force_crota_0=False, unit_lag="arcsec", cdelt_semantics="reference", strict_arithmetic=False,
                 arithmetic=None):
        self.large_fov_known_pointing = large_fov_known_pointing
        self.list_small = list(list_small_fov_to_correct)
        self.kw = dict(lag_crval1=lag_crval1, lag_crval2=lag_crval2, lag_cdelt1=lag_cdelt1, lag_cdelt2=lag_cdelt2,
                       lag_crota=lag_crota, small_fov_value_min=small_fov_value_min,
                       small_fov_value_max=small_fov_value_max, large_fov_window=large_fov_window,
                       small_fov_window=small_fov_window, reprojection_order=reprojection_order,
                       force_crota_0=force_crota_0, unit_lag=unit_lag, cdelt_semantics=cdelt_semantics,
                       strict_arithmetic=strict_arithmetic, parallelism=True, arithmetic=arithmetic)
        self.engine = None
        self.frames_per_s = None

    def _host_prepare(self, path_small):
        """Everything `Alignment` does on the host before the device search, for one frame."""
        a = Alignment(self.large_fov_known_pointing, path_small, **self.kw)
        a.method, a.coordinate_frame = "correlation", "final_helioprojective"
        a.lon_ctype, a.lat_ctype, a.ang2pipi = "HPLN-TAN", "HPLT-TAN", True
        f_small = a._open_small()
        a.hdr_small = f_small[a.small_fov_window].header.copy()
        a._check_ant_create_pcij_matrix(a.hdr_small)
        h_small = f_small[a.small_fov_window]
        masking = a.small_fov_value_min is not None or a.small_fov_value_max is not None
        raw = h_small.raw_big_endian() if (not masking and hasattr(h_small, "raw_big_endian")) else None
        if raw is not None and raw.dtype == np.dtype(">f4") and raw.ndim == 2:
            a.data_small = raw      # as stored: swapped on the device; the all-NaN check happens there too (`finish`)
        else:
            a.data_small = a._float_image(h_small.data)
            a._set_removed_values_to_nan_in_datasmall(fov_limits=None, remove_fov_limits=None)
            if np.isnan(a.data_small).all():
                raise ValueError("minimum or maximum value have set all small FOV to nan")
        a._set_initial_header_values(True)
        if a.unit_lag != a.hdr_small["CUNIT1"] or a.unit_lag != a.hdr_small["CUNIT2"]:
            raise ValueError("lag.unit and cUNIT are not the same")
        return a

    def align_using_helioprojective(self, return_type="AlignmentResults"):
        """-> list (one entry per frame, input order) of `AlignmentResults` or, with return_type="corr", of cubes
        float64 [n_crval1, n_crval2, n_cdelt1, n_cdelt2, n_crota, 1]."""
        import time
        torch = _engine._torch()
        dist, rank, world = _engine._dist_info()
        n_frames = len(self.list_small)
        mine = frames_of_rank(n_frames, rank, world)
        # two engines = two sets of per-frame device buffers (small image, cut of the large image, pivots, workspace)
        # sharing the one resident large image: frame k+1 is uploaded and prepared on a side stream while frame k is
        # being searched on the main stream
        engs = [_engine.LagSearchEngine(order=self.kw["reprojection_order"], strict=self.kw["strict_arithmetic"],
                                        arithmetic=self.kw["arithmetic"])
                for _ in range(2)]
        eng = engs[0]
        self.engine = eng
        # the reference image: read, checked and uploaded once
        a0 = Alignment(self.large_fov_known_pointing, self.list_small[0] if self.list_small else "", **self.kw)
        f_large = a0._open_large()
        hdr_large = f_large[a0.large_fov_window].header.copy()
        a0._check_ant_create_pcij_matrix(hdr_large)
        eng.set_large(a0._float_image(f_large[a0.large_fov_window].data), TanWcs.from_header(hdr_large))
        engs[1].d_large, engs[1].wcs_large = eng.d_large, eng.wcs_large
        # the cube shape depends on the lag arrays only: every rank needs it for the all-gather, also one that gets no
        # frame (more ranks than frames)
        shape5 = None
        if n_frames and not mine:
            a = self._host_prepare(self.list_small[0])
            shape5 = (len(a.lag_crval1), len(a.lag_crval2), len(a.lag_cdelt1), len(a.lag_cdelt2), len(a.lag_crota))
            del a
        pending = None                                      # (frame index, device cube, dead mask, Alignment, ...)
        cubes = {}
        aligns = {}

        def finish(item):
            # fetch on the side stream, ordered after THIS frame's search only: the main stream may already be busy
            # with the next frame, and a copy enqueued there would make the host wait for that one too
            k, out_dev, dead, a, evt, e, tab_dev = item
            with torch.cuda.device(out_dev.device), torch.cuda.stream(side):
                side.wait_event(evt)
                # mixed arithmetic (opt-in): lags that tripped the kernel's guard are redone in FP64, here on the side
                # stream; this engine's buffers are not touched again before the side stream reaches its next frame
                e.resolve_flags(tab_dev, out_dev)
                host = out_dev.cpu()
                if e.small_count() == 0:     # (this engine's statistics still describe frame k: see above)
                    raise ValueError("minimum or maximum value have set all small FOV to nan")
            cubes[k] = np.where(dead, 0.0, host.numpy())
            aligns[k] = a

        with torch.cuda.device(eng.device):
            main = torch.cuda.current_stream()
            side = _side_stream(eng.device)
            searched = [None, None]                         # per engine: event after its last search on `main`
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for n, k in enumerate(mine):
            e = engs[n & 1]
            a = self._host_prepare(self.list_small[k])       # overlaps the device search of the previous frame
            d = _engine.flat_lag_grid(a.lag_crval1, a.lag_crval2, a.lag_cdelt1, a.lag_cdelt2, a.lag_crota)
            refs = _Refs()
            refs.crval1_ref, refs.crval2_ref, refs.crota_ref = a.crval1_ref, a.crval2_ref, a.crota_ref
            refs.cdelt1_ref, refs.cdelt2_ref = a.cdelt1_ref, a.cdelt2_ref
            shape5 = (len(a.lag_crval1), len(a.lag_crval2), len(a.lag_cdelt1), len(a.lag_cdelt2), len(a.lag_crota))
            with torch.cuda.device(e.device), torch.cuda.stream(side):
                if searched[n & 1] is not None:
                    side.wait_event(searched[n & 1])         # this engine's buffers are free again
                e.set_small(a.data_small, pinned=True)
                e.cut_large(TanWcs.from_header(a.hdr_small))
                table, dead = e.hpc_lag_table(a.hdr_small, refs, *d, a.cdelt_semantics)
                tab_dev = e._upload(table, pinned=True)
                out_dev = torch.empty(table.shape[0], dtype=torch.float64, device=e.device)
                ready = side.record_event()
            with torch.cuda.device(e.device):
                main.wait_event(ready)
                e.evaluate(tab_dev, out_dev)                  # asynchronous: returns once the kernels are enqueued
                searched[n & 1] = main.record_event()
            a.data_small = None
            if pending is not None:
                finish(pending)
            pending = (k, out_dev, dead, a, searched[n & 1], e, tab_dev)
        if pending is not None:
            finish(pending)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        self.frames_per_s = len(mine) / dt if dt > 0 and mine else None
        # assemble: every rank ends with all cubes (one all-gather of [frames per rank, n_lags])
        n_lags = int(np.prod(shape5)) if shape5 is not None else 0
        cubes = gather_frame_cubes(cubes, n_frames, n_lags, eng.device)
        out = []
        for k in range(n_frames):
            cube = cubes[k].reshape(shape5 + (1,))
            if return_type == "corr":
                out.append(cube)
                continue
            a = aligns.get(k)
            if a is None:                                   # frame searched by another rank: host prep only
                a = self._host_prepare(self.list_small[k])
                a.data_small = None
            out.append(a._wrap_results(cube, "AlignmentResults"))
        return out
