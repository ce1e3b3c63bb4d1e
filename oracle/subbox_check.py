"""Oracle recomputation of sub-boxes of a large stack (parity at the named sizes).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): used by ``tests/`` and by
``bench.py``'s parity counters, outside every timed region.

The oracle cannot filter a 512x2048x2048 stack in the time a test or a bench
run has, but the reference's result inside a small box depends only on a bounded
neighbourhood of it:

* preprocessing works on whole ``denoise_max_shape`` blocks anchored at the chunk
  origin (``magmap/cv/stack_detect.py:122-150``), so a region whose faces lie on
  that block grid (or on a chunk face) reproduces the chunk's blocks exactly;
* a LoG value depends on the preprocessed voxels within ``r_max = int(4 sigma_max
  + 0.5)`` of it and a 4-D local maximum on the LoG values one voxel around it, so
  peaks farther than ``r_max + 1`` from a region face that is NOT a chunk face are
  exact (faces that are chunk faces see the same 'reflect' as the chunk);
* ``_prune_blobs`` couples blobs closer than ``2 sigma_max sqrt(3)``; two such hops
  are left between the compared core and the first inexact peak.

``check_chunk_core`` recomputes such a region with the restated reference
(``oracle.magmap_restated`` / ``oracle.skimage_restated``) and compares the blobs
of the core with the GPU's table of the same chunk.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence, Tuple

import numpy as np

from oracle import magmap_restated as mm
from oracle import skimage_restated as ski

Range3 = Sequence[Tuple[int, int]]


def region_for_core(core: Range3, chunk_shape: Sequence[int], block: Sequence[int],
                    r_max: int, sigma_max: float, hops: int = 2):
    """``(region, valid)`` for a core box, all chunk-relative ``[(lo, hi)] * 3``:
    ``region`` is the box the oracle must filter (faces on the block grid or on a
    chunk face), ``valid`` the part of it whose peaks are exact."""
    reach = int(math.ceil(2.0 * sigma_max * math.sqrt(3.0))) + 1
    need = (r_max + 1) + hops * reach
    region, valid = [], []
    for (a, b), n, bs in zip(core, chunk_shape, block):
        lo = max(0, (a - need) // bs * bs)
        hi = min(int(n), -(-(b + need) // bs) * bs)
        region.append((lo, hi))
        valid.append((lo if lo == 0 else lo + r_max + 1, hi if hi == n else hi - (r_max + 1)))
    return region, valid


def _ladder(prof: mm.Profile, resolution):
    scale = mm.calc_scaling_factor(resolution)[2]
    return ski.sigma_list(prof.min_sigma_factor * scale, prof.max_sigma_factor * scale,
                          prof.num_sigma)


def plan_region(core: Range3, chunk_shape: Sequence[int], prof: mm.Profile, resolution,
                block: Optional[Sequence[int]] = (25, 25, 25)):
    """The chunk-relative box ``check_chunk_core`` will ask ``fetch`` for."""
    sig = _ladder(prof, resolution)
    bs = tuple(int(b) for b in block) if block is not None else (1, 1, 1)
    return region_for_core(core, chunk_shape, bs, int(4.0 * float(sig.max()) + 0.5),
                           float(sig.max()))[0]


def check_chunk_core(fetch, chunk_shape: Sequence[int],
                     core: Range3, gpu_zyxr: np.ndarray, prof: mm.Profile,
                     resolution: Sequence[float], near_max: float,
                     block: Optional[Sequence[int]] = (25, 25, 25), tol: float = 1e-4) -> Dict:
    """Compare the GPU's blobs of one chunk with the oracle inside ``core``.

    Args:
        fetch: ``fetch(region)`` returns the raw voxels of the chunk-relative box
            ``region`` as a numpy array; or that array itself (fetched beforehand
            for the box ``plan_region`` names).
        gpu_zyxr: ``(n, >=4)`` rows ``z, y, x, radius`` of the GPU's table for this
            chunk, chunk-relative integer coordinates.
        block: preprocessing block shape, None for raw detection (no
            ``saturate_roi`` / ``denoise_roi``).

    Returns counts: ``gpu``, ``oracle``, ``matched``, ``f1``, ``near_threshold``
    (differences whose oracle LoG response is within ``tol`` of the threshold),
    ``order_dependent`` (differences inside the set whose survival depends on
    scikit-image's pair iteration order), ``unexplained``.
    """
    scale = mm.calc_scaling_factor(resolution)[2]
    sig = ski.sigma_list(prof.min_sigma_factor * scale, prof.max_sigma_factor * scale,
                         prof.num_sigma)
    r_max = int(4.0 * float(sig.max()) + 0.5)
    bs = tuple(int(b) for b in block) if block is not None else (1, 1, 1)
    region, valid = region_for_core(core, chunk_shape, bs, r_max, float(sig.max()))
    for (a, b), (va, vb) in zip(core, valid):
        if a < va or b > vb:
            raise ValueError(f"core {core} is not inside the exact part {valid} of its region")
    raw = np.ascontiguousarray(fetch(region) if callable(fetch) else fetch)
    if raw.shape[:3] != tuple(b - a for a, b in region):
        raise ValueError(f"fetched box {raw.shape} is not the region {region}")
    img = mm.preprocess_blocks(raw, bs, prof, near_max) if block is not None else raw
    res = ski.blob_log(img, prof.min_sigma_factor * scale, prof.max_sigma_factor * scale,
                       prof.num_sigma, prof.detection_threshold, prof.overlap, full=True,
                       keep_cube=True)
    off = np.array([r[0] for r in region])
    thr = prof.detection_threshold

    def in_core(zyx):
        m = np.ones(len(zyx), dtype=bool)
        for ax, (a, b) in enumerate(core):
            m &= (zyx[:, ax] >= a) & (zyx[:, ax] < b)
        return m

    want = set()
    if len(res.blobs):
        zyx = res.blobs[:, :3].astype(np.int64) + off
        sidx = np.argmin(np.abs(res.blobs[:, 3:4] - sig[None, :]), axis=1)
        for (z, y, x), s in zip(zyx[in_core(zyx)], sidx[in_core(zyx)]):
            want.add((int(z), int(y), int(x), int(s)))
    got = set()
    if gpu_zyxr is not None and len(gpu_zyxr):
        g = np.asarray(gpu_zyxr)
        zyx = g[:, :3].astype(np.int64)
        sidx = np.argmin(np.abs(g[:, 3:4] / math.sqrt(3) - sig[None, :]), axis=1)
        for (z, y, x), s in zip(zyx[in_core(zyx)], sidx[in_core(zyx)]):
            got.add((int(z), int(y), int(x), int(s)))
    diff = got ^ want
    near, od = set(), set()
    if diff:
        od_set = set()
        if res.trace is not None:
            od_set = {tuple(int(v) for v in (res.peaks[i, :3] + off)) + (int(res.peaks[i, 3]),)
                      for i in res.trace.order_dependent}
        cube = res.cube
        for key in diff:
            z, y, x, s = key
            v = float(cube[z - off[0], y - off[1], x - off[2], s])
            if abs(v - thr) < tol:
                near.add(key)
            elif key in od_set:
                od.add(key)
    unexplained = diff - near - od
    tp = len(got & want)
    return {"core": [list(c) for c in core], "region": [list(r) for r in region],
            "gpu": len(got), "oracle": len(want), "matched": tp,
            "f1": 2.0 * tp / max(len(got) + len(want), 1) if (got or want) else 1.0,
            "near_threshold": len(near), "order_dependent": len(od),
            "unexplained": len(unexplained),
            "unexplained_rows": sorted(unexplained)[:8]}


def summarize(results: Sequence[Dict]) -> Dict:
    """Totals over several cores (the bench line's ``config.parity``)."""
    tp = sum(r["matched"] for r in results)
    ng = sum(r["gpu"] for r in results)
    no = sum(r["oracle"] for r in results)
    return {"boxes": len(results), "gpu_blobs": ng, "oracle_blobs": no,
            "f1": 2.0 * tp / max(ng + no, 1) if (ng or no) else 1.0,
            "near_threshold": sum(r["near_threshold"] for r in results),
            "order_dependent": sum(r["order_dependent"] for r in results),
            "unexplained_diff": sum(r["unexplained"] for r in results)}


def _job(args):
    return check_chunk_core(*args)


def check_cores(jobs, processes: int = 1):
    """``check_chunk_core(*job)`` for every job, in a fork pool when ``processes`` > 1
    (the boxes must then be fetched arrays, not callables that touch a device)."""
    if processes <= 1 or len(jobs) <= 1:
        return [_job(j) for j in jobs]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(processes=min(processes, len(jobs))) as pool:
        return pool.map(_job, jobs, chunksize=1)


def default_cores(blocks, shape, size: int = 48):
    """Cores that exercise what differs between chunks of a stack: the chunk at the
    volume origin (three volume faces), a chunk-interior box, a box on a chunk's high
    faces next to the seams, and a box in the last (often thin) chunk of the grid.
    Returns ``[(chunk coord, chunk shape, core)]``."""
    grid = blocks.sub_roi_slices.shape
    out = []

    def add(coord, place):
        sl = blocks.sub_roi_slices[coord]
        cs = tuple(s.stop - s.start for s in sl)
        core = []
        for n, where in zip(cs, place):
            w = min(size, n)
            lo = {"lo": 0, "mid": max(0, (n - w) // 2), "hi": n - w}[where]
            core.append((lo, lo + w))
        out.append((tuple(int(c) for c in coord), cs, tuple(core)))
    add((0, 0, 0), ("lo", "lo", "lo"))
    add((0, 0, 0), ("mid", "mid", "mid"))
    add((0, min(1, grid[1] - 1), min(1, grid[2] - 1)), ("hi", "hi", "hi"))
    last = tuple(g - 1 for g in grid)
    add(last, ("lo", "mid", "lo"))
    add((last[0], 0, min(1, grid[2] - 1)), ("lo", "hi", "mid"))
    return out


def check_stack(fetch_abs, shape, prof: mm.Profile, resolution, near_max: float, seg_rois,
                final_blobs, processes: int = 1, cores=None, core_size: int = 48) -> Dict:
    """Full-size parity of one chunked stack against the oracle.

    Args:
        fetch_abs: ``fetch_abs(z0, z1, y0, y1, x0, x1)`` -> raw voxels of that box of
            the stack (numpy).
        seg_rois: the GPU's per-chunk blob tables before seam pruning (object array
            shaped like the chunk grid, absolute coordinates; the layout of
            ``StackDetector.detect_blobs_sub_rois``).
        final_blobs: the GPU's final ``(N, 8)`` table of the same stack.

    Two checks: (1) detection inside several cores recomputed by the oracle from the
    raw voxels (``check_chunk_core``); (2) the oracle's ``merge_blobs`` ->
    ``prune_blobs_mp`` -> final layout applied to the GPU's per-chunk tables must
    reproduce the GPU's final table row for row (seam pruning at full size).
    """
    blocks = mm.setup_blocks(prof, shape, resolution)
    block = blocks.denoise_max_shape
    jobs, meta = [], []
    for coord, cs, core in (cores if cores is not None else default_cores(blocks, shape,
                                                                          core_size)):
        sl = blocks.sub_roi_slices[coord]
        o = [s.start for s in sl]
        region = plan_region(core, cs, prof, resolution, block)
        raw = fetch_abs(o[0] + region[0][0], o[0] + region[0][1], o[1] + region[1][0],
                        o[1] + region[1][1], o[2] + region[2][0], o[2] + region[2][1])
        tab = seg_rois[coord]
        rel = None
        if tab is not None and len(tab):
            rel = np.array(tab[:, :4], dtype=np.float64)
            rel[:, :3] -= np.asarray(o, dtype=np.float64)
        jobs.append((np.ascontiguousarray(raw), cs, core, rel, prof, resolution, near_max, block))
        meta.append(coord)
    results = check_cores(jobs, processes)
    for r, coord in zip(results, meta):
        r["chunk"] = list(coord)
    out = summarize(results)
    out["box_results"] = results
    pruned = mm.prune_blobs_mp(shape, seg_rois, blocks.overlap, blocks.tol,
                               blocks.sub_roi_slices, blocks.sub_rois_offsets, (0,),
                               blocks.overlap_padding)
    want = mm.finalize_blobs(pruned[0] if isinstance(pruned, tuple) else pruned)
    n_want = 0 if want is None else len(want)
    n_got = 0 if final_blobs is None else len(final_blobs)
    if n_want == n_got and n_want:
        bad = int(np.count_nonzero(np.any(np.asarray(final_blobs) != want, axis=1)))
    else:
        bad = abs(n_want - n_got) if n_want != n_got else 0
    out["seam_rows_oracle"] = n_want
    out["seam_rows_gpu"] = n_got
    out["seam_rows_differing"] = bad
    return out
