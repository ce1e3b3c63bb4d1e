"""Blob detection over several GPUs of one box: one process per GPU,
``torch.distributed`` (NCCL over NVLink on the GPUs, gloo in the CPU tests).

The reference spreads sub-ROIs over a ``multiprocessing.Pool``
(``magmap/cv/stack_detect.py:175-257``, ``chunking.py:143-167``); here the same
sub-ROIs are spread over GPUs.  Two shardings:

* **chunk-faithful z-slabs** (``detect_blobs_blocks_slabs``): the volume lives
  as contiguous z-slabs, one per rank.  The reference's chunk grid
  (``chunking.stack_splitter``) is laid over the WHOLE volume; contiguous runs of
  chunk z-rows are dealt to the ranks so that the largest run is as small as
  possible, the planes a run needs from other slabs (the 5-voxel overlap, and
  whatever the 500-voxel chunk pitch leaves on the far side of a slab face)
  arrive by ``send/recv`` between slab neighbours, chunks run independently, the per-chunk tables are gathered
  to rank 0 (sizes, then payload) and ``StackPruner.prune_blobs_mp`` removes
  the seam duplicates there.  Row for row equal to the single-GPU result.
* **seamless z-slabs** (``detect_seamless``): the volume is one chunk.  Each
  rank filters its slab extended by a halo of ``halo_planes`` (>= r_max + 1
  planes of real neighbour data, rounded to whole preprocessing block layers),
  keeps the local maxima of the planes it owns, the candidates are gathered to
  rank 0 and ``_prune_blobs`` runs once over all of them.  Equal to one chunk
  covering the whole volume.

Nothing here touches voxel arithmetic: that is the CUDA library's.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

Range = Tuple[int, int]


# ----------------------------------------------------------------------------
# geometry (pure integer host logic; covered by the CPU tests)
# ----------------------------------------------------------------------------

def slab_bounds(n_planes: int, world: int, align: int = 1) -> List[Range]:
    """Contiguous z-slabs [z0, z1), one per rank, faces on multiples of
    ``align`` (the preprocessing block depth in seamless mode); trailing ranks
    may be empty when there are fewer aligned layers than ranks."""
    n_layers = -(-int(n_planes) // int(align))
    cuts = [min(int(n_planes), (n_layers * r // world) * align) for r in range(world)]
    cuts.append(int(n_planes))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def owner_of(z: int, held: Sequence[Range]) -> int:
    for r, (a, b) in enumerate(held):
        if a <= z < b:
            return r
    raise ValueError(f"plane {z} is held by no rank: {held}")


def assign_chunk_rows(z_starts: Sequence[int], held: Sequence[Range]) -> List[List[int]]:
    """Chunk z-row ``k`` (first plane ``z_starts[k]``) goes to the rank whose
    slab holds that plane, so a rank fetches planes only from slabs after its
    own."""
    rows: List[List[int]] = [[] for _ in held]
    for k, z in enumerate(z_starts):
        rows[owner_of(int(z), held)].append(k)
    return rows


def assign_chunk_rows_balanced(z_bounds: Sequence[Range], world: int) -> List[List[int]]:
    """Contiguous runs of chunk z-rows, one run per rank, minimising the largest
    number of planes any rank filters (the chunk pitch rarely divides the slab
    depth: 512-plane slabs against 500-plane chunks would otherwise leave the last
    rank with the 24-plane remainder and the first with two full rows).  Runs are
    contiguous and in rank order so halo planes only ever travel between slab
    neighbours."""
    n = len(z_bounds)
    cost = [b - a for a, b in z_bounds]
    pre = [0]
    for c in cost:
        pre.append(pre[-1] + c)
    INF = float("inf")
    # best[k][i] = minimal achievable maximum when the first i rows go to k ranks
    best = [[INF] * (n + 1) for _ in range(world + 1)]
    cut = [[0] * (n + 1) for _ in range(world + 1)]
    best[0][0] = 0
    for k in range(1, world + 1):
        for i in range(0, n + 1):
            for j in range(0, i + 1):
                if best[k - 1][j] == INF:
                    continue
                v = max(best[k - 1][j], pre[i] - pre[j])
                if v < best[k][i] or (v == best[k][i] and j > cut[k][i]):
                    best[k][i], cut[k][i] = v, j
    rows: List[List[int]] = [[] for _ in range(world)]
    i = n
    for k in range(world, 0, -1):
        j = cut[k][i]
        rows[k - 1] = list(range(j, i))
        i = j
    return rows


def wanted_range(rows: Sequence[int], z_bounds: Sequence[Range], held: Range) -> Range:
    """Planes a rank must see: its own slab plus every plane of its chunk rows."""
    lo, hi = held
    for k in rows:
        lo, hi = min(lo, z_bounds[k][0]), max(hi, z_bounds[k][1])
    return lo, hi


def transfer_plan(held: Sequence[Range], wanted: Sequence[Range]) -> List[Tuple[int, int, int, int]]:
    """Every (src, dst, z0, z1) with src != dst: planes [z0, z1) held by ``src``
    that ``dst`` wants.  Deterministic order, identical on every rank."""
    plan = []
    for dst, (w0, w1) in enumerate(wanted):
        for src, (h0, h1) in enumerate(held):
            if src == dst:
                continue
            z0, z1 = max(w0, h0), min(w1, h1)
            if z0 < z1:
                plan.append((src, dst, z0, z1))
    return plan


# ----------------------------------------------------------------------------
# collectives
# ----------------------------------------------------------------------------

def _world(group=None) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _comm_device(group=None) -> torch.device:
    """NCCL moves device memory, gloo host memory."""
    if dist.is_initialized() and dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def exchange_planes(local: torch.Tensor, held: Sequence[Range], wanted: Sequence[Range],
                    group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Halo exchange.  ``local`` holds planes ``held[rank]`` of the volume
    (dim 0 = z); returns a tensor holding ``wanted[rank]`` (a superset), the
    missing planes received from the ranks that hold them with grouped
    ``isend/irecv`` (NCCL send/recv over NVLink on the GPUs).  ``out``: a
    preallocated tensor for ``wanted[rank]``; when ``local`` is already the view of
    ``out`` that its planes belong to (a slab generated or loaded in place, with room
    for the halo around it) nothing is copied - at whole-brain size a second copy of
    the slab would not fit next to the first."""
    rank, world = _world(group)
    h0, h1 = held[rank]
    w0, w1 = wanted[rank]
    if local.shape[0] != h1 - h0:
        raise ValueError(f"rank {rank} holds {local.shape[0]} planes, expected {h1 - h0}")
    if out is not None and out.shape[0] != max(0, w1 - w0):
        raise ValueError(f"out holds {out.shape[0]} planes, expected {w1 - w0}")
    if (w0, w1) == (h0, h1) and out is None:
        ext = local
    else:
        ext = out if out is not None else torch.empty(
            (max(0, w1 - w0),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        a, b = max(w0, h0), min(w1, h1)          # the part of the wanted range held here
        if a < b:
            dst = ext[a - w0:b - w0]
            if dst.data_ptr() != local[a - h0:b - h0].data_ptr():
                dst.copy_(local[a - h0:b - h0])
    if world == 1:
        if not (h0 <= w0 and w1 <= h1) and w1 > w0:
            raise ValueError(f"planes {wanted[rank]} are not all held ({held[rank]})")
        return ext
    ops, keep = [], []
    for src, dst, z0, z1 in transfer_plan(held, wanted):
        if src == rank:
            # planes travel as raw bytes: NCCL has no 16-bit integer type
            buf = local[z0 - h0:z1 - h0].contiguous().view(torch.uint8)
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, dist.get_global_rank(group, dst)
                                  if group is not None else dst, group))
        elif dst == rank:
            view = ext[z0 - w0:z1 - w0].view(torch.uint8)      # contiguous: whole planes
            ops.append(dist.P2POp(dist.irecv, view, dist.get_global_rank(group, src)
                                  if group is not None else src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return ext


def exchange_edges(host_slab: np.ndarray, held: Sequence[Range], wanted: Sequence[Range],
                   group=None, device=None):
    """Halo exchange for a slab that still lives in HOST memory: only the planes other
    ranks want are uploaded and sent; returns ``(prefix, own, suffix)`` where
    ``prefix`` / ``suffix`` are device tensors with the wanted planes below / above
    this rank's slab and ``own`` is the host view of the wanted part of the slab
    itself (to be streamed by ``gpu.StripFeeder``)."""
    rank, world = _world(group)
    dev = device or torch.device("cuda", torch.cuda.current_device())
    h0, h1 = held[rank]
    w0, w1 = wanted[rank]
    if host_slab.shape[0] != h1 - h0:
        raise ValueError(f"rank {rank} holds {host_slab.shape[0]} planes, expected {h1 - h0}")
    as_t = (lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a))
    tail = tuple(host_slab.shape[1:])
    tdtype = as_t(host_slab[:0]).dtype
    n_pre = max(0, min(w1, h0) - w0)
    n_suf = max(0, w1 - max(w0, h1))
    prefix = torch.empty((n_pre,) + tail, dtype=tdtype, device=dev)
    suffix = torch.empty((n_suf,) + tail, dtype=tdtype, device=dev)
    a, b = max(w0, h0), min(w1, h1)
    own = host_slab[max(0, a - h0):max(0, b - h0)] if a < b else host_slab[:0]
    if world > 1:
        ops, keep = [], []
        g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        for src, dst, z0, z1 in transfer_plan(held, wanted):
            if src == rank:
                buf = as_t(np.ascontiguousarray(host_slab[z0 - h0:z1 - h0])).to(dev).view(torch.uint8)
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, g(dst), group))
            elif dst == rank:
                if z1 <= h0:
                    view = prefix[z0 - w0:z1 - w0]
                else:
                    view = suffix[z0 - max(w0, h1):z1 - max(w0, h1)]
                ops.append(dist.P2POp(dist.irecv, view.view(torch.uint8), g(src), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    elif n_pre or n_suf:
        raise ValueError(f"planes {wanted[rank]} are not all held ({held[rank]})")
    return prefix, own, suffix


def gather_rows(rows: Optional[np.ndarray], n_cols: int, group=None, dst: int = 0,
                dtype=np.float64) -> Optional[List[np.ndarray]]:
    """Variable-length gather of ``(n_r, n_cols)`` tables to ``dst``: row counts
    first (``all_gather``), then each payload with ``send/recv``.  Returns the
    list of per-rank tables on ``dst`` and None elsewhere."""
    rank, world = _world(group)
    mine = np.zeros((0, n_cols), dtype=dtype) if rows is None else np.ascontiguousarray(
        rows, dtype=dtype).reshape(-1, n_cols)
    if world == 1:
        return [mine]
    dev = _comm_device(group)
    tdtype = torch.from_numpy(np.zeros(1, dtype=dtype)).dtype
    count = torch.tensor([mine.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    if rank == dst:
        out = []
        for r in range(world):
            if r == rank:
                out.append(mine)
            elif counts[r] == 0:
                out.append(np.zeros((0, n_cols), dtype=dtype))
            else:
                buf = torch.empty((counts[r], n_cols), dtype=tdtype, device=dev)
                dist.recv(buf, src=g(r), group=group)
                out.append(buf.cpu().numpy())
        return out
    if mine.shape[0]:
        dist.send(torch.from_numpy(mine).to(dev), dst=g(dst), group=group)
    return None


def gather_tensor_rows(rows: Optional[torch.Tensor], n_cols: int, group=None, dst: int = 0,
                       dtype=torch.float64, device=None) -> Optional[List[torch.Tensor]]:
    """``gather_rows`` for tables that already live on the communication device
    (CUDA tensors under NCCL): no host staging on either side.  Row counts travel by
    ``all_gather``; the payloads are ONE batch of point-to-point operations
    (``batch_isend_irecv``: every receive of ``dst`` is posted at once, so the
    transfers of all ranks overlap on NVLink instead of queueing rank by rank)."""
    rank, world = _world(group)
    dev = device if device is not None else (rows.device if rows is not None else
                                             _comm_device(group))
    mine = torch.zeros((0, n_cols), dtype=dtype, device=dev) if rows is None else \
        rows.to(dtype).reshape(-1, n_cols).contiguous()
    if world == 1:
        return [mine]
    count = torch.tensor([mine.shape[0]], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, count, group=group)
    counts = [int(c) for c in counts.cpu()]
    g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    if rank == dst:
        out, ops = [], []
        for r in range(world):
            if r == rank:
                out.append(mine)
                continue
            buf = torch.empty((counts[r], n_cols), dtype=dtype, device=dev)
            out.append(buf)
            if counts[r]:
                ops.append(dist.P2POp(dist.irecv, buf, g(r), group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out
    if mine.shape[0]:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, g(dst), group)]):
            req.wait()
    return None


def pack_tables(seg_rois: np.ndarray) -> Optional[np.ndarray]:
    """Flatten a chunk-grid object array of blob tables into one table with the
    chunk coordinate in three trailing columns (``chunking.merge_blobs``)."""
    from .cv import chunking
    return chunking.merge_blobs(seg_rois)


def unpack_tables(merged_parts: Sequence[np.ndarray], grid_shape: Sequence[int]) -> np.ndarray:
    """Inverse of ``pack_tables`` over the tables of every rank."""
    seg_rois = np.empty(tuple(grid_shape), dtype=object)
    for part in merged_parts:
        if part is None or len(part) == 0:
            continue
        tags = part[:, -3:].astype(np.int64)
        # rows of one chunk are contiguous in a packed table
        change = np.flatnonzero(np.any(np.diff(tags, axis=0) != 0, axis=1)) + 1
        for a, b in zip(np.concatenate(([0], change)), np.concatenate((change, [len(part)]))):
            seg_rois[tuple(tags[a])] = part[a:b, :-3]
    return seg_rois


# ----------------------------------------------------------------------------
# chunk-faithful z-slabs
# ----------------------------------------------------------------------------

def chunk_row_plan(global_shape: Sequence[int], blocks, held: Sequence[Range]):
    """(rows per rank, z extent of every chunk row, wanted plane range per rank)."""
    grid = blocks.sub_roi_slices.shape
    z_bounds = [(blocks.sub_roi_slices[k, 0, 0][0].start, blocks.sub_roi_slices[k, 0, 0][0].stop)
                for k in range(grid[0])]
    rows = assign_chunk_rows_balanced(z_bounds, len(held))
    # a rank asks only for the planes of its chunk rows (not for its whole slab)
    wanted = []
    for r in range(len(held)):
        if rows[r]:
            wanted.append((min(z_bounds[k][0] for k in rows[r]),
                           max(z_bounds[k][1] for k in rows[r])))
        else:
            wanted.append((held[r][0], held[r][0]))
    return rows, z_bounds, wanted


#: fixed cost of one chunk in voxel equivalents (launch gaps and kernel tails of about
#: 45 launches): on a B200 the 34 thin chunks of config 2 (86 MVoxel) take 47 ms where
#: 0.141 ms/MVoxel would predict 12 ms, i.e. about 1 ms = 7 MVoxel per chunk
CHUNK_OVERHEAD_VOXELS = 7.0e6


def loan_units(rows: Sequence[Sequence[int]], z_bounds: Sequence[Range],
               y_bounds: Sequence[Range], x_bounds: Optional[Sequence[Range]] = None
               ) -> List[Tuple[int, int, int, int]]:
    """Even out the row-wise dealing by lending single (chunk z-row k, chunk y-column j)
    units - the five or so chunks of one row that share a y range - from the most to
    the least loaded rank while that lowers the maximum load.  ceil(N*512/500) chunk
    rows never split evenly over N ranks (601 vs 505 planes at N = 8); whole rows keep
    the plane traffic between slab neighbours, and only the odd row travels as
    sub-boxes.  The load of a unit is its voxel count plus, when ``x_bounds`` names its
    chunks, ``CHUNK_OVERHEAD_VOXELS`` per chunk.  Returns ``(k, j, owner, worker)``
    tuples, deterministic."""
    world = len(rows)

    def weight(k, j):
        area = (z_bounds[k][1] - z_bounds[k][0]) * (y_bounds[j][1] - y_bounds[j][0])
        if x_bounds is None:
            return float(area)
        return sum(area * (b - a) + CHUNK_OVERHEAD_VOXELS for a, b in x_bounds)
    units = {r: [(k, j) for k in rows[r] for j in range(len(y_bounds))] for r in range(world)}
    owner = {u: r for r in range(world) for u in units[r]}
    load = [sum(weight(*u) for u in units[r]) for r in range(world)]
    loans = []
    for _ in range(4 * world * len(y_bounds)):
        hi = max(range(world), key=lambda r: (load[r], -r))
        lo = min(range(world), key=lambda r: (load[r], r))
        gap = load[hi] - load[lo]
        # the unit whose move brings the two loads closest; moving w helps iff w < gap
        cands = [u for u in units[hi] if weight(*u) < gap]
        if not cands:
            break
        u = min(cands, key=lambda u: (abs(gap - 2 * weight(*u)), u))
        units[hi].remove(u)
        units[lo].append(u)
        load[hi] -= weight(*u)
        load[lo] += weight(*u)
        loans = [l for l in loans if (l[0], l[1]) != u]
        if owner[u] != lo:
            loans.append((u[0], u[1], owner[u], lo))
    return sorted(loans)


def box_transfer_plan(loans, z_bounds, y_bounds, held):
    """``(src, dst, k, j, z0, z1)`` for every piece of a lent unit's box: the planes
    [z0, z1) of chunk row k held by ``src``, restricted to chunk column j's y range,
    needed by worker ``dst`` (src == dst: the worker holds those planes itself)."""
    plan = []
    for k, j, _, worker in loans:
        for src, (h0, h1) in enumerate(held):
            z0, z1 = max(z_bounds[k][0], h0), min(z_bounds[k][1], h1)
            if z0 < z1:
                plan.append((src, worker, k, j, z0, z1))
    return plan


#: stage times of the last ``detect_blobs_blocks_slabs`` call of this rank, in seconds, when
#: the environment sets MMB_STAGE_SYNC (every stage then ends with a device synchronisation,
#: which the normal path avoids): plane exchange, box exchange, detection, gather, tables
LAST_STAGE_S: Dict[str, float] = {}


def _stage(name: str, t0: float) -> float:
    import os
    import time
    if os.environ.get("MMB_STAGE_SYNC"):
        torch.cuda.synchronize()
        now = time.perf_counter()
        LAST_STAGE_S[name] = LAST_STAGE_S.get(name, 0.0) + now - t0
        return now
    return t0


def detect_blobs_blocks_slabs(filename_base: str, slab, held: Sequence[Range],
                              global_shape: Sequence[int],
                              channels: Optional[Sequence[int]] = None, group=None,
                              save_dfs: bool = False, balance_units: bool = True):
    """``stack_detect.detect_blobs_blocks`` over a volume sharded as z-slabs.

    Args:
        slab: this rank's planes ``held[rank]`` of the (z, y, x[, c]) volume: a
            CUDA tensor (uint16 as int16 bits is accepted like everywhere), or a
            C-contiguous HOST array, which is then streamed to the device strip by
            strip under the kernels (only neighbours' halo planes move up front).
        held: the slab [z0, z1) of every rank, in rank order.
        global_shape: (Z, Y, X) of the whole volume.

    Returns ``(stats, fdbk, Blobs)`` on rank 0 (identical to the single-GPU
    call on the whole volume) and ``(None, None, None)`` on the other ranks.
    """
    from .cv import detector, stack_detect
    from .plot import plot_3d
    from .settings import config
    from .io import libmag, np_io
    rank, world = _world(group)
    if channels is None:
        _, channels = plot_3d.setup_channels(slab, channels, 3)
    settings = config.get_roi_profile(channels[0])
    shape = tuple(int(v) for v in global_shape[:3]) + tuple(slab.shape[3:])
    blocks = stack_detect.setup_blocks(settings, shape)
    rows, z_bounds, wanted = chunk_row_plan(shape, blocks, held)
    grid = blocks.sub_roi_slices.shape
    y_bounds = [(blocks.sub_roi_slices[0, j, 0][1].start, blocks.sub_roi_slices[0, j, 0][1].stop)
                for j in range(grid[1])]
    x_bounds = [(blocks.sub_roi_slices[0, 0, i][2].start, blocks.sub_roi_slices[0, 0, i][2].stop)
                for i in range(grid[2])]
    loans = loan_units(rows, z_bounds, y_bounds, x_bounds) if balance_units else []
    lent = {(k, j) for k, j, _, _ in loans}
    host_slab = isinstance(slab, np.ndarray)
    device_route = (blocks.exclude_border is None and stack_detect.DEVICE_TABLES
                    and settings["isotropic"] is None)
    streamed = host_slab and slab.flags.c_contiguous and device_route
    prefix = suffix = None
    import time as _time
    LAST_STAGE_S.clear()
    _t = _time.perf_counter()
    if streamed:
        # the slab stays on the host and is streamed strip by strip under the kernels;
        # only the halo planes of the neighbours are exchanged up front
        prefix, ext, suffix = exchange_edges(slab, held, wanted, group)
    else:
        if host_slab:
            slab = torch.from_numpy(np.ascontiguousarray(
                slab.view(np.int16) if slab.dtype == np.uint16 else slab)).cuda()
        ext = exchange_planes(slab, held, wanted, group)
    w0 = wanted[rank][0]
    final_on_device = False

    local_slices = np.empty(grid, dtype=object)
    coords = []
    for k in rows[rank]:
        for j in range(grid[1]):
            if (k, j) in lent:
                continue                      # worked on by another rank (loan_units)
            for i in range(grid[2]):
                sz, sy, sx = blocks.sub_roi_slices[k, j, i]
                local_slices[k, j, i] = (slice(sz.start - w0, sz.stop - w0), sy, sx)
                coords.append((k, j, i))
    # cells of other ranks still need a shape for the workspace sizing
    for c in np.ndindex(*grid):
        if local_slices[c] is None:
            local_slices[c] = (slice(0, 0), slice(0, 0), slice(0, 0))
    if device_route:
        # device-resident tables: the gather moves CUDA tensors over NVLink and rank 0
        # prunes the seams on its GPU
        from .cv import device_tables
        dev = torch.device("cuda", torch.cuda.current_device()) if host_slab else slab.device
        _t = _stage("exchange_planes", _t)
        boxes = _exchange_boxes(slab, held, loans, z_bounds, y_bounds, group, dev)
        _t = _stage("exchange_boxes", _t)
        tables = device_tables.ChunkTables(dev, channels)
        if coords:
            stack_detect.StackDetector.detect_blobs_sub_rois_device(
                ext, local_slices, blocks.sub_rois_offsets, blocks.denoise_max_shape,
                channels, coords=coords, prefix=prefix, suffix=suffix, tables=tables)
        for (k, j), box in boxes.items():
            # a lent unit: the chunks (k, j, :) cut from their own small box
            box_slices = np.empty(grid, dtype=object)
            for c in np.ndindex(*grid):
                box_slices[c] = (slice(0, 0), slice(0, 0), slice(0, 0))
            unit = []
            for i in range(grid[2]):
                sz, sy, sx = blocks.sub_roi_slices[k, j, i]
                box_slices[k, j, i] = (slice(0, sz.stop - sz.start), slice(0, sy.stop - sy.start), sx)
                unit.append((k, j, i))
            stack_detect.StackDetector.detect_blobs_sub_rois_device(
                box, box_slices, blocks.sub_rois_offsets, blocks.denoise_max_shape, channels,
                coords=unit, tables=tables)
        # 32-byte rows travel; rank 0 orders them by chunk (whoever worked on it), prunes
        # the seams and formats the table with the library's table kernels
        _t = _stage("detect", _t)
        parts = gather_tensor_rows(tables.rows(), device_tables.ROW_INTS, group,
                                   dtype=torch.int32, device=dev)
        _t = _stage("gather", _t)
        if rank != 0:
            return None, None, None
        allr = parts[0] if len(parts) == 1 else torch.cat(parts)
        segments_all, df_pruning = device_tables.prune_rows(
            allr, stack_detect.channel_ladders(slab, blocks.denoise_max_shape, channels),
            blocks.overlap, blocks.tol, blocks.sub_roi_slices, channels,
            blocks.overlap_padding, final_layout=True)
        _t = _stage("tables", _t)
        final_on_device = True
    else:
        seg_rois = None
        if coords:
            seg_rois = stack_detect.StackDetector.detect_blobs_sub_rois(
                None, ext, local_slices, blocks.sub_rois_offsets, blocks.denoise_max_shape,
                blocks.exclude_border, False, channels, coords=coords)
        packed = pack_tables(seg_rois) if seg_rois is not None else None
        n_cols = _agree_max(14 if packed is None else packed.shape[1], group)
        parts = gather_rows(packed, n_cols, group)
        if rank != 0:
            return None, None, None
        seg_all = unpack_tables(parts, grid)
        segments_all, df_pruning = stack_detect.StackPruner.prune_blobs_mp(
            None, seg_all, blocks.overlap, blocks.tol, blocks.sub_roi_slices,
            blocks.sub_rois_offsets, channels, blocks.overlap_padding)
    filename_blobs = libmag.combine_paths(filename_base, config.SUFFIX_BLOBS)
    if final_on_device:
        from .cv import device_tables as _dt
        blobs = detector.Blobs(segments_all, path=filename_blobs, cols=list(_dt.FINAL_COLS))
    else:
        blobs = detector.Blobs(segments_all, path=filename_blobs)
    if segments_all is not None and not final_on_device:
        blobs.replace_rel_with_abs_blob_coords(segments_all)
        blobs.blobs = segments_all
        segments_all = blobs.remove_abs_blob_coords(True)
    blobs.blobs = segments_all
    blobs.colocalizations = None
    blobs.resolutions = config.resolutions
    blobs.roi_offset = (0, 0, 0)
    blobs.roi_size = shape[:3]
    if save_dfs and df_pruning is not None and len(df_pruning.columns):
        df_pruning.to_csv("blob_ratios.csv", index=False)
    return None, None, blobs


def _exchange_boxes(slab, held, loans, z_bounds, y_bounds, group, dev) -> Dict[tuple, torch.Tensor]:
    """Move the sub-boxes of lent units to their workers (NCCL send/recv of dense
    copies; from a host slab the pieces go up with one pitched DMA each).  Returns
    ``{(k, j): box}`` for the units this rank works on."""
    rank, world = _world(group)
    if not loans:
        return {}
    host_slab = isinstance(slab, np.ndarray)
    h0 = held[rank][0]
    if host_slab:
        tdtype = torch.int16 if slab.dtype == np.uint16 else torch.from_numpy(slab[:0]).dtype
    else:
        tdtype = slab.dtype
    tail = tuple(slab.shape[2:])
    boxes = {}
    for k, j, _, worker in loans:
        if worker == rank:
            boxes[(k, j)] = torch.empty(
                (z_bounds[k][1] - z_bounds[k][0], y_bounds[j][1] - y_bounds[j][0]) + tail,
                dtype=tdtype, device=dev)

    def piece(z0, z1, y0, y1):
        if not host_slab:
            return slab[z0 - h0:z1 - h0, y0:y1].contiguous()
        from . import gpu, _lib
        import ctypes as C
        out = torch.empty((z1 - z0, y1 - y0) + tail, dtype=tdtype, device=dev)
        row_bytes = int(np.prod(tail)) * slab.itemsize
        src = slab.ctypes.data + ((z0 - h0) * slab.shape[1] + y0) * row_bytes
        _lib.check(_lib.load().mmb_upload_pieces(
            C.c_void_p(out.data_ptr()), C.c_void_p(src), z1 - z0, (y1 - y0) * row_bytes,
            slab.shape[1] * row_bytes, gpu._stream()))
        return out

    g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    ops, keep = [], []
    for src, dst, k, j, z0, z1 in box_transfer_plan(loans, z_bounds, y_bounds, held):
        y0, y1 = y_bounds[j]
        if src == rank and dst == rank:
            boxes[(k, j)][z0 - z_bounds[k][0]:z1 - z_bounds[k][0]].copy_(piece(z0, z1, y0, y1))
        elif src == rank:
            buf = piece(z0, z1, y0, y1).view(torch.uint8)
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, g(dst), group))
        elif dst == rank:
            view = boxes[(k, j)][z0 - z_bounds[k][0]:z1 - z_bounds[k][0]]   # whole planes of the box
            ops.append(dist.P2POp(dist.irecv, view.view(torch.uint8), g(src), group))
    if ops:
        if host_slab:
            torch.cuda.current_stream().synchronize()      # uploads done before NCCL reads them
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return boxes


def _agree_max(v: int, group=None) -> int:
    rank, world = _world(group)
    if world == 1:
        return int(v)
    t = torch.tensor([int(v)], dtype=torch.int64, device=_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


# ----------------------------------------------------------------------------
# seamless z-slabs
# ----------------------------------------------------------------------------

def seamless_plan(n_planes: int, world: int, block_depth: int, halo_planes: int):
    """Owned slabs with faces on multiples of ``block_depth`` and the extended
    range each rank filters: the halo is rounded up to whole block layers so the
    25^3 preprocessing blocks of every rank coincide with those of one chunk
    anchored at the volume origin."""
    own = slab_bounds(n_planes, world, block_depth)
    layers = -(-int(halo_planes) // int(block_depth)) * int(block_depth)
    ext = [(max(0, a - layers), min(int(n_planes), b + layers)) if b > a else (a, b)
           for a, b in own]
    return own, ext


def _seamless_setup(global_shape, channel, image_is_f32: bool = False):
    from .cv import detector, stack_detect
    from .plot import plot_3d
    from .settings import config
    settings = config.get_roi_profile(channel)
    Z, Y, X = (int(v) for v in global_shape[:3])
    blocks = stack_detect.setup_blocks(settings, (Z, Y, X))
    scale = detector.calc_scaling_factor()[2]
    dms = blocks.denoise_max_shape
    pre = plot_3d.preproc_params(settings, channel) if dms is not None else None
    # a raw float32 image gets scikit-image's float32-rounded ladder; preprocessing yields
    # float64 in the reference
    sigmas = detector.sigma_ladder(settings, scale, image_is_f32 and dms is None)
    halo = int(4.0 * float(np.max(sigmas)) + 0.5) + 1          # r_max + 1
    bd = (int(dms[0]), int(dms[1]), int(dms[2])) if dms is not None else (1, 1, 1)
    return settings, pre, sigmas, halo, bd


def seamless_candidates(ext, ext_range: Range, own_range: Range, global_shape: Sequence[int],
                        channel: int = 0, tile_yx: Optional[Sequence[int]] = None,
                        capacity: Optional[int] = None) -> torch.Tensor:
    """The local part of ``detect_seamless``: local maxima (no pruning) of the
    planes ``own_range`` given the planes ``ext_range`` (own + halo) of the volume,
    in GLOBAL coordinates, as an ``(n, 5)`` int32 CUDA tensor of ``mmb_cand`` records
    (arbitrary order).  y and x - and z, with a ``(z, y, x)`` triple - are tiled with the
    same halo when ``tile_yx`` is given;
    every tile's chunk is enqueued asynchronously and its owned candidates are appended
    to one device list (``mmb_cands_append``), so nothing but three counters per tile
    crosses to the host."""
    import ctypes as C
    from collections import deque
    from . import gpu, _lib
    lib = _lib.load()
    settings, pre, sigmas, halo, bd = _seamless_setup(global_shape, channel,
                                                      ext.dtype == torch.float32)
    Z, Y, X = (int(v) for v in global_shape[:3])
    e0, e1 = ext_range
    z0, z1 = own_range
    dev = ext.device
    if z1 <= z0:
        return torch.zeros((0, 5), dtype=torch.int32, device=dev)
    if e0 % bd[0] != 0 or (e1 % bd[0] != 0 and e1 != Z):
        raise ValueError(f"extended range {ext_range} is not aligned to the block depth {bd[0]}")
    if (z0 - e0 < halo and e0 > 0) or (e1 - z1 < halo and e1 < Z):
        raise ValueError(f"halo of {ext_range} around {own_range} is thinner than {halo} planes")
    in_scale = 1.0
    if pre is None:
        probe = gpu.as_source(ext, channel if ext.dim() == 4 else None)
        in_scale = {gpu._lib.MMB_U8: 1 / 255.0, gpu._lib.MMB_U16: 1 / 65535.0}.get(probe.dtype, 1.0)
    # tiles: (y, x) or (z, y, x) voxels of OWNED volume per tile, rounded up to whole
    # preprocessing blocks; every tile is filtered with a halo of whole block layers
    tz = z1 - z0
    if tile_yx is None:
        ty, tx = Y, X
    elif len(tile_yx) == 3:
        tz, ty, tx = (int(v) for v in tile_yx)
    else:
        ty, tx = int(tile_yx[0]), int(tile_yx[1])
    tz = -(-tz // bd[0]) * bd[0]
    ty = -(-ty // bd[1]) * bd[1]
    tx = -(-tx // bd[2]) * bd[2]
    hz = -(-halo // bd[0]) * bd[0]
    hy = -(-halo // bd[1]) * bd[1]
    hx = -(-halo // bd[2]) * bd[2]
    det = gpu.ChunkDetector((min(e1 - e0, tz + 2 * hz), min(Y, ty + 2 * hy),
                             min(X, tx + 2 * hx)))
    own_vox = (z1 - z0) * Y * X
    cap = int(capacity) if capacity else max(1 << 16, own_vox // 512)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    while True:
        out = torch.empty((cap, 5), dtype=torch.int32, device=dev)
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        pending = deque()

        def finish_oldest():
            ticket, (za, ya, xa, t0, t1, y0, x0) = pending.popleft()
            got, _ = det.collect_device(ticket)
            n = int(got.shape[0])
            if n:
                _lib.check(lib.mmb_cands_append(
                    C.c_void_p(got.data_ptr()), n, None, _lib._I32x3(za, ya, xa),
                    _lib._I32x3(t0, y0, x0),
                    _lib._I32x3(t1, min(Y, y0 + ty), min(X, x0 + tx)),
                    C.c_void_p(out.data_ptr()), cap, C.c_void_p(counter.data_ptr()), stream))

        for t0 in range(z0, z1, tz):
            t1 = min(z1, t0 + tz)
            za, zb = max(e0, t0 - hz), min(e1, t1 + hz)
            for y0 in range(0, Y, ty):
                for x0 in range(0, X, tx):
                    ya, yb = max(0, y0 - hy), min(Y, y0 + ty + hy)
                    xa, xb = max(0, x0 - hx), min(X, x0 + tx + hx)
                    src = gpu.as_source(ext[za - e0:zb - e0, ya:yb, xa:xb],
                                        channel if ext.dim() == 4 else None)
                    while pending and det.free_slots() < 1:
                        finish_oldest()
                    # overlap 1.0 = no pruning here: _prune_blobs runs once over all slabs
                    pending.append((det.enqueue(src, sigmas, settings["detection_threshold"],
                                                1.0, scale=in_scale, pre=pre, block_shape=bd,
                                                z_lo=t0 - za, z_hi=t1 - za),
                                    (za, ya, xa, t0, t1, y0, x0)))
        while pending:
            finish_oldest()
        n = int(counter.item())
        if n <= cap:
            return out[:n]
        cap = int(n * 1.1) + 1024          # the list overflowed: redo with room for all


def detect_seamless(slab, held: Sequence[Range], global_shape: Sequence[int],
                    channel: int = 0, group=None, tile_yx: Optional[Sequence[int]] = None,
                    ext_out: Optional[torch.Tensor] = None):
    """Detect blobs as if the whole volume were ONE chunk (no chunk seams).

    Each rank: exchange halo planes, run the fused chunk driver on its extended
    slab without pruning and keep the local maxima of the planes it owns
    (``z_lo``/``z_hi`` of ``mmb_detect_chunk_enqueue``); rank 0: gather the
    candidates (device to device, one batch of NCCL send/recv) and run
    ``_prune_blobs`` once over all of them (``mmb_prune_within``: cell-bucketed
    pair search, any listing order).  ``tile_yx`` additionally tiles y and x inside
    a rank (tile + halo must fit the workspace of eight float volumes).  ``ext_out``:
    preallocated room for this rank's slab plus halo (``seamless_plan``'s extended
    range) of which ``slab`` is already the owned view - see ``exchange_planes``.

    Returns on rank 0 the ``(n, 11)`` blob table of ``detector.detect_blobs`` in
    ``peak_local_max`` order (None if empty), None elsewhere.
    """
    rank, world = _world(group)
    settings, pre, sigmas, halo, bd = _seamless_setup(global_shape, channel,
                                                      slab.dtype == torch.float32)
    Z, Y, X = (int(v) for v in global_shape[:3])
    own, ext_ranges = seamless_plan(Z, world, bd[0], halo)
    # the caller's slabs need not coincide with the block-aligned owned slabs
    ext = exchange_planes(slab, held, ext_ranges, group, out=ext_out)
    mine = seamless_candidates(ext, ext_ranges[rank], own[rank], (Z, Y, X), channel, tile_yx)
    del ext
    parts = gather_tensor_rows(mine, 5, group, dtype=torch.int32, device=mine.device)
    if rank != 0:
        return None
    allc = parts[0] if len(parts) == 1 else torch.cat(parts)
    return prune_global(allc, sigmas, settings["overlap"], (Z, Y, X), channel)


def prune_global(cands, sigmas, overlap: float, shape: Sequence[int], channel: int):
    """``_prune_blobs`` over the candidates of the whole volume (``(n, 5)`` int32 CUDA
    tensor, or ``gpu.CAND_DTYPE`` records on the host), then the blob table of
    ``detector.detect_blobs`` in ``peak_local_max`` order - all on the device."""
    import ctypes as C
    from . import gpu, _lib
    from .cv import device_tables
    if isinstance(cands, np.ndarray):
        cands = gpu.cands_from_numpy(cands)
    n = int(cands.shape[0])
    if n == 0:
        return None
    Z, Y, X = (int(v) for v in shape)
    cands = cands.contiguous()
    keep = gpu.prune_within(cands, n, sigmas, overlap, Y, X)
    kept = torch.empty_like(cands)
    counter = torch.zeros(1, dtype=torch.int32, device=cands.device)
    _lib.check(_lib.load().mmb_cands_append(
        C.c_void_p(cands.data_ptr()), n, C.c_void_p(keep.data_ptr()), _lib._I32x3(0, 0, 0),
        _lib._I32x3(0, 0, 0), _lib._I32x3(Z, Y, X), C.c_void_p(kept.data_ptr()), n,
        C.c_void_p(counter.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return device_tables.cands_to_table(kept[:int(counter.item())], sigmas, (Z, Y, X), channel)
