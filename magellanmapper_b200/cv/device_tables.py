"""Device-resident blob tables: the per-chunk survivors stay in HBM, are merged
and seam-pruned there by the library's table kernels, and reach the host once as
the final table.

Same results, row for row, as the host-side route through
``StackDetector.detect_blobs_sub_rois`` -> ``chunking.merge_blobs`` ->
``StackPruner.prune_blobs_mp`` (``magmap/cv/stack_detect.py:175-257, 680-861``,
``chunking.py:410-445``); ``tests/test_gpu_api.py`` compares the two.  A blob
travels as a 32-byte ``mmb_row`` (``include/mmb200.h``); merge ordering, seam
classification, the box match of ``remove_close_blobs``, the coordinate
averaging and the final column layout are ``mmb_stack_tables``
(``csrc/tables.cu``: hand-written radix sort, counting-sort buckets, match and
gather kernels - no library sort or select, no host synchronisation).  torch only
holds the buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import pandas as pd
import torch

from . import detector
from .. import _lib

N_COLS = 11          # Blobs.Cols
ROW_INTS = 8         # mmb_row as int32 words
_MAX_SEC = 128       # seam_counts pitch of mmb_stack_tables

#: columns of the table ``detect_blobs_blocks`` returns (stack_detect.py:458-467):
#: relative coordinates replaced by the seam-averaged absolute ones, abs columns dropped
FINAL_COLS = ["z", "y", "x", "radius", "confirmed", "truth", "channel", "region"]


class ChunkTables:
    """Survivors of every chunk as ``mmb_row`` records, kept on the device in
    arrival order (the table kernels restore the chunk-grid order)."""

    def __init__(self, device, channels: Sequence[int]):
        self.device = device
        self.channels = [int(c) for c in channels]
        self.parts: List[torch.Tensor] = []          # (n, 8) int32 = mmb_row records
        self.sigmas = {}                             # channel -> ladder

    def append(self, cand: torch.Tensor, chunk: int, sigmas, channel: int) -> None:
        """``cand``: ``(n, 5)`` int32 candidate records of one detection (the
        survivors of ``mmb_detect_chunk_enqueue``); ``chunk``: position of the chunk
        in the C-ordered chunk grid.  Converted on the current stream."""
        self.sigmas.setdefault(int(channel), np.asarray(sigmas, dtype=np.float64))
        n = int(cand.shape[0])
        if n == 0:
            return
        rows = torch.empty((n, ROW_INTS), dtype=torch.int32, device=self.device)
        _lib.check(_lib.load().mmb_rows_from_cands(
            C.c_void_p(cand.data_ptr()), n, int(chunk), self.channels.index(int(channel)),
            C.c_void_p(rows.data_ptr()),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.parts.append(rows)

    def rows(self) -> torch.Tensor:
        if not self.parts:
            return torch.zeros((0, ROW_INTS), dtype=torch.int32, device=self.device)
        return self.parts[0] if len(self.parts) == 1 else torch.cat(self.parts)

    def ladders(self, num_sigma: Optional[int] = None) -> np.ndarray:
        """``(n_channels, num_sigma)`` ladders in channel-list order (zeros where a
        channel found nothing and never reported its ladder)."""
        n_sig = num_sigma or max([len(v) for v in self.sigmas.values()] + [1])
        out = np.zeros((len(self.channels), n_sig))
        for i, c in enumerate(self.channels):
            if c in self.sigmas:
                out[i, :len(self.sigmas[c])] = self.sigmas[c]
        return out


_pinned: List[Optional[torch.Tensor]] = [None]
_pool = [None]
_CHUNK_BYTES = 8 << 20


def _to_host(t: torch.Tensor) -> np.ndarray:
    """Device table -> fresh numpy array through a cached pinned staging buffer (a
    pageable copy of a config-2 table costs more than pruning it).  Tables above a few
    chunks (2 M rows = 132 MB at eight GPUs) are moved as a pipeline: 8 MB pieces cross
    PCIe on the current stream while a small thread pool copies the pieces that have
    landed out of the staging buffer (numpy releases the GIL in large copies), so the
    single-threaded 132 MB host copy - 25 ms, the longest part of rank 0's tail - hides
    behind the transfer."""
    n = t.numel()
    if n == 0:
        return np.zeros(tuple(t.shape), dtype=np.float64)
    buf = _pinned[0]
    if buf is None or buf.numel() < n:
        buf = _pinned[0] = torch.empty(max(n, 1 << 20), dtype=t.dtype).pin_memory()
    flat = t.reshape(-1)
    stage = buf[:n]
    per = max(1, _CHUNK_BYTES // t.element_size())
    if n <= 2 * per:
        stage.copy_(flat, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return stage.view(t.shape).numpy().copy()
    from concurrent.futures import ThreadPoolExecutor
    if _pool[0] is None:
        _pool[0] = ThreadPoolExecutor(max_workers=4)
    out = np.empty(n, dtype=stage.numpy().dtype)
    src = stage.numpy()
    events, jobs = [], []
    for a in range(0, n, per):
        b = min(n, a + per)
        stage[a:b].copy_(flat[a:b], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        events.append((a, b, ev))
    for a, b, ev in events:
        ev.synchronize()
        jobs.append(_pool[0].submit(np.copyto, out[a:b], src[a:b]))
    for j in jobs:
        j.result()
    return out.reshape(tuple(t.shape))


def _axis_sections(sub_roi_slices, axis: int):
    n = sub_roi_slices.shape[axis]
    start = np.zeros(n, dtype=np.int32)
    size = np.zeros(n, dtype=np.int32)
    for j in range(n):
        coord = [0, 0, 0]
        coord[axis] = j
        sl = sub_roi_slices[tuple(coord)][axis]
        start[j], size[j] = sl.start, sl.stop - sl.start
    return start, size


def prune_rows(rows: torch.Tensor, ladders: np.ndarray, overlap, tol, sub_roi_slices,
               channels: Sequence[int], overlap_padding=None, final_layout: bool = False):
    """``chunking.merge_blobs`` + ``StackPruner.prune_blobs_mp`` (+ the final column
    layout of ``detect_blobs_blocks`` with ``final_layout``) on device-resident rows.

    Returns ``(float64 numpy table, DataFrame of pruning ratios)``: ``(N', 8)`` in
    ``FINAL_COLS`` order with ``final_layout`` (what ``Blobs.replace_rel_with_abs_
    blob_coords`` + ``remove_abs_blob_coords(True)`` leave), else the ``(N', 11)``
    ``Blobs.Cols`` table.  The only device-to-host transfer of blob rows is that
    table; ``(None, None)`` when there are no rows."""
    n = int(rows.shape[0])
    if n == 0:
        return None, None
    lib = _lib.load()
    if overlap_padding is None:
        overlap_padding = tol
    dev = rows.device
    rows = rows.contiguous()
    geom = _lib.MmbStackGeom()
    keep = []
    for a in range(3):
        start, size = _axis_sections(sub_roi_slices, a)
        keep += [start, size]
        geom.grid[a] = int(sub_roi_slices.shape[a])
        geom.overlap[a] = int(np.broadcast_to(overlap, (3,))[a])
        geom.tol[a] = int(np.broadcast_to(tol, (3,))[a])
        geom.pad[a] = int(np.broadcast_to(overlap_padding, (3,))[a])
        geom.start[a] = start.ctypes.data_as(C.POINTER(C.c_int32))
        geom.size[a] = size.ctypes.data_as(C.POINTER(C.c_int32))
    lad = np.ascontiguousarray(ladders, dtype=np.float64)
    ids = np.ascontiguousarray([int(c) for c in channels], dtype=np.int32)
    geom.n_channels = len(ids)
    geom.num_sigma = int(lad.shape[1])
    geom.sigmas = lad.ctypes.data_as(C.POINTER(C.c_double))
    geom.channel_ids = ids.ctypes.data_as(C.POINTER(C.c_int32))
    ncols = len(FINAL_COLS) if final_layout else N_COLS
    out = torch.empty((n, ncols), dtype=torch.float64, device=dev)
    n_out = torch.zeros(1, dtype=torch.int32, device=dev)
    counts = torch.zeros((len(ids), 3, _MAX_SEC, 4), dtype=torch.int32, device=dev)
    work = torch.empty(lib.mmb_stack_tables_work_bytes(n), dtype=torch.uint8, device=dev)
    _lib.check(lib.mmb_stack_tables(
        C.c_void_p(rows.data_ptr()), n, C.byref(geom), 1 if final_layout else 0,
        C.c_void_p(out.data_ptr()), C.c_void_p(n_out.data_ptr()), C.c_void_p(counts.data_ptr()),
        C.c_void_p(work.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    n_final = int(n_out.item())                       # the one synchronisation
    table = _to_host(out[:n_final])
    cnt = counts.cpu().numpy()
    cols = ("blobs", "ratio_pruning", "ratio_adjacent")
    ratios_out = {}
    for ci in range(len(ids)):
        for axis in range(3):
            for j in range(int(sub_roi_slices.shape[axis]) - 1):
                in_slab, after, n_next, _ = (int(v) for v in cnt[ci, axis, j])
                r = detector.meas_pruning_ratio(in_slab, after, n_next)
                if r:
                    for c, v in zip(cols, r):
                        ratios_out.setdefault(c, []).append(v)
    return table, pd.DataFrame(ratios_out)


def cands_to_table(cand: torch.Tensor, sigmas, shape_zyx: Sequence[int], channel: int
                   ) -> Optional[np.ndarray]:
    """The ``(n, 11)`` table of ``detector.detect_blobs`` for the survivors of ONE
    detection over a volume of ``shape_zyx`` (``peak_local_max`` order: descending
    response, ties in C order), built by the table kernels."""
    if int(cand.shape[0]) == 0:
        return None
    tables = ChunkTables(cand.device, [channel])
    tables.append(cand, 0, sigmas, channel)
    slices = np.empty((1, 1, 1), dtype=object)
    slices[0, 0, 0] = tuple(slice(0, int(s)) for s in shape_zyx[:3])
    table, _ = prune_rows(tables.rows(), tables.ladders(), (0, 0, 0), (0, 0, 0), slices,
                          [channel])
    return table
