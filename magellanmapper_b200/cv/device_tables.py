"""Device-resident blob tables: the per-chunk survivors stay in HBM, are merged
and seam-pruned there, and reach the host once as the final table.

Same results, row for row, as the host-side route through
``StackDetector.detect_blobs_sub_rois`` -> ``chunking.merge_blobs`` ->
``StackPruner.prune_blobs_mp`` (``magmap/cv/stack_detect.py:175-257, 680-861``,
``chunking.py:410-445``); ``tests/test_gpu_api.py`` compares the two.  What moves
to the device is index bookkeeping (sorting rows into ``peak_local_max`` order,
selecting the blobs of a seam slab, concatenating survivors) done with torch
tensor ops, plus the box match of ``remove_close_blobs`` which is the library's
``mmb_prune_seams`` kernel.  The host route copies a (N, 14) float64 table several
times per seam; at config-2 scale (2.6e5 blobs) that is 50-80 ms of a 430 ms
step and, with several GPUs, serial work on rank 0.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import pandas as pd
import torch

from . import detector

N_COLS = 11          # Blobs.Cols
N_MERGED = 14        # + chunk coordinate (chunking.merge_blobs)


class ChunkTables:
    """Survivors of every chunk, kept on the device in arrival order."""

    def __init__(self, device):
        self.device = device
        self.parts: List[torch.Tensor] = []      # (n, 5) int32 candidate records
        self.meta: List[tuple] = []              # (n, coord, offset, (Y, X), sigmas, channel)

    def append(self, cand: torch.Tensor, coord, offset, shape_yx, sigmas, channel: int,
               grid_rank: Optional[int] = None) -> None:
        """``grid_rank`` = position of the chunk in the C-ordered chunk grid; chunks
        may arrive in any order (strip-wise streaming walks y first), the merged
        table is always in grid order.  Default: arrival order."""
        if cand.shape[0] == 0:
            return
        self.parts.append(cand)
        self.meta.append((int(cand.shape[0]), tuple(int(c) for c in coord),
                          tuple(float(o) for o in offset), (int(shape_yx[0]), int(shape_yx[1])),
                          np.asarray(sigmas, dtype=np.float64), int(channel),
                          len(self.meta) if grid_rank is None else int(grid_rank)))

    def merged(self) -> Optional[torch.Tensor]:
        """(N, 14) float64 device table in the layout and row order of
        ``chunking.merge_blobs`` over ``Blobs.format_blobs`` tables: chunks in
        grid order, channels in request order inside a chunk, rows of one
        detection in ``peak_local_max`` order (descending response, ties in C
        order of (z, y, x, scale))."""
        if not self.parts:
            return None
        dev = self.device
        cand = torch.cat(self.parts)
        counts = torch.tensor([m[0] for m in self.meta], device=dev)
        T = len(self.meta)
        tix = torch.repeat_interleave(torch.arange(T, device=dev), counts)
        n_sig = max(len(m[4]) for m in self.meta)
        sig = np.zeros((T, n_sig))
        for i, m in enumerate(self.meta):
            sig[i, :len(m[4])] = m[4]
        # sort key of a detection: chunk position in the grid, then arrival (channels of
        # one chunk arrive in request order)
        seq = sorted(range(T), key=lambda i: (self.meta[i][6], i))
        place = [0] * T
        for pos, i in enumerate(seq):
            place[i] = pos
        per = torch.tensor(
            [list(m[1]) + list(m[2]) + [m[3][0], m[3][1], len(m[4]), m[5], place[i]]
             for i, m in enumerate(self.meta)],
            dtype=torch.float64, device=dev)                     # (T, 11)
        sig_t = torch.from_numpy(sig).to(dev)
        row = per[tix]
        z, y, x, s = (cand[:, k].long() for k in range(4))
        resp = cand[:, 4].contiguous().view(torch.float32)
        Y, X, S = row[:, 6].long(), row[:, 7].long(), row[:, 8].long()
        lin = ((z * Y + y) * X + x) * S + s
        # three stable sorts, least significant key first
        o = torch.sort(lin, stable=True).indices
        o = o[torch.sort(resp[o], descending=True, stable=True).indices]
        o = o[torch.sort(row[:, 10].long()[o], stable=True).indices]
        out = torch.empty((cand.shape[0], N_MERGED), dtype=torch.float64, device=dev)
        zyx = torch.stack((z, y, x), dim=1).double() + row[:, 3:6]
        out[:, 0:3] = zyx
        out[:, 3] = sig_t[tix, s] * math.sqrt(3)
        out[:, 4] = -1.0
        out[:, 5] = -1.0
        out[:, 6] = row[:, 9]
        out[:, 7:10] = zyx
        out[:, 10] = -1.0
        out[:, 11:14] = row[:, 0:3]
        return out[o]


#: columns of the table ``detect_blobs_blocks`` returns (stack_detect.py:458-467):
#: relative coordinates replaced by the seam-averaged absolute ones, abs columns dropped
FINAL_COLS = ["z", "y", "x", "radius", "confirmed", "truth", "channel", "region"]


def prune_merged(merged: torch.Tensor, overlap, tol, sub_roi_slices, sub_rois_offsets,
                 channels: Sequence[int], overlap_padding=None, final_layout: bool = False):
    """``StackPruner.prune_blobs_mp`` on a device-resident merged table.

    Returns ``((N', 11) float64 numpy table, DataFrame of pruning ratios)``; the
    only device-to-host transfer of blob rows is the final table.  With
    ``final_layout`` the table is already what ``detect_blobs_blocks`` makes of it
    afterwards (``Blobs.replace_rel_with_abs_blob_coords`` then
    ``remove_abs_blob_coords(True)``): ``(N', 8)`` in ``FINAL_COLS`` order, built on
    the device so that the host never copies the wide table."""
    from .. import gpu
    if merged is None or merged.shape[0] == 0:
        return None, None
    if overlap_padding is None:
        overlap_padding = tol
    cols = ("blobs", "ratio_pruning", "ratio_adjacent")
    ratios_out = {}
    rel = merged[:, 0:3].contiguous()
    rel_i = rel.to(torch.int32)          # the reference matches on integer-cast coordinates
    abs_zyx = merged[:, 7:10].contiguous()
    tags = merged[:, 11:14].long()
    chl_col = merged[:, 6]
    tol_i = [int(t) for t in np.broadcast_to(tol, (3,))]
    last = tuple(np.subtract(sub_roi_slices.shape, 1))
    order = []
    for chl in channels:
        cur = torch.nonzero(chl_col == float(chl)).flatten()
        for axis in range(3):
            n_sec = sub_rois_offsets.shape[axis]
            if n_sec <= 1:
                continue
            pos = rel[cur, axis]
            tag = tags[cur, axis]
            keep_parts, seam_parts = [], []
            for j in range(n_sec):
                coord = [0, 0, 0]
                coord[axis] = j
                coord = tuple(coord)
                start = float(sub_rois_offsets[coord][axis])
                sl = sub_roi_slices[coord]
                size = sl[axis].stop - sl[axis].start
                end = start + size
                shift = float(overlap[axis] + overlap_padding[axis])
                if j < n_sec - 1:
                    lo, hi = end - shift, end + float(overlap_padding[axis])
                    in_slab = torch.nonzero((pos >= lo) & (pos < hi)).flatten()
                    n_next = None
                    nlo = end + float(tol[axis])
                    nhi = nlo + float(overlap[axis]) + 2 * float(overlap_padding[axis])
                    total = float(sub_rois_offsets[last][axis]) + size
                    if nlo < total and nhi < total:
                        n_next = int(((pos >= nlo) & (pos < nhi)).sum().item())
                    t = tag[in_slab]
                    master = cur[in_slab[t == j]]
                    check = cur[in_slab[t == j + 1]]
                    if master.shape[0] and check.shape[0]:
                        m_last, hit = gpu.prune_seams(rel_i[master].contiguous(),
                                                      rel_i[check].contiguous(), tol_i)
                        sel = m_last >= 0
                        ms = master[sel]
                        if ms.shape[0]:
                            abs_zyx[ms] = torch.round(
                                (abs_zyx[ms] + abs_zyx[check[m_last[sel].long()]]) / 2)
                        check = check[~hit.bool()]
                    seam_parts.append(master)
                    seam_parts.append(check)
                    if n_next is not None:
                        ratios = detector.meas_pruning_ratio(
                            int(in_slab.shape[0]), int(master.shape[0] + check.shape[0]), n_next)
                        if ratios:
                            for c, v in zip(cols, ratios):
                                ratios_out.setdefault(c, []).append(v)
                    upper = lo
                else:
                    upper = end
                lower = start + (shift if j > 0 else 0.0)
                keep_parts.append(cur[torch.nonzero((pos < upper) & (pos >= lower)).flatten()])
            cur = torch.cat(keep_parts + seam_parts)
        order.append(cur)
    order = order[0] if len(order) == 1 else torch.cat(order)
    if final_layout:
        out = torch.empty((order.shape[0], len(FINAL_COLS)), dtype=torch.float64,
                          device=merged.device)
        out[:, 0:3] = abs_zyx[order]
        out[:, 3:7] = merged[order, 3:7]
        out[:, 7] = merged[order, 10]
        return out.cpu().numpy(), pd.DataFrame(ratios_out)
    out = merged[order, :N_COLS]
    out[:, 7:10] = abs_zyx[order]
    return out.cpu().numpy(), pd.DataFrame(ratios_out)
