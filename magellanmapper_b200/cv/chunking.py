"""Chunk geometry of a stack (mirror of ``magmap/cv/chunking.py:170-445``).

Pure integer bookkeeping on the host: which voxels belong to which overlapping
sub-ROI, how to stitch sub-ROIs back together, and how per-chunk blob tables
are merged.  The multiprocessing pool helpers of the reference
(``chunking.py:26-167``) have no role on the GPU path and are not mirrored.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def _num_units(size: Sequence[int], max_pixels: Sequence[int]) -> np.ndarray:
    """Number of sub-regions per axis = ceil(size / max_pixels)."""
    size = np.asarray(size)
    max_pixels = np.asarray(max_pixels)
    return np.ceil(size / max_pixels).astype(int)


def _bounds_side(size, max_pixels, overlap, coord, axis) -> Tuple[int, int]:
    """[start, end) of sub-region ``coord`` along ``axis``: ``max_pixels`` long
    plus the overlap into the next region, clipped at the stack's end."""
    start = int(coord[axis]) * int(max_pixels[axis])
    end = start + int(max_pixels[axis])
    if overlap is not None:
        end += int(overlap[axis])
    return start, min(end, int(size[axis]))


def stack_splitter(shape: Sequence[int], max_pixels: Sequence[int],
                   overlap: Optional[Sequence[int]] = None
                   ) -> Tuple[np.ndarray, np.ndarray]:
    """Split a stack into sub-regions.

    Returns ``(sub_roi_slices, sub_rois_offsets)``: an object array indexed by
    the (z, y, x) chunk coordinate holding a tuple of three slices, and a float
    array of the same grid shape + (3,) with each chunk's start corner.
    """
    grid = _num_units(shape[:3], max_pixels)
    slices = np.empty(tuple(grid), dtype=object)
    offsets = np.zeros(tuple(grid) + (3,))
    edges = []
    for ax in range(3):
        starts = np.arange(grid[ax]) * int(max_pixels[ax])
        ends = starts + int(max_pixels[ax]) + (0 if overlap is None else int(overlap[ax]))
        edges.append((starts, np.minimum(ends, int(shape[ax]))))
    for c in np.ndindex(*grid):
        slices[c] = tuple(slice(int(edges[a][0][c[a]]), int(edges[a][1][c[a]]))
                          for a in range(3))
        offsets[c] = [edges[a][0][c[a]] for a in range(3)]
    return slices, offsets


def _core_extent(sub_shape, coord, grid, overlap, max_pixels=None):
    """Extent of a sub-ROI with its overlap removed (not on the last chunk)."""
    ext = list(sub_shape[:3])
    for a in range(3):
        if coord[a] == grid[a] - 1:
            continue
        if max_pixels is not None:
            # a chunk shorter than max_pixels + overlap was clipped at the end of
            # the stack; its core is still max_pixels long
            if ext[a] < max_pixels[a] + overlap[a]:
                ext[a] = int(max_pixels[a])
            else:
                ext[a] -= int(overlap[a])
        elif overlap is not None:
            ext[a] -= int(overlap[a])
    return ext


def merge_split_stack(sub_rois: np.ndarray, max_pixels: Sequence[int],
                      overlap: np.ndarray) -> np.ndarray:
    """Inverse of :func:`stack_splitter` on an object array of sub-ROI arrays."""
    overlap = np.asarray(overlap).astype(int)
    grid = sub_rois.shape
    planes = []
    for z in range(grid[0]):
        rows = []
        for y in range(grid[1]):
            cols = []
            for x in range(grid[2]):
                sub = sub_rois[z, y, x]
                e = _core_extent(sub.shape, (z, y, x), grid, overlap, max_pixels)
                cols.append(sub[:e[0], :e[1], :e[2]])
            rows.append(np.concatenate(cols, axis=2))
        planes.append(np.concatenate(rows, axis=1))
    return np.concatenate(planes, axis=0)


def get_split_stack_total_shape(sub_rois: np.ndarray, overlap=None) -> np.ndarray:
    """Shape of the stack that :func:`merge_split_stack2` would fill."""
    grid = sub_rois.shape
    first = sub_rois[0, 0, 0].shape
    total = np.zeros(len(first), dtype=int)
    for a in range(3):
        idx = [0, 0, 0]
        for i in range(grid[a]):
            idx[a] = i
            sub = sub_rois[tuple(idx)]
            total[a] += _core_extent(sub.shape, tuple(idx), grid, overlap)[a]
    if len(first) > 3:
        total[3] = first[3]
    return total


def merge_split_stack2(sub_rois: np.ndarray, overlap, offset: int, output) -> None:
    """Write sub-ROIs into a preallocated ``output`` (e.g. a memmap).

    As in the reference, the write cursor advances by the FIRST chunk's shape on
    each axis, so the chunk grid is assumed regular except for the last chunk.
    """
    grid = sub_rois.shape
    step = sub_rois[0, 0, 0].shape
    if offset > 0:
        output = output[0]
    for c in np.ndindex(*grid):
        sub = sub_rois[c]
        e = _core_extent(sub.shape, c, grid, overlap)
        o = [c[a] * step[a] for a in range(3)]
        output[o[0]:o[0] + e[0], o[1]:o[1] + e[1], o[2]:o[2] + e[2]] = \
            sub[:e[0], :e[1], :e[2]]


def merge_blobs(blob_rois: np.ndarray) -> Optional[np.ndarray]:
    """Stack every chunk's blob table, tagging rows with the chunk coordinate in
    three extra trailing columns; ``None`` when no chunk has blobs."""
    cells = [(c, blob_rois[c]) for c in np.ndindex(*blob_rois.shape) if blob_rois[c] is not None]
    if not cells:
        return None
    n_rows = sum(b.shape[0] for _, b in cells)
    n_cols = cells[0][1].shape[1]
    # one allocation, filled chunk by chunk (no per-chunk temporaries, no vstack copy)
    out = np.empty((n_rows, n_cols + 3), dtype=np.result_type(*[b.dtype for _, b in cells], int))
    at = 0
    for c, blobs in cells:
        n = blobs.shape[0]
        out[at:at + n, :-3] = blobs
        out[at:at + n, -3:] = c
        at += n
    return out
