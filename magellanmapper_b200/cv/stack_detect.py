"""Whole-stack blob detection on the GPU behind ``magmap.cv.stack_detect``.

Mirror of ``magmap/cv/stack_detect.py``: block setup (``:260-335``), the
per-sub-ROI worker (``:82-172``), the fan-out over sub-ROIs (``:175-257``, a
``multiprocessing.Pool`` in the reference, a loop of fused GPU chunk launches
here), seam pruning (``:618-861``) and the two drivers ``detect_blobs_blocks``
(``:338-517``) and ``detect_blobs_stack`` (``:520-615``).

Chunk-faithful by construction: the same chunk grid, 'reflect' filtering at
every chunk face, preprocessing blocks anchored at each chunk's origin and the
same seam pruning, so results are comparable blob for blob with the reference.
"""
from __future__ import annotations

import os
from enum import Enum
from time import time
from typing import NamedTuple, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

from . import chunking, detector
from ..io import libmag, np_io
from ..plot import plot_3d
from ..settings import config, roi_prof

_logger = config.logger.getChild(__name__)


#: keep per-chunk blob tables on the device and seam-prune them there
#: (``device_tables``); False routes everything through the host tables of the
#: reference's structure (``detect_blobs_sub_rois`` + ``prune_blobs_mp``)
DEVICE_TABLES = True


#: Chunks with at most this fraction of the largest chunk's voxels (the thin trailing chunks
#: of a grid: 4-7 % of the voxels, 12 % of the step when run alone) go to a side stream with
#: their own small workspace, under the kernels of the full chunks.  Used for a
#: device-resident image on one GPU (the streamed-host and multi-GPU calls keep one stream).
#: Off until round 2's fix of the TMA stage refill (csrc/log_xy.cu, log_x.cu): the run-to-run
#: differences this route showed were that race, which any concurrent work exposed
#: (tools/concurrency_probe.py, tools/side_stream_check.py: 12 of 12 runs identical now).
THIN_CHUNK_FRACTION = 0.3
#: The same for a host image that is streamed strip by strip: results equal (8 of 8 runs), but
#: 320 instead of 278 ms per config-2 stack, because every strip hand-over joins the two
#: streams (tools/side_stream_host_check.py) - off.
THIN_SIDE_FOR_HOST_IMAGES = False


class StackTimes(Enum):
    """Keys of ``stack_detection_times.csv``."""
    DETECTION = "Detection"
    PRUNING = "Pruning"
    TOTAL = "Total_stack"


class Blocks(NamedTuple):
    """Block-processing geometry (stack_detect.py:260-279)."""
    sub_roi_slices: np.ndarray
    sub_rois_offsets: np.ndarray
    denoise_max_shape: Optional[np.ndarray]
    exclude_border: Optional[Sequence[int]]
    tol: np.ndarray
    overlap_base: np.ndarray
    overlap: np.ndarray
    overlap_padding: np.ndarray
    max_pixels: np.ndarray


def setup_blocks(settings, shape: Sequence[int]) -> Blocks:
    """Derive chunk and preprocessing-block geometry from a profile and the
    image resolution (stack_detect.py:282-335)."""
    scaling = detector.calc_scaling_factor()
    denoise_max_shape = None
    if settings["denoise_size"]:
        denoise_max_shape = np.ceil(scaling * settings["denoise_size"]).astype(int)
    overlap_base = detector.calc_overlap()
    tol = np.multiply(overlap_base, settings["prune_tol_factor"]).astype(int)
    overlap = overlap_base.copy()
    overlap_padding = tol.copy()
    exclude_border = settings["exclude_border"]
    if exclude_border is not None:
        # the overlap must exceed twice the excluded border so that no plane is
        # excluded from both neighbouring chunks; no padding past an excluded border
        twice = np.multiply(2, exclude_border)
        overlap = np.where(overlap < twice, twice, overlap)
        excluded = np.greater(exclude_border, 0)
        overlap[excluded] += 1
        overlap_padding[excluded] = 0
    max_pixels = np.ceil(scaling * settings["segment_size"]).astype(int)
    slices, offsets = chunking.stack_splitter(shape, max_pixels, overlap)
    return Blocks(slices, offsets, denoise_max_shape, exclude_border, tol, overlap_base,
                  overlap, overlap_padding, max_pixels)


def channel_ladders(img, denoise_max_shape, channels: Sequence[int]) -> np.ndarray:
    """``(n_channels, num_sigma)`` sigma ladders of a pass over ``img`` (numpy array or
    tensor; only its dtype matters): what ``StackDetector.enqueue_sub_roi`` hands to the
    library for every chunk, without needing a chunk to have run (ranks of a multi-GPU
    job that own no chunk still format the gathered table)."""
    import torch
    scale = detector.calc_scaling_factor()[2]
    is_f32 = (img.dtype == np.float32) if isinstance(img, np.ndarray) else (
        img.dtype == torch.float32)
    lads = [detector.sigma_ladder(config.get_roi_profile(c), scale,
                                  is_f32 and denoise_max_shape is None) for c in channels]
    out = np.zeros((len(lads), max(len(v) for v in lads)))
    for i, v in enumerate(lads):
        out[i, :len(v)] = v
    return out


class StackDetector(object):
    """Detects blobs sub-ROI by sub-ROI.  Class attributes carry the shared
    state like the reference's fork-friendly design; here they also cache the
    GPU workspace between sub-ROIs."""
    img5d: Optional[np_io.Image5d] = None
    img = None
    last_coord = None
    denoise_max_shape = None
    exclude_border = None
    coloc = False
    channel = None
    _gpu_detector = None
    _gpu_detector_thin = None      # second workspace: thin chunks run on a side stream
    _side_stream = None

    @classmethod
    def _workspace(cls, shape):
        from .. import gpu
        det = cls._gpu_detector
        if det is None or any(s > m for s, m in zip(shape, det.max_shape)):
            grow = shape if det is None else tuple(max(s, m) for s, m in zip(shape, det.max_shape))
            cls._gpu_detector = det = gpu.ChunkDetector(grow)
        return det

    @classmethod
    def release_workspace(cls):
        cls._gpu_detector = None
        cls._gpu_detector_thin = None
        cls._side_stream = None

    @classmethod
    def detect_sub_roi_from_data(cls, coord, sub_roi_slices, offset):
        return cls.detect_sub_roi(coord, offset, cls.last_coord, cls.denoise_max_shape,
                                  cls.exclude_border, cls.img5d, cls.img[sub_roi_slices],
                                  cls.channel, coloc=cls.coloc)

    @classmethod
    def detect_sub_roi(cls, coord: Sequence[int], offset: Sequence[int],
                       last_coord: Sequence[int], denoise_max_shape: Optional[Sequence[int]],
                       exclude_border, img5d, sub_roi, channel: Optional[Sequence[int]],
                       img_path: Optional[str] = None, coloc: bool = False
                       ) -> Tuple[Sequence[int], Optional[np.ndarray]]:
        """Preprocess (per ``denoise_max_shape`` block) and detect one sub-ROI in
        ONE fused GPU call per channel, then shift the blobs to the sub-ROI's
        offset (stack_detect.py:82-172)."""
        pending = cls.enqueue_sub_roi(coord, offset, last_coord, denoise_max_shape,
                                      exclude_border, sub_roi, channel, coloc)
        return cls.finish_sub_roi(pending)

    @classmethod
    def enqueue_sub_roi(cls, coord, offset, last_coord, denoise_max_shape, exclude_border,
                        sub_roi, channel, coloc: bool = False, det=None):
        """Launch the GPU work of one sub-ROI (every channel) without waiting
        for it; ``finish_sub_roi`` turns the returned handle into the blob
        table.  Splitting the two lets the table assembly of one sub-ROI
        overlap the kernels of the next."""
        shape = tuple(sub_roi.shape)
        multichannel, channels = plot_3d.setup_channels(sub_roi, channel, 3)
        channels = list(channels)
        if det is None:
            det = cls._workspace(detector.detection_shape(shape, channels))
        coloc_roi = None
        if coloc:
            # co-localisation reads the preprocessed intensities of EVERY channel after the
            # detection (stack_detect.py:153-156): preprocess once into a tensor that both
            # steps use instead of fusing the preprocessing into each channel's launch
            coloc_roi = sub_roi
            if denoise_max_shape is not None:
                coloc_roi = detector.preprocessed_roi(sub_roi, multichannel, denoise_max_shape)
            tickets = detector.enqueue_detection(det, coloc_roi, channels, multichannel, None,
                                                 as_float64=denoise_max_shape is not None)
        else:
            tickets = detector.enqueue_detection(det, sub_roi, channels, multichannel,
                                                 denoise_max_shape)
        return (coord, offset, last_coord, exclude_border, shape, det, tickets, coloc_roi)

    @classmethod
    def finish_sub_roi(cls, pending) -> Tuple[Sequence[int], Optional[np.ndarray]]:
        coord, offset, last_coord, exclude_border, shape, det, tickets, coloc_roi = pending
        tables = []
        chls = [t[0] for t in tickets]
        det_shape = detector.detection_shape(shape, chls) if chls else shape
        for chl, sigmas, ticket in tickets:
            cands, _ = det.collect(ticket)
            if len(cands):
                tables.append(detector.cands_to_blobs(cands, sigmas, det_shape[1:3], chl))
        segments = np.vstack(tables) if tables else None
        if segments is not None:
            segments = detector.scale_back_isotropic(segments, chls)
        if segments is not None and exclude_border is not None:
            exclude = np.array([exclude_border, exclude_border])
            exclude[0, np.equal(coord, 0)] = 0
            exclude[1, np.equal(coord, last_coord)] = 0
            segments = detector.get_blobs_interior(segments, shape, *exclude)
        if coloc_roi is not None and segments is not None:
            from . import colocalizer
            colocs = colocalizer.colocalize_blobs(coloc_roi, segments)
            if colocs is not None:
                segments = np.hstack((segments, colocs))
        if segments is not None:
            detector.Blobs.shift_blob_rel_coords(segments, offset)
            detector.Blobs.shift_blob_abs_coords(segments, offset)
        return coord, segments

    @classmethod
    def detect_blobs_sub_rois_device(cls, img, sub_roi_slices, sub_rois_offsets,
                                     denoise_max_shape, channel, coords=None,
                                     prefix=None, suffix=None, tables=None):
        """``detect_blobs_sub_rois`` with the per-chunk tables left on the device:
        returns a ``device_tables.ChunkTables`` (the survivors of every chunk as
        ``mmb_row`` records in HBM) instead of an object array of host tables;
        ``device_tables.prune_rows`` turns it into the final table.  No
        ``exclude_border`` support (the caller takes the host route for that).
        ``prefix`` / ``suffix`` (device tensors of whole planes around a HOST ``img``,
        see ``gpu.StripFeeder``) let ``multi_gpu`` stream a rank's own planes from host
        memory while the halo planes of its neighbours are already in HBM.  With
        ``tables`` the survivors are appended to an existing ``ChunkTables``, so
        several calls on different pieces of a volume can feed one table."""
        from collections import deque
        from .. import gpu
        from . import device_tables
        last_coord = np.subtract(sub_roi_slices.shape, 1)
        grid = sub_roi_slices.shape
        todo = list(np.ndindex(*grid)) if coords is None else [tuple(c) for c in coords]
        # a host image is streamed strip by strip (one strip = every chunk of one y
        # column) so that uploads overlap the kernels; chunks are then walked in
        # (y, z, x) order and the table restores the grid order
        feeder = None
        if (isinstance(img, np.ndarray) and img.flags.c_contiguous and img.dtype in gpu._NP2MMB
                and todo):
            cols = sorted({c[1] for c in todo})
            # the narrowest strip goes first: its upload is the only one nothing can
            # hide, and its kernels then cover part of the first wide strip's upload
            if len(cols) > 1:
                width = {j: (lambda sy: sy.stop - sy.start)(
                    sub_roi_slices[next(c for c in todo if c[1] == j)][1]) for j in cols}
                first = min(cols, key=lambda j: (width[j], j))
                cols = [first] + [j for j in cols if j != first]
            y_ranges = []
            for j in cols:
                # every chunk of a column has the same y range; take it from one that is
                # on this call's list (other cells of the grid may be placeholders)
                sy = sub_roi_slices[next(c for c in todo if c[1] == j)][1]
                y_ranges.append((sy.start, sy.stop))
            feeder = gpu.StripFeeder(img, y_ranges, prefix=prefix, suffix=suffix)
            todo.sort(key=lambda c: (cols.index(c[1]), c[0], c[2]))
        else:
            if prefix is not None or suffix is not None:
                raise ValueError("prefix/suffix planes need a C-contiguous host image")
            img = gpu.upload_if_fits(img)
        largest = [max(s[a].stop - s[a].start for s in sub_roi_slices.flat) for a in range(3)]
        det = cls._workspace(tuple(largest))
        n_chl = len(plot_3d.setup_channels(img, channel, 3)[1])
        if n_chl > det.n_slots - 1:
            cls._gpu_detector = None
            cls._gpu_detector = det = gpu.ChunkDetector(det.max_shape, n_slots=2 * n_chl)
        if tables is None:
            tables = device_tables.ChunkTables(
                det.device, plot_3d.setup_channels(img, channel, 3)[1])
        pending = deque()

        # The trailing chunks of a grid are thin (12 planes, or 48 voxels wide in
        # config 2): 7 % of the voxels but 12 % of the step when they run alone,
        # because their launches cannot fill the GPU.  They get their own small
        # workspace and run on a side stream, under the kernels of the full chunks.
        import torch
        nvox = {c: int(np.prod([s.stop - s.start for s in sub_roi_slices[c]])) for c in todo}
        big = max(nvox.values()) if nvox else 0
        use_side = (THIN_CHUNK_FRACTION > 0 and coords is None
                    and (feeder is None or THIN_SIDE_FOR_HOST_IMAGES))
        thin = {c for c in todo if use_side and nvox[c] <= THIN_CHUNK_FRACTION * big}
        thin_det, side, main = None, None, torch.cuda.current_stream()
        if thin and len(thin) < len(todo):
            tshape = tuple(max(sub_roi_slices[c][a].stop - sub_roi_slices[c][a].start for c in thin)
                           for a in range(3))
            if cls._side_stream is None:
                cls._side_stream = torch.cuda.Stream(device=det.device)
            side = cls._side_stream
            thin_det = cls._gpu_detector_thin
            if (thin_det is None or thin_det.n_slots < 2 * n_chl + 2
                    or any(s_ > m for s_, m in zip(tshape, thin_det.max_shape))):
                with torch.cuda.stream(side):     # its buffers belong to the side stream
                    cls._gpu_detector_thin = thin_det = gpu.ChunkDetector(
                        tshape, n_slots=2 * n_chl + 2)
            side.wait_stream(main)           # the image (if resident) was produced on `main`
        else:
            thin = set()

        # strips whose chunks are all enqueued but not all collected: a strip's device
        # buffer may only be handed back to the feeder (which at once starts overwriting
        # it with strip j + 2) when none of its chunks can still need a redo - a chunk
        # whose candidate or edge buffer overflowed is re-run from the same device view
        in_flight = {}                 # strip -> chunks enqueued and not yet collected
        closed = set()                 # strips with nothing left to enqueue

        def finish_oldest():
            (coord, offset, _, _, shape, det_, tickets, _), strip = pending.popleft()
            rank_in_grid = int(np.ravel_multi_index(coord, grid))
            for chl, sigmas, ticket in tickets:
                cand, _ = det_.collect_device(ticket)
                # the survivors were copied on the stream the chunk ran on: tag them there
                with torch.cuda.stream(ticket.stream or main):
                    tables.append(cand, rank_in_grid, sigmas, chl)
            if strip is not None:
                in_flight[strip] -= 1
                if in_flight[strip] == 0 and strip in closed:
                    if side is not None:
                        main.wait_stream(side)          # side-stream readers of the strip
                    feeder.release(strip)

        def close_strip(strip):
            closed.add(strip)
            if in_flight.get(strip, 0) == 0:
                if side is not None:
                    main.wait_stream(side)
                feeder.release(strip)

        strip_of, strip_dev = None, None
        for n_done, coord in enumerate(todo):
            use = thin_det if coord in thin else det
            while pending and use.free_slots() < n_chl:
                finish_oldest()
            if feeder is not None:
                j = cols.index(coord[1])
                if j != strip_of:
                    if strip_of is not None:
                        close_strip(strip_of)
                    # strip j reuses the buffer of strip j - 2: its upload starts when
                    # the last chunk of that strip has been collected
                    while feeder.uploaded[j] is None and pending:
                        finish_oldest()
                    strip_of, strip_dev = j, feeder.strip(j)
                    if side is not None:
                        side.wait_event(feeder.uploaded[j])
                sz, sy, sx = sub_roi_slices[coord]
                y0 = feeder.ranges[j][0]
                sub = strip_dev[sz, sy.start - y0:sy.stop - y0, sx]
            else:
                sub = img[sub_roi_slices[coord]]
            if strip_of is not None:
                in_flight[strip_of] = in_flight.get(strip_of, 0) + 1
            if coord in thin:
                with torch.cuda.stream(side):
                    pending.append((cls.enqueue_sub_roi(
                        coord, sub_rois_offsets[coord], last_coord, denoise_max_shape, None,
                        sub, channel, False, det=thin_det), strip_of))
            else:
                pending.append((cls.enqueue_sub_roi(
                    coord, sub_rois_offsets[coord], last_coord, denoise_max_shape, None,
                    sub, channel, False), strip_of))
        if side is not None:
            main.wait_stream(side)
        if feeder is not None and strip_of is not None:
            close_strip(strip_of)
        while pending:
            finish_oldest()
        if side is not None:
            main.wait_stream(side)           # the side stream's table copies
        return tables

    @classmethod
    def detect_blobs_sub_rois(cls, img5d, img, sub_roi_slices, sub_rois_offsets,
                              denoise_max_shape, exclude_border, coloc, channel, coords=None):
        """Run every sub-ROI through the GPU in z, y, x order and collect the
        blob tables in an object array shaped like the chunk grid
        (stack_detect.py:175-257).  The reference fans out over a process pool;
        here sub-ROIs are enqueued back to back on one stream and their tables
        are assembled on the host while later sub-ROIs compute.

        ``coords`` (not in the reference) restricts the work to a subset of the
        chunk grid; the other cells stay None.  ``multi_gpu`` uses it to give
        every rank its share of the chunks."""
        from collections import deque
        from .. import gpu
        last_coord = np.subtract(sub_roi_slices.shape, 1)
        # a host image that fits is moved to the device once (one large DMA, fast
        # from pinned memory); sub-ROIs are then strided views of it
        img = gpu.upload_if_fits(img)
        cls.img5d, cls.img, cls.last_coord = img5d, img, last_coord
        cls.denoise_max_shape, cls.exclude_border = denoise_max_shape, exclude_border
        cls.coloc, cls.channel = coloc, channel
        seg_rois = np.zeros(sub_roi_slices.shape, dtype=object)
        # size the workspace once for the largest chunk (in the shape blob_log sees: the
        # isotropic one when the profile resizes)
        largest = [max(s[a].stop - s[a].start for s in sub_roi_slices.flat) for a in range(3)]
        largest = list(detector.detection_shape(
            largest, list(plot_3d.setup_channels(img, channel, 3)[1])))
        det = cls._workspace(tuple(largest))
        n_chl = len(plot_3d.setup_channels(img, channel, 3)[1])
        if n_chl > det.n_slots - 1:
            cls._gpu_detector = None
            cls._gpu_detector = det = gpu.ChunkDetector(det.max_shape, n_slots=2 * n_chl)
        pending = deque()

        def finish_oldest():
            coord, segments = cls.finish_sub_roi(pending.popleft())
            seg_rois[coord] = segments

        todo = np.ndindex(*sub_roi_slices.shape) if coords is None else [tuple(c) for c in coords]
        for coord in todo:
            while pending and cls._workspace(tuple(largest)).free_slots() < n_chl:
                finish_oldest()
            pending.append(cls.enqueue_sub_roi(
                coord, sub_rois_offsets[coord], last_coord, denoise_max_shape, exclude_border,
                img[sub_roi_slices[coord]], channel, coloc))
        while pending:
            finish_oldest()
        # a table with zero rows is stored as None, like the reference
        for coord in np.ndindex(*seg_rois.shape):
            if isinstance(seg_rois[coord], int):       # cells never filled keep the zeros() default
                seg_rois[coord] = None
            elif seg_rois[coord] is not None and len(seg_rois[coord]) == 0:
                seg_rois[coord] = None
        return seg_rois


class StackPruner(object):
    """Removes duplicate blobs detected twice in the overlap between
    neighbouring sub-ROIs (stack_detect.py:618-861)."""
    blobs_to_prune = None

    @classmethod
    def prune_overlap_by_index(cls, i):
        return cls.prune_overlap(i, cls.blobs_to_prune[i])

    @classmethod
    def prune_overlap(cls, i, pruner):
        """Within one overlap slab, blobs tagged with chunk ``i`` along ``axis``
        are the masters and blobs tagged ``i + 1`` are checked against them;
        blobs of any other chunk in the slab are dropped (stack_detect.py:644-677)."""
        blobs, axis, tol, n_next = pruner
        if blobs is None:
            return None, None
        if n_next is not None and not np.isscalar(n_next):
            n_next = len(n_next)           # the reference's tuple carries the blobs themselves
        tag_col = blobs.shape[1] - 3 + axis
        n_orig = len(blobs)
        master = blobs[blobs[:, tag_col] == i]
        check = blobs[blobs[:, tag_col] == i + 1]
        pruned, master = detector.remove_close_blobs(check, master, tol)
        merged = np.concatenate((master, pruned))
        ratios = None
        if n_next is not None:
            # the reference passes the blobs of the adjacent region; only their
            # count enters the ratio (detector.py:1122-1144)
            ratios = detector.meas_pruning_ratio(n_orig, len(merged), n_next)
        return merged, ratios

    @classmethod
    def prune_blobs_mp(cls, img, seg_rois, overlap, tol, sub_roi_slices, sub_rois_offsets,
                       channels, overlap_padding=None):
        """Prune axis by axis.  For each seam the slab ``[end - (overlap + pad),
        end + pad)`` of chunk ``j`` is pruned (chunk ``j`` = master, ``j + 1`` =
        check, ``prune_overlap``); blobs outside every slab pass through; the next
        axis works on the recombined table.  Returns ``(table without chunk tags,
        DataFrame of pruning ratios)`` or ``(None, None)``.

        Same arithmetic and row order as ``prune_overlap`` applied seam by seam
        (``tests/test_host_mirror.py`` checks that), but carried out on row
        INDICES plus the three narrow arrays that matter (positions, chunk tags,
        absolute coordinates): the wide table is gathered once at the end instead
        of being copied for every seam of every axis."""
        merged = chunking.merge_blobs(seg_rois)
        if merged is None:
            return None, None
        if overlap_padding is None:
            overlap_padding = tol
        cols = ("blobs", "ratio_pruning", "ratio_adjacent")
        ratios_out = {}
        abs_inds = detector.Blobs._get_abs_inds()
        rel = np.ascontiguousarray(merged[:, :3])
        abs_zyx = np.ascontiguousarray(merged[:, abs_inds])
        tags = np.ascontiguousarray(merged[:, -3:]).astype(np.int64)
        chl_col = detector.Blobs.get_blobs_channel(merged)
        last = tuple(np.subtract(sub_roi_slices.shape, 1))
        order = []
        for chl in channels:
            cur = np.flatnonzero(np.isin(chl_col, chl))
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
                    start = sub_rois_offsets[coord][axis]
                    sl = sub_roi_slices[coord]
                    size = sl[axis].stop - sl[axis].start
                    end = start + size
                    shift = overlap[axis] + overlap_padding[axis]
                    if j < n_sec - 1:
                        lo, hi = end - shift, end + overlap_padding[axis]
                        in_slab = np.flatnonzero((pos >= lo) & (pos < hi))
                        # same-sized region just past the slab, for the ratio metric
                        n_next = None
                        nlo = end + tol[axis]
                        nhi = nlo + overlap[axis] + 2 * overlap_padding[axis]
                        total = sub_rois_offsets[last][axis] + size
                        if nlo < total and nhi < total:
                            n_next = int(np.count_nonzero((pos >= nlo) & (pos < nhi)))
                        # chunk j = master, chunk j + 1 = check, any other chunk dropped
                        master = cur[in_slab[tag[in_slab] == j]]
                        check = cur[in_slab[tag[in_slab] == j + 1]]
                        if len(master) and len(check):
                            m_last, hit = detector._find_close_blobs(rel[check], rel[master], tol)
                            sel = m_last >= 0
                            if np.any(sel):
                                abs_zyx[master[sel]] = np.around(
                                    (abs_zyx[master[sel]] + abs_zyx[check[m_last[sel]]]) / 2)
                            check = check[~hit]
                        seam_parts.append(master)
                        seam_parts.append(check)
                        if n_next is not None:
                            ratios = detector.meas_pruning_ratio(
                                len(in_slab), len(master) + len(check), n_next)
                            if ratios:
                                for c, v in zip(cols, ratios):
                                    ratios_out.setdefault(c, []).append(v)
                        upper = lo
                    else:
                        upper = end
                    lower = start + (shift if j > 0 else 0)
                    keep_parts.append(cur[np.flatnonzero((pos < upper) & (pos >= lower))])
                cur = np.concatenate(keep_parts + seam_parts)
            order.append(cur)
        order = order[0] if len(order) == 1 else np.concatenate(order)
        out = merged[order, :-3]
        out[:, abs_inds] = abs_zyx[order]
        return out, pd.DataFrame(ratios_out)


def detect_blobs_blocks(filename_base: str, img5d: np_io.Image5d,
                        offset: Optional[Sequence[int]] = None,
                        size: Optional[Sequence[int]] = None,
                        channels: Optional[Sequence[int]] = None, verify: bool = False,
                        save_dfs: bool = True, full_roi: bool = False, coloc: bool = False):
    """Detect blobs in a large image by block processing.

    Returns ``(stats_detection, fdbk, blobs)`` like the reference
    (stack_detect.py:338-517); verification against a truth database is outside
    the accelerated path, so the first two are always None.

    Raises:
        ValueError: if ``img5d.img`` is None.
    """
    t_start = time()
    if img5d.img is None:
        raise ValueError("Image data is None")
    if verify:
        raise NotImplementedError("truth-database verification is outside the accelerated path")
    image5d = img5d.img
    subimg_base = filename_base
    if size is None or offset is None:
        size = image5d.shape[1:4]
        offset = (0, 0, 0)
    else:
        # named x,y,z as naming.make_subimage_name does (magmap/io/naming.py:9-37)
        roi_site = "{}x{}".format(tuple(int(v) for v in offset)[::-1],
                                  tuple(int(v) for v in size)[::-1]).replace(" ", "")
        subimg_base = libmag.insert_before_ext(filename_base, roi_site, "_")
    filename_blobs = libmag.combine_paths(subimg_base, config.SUFFIX_BLOBS)

    if full_roi:
        roi = image5d[0]
    else:
        z0, y0, x0 = (int(v) for v in offset)
        roi = image5d[0, z0:z0 + int(size[0]), y0:y0 + int(size[1]), x0:x0 + int(size[2])]
    n_chl = 1 if roi.ndim < 4 else roi.shape[3]
    if n_chl < 2:
        coloc = False

    t_det = time()
    final_on_device = False
    if channels is None:
        _, channels = plot_3d.setup_channels(roi, channels, 3)
    settings = config.get_roi_profile(channels[0])
    blocks = setup_blocks(settings, roi.shape)
    if (blocks.exclude_border is None and DEVICE_TABLES and not coloc
            and settings["isotropic"] is None):
        # per-chunk tables stay in HBM; merged, seam-pruned there, one copy back
        from . import device_tables
        tables = StackDetector.detect_blobs_sub_rois_device(
            roi, blocks.sub_roi_slices, blocks.sub_rois_offsets, blocks.denoise_max_shape,
            channels)
        detection_time = time() - t_det
        t_prune = time()
        segments_all, df_pruning = device_tables.prune_rows(
            tables.rows(), tables.ladders(), blocks.overlap, blocks.tol, blocks.sub_roi_slices,
            channels, blocks.overlap_padding, final_layout=True)
        final_on_device = True
        pruning_time = time() - t_prune
    else:
        seg_rois = StackDetector.detect_blobs_sub_rois(
            img5d, roi, blocks.sub_roi_slices, blocks.sub_rois_offsets,
            blocks.denoise_max_shape, blocks.exclude_border, coloc, channels)
        detection_time = time() - t_det
        t_prune = time()
        segments_all, df_pruning = StackPruner.prune_blobs_mp(
            roi, seg_rois, blocks.overlap, blocks.tol, blocks.sub_roi_slices,
            blocks.sub_rois_offsets, channels, blocks.overlap_padding)
        pruning_time = time() - t_prune

    if df_pruning is not None and save_dfs and len(df_pruning.columns):
        df_pruning.to_csv("blob_ratios.csv", index=False)
        if "blobs" in df_pruning.columns:
            w = df_pruning["blobs"]
            means = {f"mean_{c}": [float(np.sum(df_pruning[c] * w) / np.sum(w))]
                     for c in df_pruning.columns[1:]}
            pd.DataFrame(means).to_csv("blob_ratios_means.csv", index=False)

    if final_on_device:
        # the device route already delivered the final layout
        from . import device_tables
        blobs = detector.Blobs(segments_all, path=filename_blobs,
                               cols=list(device_tables.FINAL_COLS))
    else:
        blobs = detector.Blobs(segments_all, path=filename_blobs)
    colocs = None
    if segments_all is not None and not final_on_device:
        # the abs columns carried the seam-averaged positions; they become the
        # coordinates and the helper columns go away (stack_detect.py:458-467)
        blobs.replace_rel_with_abs_blob_coords(segments_all)
        blobs.blobs = segments_all
        if coloc:
            # the reference slices from column 10 (stack_detect.py:463), i.e. the `region`
            # column and all but the last co-localisation column: kept as it is
            colocs = segments_all[:, 10:10 + n_chl].astype(np.uint8)
        segments_all = blobs.remove_abs_blob_coords(True)

    blobs.blobs = segments_all
    blobs.colocalizations = colocs
    blobs.resolutions = config.resolutions
    blobs.basename = os.path.basename(config.filename) if config.filename else None
    blobs.roi_offset = offset
    blobs.roi_size = size

    times = {StackTimes.DETECTION: [detection_time], StackTimes.PRUNING: [pruning_time],
             StackTimes.TOTAL: [time() - t_start]}
    if save_dfs:
        pd.DataFrame({k.value: v for k, v in times.items()}).to_csv(
            "stack_detection_times.csv", index=False)
    _logger.info("Blob detection %.3f s, pruning %.3f s, blobs %s", detection_time,
                 pruning_time, 0 if segments_all is None else len(segments_all))
    blobs.times = times
    return None, None, blobs


def detect_blobs_stack(filename_base: str, img5d: Optional[np_io.Image5d],
                       subimg_offset: Optional[Sequence[int]] = None,
                       subimg_size: Optional[Sequence[int]] = None, coloc: bool = False):
    """Detect blobs in a whole stack, grouping channels whose profiles share the
    same block settings into one pass, and save the ``*_blobs.npz`` archive
    (stack_detect.py:520-615).

    Raises:
        IOError: if there is no image.
    """
    if img5d is None or img5d.img is None:
        raise IOError("No image data available for blob detection")
    channels = plot_3d.setup_channels(img5d.img, config.channel, 4)[1]
    channels = list(channels)
    if roi_prof.ROIProfile.is_identical_settings(
            [config.get_roi_profile(c) for c in channels], roi_prof.ROIProfile.BLOCK_SIZES):
        groups = [channels]
    else:
        groups = [[c] for c in channels]

    outs = []
    for chl in groups:
        _, _, blobs = detect_blobs_blocks(
            filename_base, img5d, subimg_offset, subimg_size, chl, False,
            not config.grid_search_profile, img5d.is_roi, coloc)
        outs.append(blobs)
    blobs_all = None
    if outs:
        blobs_all = outs[0]
        blobs_all.blobs = libmag.combine_arrs([b.blobs for b in outs if b.blobs is not None])
        blobs_all.colocalizations = None
        blobs_all.save_archive()
    return None, None, blobs_all
