"""Restatement of MagellanMapper's own Python on the blob-detection path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the
unmodified reference by ``oracle/make_golden.py`` -> ``tests/golden/`` ->
``tests/test_oracle_vs_reference.py``.

Everything takes explicit parameters (``Profile``, ``resolution``,
``near_max``) where the reference reads ``magmap.settings.config`` globals.
Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
import multiprocessing as mp
import os
from dataclasses import dataclass, field, replace
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from oracle import skimage_restated as ski

#: Column order of the blob table, ``magmap/cv/detector.py:88-113``.
COLS = ("z", "y", "x", "radius", "confirmed", "truth", "channel",
        "abs_z", "abs_y", "abs_x", "region")
OVERLAP_FACTOR = 5                      # magmap/cv/detector.py:41


@dataclass
class Profile:
    """The ROI-profile keys the path reads (``magmap/settings/roi_prof.py:72-134``)
    with the ``roi_blobs`` values (``profiles/roi_blobs.yaml``), which equal the
    ROIProfile defaults."""
    clip_vmin: float = 5
    clip_vmax: float = 99.5
    clip_min: float = 0.2
    clip_max: float = 1.0
    max_thresh_factor: float = 0.5
    tot_var_denoise: Optional[float] = None
    unsharp_strength: float = 0.3
    erosion_threshold: float = 0.2
    min_sigma_factor: float = 3
    max_sigma_factor: float = 5
    num_sigma: int = 10
    detection_threshold: float = 0.1
    overlap: float = 0.5
    exclude_border: Optional[Sequence[int]] = None
    segment_size: float = 500
    denoise_size: Optional[float] = 25
    prune_tol_factor: Sequence[float] = (1, 1, 1)
    isotropic: Optional[Sequence[float]] = None
    #: ``{channel: {channel to subtract: factor}}`` (roi_prof.py:139-142)
    spectral_unmixing: Optional[Dict[int, Dict[int, float]]] = None


# --------------------------------------------------------------------------
# chunk geometry: magmap/cv/chunking.py:170-256
# --------------------------------------------------------------------------

def num_units(size: Sequence[int], max_pixels: Sequence[int]) -> np.ndarray:
    """chunking.py:170-185 - ceil division per axis."""
    size = np.asarray(size)[:3]
    mp_ = np.asarray(max_pixels)
    return (-(-size // mp_)).astype(int)


def stack_splitter(shape: Sequence[int], max_pixels: Sequence[int],
                   overlap: Optional[Sequence[int]] = None
                   ) -> Tuple[np.ndarray, np.ndarray]:
    """chunking.py:214-256.  Chunk ``c`` on an axis starts at ``c*max_pixels``
    and ends ``max_pixels (+ overlap)`` later, clipped to the axis length.
    Returns the object array of slice triples and the float (…,3) offsets."""
    n = num_units(shape, max_pixels)
    slices = np.zeros(n, dtype=object)
    offsets = np.zeros(tuple(n) + (3,))
    for c in np.ndindex(*n):
        sl = []
        for ax in range(3):
            lo = int(c[ax] * max_pixels[ax])
            hi = lo + int(max_pixels[ax]) + (0 if overlap is None else int(overlap[ax]))
            hi = min(hi, int(shape[ax]))
            sl.append(slice(lo, hi))
        slices[c] = tuple(sl)
        offsets[c] = [s.start for s in sl]
    return slices, offsets


def calc_scaling_factor(resolution: Sequence[float]) -> np.ndarray:
    """detector.py:810-825 - pixels per micron, ``1 / resolutions[0]``."""
    return np.divide(1.0, resolution)


def calc_overlap(resolution: Sequence[float], factor: Optional[int] = None) -> np.ndarray:
    """detector.py:828-841."""
    if factor is None:
        factor = OVERLAP_FACTOR
    return np.ceil(np.multiply(calc_scaling_factor(resolution), factor)).astype(int)


@dataclass
class Blocks:
    """stack_detect.py:260-279."""
    sub_roi_slices: np.ndarray
    sub_rois_offsets: np.ndarray
    denoise_max_shape: Optional[np.ndarray]
    exclude_border: Optional[Sequence[int]]
    tol: np.ndarray
    overlap_base: np.ndarray
    overlap: np.ndarray
    overlap_padding: np.ndarray
    max_pixels: np.ndarray


def setup_blocks(prof: Profile, shape: Sequence[int],
                 resolution: Sequence[float]) -> Blocks:
    """stack_detect.py:282-335."""
    scaling = calc_scaling_factor(resolution)
    denoise_max_shape = None
    if prof.denoise_size:
        denoise_max_shape = np.ceil(np.multiply(scaling, prof.denoise_size)).astype(int)
    overlap_base = calc_overlap(resolution)
    tol = np.multiply(overlap_base, prof.prune_tol_factor).astype(int)
    overlap_padding = np.copy(tol)
    overlap = np.copy(overlap_base)
    if prof.exclude_border is not None:
        thresh = np.multiply(2, prof.exclude_border)
        less = np.less(overlap, thresh)
        overlap[less] = thresh[less]
        excluded = np.greater(prof.exclude_border, 0)
        overlap[excluded] += 1
        overlap_padding[excluded] = 0
    max_pixels = np.ceil(np.multiply(scaling, prof.segment_size)).astype(int)
    sl, off = stack_splitter(shape, max_pixels, overlap)
    return Blocks(sl, off, denoise_max_shape, prof.exclude_border, tol,
                  overlap_base, overlap, overlap_padding, max_pixels)


# --------------------------------------------------------------------------
# preprocessing: magmap/plot/plot_3d.py:55-172 (single channel)
# --------------------------------------------------------------------------

def saturate_roi(roi: np.ndarray, prof: Profile, near_max: float = -1.0) -> np.ndarray:
    """plot_3d.py:55-111 for one channel.  A block whose two percentiles
    coincide is returned UNCHANGED (raw dtype), ``plot_3d.py:95-96``."""
    vmin, vmax = np.percentile(roi, (prof.clip_vmin, prof.clip_vmax))
    if vmin == vmax:
        return roi
    max_thresh = near_max * prof.max_thresh_factor
    if vmax < max_thresh:
        vmax = max_thresh
    sat = np.clip(roi, vmin, vmax)
    return (sat - vmin) / (vmax - vmin)


def denoise_roi(roi: np.ndarray, prof: Profile) -> np.ndarray:
    """plot_3d.py:114-172 for one channel, ``tot_var_denoise`` off (the
    default and ``roi_blobs`` value; total-variation denoising is outside the
    accelerated path)."""
    if prof.tot_var_denoise:
        raise NotImplementedError("tot_var_denoise is outside the hot path")
    mean = np.mean(roi)
    den = np.clip(roi, prof.clip_min, prof.clip_max)
    if prof.unsharp_strength:
        blurred = ski.filters_gaussian(den, 8)
        high_pass = den - prof.unsharp_strength * blurred
        den = den + high_pass
    if prof.erosion_threshold and mean > prof.erosion_threshold:
        den = ski.erosion_octahedron1(den)
    return den


def preprocess_blocks(sub_roi: np.ndarray, denoise_max_shape: Sequence[int],
                      prof: Profile, near_max: float = -1.0) -> np.ndarray:
    """stack_detect.py:122-150: split the chunk into non-overlapping blocks of
    ``denoise_max_shape`` anchored at the chunk origin, saturate + denoise each
    on its own, and reassemble."""
    sl, _ = stack_splitter(sub_roi.shape, denoise_max_shape)
    out = None
    for c in np.ndindex(*sl.shape):
        blk = denoise_roi(saturate_roi(sub_roi[sl[c]], prof, near_max), prof)
        if out is None:
            # merged array takes the dtype of the first block (:147-148)
            out = np.zeros(sub_roi.shape[:3], dtype=blk.dtype)
        out[sl[c]] = blk
    return out


# --------------------------------------------------------------------------
# detection: magmap/cv/detector.py:874-957
# --------------------------------------------------------------------------

def format_blobs(blobs4: np.ndarray, channel: int) -> np.ndarray:
    """detector.py:325-364: pad ``[z,y,x,radius]`` to all ``COLS`` with -1,
    copy relative into absolute coordinates, set the channel."""
    n = len(blobs4)
    out = np.full((n, len(COLS)), -1.0)
    out[:, :4] = blobs4
    out[:, 7:10] = out[:, 0:3]
    out[:, 6] = channel
    return out


def get_blobs_interior(blobs, shape, pad_start, pad_end):
    """detector.py:1248-1268."""
    m = np.ones(len(blobs), dtype=bool)
    for ax in range(3):
        m &= blobs[:, ax] >= pad_start[ax]
        m &= blobs[:, ax] < shape[ax] - pad_end[ax]
    return blobs[m]


def calc_isotropic_factor(scale, resolution: Sequence[float]) -> np.ndarray:
    """cv_nd.py:1040-1068."""
    resize_factor = np.divide(resolution, np.amin(resolution))
    return resize_factor * scale


def make_isotropic(roi: np.ndarray, scale, resolution: Sequence[float]) -> np.ndarray:
    """cv_nd.py:1071-1167: ``transform.resize`` to ``(shape[:3] * factor).astype(int)``
    with ``preserve_range=True``, 'reflect' boundaries ('edge' when any axis is one voxel
    thick), cast back to the ROI's dtype."""
    factor = calc_isotropic_factor(scale, resolution)
    iso = np.array(roi.shape)
    iso[:3] = (iso[:3] * factor).astype(int)
    mode = "edge" if np.any(np.array(roi.shape) == 1) else "reflect"
    out = ski.transform_resize(roi, iso, mode=mode, preserve_range=True)
    return out.astype(roi.dtype)


def detect_blobs(roi: np.ndarray, prof, resolution: Sequence[float],
                 channel=0, exclude_border=None, full: bool = False):
    """detector.py:874-957.  ``prof`` is one ``Profile`` or one per channel of a
    ``(z, y, x, c)`` ROI; ``channel`` an index or a list of them.  ``isotropic`` resize
    with the FIRST channel's setting (:893-897), spectral unmixing per channel
    (:910-921), sigma range = factor x x-axis scaling (:903-927), radius =
    sigma * sqrt(3) (:937), coordinates scaled back and cast to int (:944-951)."""
    shape = roi.shape
    multichannel = roi.ndim > 3
    channels = list(channel) if isinstance(channel, (list, tuple, range)) else [channel]
    profs = list(prof) if isinstance(prof, (list, tuple)) else None
    get = (lambda c: profs[c]) if profs is not None else (lambda c: prof)
    isotropic = get(channels[0]).isotropic
    if isotropic is not None:
        roi = make_isotropic(roi, isotropic, resolution)
    scale = calc_scaling_factor(resolution)[2]
    tables, last = [], None
    for chl in channels:
        p = get(chl)
        roi_detect = roi[..., chl] if multichannel else roi
        if p.spectral_unmixing is not None:
            for spec_chl, spec_subtr in p.spectral_unmixing.items():
                if spec_chl != chl:
                    continue
                for subt_chl, subt_fac in spec_subtr.items():
                    roi_detect = np.subtract(roi_detect, subt_fac * roi[..., subt_chl])
                    roi_detect[roi_detect < 0] = 0
        res = ski.blob_log(
            roi_detect, min_sigma=p.min_sigma_factor * scale,
            max_sigma=p.max_sigma_factor * scale, num_sigma=p.num_sigma,
            threshold=p.detection_threshold, overlap=p.overlap, full=True)
        last = res
        if res.blobs.size < 1:
            continue
        b = res.blobs.copy()
        b[:, 3] = b[:, 3] * math.sqrt(3)
        tables.append(format_blobs(b, chl))
    if not tables:
        return (None, last) if full else None
    out = np.vstack(tables)
    if isotropic is not None:
        f = 1 / calc_isotropic_factor(isotropic, resolution)
        out[:, 0:3] = np.multiply(out[:, 0:3], f).astype(int)
        out[:, 7:10] = np.multiply(out[:, 7:10], f).astype(int)
    if exclude_border is not None:
        out = get_blobs_interior(out, shape, *exclude_border)
    return (out, last) if full else out


# --------------------------------------------------------------------------
# per-chunk worker: magmap/cv/stack_detect.py:82-172
# --------------------------------------------------------------------------

def detect_sub_roi(coord, offset, last_coord, denoise_max_shape, exclude_border,
                   sub_roi, prof: Profile, resolution, near_max=-1.0, channel=0):
    """stack_detect.py:82-172."""
    if denoise_max_shape is not None:
        sub_roi = preprocess_blocks(sub_roi, denoise_max_shape, prof, near_max)
    exclude = None
    if exclude_border is not None:
        exclude = np.array([exclude_border, exclude_border])
        exclude[0, np.equal(coord, 0)] = 0
        exclude[1, np.equal(coord, last_coord)] = 0
    seg = detect_blobs(sub_roi, prof, resolution, channel, exclude)
    if seg is not None:
        seg[:, 0:3] += offset
        seg[:, 7:10] += offset
    return coord, seg


# --------------------------------------------------------------------------
# seam pruning: detector.py:1000-1085, chunking.py:410-445,
# stack_detect.py:644-861
# --------------------------------------------------------------------------

def merge_blobs(seg_rois: np.ndarray) -> Optional[np.ndarray]:
    """chunking.py:410-445: stack every chunk's table, tagging rows with the
    chunk coordinate in three trailing columns."""
    parts = []
    for c in np.ndindex(*seg_rois.shape):
        b = seg_rois[c]
        if b is None:
            continue
        tag = np.zeros((len(b), 3), dtype=int)
        tag[:] = c
        parts.append(np.concatenate((b, tag), axis=1))
    return np.vstack(parts) if parts else None


def _int_dtype_for(max_val: float):
    """libmag.py:1116-1153 with integer=True, signed=True."""
    for dt in (np.int8, np.int16, np.int32, np.int64):
        if np.iinfo(dt).min <= 0 and np.iinfo(dt).max >= max_val:
            return dt
    raise TypeError("no integer dtype holds the coordinate range")


def remove_close_blobs(blobs: np.ndarray, blobs_master: np.ndarray, tol,
                       chunk_size: int = 1000):
    """detector.py:1009-1085.  Box test ``|d| <= tol`` on coordinates cast to
    the smallest signed integer dtype; every matched check blob is dropped;
    matched masters get abs coords = round-half-even mean with the check blob,
    the LAST match in (master tile, check tile, master row, check row) order
    winning through repeated fancy-index assignment."""
    if len(blobs) < 1 or len(blobs_master) < 1:
        return blobs, blobs_master
    dt = _int_dtype_for(max(np.amax(blobs[:, :3]), np.amax(blobs_master[:, :3])))
    mc, mm = [], []
    for i0 in range(0, len(blobs_master), chunk_size):
        ref = blobs_master[i0:i0 + chunk_size, :3].astype(dt)
        for j0 in range(0, len(blobs), chunk_size):
            chk = blobs[j0:j0 + chunk_size].astype(dt)
            diffs = np.abs(ref[:, None, :3] - chk[:, :3])
            cm, cc = np.nonzero((diffs <= tol).all(2))
            mc.append(cc + j0)
            mm.append(cm + i0)
    match_check = np.concatenate(mc)
    match_master = np.concatenate(mm)
    pruned = np.delete(blobs, match_check, axis=0)
    between = np.around((blobs_master[match_master][:, 7:10]
                         + blobs[match_check][:, 7:10]) / 2)
    upd = blobs_master[match_master]
    upd[:, 7:10] = between
    blobs_master[match_master] = upd
    return pruned, blobs_master


def prune_overlap(i: int, blobs: Optional[np.ndarray], axis: int, tol):
    """stack_detect.py:644-677 (without the ratio diagnostics): rows tagged
    with chunk ``i`` on ``axis`` are masters, rows tagged ``i+1`` are checked,
    rows from any other chunk are dropped."""
    if blobs is None:
        return None
    col = blobs.shape[1] - 3 + axis
    master = blobs[blobs[:, col] == i]
    check = blobs[blobs[:, col] == i + 1]
    pruned, master = remove_close_blobs(check, master, tol)
    return np.concatenate((master, pruned))


def prune_blobs_mp(roi_shape, seg_rois, overlap, tol, sub_roi_slices,
                   sub_rois_offsets, channels=(0,), overlap_padding=None):
    """stack_detect.py:680-861, run serially.  Returns the (N, 11) table (chunk
    tags stripped) or ``None``."""
    merged = merge_blobs(seg_rois)
    if merged is None:
        return None
    if overlap_padding is None:
        overlap_padding = tol
    out_all = []
    for chl in channels:
        blobs = merged[np.isin(merged[:, 6], chl)]
        for axis in range(3):
            nsec = sub_rois_offsets.shape[axis]
            if nsec <= 1:
                continue
            non_ol_all = None
            work = []
            for j in range(nsec):
                coord = [0, 0, 0]
                coord[axis] = j
                off = sub_rois_offsets[tuple(coord)]
                sl = sub_roi_slices[tuple(coord)]
                size = [s.stop - s.start for s in sl]
                shift = overlap[axis] + overlap_padding[axis]
                o = off[axis]
                conds = []
                blobs_ol = None
                if j < nsec - 1:
                    b0 = o + size[axis] - shift
                    b1 = o + size[axis] + overlap_padding[axis]
                    blobs_ol = blobs[(blobs[:, axis] >= b0) & (blobs[:, axis] < b1)]
                    conds.append(blobs[:, axis] < b0)
                else:
                    conds.append(blobs[:, axis] < o + size[axis])
                start = o + (shift if j > 0 else 0)
                conds.append(blobs[:, axis] >= start)
                non_ol = blobs[np.all(conds, axis=0)]
                non_ol_all = non_ol if non_ol_all is None else np.concatenate(
                    (non_ol_all, non_ol))
                work.append(blobs_ol)
            ol_all = None
            for j, b in enumerate(work):
                res = prune_overlap(j, b, axis, tol)
                if ol_all is None:
                    ol_all = res
                elif res is not None:
                    ol_all = np.concatenate((ol_all, res))
            if ol_all is None:
                blobs = non_ol_all
            elif non_ol_all is None:
                blobs = ol_all
            else:
                blobs = np.concatenate((non_ol_all, ol_all))
        out_all.append(blobs)
    return np.vstack(out_all)[:, :-3]


def finalize_blobs(segments_all: Optional[np.ndarray]) -> Optional[np.ndarray]:
    """stack_detect.py:458-467: overwrite relative with absolute coordinates,
    then drop the three abs columns -> ``z,y,x,radius,confirmed,truth,channel,
    region``."""
    if segments_all is None:
        return None
    segments_all = segments_all.copy()
    segments_all[:, 0:3] = segments_all[:, 7:10]
    keep = [i for i in range(len(COLS)) if i not in (7, 8, 9)]
    return segments_all[:, keep]


# --------------------------------------------------------------------------
# intensity co-localisation: magmap/cv/colocalizer.py:340-441
# --------------------------------------------------------------------------

def colocalize_blobs(roi: np.ndarray, blobs: Optional[np.ndarray], thresh=None):
    """colocalizer.py:340-441: per channel, label a mask with the indices of that
    channel's blobs, grey-dilate it with ``ball(2)``, average the ROI under every label
    in every channel, and flag the blobs whose average reaches the other channel's
    threshold (the smallest such average of that channel's own blobs, or a percentile)."""
    if blobs is None or roi is None or roi.ndim < 4:
        return None
    if thresh is None:
        thresh = "min"
    selem = ski.ball(2)
    shape = roi.shape[:3]
    in_roi = np.all([(blobs[:, a] >= 0) & (blobs[:, a] < shape[a]) for a in range(3)], axis=0)
    blobs_roi = blobs[in_roi]
    blobs_chl = blobs_roi[:, 6]
    threshs, masks, ranges = [], [], []
    for chl in range(roi.shape[3]):
        sel = np.isin(blobs_chl, chl)
        rng_ = np.where(sel)[0]
        ranges.append(rng_)
        mask = np.ones(shape, dtype=int) * -1
        zyx = blobs_roi[sel, :3].astype(int)
        mask[tuple(zyx.T)] = rng_
        mask = ski.dilation(mask, selem)
        masks.append(mask)
        if thresh == "min":
            threshs.append(None if len(rng_) == 0 else np.amin(
                [np.mean(roi[mask == b, chl]) for b in rng_]))
        else:
            mb = mask >= 0
            threshs.append(np.percentile(roi if np.sum(mb) < 1 else roi[mb, chl], thresh))
    channels = np.unique(blobs_chl).astype(int)
    colocs_roi = np.zeros((len(blobs_roi), roi.shape[3]), dtype=np.uint8)
    for chl in channels:
        for other in channels:
            if threshs[other] is None:
                continue
            for b in ranges[chl]:
                if np.mean(roi[masks[chl] == b, other]) >= threshs[other]:
                    colocs_roi[b, other] = 1
    colocs = np.zeros((len(blobs), roi.shape[3]), dtype=np.uint8)
    colocs[in_roi] = colocs_roi
    return colocs


# --------------------------------------------------------------------------
# whole-stack driver: stack_detect.py:338-517 (single channel, no coloc/verify)
# --------------------------------------------------------------------------

_WORK: Dict[str, object] = {}


def _pool_task(coord):
    w = _WORK
    blocks: Blocks = w["blocks"]
    sub = w["img"][blocks.sub_roi_slices[coord]]
    return detect_sub_roi(coord, blocks.sub_rois_offsets[coord], w["last"],
                          blocks.denoise_max_shape, blocks.exclude_border, sub,
                          w["prof"], w["res"], w["near_max"], w["chl"])


def detect_blobs_blocks(img: np.ndarray, prof: Profile, resolution,
                        near_max: float = -1.0, channel: int = 0,
                        processes: Optional[int] = 1, return_parts: bool = False):
    """stack_detect.py:338-517.  ``processes=1`` runs in-process; otherwise a
    fork ``multiprocessing.Pool`` with one task per chunk in z,y,x order, as
    ``detect_blobs_sub_rois`` (:175-257) with the image shared copy-on-write."""
    blocks = setup_blocks(prof, img.shape, resolution)
    last = np.subtract(blocks.sub_roi_slices.shape, 1)
    _WORK.update(blocks=blocks, img=img, last=last, prof=prof, res=resolution,
                 near_max=near_max, chl=channel)
    coords = list(np.ndindex(*blocks.sub_roi_slices.shape))
    seg_rois = np.zeros(blocks.sub_roi_slices.shape, dtype=object)
    if processes == 1:
        results = [_pool_task(c) for c in coords]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(processes=processes) as pool:
            results = [r.get() for r in [pool.apply_async(_pool_task, (c,)) for c in coords]]
    for coord, seg in results:
        seg_rois[coord] = seg
    pruned = prune_blobs_mp(img.shape, seg_rois, blocks.overlap, blocks.tol,
                            blocks.sub_roi_slices, blocks.sub_rois_offsets,
                            (channel,), blocks.overlap_padding)
    final = finalize_blobs(pruned)
    if return_parts:
        return final, seg_rois, blocks
    return final


# ---- import metadata: intensity bounds (magmap/io/importer.py) ---------------------

def calc_intensity_bounds(image, lower=0.5, upper=99.5, channel_axis=None):
    """``importer.calc_intensity_bounds`` (importer.py:1415-1444): np.percentile of the
    whole array, per channel when ``channel_axis`` is given."""
    if channel_axis is None:
        low, high = np.percentile(image, (lower, upper))
        return [low], [high]
    lows, highs = [], []
    for c in range(image.shape[channel_axis]):
        low, high = np.percentile(np.take(image, c, axis=channel_axis), (lower, upper))
        lows.append(low)
        highs.append(high)
    return lows, highs


def calc_near_bounds(image, multichannel=False):
    """Per-plane bounds reduced over the planes (importer.py:571-583, 1447-1468) of a
    (z, y, x[, c]) image: ``(near_mins, near_maxs)`` as 1-D float64 arrays per channel."""
    lows, highs = [], []
    for z in range(image.shape[0]):
        lo, hi = calc_intensity_bounds(image[z], channel_axis=2 if multichannel else None)
        lows.append(lo)
        highs.append(hi)
    return np.amin(np.array(lows), 0), np.amax(np.array(highs), 0)
