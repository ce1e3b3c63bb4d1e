"""Intensity-based co-localisation of blobs across channels (mirror of
``magmap/cv/colocalizer.py:340-441``).  The per-voxel work - labelling every blob's
``ball(2)`` neighbourhood and summing the ROI under it in every channel - runs on the GPU
(``mmb_coloc_sums``); thresholds and comparisons follow the reference's numpy."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import detector


def blob_surround_means(roi, blobs_roi: np.ndarray):
    """``(means, counts)``: ``means[b, c] = np.mean(roi[mask_chl(b) == b, c])`` where
    ``mask_chl`` is the reference's dilated label mask of blob ``b``'s own channel;
    ``counts[b]`` the number of voxels blob ``b`` owns (0 for a blob that another blob of
    its channel with a larger index covers completely)."""
    import torch
    from .. import gpu, _lib
    lib = _lib.load()
    dev = gpu.require_cuda()
    n = len(blobs_roi)
    Z, Y, X, nc = (int(v) for v in roi.shape[:4])
    if isinstance(roi, np.ndarray):
        dt = gpu._NP2MMB.get(roi.dtype)
        a = np.ascontiguousarray(roi if dt is not None else roi.astype(np.float64))
        dt = dt if dt is not None else _lib.MMB_F64
        t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to(dev)
    else:
        t = roi if roi.is_cuda else roi.to(dev)
        dt = gpu._T2MMB[t.dtype]
    zyxc = np.column_stack([blobs_roi[:, :3].astype(int),
                            detector.Blobs.get_blobs_channel(blobs_roi).astype(int)])
    d_blobs = torch.from_numpy(np.ascontiguousarray(zyxc, dtype=np.int32)).to(dev)
    sums = torch.empty((max(n, 1), nc), dtype=torch.float64, device=dev)
    counts = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    work = torch.empty(lib.mmb_coloc_work_bytes(Z, Y, X), dtype=torch.uint8, device=dev)
    _lib.check(lib.mmb_coloc_sums(
        C.c_void_p(t.data_ptr()), dt, (C.c_int64 * 4)(*[int(s) for s in t.stride()]), Z, Y, X, nc,
        C.c_void_p(d_blobs.data_ptr()), n, C.c_void_p(sums.data_ptr()),
        C.c_void_p(counts.data_ptr()), C.c_void_p(work.data_ptr()), gpu._stream()))
    cnt = counts[:n].cpu().numpy()
    with np.errstate(invalid="ignore", divide="ignore"):
        means = sums[:n].cpu().numpy() / cnt[:, None]
    return means, cnt


def colocalize_blobs(roi, blobs: Optional[np.ndarray], thresh=None) -> Optional[np.ndarray]:
    """Flag, per blob and channel, whether the intensity around the blob in that channel
    reaches the channel's threshold.

    Args:
        roi: ``(z, y, x, c)`` region of interest (numpy or CUDA tensor).
        blobs: ``(n, >=7)`` blob table; blobs outside the ROI get zeros.
        thresh: percentile of the intensities around all blobs of a channel, or "min"
            (the default) for the smallest per-blob average of the channel.

    Returns:
        ``(n, c)`` uint8 array, or None without blobs or without a channel axis.
    """
    if blobs is None or roi is None or len(roi.shape) < 4:
        return None
    if thresh is None:
        thresh = "min"
    blobs_roi, blobs_roi_mask = detector.get_blobs_in_roi(
        blobs, (0, 0, 0), roi.shape[:3], reverse=False)
    blobs_chl = detector.Blobs.get_blobs_channel(blobs_roi)
    n_chl = int(roi.shape[3])
    means, counts = blob_surround_means(roi, blobs_roi)
    threshs = []
    for chl in range(n_chl):
        idx = np.where(np.isin(blobs_chl, chl))[0]
        if thresh == "min":
            # np.mean of an empty selection is nan, which np.amin propagates
            threshs.append(None if len(idx) == 0 else np.amin(means[idx, chl]))
        else:
            threshs.append(_percentile_under_blobs(roi, blobs_roi, idx, chl, thresh))
    channels = np.unique(blobs_chl).astype(int)
    colocs_roi = np.zeros((blobs_roi.shape[0], n_chl), dtype=np.uint8)
    for chl in channels:
        idx = np.where(np.isin(blobs_chl, chl))[0]
        for chl_other in channels:
            if threshs[chl_other] is None:
                continue
            with np.errstate(invalid="ignore"):
                colocs_roi[idx, chl_other] = means[idx, chl_other] >= threshs[chl_other]
    colocs = np.zeros((blobs.shape[0], n_chl), dtype=np.uint8)
    colocs[blobs_roi_mask] = colocs_roi
    return colocs


def _percentile_under_blobs(roi, blobs_roi, idx, chl, thresh):
    """``np.percentile(roi[mask >= 0, chl], thresh)`` (colocalizer.py:405-410): the voxels
    under the balls of the channel's blobs, gathered on the host (a few tens of voxels per
    blob); the whole ROI when the channel has no blob."""
    host = roi if isinstance(roi, np.ndarray) else roi.cpu().numpy()
    if len(idx) == 0:
        return np.percentile(host, thresh)
    Z, Y, X = host.shape[:3]
    r = np.arange(-2, 3)
    off = np.array([(a, b, c) for a in r for b in r for c in r if a * a + b * b + c * c <= 4])
    pos = (blobs_roi[idx, :3].astype(int)[:, None, :] + off[None]).reshape(-1, 3)
    ok = np.all((pos >= 0) & (pos < [Z, Y, X]), axis=1)
    lin = np.unique(np.ravel_multi_index(pos[ok].T, (Z, Y, X)))
    return np.percentile(host[..., chl].reshape(-1)[lin], thresh)
