"""Preprocessing entry points (mirror of ``magmap/plot/plot_3d.py:24-172``).

``saturate_roi`` and ``denoise_roi`` keep the reference signatures and read the
same profile keys, but run on the GPU through ``mmb_preprocess_blocks``.  Called
on their own they treat the whole ROI as one block, as the reference does
(the GUI path, ``magmap/gui/visualizer.py:2739-2743``); inside the stack
detector the two are fused per 25^3 block in one kernel.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from ..settings import config


def setup_channels(roi, channel: Optional[Sequence[int]], dim_channel: int
                   ) -> Tuple[bool, Sequence[int]]:
    """``(multichannel, channels)`` for an ROI whose channel axis, if present,
    is ``dim_channel`` (plot_3d.py:24-52)."""
    multichannel = roi.ndim > dim_channel
    if not multichannel:
        return False, [0]
    if channel is None:
        return True, range(roi.shape[dim_channel])
    return True, channel


def preproc_params(settings, chl: int, clip_vmin=-1, clip_vmax=-1, max_thresh_factor=-1,
                   saturate: bool = True, denoise: bool = True):
    """Profile keys -> ``mmb_preproc_params``.  ``saturate=False`` /
    ``denoise=False`` switch the corresponding half off."""
    from .._lib import MmbPreprocParams
    if settings["tot_var_denoise"] and denoise:
        raise NotImplementedError(
            "tot_var_denoise (total-variation denoising) is outside the accelerated path")
    vmin = settings["clip_vmin"] if clip_vmin == -1 else clip_vmin
    vmax = settings["clip_vmax"] if clip_vmax == -1 else clip_vmax
    mtf = settings["max_thresh_factor"] if max_thresh_factor == -1 else max_thresh_factor
    return MmbPreprocParams(
        clip_vmin=float(vmin), clip_vmax=float(vmax),
        max_thresh=float(config.near_max_for(chl) * mtf),
        clip_min=float(settings["clip_min"]), clip_max=float(settings["clip_max"]),
        unsharp_strength=float(settings["unsharp_strength"] or 0.0) if denoise else 0.0,
        erosion_threshold=float(settings["erosion_threshold"] or 0.0) if denoise else 0.0)


def _run_whole_roi(roi, channel, make_params) -> np.ndarray:
    from .. import gpu
    multichannel, channels = setup_channels(roi, channel, 3)
    out = None
    for chl in channels:
        settings = config.get_roi_profile(chl)
        src = gpu.as_source(roi, chl if multichannel else None)
        res = gpu.whole_roi_preprocess(src, make_params(settings, chl))
        res = res.astype(np.float64)
        if multichannel:
            if out is None:
                out = np.zeros(roi.shape, dtype=res.dtype)
            out[..., chl] = res
        else:
            out = res
    return out


def saturate_roi(roi, clip_vmin: float = -1, clip_vmax: float = -1,
                 max_thresh_factor: float = -1,
                 channel: Optional[Sequence[int]] = None) -> np.ndarray:
    """Clip to the profile's percentiles and stretch to [0, 1] (plot_3d.py:55-111).
    -1 takes the value from each channel's profile."""
    def mk(settings, chl):
        p = preproc_params(settings, chl, clip_vmin, clip_vmax, max_thresh_factor,
                           denoise=False)
        p.clip_min, p.clip_max = -np.inf, np.inf
        return p
    return _run_whole_roi(roi, channel, mk)


def denoise_roi(roi, channel: Optional[Sequence[int]] = None) -> np.ndarray:
    """Clip, unsharp-mask and conditionally erode (plot_3d.py:114-172)."""
    def mk(settings, chl):
        p = preproc_params(settings, chl, saturate=False)
        # percentiles 0 and 0 make vmin == vmax: the stretch is skipped
        p.clip_vmin = p.clip_vmax = 0.0
        return p
    return _run_whole_roi(roi, channel, mk)
