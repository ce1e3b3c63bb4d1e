"""Intensity-bound metadata of an image (mirror of the part of
``magmap/io/importer.py`` that the blob-detection path depends on).

``saturate_roi`` stretches every preprocessing block up to at least
``config.near_max[channel] * max_thresh_factor`` (``magmap/plot/plot_3d.py:97-100``).
The reference fills ``near_min`` / ``near_max`` at import time: per channel, the
minimum over z-planes of each plane's 0.5th percentile and the maximum of its
99.5th (``importer.py:1368-1377``, ``:571-583``, ``calc_near_intensity_bounds``
``:1447-1468``).  Volumes that reach the detector without that metadata (raw
stacks, device-resident data) get it here from the GPU: two histogram passes over
the image (``mmb_percentiles``), no host copy of the voxels.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from ..plot import plot_3d


def _channel_sources(image, dim_channel: int):
    """(z, y, x) device views, one per channel, of an array whose channel axis (if
    any) is ``dim_channel``; leading singleton axes (time) are dropped and a 2-D
    plane becomes one z-plane."""
    from .. import gpu
    multichannel, channels = plot_3d.setup_channels(image, None, dim_channel)
    arr = image
    while arr.ndim - (1 if multichannel else 0) > 3:
        if arr.shape[0] != 1:
            raise ValueError(f"cannot reduce an array of shape {tuple(image.shape)} to z, y, x")
        arr = arr[0]
    while arr.ndim - (1 if multichannel else 0) < 3:
        arr = arr[None]
    return [gpu.as_source(arr, c if multichannel else None) for c in channels]


def calc_intensity_bounds(image5d, lower: float = 0.5, upper: float = 99.5,
                          dim_channel: int = 4) -> Tuple[List[float], List[float]]:
    """Percentile bounds of a whole image, one pair per channel
    (``importer.py:1415-1444``): ``(lows, highs)`` lists of float64."""
    from .. import gpu
    lows, highs = [], []
    for src in _channel_sources(image5d, dim_channel):
        lo, hi = gpu.percentiles(src, (lower, upper), per_plane=False)
        lows.append(np.float64(lo))
        highs.append(np.float64(hi))
    return lows, highs


def calc_near_intensity_bounds(near_mins, near_maxs, lows, highs):
    """Extremes over a list of per-plane bounds (``importer.py:1447-1468``): with one
    channel the minimum / maximum are APPENDED to the given lists, with several they
    REPLACE them by per-channel arrays - the reference's behaviour, kept as is."""
    if lows:
        num_channels = len(lows[0])
        if num_channels <= 1:
            near_mins.append(min(lows)[0])
            near_maxs.append(max(highs)[0])
        else:
            near_mins = np.amin(np.array(lows), 0)
            near_maxs = np.amax(np.array(highs), 0)
    return near_mins, near_maxs


def calc_plane_bounds(image, lower: float = 0.5, upper: float = 99.5, dim_channel: int = 3
                      ) -> Tuple[List[List[float]], List[List[float]]]:
    """The reference's per-plane loop (``importer.py:575-581``:
    ``calc_intensity_bounds(image5d[0, i], dim_channel=2)`` for every plane ``i``) in
    one launch per channel: ``(lows, highs)``, each a list over planes of a list over
    channels, ready for ``calc_near_intensity_bounds``."""
    from .. import gpu
    per_chl = [gpu.percentiles(src, (lower, upper), per_plane=True)
               for src in _channel_sources(image, dim_channel)]
    n_planes = per_chl[0].shape[0]
    lows = [[np.float64(p[z, 0]) for p in per_chl] for z in range(n_planes)]
    highs = [[np.float64(p[z, 1]) for p in per_chl] for z in range(n_planes)]
    return lows, highs


def calc_near_bounds(image, dim_channel: int = 3):
    """``(near_mins, near_maxs)`` of a (z, y, x[, c]) image the way the importer
    records them: per-plane 0.5 / 99.5 percentiles reduced over the planes."""
    lows, highs = calc_plane_bounds(image, dim_channel=dim_channel)
    return calc_near_intensity_bounds([], [], lows, highs)
