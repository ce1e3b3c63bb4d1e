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

from typing import List, Optional, Sequence, Tuple

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


# ---- image + metadata files (importer.py:272-301, 482-522, 606-745) -----------------

#: version number written into the metadata file (importer.py:69)
IMAGE5D_NP_VER = 15


def filename_to_base(filename: str, series: Optional[int] = None, modifier: str = "",
                     keep_ext: bool = False) -> str:
    """Image path -> base path: the extension dropped unless ``keep_ext``, ``modifier``
    appended after ``_``; ``series`` is ignored, as in the reference (importer.py:304-326)."""
    from . import libmag
    base = filename if keep_ext else libmag.splitext(filename)[0]
    return libmag.combine_paths(base, modifier, keep_ext=True) if modifier else base


def make_filenames(filename: str, series: Optional[int] = None, modifier: str = "",
                   keep_ext: bool = False) -> Tuple[str, str]:
    """``(<base>_image5d.npy, <base>_meta.yml)`` for an image path (importer.py:272-301)."""
    from . import libmag
    from ..settings import config
    base = filename_to_base(filename, series, modifier, keep_ext)
    return (libmag.combine_paths(base, config.SUFFIX_IMAGE5D, keep_ext=True),
            libmag.combine_paths(base, config.SUFFIX_META, keep_ext=True))


def _primitive(val):
    """Numpy scalars / arrays / tuples -> Python primitives and lists (yaml_io.save_yaml
    with ``use_primitives``)."""
    if isinstance(val, dict):
        return {k: _primitive(v) for k, v in val.items()}
    if isinstance(val, (list, tuple, np.ndarray)):
        return [_primitive(v) for v in val]
    try:
        return val.item()
    except AttributeError:
        return val


def save_image_info(filename_info_npz: str, names, sizes, resolutions, magnification, zoom,
                    near_min, near_max, scaling=None, plane=None) -> dict:
    """Write the image metadata as the reference's ``*_meta.yml`` (importer.py:482-522)."""
    import yaml
    data = _primitive({
        "ver": IMAGE5D_NP_VER, "names": names, "sizes": sizes, "resolutions": resolutions,
        "magnification": magnification, "zoom": zoom, "near_min": near_min,
        "near_max": near_max, "scaling": scaling, "plane": plane})
    with open(filename_info_npz, "w") as f:
        yaml.dump(data, f)
    return data


def assign_metadata(img5d, md: dict) -> None:
    """Metadata dictionary -> ``img5d.shapes`` and the ``config`` globals the path reads
    (resolutions, magnification, zoom, ``near_min``, ``near_max``), with the first series'
    values under the ``config.MetaKeys`` entries (importer.py:671-745)."""
    from ..settings import config

    def first(val):
        try:
            return val[0]
        except (TypeError, IndexError, KeyError):
            return None
    if "sizes" in md:
        img5d.shapes = md["sizes"]
        md[config.MetaKeys.SHAPE] = first(img5d.shapes)
    if "resolutions" in md:
        config.resolutions = np.array(md["resolutions"])
        md[config.MetaKeys.RESOLUTIONS] = first(config.resolutions) \
            if np.ndim(config.resolutions) else None
    if "magnification" in md:
        config.magnification = md["magnification"]
        md[config.MetaKeys.MAGNIFICATION] = first(config.magnification)
    if "zoom" in md:
        config.zoom = md["zoom"]
        md[config.MetaKeys.ZOOM] = first(config.zoom)
    if "near_min" in md:
        config.near_min = md["near_min"]
    if "near_max" in md:
        config.near_max = md["near_max"]


def load_metadata(path: str, check_ver: bool = False, img5d=None):
    """Read ``*_meta.yml`` (or the older ``.npz`` next to it); with ``img5d``, assign its
    values to the image object and to ``config`` - unless ``check_ver`` and the file's
    version is older than ``IMAGE5D_NP_VER`` (importer.py:606-668).
    Returns ``(metadata dict or None, version number or -1)``."""
    import os
    import yaml
    from . import np_io
    from ..settings import config
    try:
        with open(path) as f:
            docs = list(yaml.load_all(f, Loader=yaml.FullLoader))
        output = docs[0] if docs else None
        if output:
            output.update(dict.fromkeys(config.MetaKeys, None))
    except FileNotFoundError:
        try:
            output = np_io.read_np_archive(np.load(f"{os.path.splitext(path)[0]}.npz"))
        except FileNotFoundError:
            return None, -1
    if not output:
        return None, -1
    ver = output.get("ver", -1)
    ver = int(ver) if ver is not None else -1
    if img5d is not None and (not check_ver or ver >= IMAGE5D_NP_VER):
        img5d.meta = output
        assign_metadata(img5d, output)
    return output, ver
