"""Module-level settings read by the detection path at call time.

Mirror of the globals of ``magmap/settings/config.py`` that the hot path reads:
``resolutions`` (:246), ``near_max`` (:211), ``cpus`` (:79), ``channel`` (:144),
``roi_profile`` / ``roi_profiles`` / ``get_roi_profile`` (:883-902),
``filename`` (:131), ``SUFFIX_BLOBS`` (:126), ``save_subimg`` (:508),
``truth_db_mode`` (:539), ``grid_search_profile`` (:905), ``verbose`` (:108).
A caller of the reference sets these the same way and then calls
``cv.detector.detect_blobs`` / ``cv.stack_detect.detect_blobs_stack``.
"""
from __future__ import annotations

import logging
from enum import Enum, auto
from typing import Any, Dict, List, Optional, Sequence

logger = logging.getLogger("magellanmapper_b200")

#: print verbose diagnostics
verbose: bool = False
#: worker count of the reference's CPU pool; unused by the GPU path, kept so
#: that callers that set it keep working
cpus: Optional[int] = None
#: base path of the current image; only its basename is stored in archives
filename: Optional[str] = None
#: channel(s) of interest, None = all
channel: Optional[Sequence[int]] = None
#: image resolutions ``[[z, y, x], ...]`` in physical units per voxel
resolutions: Optional[Sequence[Sequence[float]]] = None
#: per-channel near-maximum intensity of the whole image (importer metadata)
near_max: List[float] = [-1.0]
near_min: List[float] = [0.0]

#: objective magnification and zoom of the loaded image (importer metadata)
magnification = 1.0
zoom = 1.0


class MetaKeys(Enum):
    """Keys of the image-import metadata dictionary (config.py:227-238), as
    ``np_io.write_npy`` takes them."""
    RESOLUTIONS = auto()
    MAGNIFICATION = auto()
    ZOOM = auto()
    SHAPE = auto()
    DTYPE = auto()


#: metadata for image import (config.py:241-242)
meta_dict: Dict[MetaKeys, Any] = dict.fromkeys(MetaKeys, None)

SUFFIX_IMAGE5D = "image5d.npy"
SUFFIX_META = "meta.yml"
SUFFIX_BLOBS = "blobs.npz"
SUFFIX_SUBIMG = "subimg.npy"
save_subimg: bool = False
truth_db_mode = None
grid_search_profile = None

#: default ROI profile and optional per-channel profiles
roi_profile = None
roi_profiles: list = []


def get_roi_profile(i: int):
    """Profile for channel ``i``; the default profile when fewer per-channel
    profiles than channels were configured (config.py:887-902)."""
    if len(roi_profiles) > i:
        return roi_profiles[i]
    return roi_profile


def near_max_for(chl: int) -> float:
    """``near_max[chl]`` with the reference's single-element broadcast."""
    if chl < len(near_max):
        return float(near_max[chl])
    return float(near_max[0]) if len(near_max) else -1.0


# ---- GPU execution settings (new surface; no reference equivalent) ----------
#: CUDA device index; None = torch's current device
gpu_device: Optional[int] = None
