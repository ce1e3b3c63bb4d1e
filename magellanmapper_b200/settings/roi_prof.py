"""ROI profile: detection, preprocessing and block-sizing keys.

Mirror of ``magmap/settings/roi_prof.py`` restricted to the keys the blob
detection path reads (defaults ``:72-134``) and the named modifiers that touch
them (``:147-334``).  Visualisation-only keys are omitted.
"""
from __future__ import annotations

from typing import Dict, Optional

from . import profiles


class ROIProfile(profiles.SettingsDict):
    PATH_PREFIX = "roi"
    BLOB_PREPROCESSING = ("clip_vmin", "clip_vmax", "clip_min", "clip_max",
                          "max_thresh_factor", "tot_var_denoise", "unsharp_strength",
                          "erosion_threshold", "adapt_hist_lim")
    BLOCK_SIZES = ("segment_size", "denoise_size", "prune_tol_factor",
                   "sub_stack_max_pixels", "isotropic")

    _DEFAULTS = dict(
        # preprocessing (roi_prof.py:72-84)
        clip_vmin=5, clip_vmax=99.5, clip_min=0.2, clip_max=1.0,
        max_thresh_factor=0.5, tot_var_denoise=None, unsharp_strength=0.3,
        erosion_threshold=0.2, adapt_hist_lim=0.1,
        # detection (roi_prof.py:86-97)
        min_sigma_factor=3, max_sigma_factor=5, num_sigma=10,
        detection_threshold=0.1, overlap=0.5, thresholding=None,
        thresholding_size=-1, exclude_border=None,
        # block processing (roi_prof.py:99-134)
        mp_start="fork", mp_max_tasks=None, segment_size=500, denoise_size=25,
        prune_tol_factor=(1, 1, 1), verify_tol_factor=(1, 1, 1),
        sub_stack_max_pixels=(1000, 1000, 1000), isotropic=None,
        isotropic_vis=(1, 1, 1), resize_blobs=None,
    )

    _MODIFIERS: Dict[str, Dict] = {
        "lightsheet": dict(
            clip_vmax=98.5, clip_min=0, clip_max=0.5, unsharp_strength=0.3,
            erosion_threshold=0.3, min_sigma_factor=2.6, max_sigma_factor=2.8,
            num_sigma=10, overlap=0.55, segment_size=150,
            prune_tol_factor=(1, 0.9, 0.9), verify_tol_factor=(3, 1.2, 1.2),
            isotropic=(0.96, 1, 1), isotropic_vis=(0.5, 1, 1),
            sub_stack_max_pixels=(1200, 800, 800), exclude_border=(1, 0, 0)),
        "minpreproc": dict(clip_vmin=0, clip_vmax=99.99, clip_max=1,
                           tot_var_denoise=0.01, unsharp_strength=0, erosion_threshold=0),
        "lowres": dict(min_sigma_factor=10, max_sigma_factor=14, isotropic=None,
                       denoise_size=2000, segment_size=1000, max_thresh_factor=1.5,
                       exclude_border=(8, 1, 1), verify_tol_factor=(3, 2, 2)),
        "2p20x": dict(clip_vmax=97, clip_min=0, clip_max=0.7, tot_var_denoise=True,
                      unsharp_strength=2.5, min_sigma_factor=2.6, max_sigma_factor=4,
                      num_sigma=20, overlap=0.1, thresholding=None, thresholding_size=64,
                      denoise_size=25, segment_size=100, prune_tol_factor=(1.5, 1.3, 1.3)),
        "zebrafish": dict(min_sigma_factor=2.5, max_sigma_factor=3),
        "cytoplasm": dict(clip_min=0.3, clip_max=0.8, min_sigma_factor=4,
                          max_sigma_factor=10, num_sigma=10, overlap=0.2),
        "binary": dict(denoise_size=None, detection_threshold=0.001),
        "4xnuc": dict(min_sigma_factor=3, max_sigma_factor=4),
        "20x": dict(segment_size=50),
        "exportdl": dict(isotropic=(0.93, 1, 1)),
        "downiso": dict(isotropic=None, resize_blobs=(.2, 1, 1)),
        "register": dict(unsharp_strength=1.5),
        "atlas": dict(clip_vmax=97),
        "spawn": dict(mp_start="spawn"),
        # display-only modifiers (roi_prof.py:223-330): nothing on the detection path reads
        # their keys, but a profile string that names them must layer and be recorded in
        # ``settings_name`` as the reference records it
        "isotropic": dict(points_3d_thresh=0.3, isotropic_vis=(1, 1, 1)),
        "contrast": dict(channel_colors=("inferno", "inferno"), scale_bar_color="w"),
        "bone": dict(channel_colors=("bone", "bone"), scale_bar_color="w"),
        "diverging": dict(channel_colors=("RdBu", "BrBG"), scale_bar_color="k",
                          colorbar=dict(shrink=0.7)),
        "randomcolors": dict(channel_colors=[]),
        "norm": dict(norm=(0.0, 1.0)),
        "rot180": dict(load_rot90=2),
    }

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in self._DEFAULTS.items():
            self[k] = v
        #: ``{channel: {channel_to_subtract: factor}}`` or None (roi_prof.py:141-145)
        self.spectral_unmixing: Optional[Dict[int, Dict[int, float]]] = None
        self.update(*args, **kwargs)
        self.profiles = {k: dict(v) for k, v in self._MODIFIERS.items()}
