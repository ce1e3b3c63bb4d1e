"""Blob detection on the GPU behind the ``magmap.cv.detector`` surface.

Mirror of ``magmap/cv/detector.py``: the ``Blobs`` table container and its
column schema (``:46-807``), ``calc_scaling_factor`` / ``calc_overlap``
(``:810-841``), ``detect_blobs`` (``:874-957``), seam pruning
``remove_close_blobs`` (``:1000-1085``) and the ROI filters
(``:1210-1268``).  Same names, argument order and array layouts; the arithmetic
(``skimage.feature.blob_log`` in the reference) runs in ``libmmb200.so``.
"""
from __future__ import annotations

import math
from enum import Enum
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from ..io import libmag, np_io, npz_writer
from ..plot import plot_3d
from ..settings import config

#: blob confirmation flags
CONFIRMATION: Dict[int, str] = {-1: "unverified", 0: "no", 1: "yes", 2: "maybe"}
#: pixels of sub-ROI overlap per unit of scaling
OVERLAP_FACTOR: int = 5

_logger = config.logger.getChild(__name__)


class Blobs:
    """Blob table ``[[z, y, x, radius, confirmed, truth, channel, abs_z, abs_y,
    abs_x, region], ...]`` plus archive metadata.

    Column positions live in the CLASS attribute ``_col_inds`` and are rewritten
    whenever ``cols`` is assigned (as in the reference, ``detector.py:116,
    142-162``), so the class-level accessors always describe the most recently
    configured table.
    """

    #: archive version (5: abs-coordinate columns removed from stored tables)
    BLOBS_NP_VER: int = 5

    class Keys(Enum):
        VER = "ver"
        BLOBS = "segments"
        COLOCS = "colocs"
        RESOLUTIONS = "resolutions"
        BASENAME = "basename"
        ROI_OFFSET = "offset"
        ROI_SIZE = "roi_size"
        COLS = "columns"

    class Cols(Enum):
        Z = "z"
        Y = "y"
        X = "x"
        RADIUS = "radius"
        CONFIRMED = "confirmed"
        TRUTH = "truth"
        CHANNEL = "channel"
        ABS_Z = "abs_z"
        ABS_Y = "abs_y"
        ABS_X = "abs_x"
        REGION = "region"

    _col_inds: Dict["Blobs.Cols", Optional[int]] = {c: i for i, c in enumerate(Cols)}

    def __init__(self, blobs=None, blob_matches=None, colocalizations=None, path=None,
                 cols=None):
        self.cols = cols
        self.blobs = blobs
        self.blob_matches = blob_matches
        self.colocalizations = colocalizations
        self.path = path
        self.ver = self.BLOBS_NP_VER
        self.roi_offset = None
        self.roi_size = None
        self.resolutions = None
        self.basename = None
        self.scaling = np.ones(3)

    # -- schema -----------------------------------------------------------------
    @property
    def cols(self):
        return self._cols

    @cols.setter
    def cols(self, cols):
        self._cols = cols
        if cols is None:
            return
        inds = {c: None for c in self.Cols}
        for i, name in enumerate(cols):
            try:
                inds[self.Cols(name)] = i
            except ValueError:
                _logger.warning("%s is not a valid Blobs column, skipping", name)
        Blobs._col_inds = inds

    @property
    def blobs(self):
        return self._blobs

    @blobs.setter
    def blobs(self, blobs):
        self._blobs = blobs
        if blobs is not None and self.cols is None:
            self.cols = [c.value for c in self.Cols][:blobs.shape[1]]

    def format_blobs(self, channel=None) -> np.ndarray:
        """Widen ``[z, y, x, radius, ...]`` to every column of ``Cols`` (new
        columns = -1), copy relative into absolute coordinates and optionally
        set the channel (detector.py:325-364)."""
        n, have = self.blobs.shape
        pad = np.full((n, len(self.Cols) - have), -1.0)
        self.blobs = np.concatenate((self.blobs, pad), axis=1)
        self.cols = [c.value for c in self.Cols]
        self.blobs[:, self._get_abs_inds()] = self.blobs[:, self._get_rel_inds()]
        if channel is not None:
            self.set_blob_channel(self.blobs, channel)
        return self.blobs

    # -- archive ------------------------------------------------------------------
    def load_blobs(self, path: Optional[str] = None) -> "Blobs":
        """Read a ``*_blobs.npz`` archive (detector.py:185-267)."""
        if path is not None:
            self.path = path
        # no pickle: keys that would need it (the None-valued metadata saved as object
        # arrays) are skipped by read_np_archive, as in the reference (np_io.py:159-177)
        with np.load(self.path) as archive:
            info = np_io.read_np_archive(archive)
        K = self.Keys
        if K.VER.value in info:
            self.ver = int(info[K.VER.value])
        if K.COLS.value in info:
            self.cols = [str(c) for c in info[K.COLS.value]]
        if K.BLOBS.value in info:
            self.blobs = info[K.BLOBS.value]
        for key, attr in ((K.COLOCS, "colocalizations"), (K.RESOLUTIONS, "resolutions"),
                          (K.BASENAME, "basename"), (K.ROI_OFFSET, "roi_offset"),
                          (K.ROI_SIZE, "roi_size")):
            if key.value in info:
                val = info[key.value]
                if isinstance(val, np.ndarray) and val.dtype == object and val.ndim == 0:
                    val = val.item()
                setattr(self, attr, val)
        if self.ver <= 4 and self.cols is not None:
            # v4 archives listed abs-coordinate names for columns that had been dropped
            self.cols = self.cols[:len(self.cols) - 3]
        self.ver = self.BLOBS_NP_VER
        return self

    def save_archive(self, to_add=None, update: bool = False):
        """Write the archive with the reference's keys, backing up an existing
        file first (detector.py:269-323)."""
        if to_add is None:
            present = {k: v for k, v in self._col_inds.items() if v is not None}
            K = self.Keys
            arc = {
                K.VER.value: self.ver, K.BLOBS.value: self.blobs,
                K.RESOLUTIONS.value: self.resolutions, K.BASENAME.value: self.basename,
                K.ROI_OFFSET.value: self.roi_offset, K.ROI_SIZE.value: self.roi_size,
                K.COLOCS.value: self.colocalizations,
                K.COLS.value: [k.value for k, _ in sorted(present.items(), key=lambda e: e[1])],
            }
        else:
            arc = to_add
        if update:
            with np.load(self.path) as archive:
                arc = np_io.read_np_archive(archive)
                arc.update(to_add)
        libmag.backup_file(self.path)
        # numpy.savez's format, written by a few threads (io/npz_writer.py): the table of a
        # whole stack is the longest host-side stage after the detection itself
        npz_writer.savez(self.path, arc, add_suffix=False)
        return arc

    # -- column accessors -----------------------------------------------------------
    @classmethod
    def _get_col_as_ind(cls, col):
        if libmag.is_seq(col):
            return [cls._col_inds[c] if isinstance(c, cls.Cols) else c for c in col]
        return cls._col_inds[col] if isinstance(col, cls.Cols) else col

    @classmethod
    def _get_rel_inds(cls) -> List[int]:
        return [cls._col_inds[c] for c in (cls.Cols.Z, cls.Cols.Y, cls.Cols.X)]

    @classmethod
    def _get_abs_inds(cls) -> List[int]:
        return [cls._col_inds[c] for c in (cls.Cols.ABS_Z, cls.Cols.ABS_Y, cls.Cols.ABS_X)]

    @classmethod
    def get_blob_col(cls, blob: np.ndarray, col):
        if col is None:
            return np.array([]) if blob.ndim > 1 else None
        col = cls._get_col_as_ind(col)
        return blob[..., col] if blob.ndim > 1 else blob[col]

    @classmethod
    def set_blob_col(cls, blob: np.ndarray, col, val, mask=np.s_[:], **kwargs) -> np.ndarray:
        col = cls._get_col_as_ind(col)
        if blob.ndim > 1:
            blob[mask, ..., col] = val
        else:
            blob[col] = val
        return blob

    @classmethod
    def get_blob_confirmed(cls, blob):
        return cls.get_blob_col(blob, cls._col_inds[cls.Cols.CONFIRMED])

    @classmethod
    def set_blob_confirmed(cls, blob, *args, **kwargs):
        return cls.set_blob_col(blob, cls._col_inds[cls.Cols.CONFIRMED], *args, **kwargs)

    @classmethod
    def get_blob_truth(cls, blob):
        return cls.get_blob_col(blob, cls._col_inds[cls.Cols.TRUTH])

    @classmethod
    def set_blob_truth(cls, blob, *args, **kwargs):
        return cls.set_blob_col(blob, cls._col_inds[cls.Cols.TRUTH], *args, **kwargs)

    @classmethod
    def get_blobs_channel(cls, blob):
        return cls.get_blob_col(blob, cls._col_inds[cls.Cols.CHANNEL])

    @classmethod
    def set_blob_channel(cls, blob, *args, **kwargs):
        return cls.set_blob_col(blob, cls._col_inds[cls.Cols.CHANNEL], *args, **kwargs)

    @classmethod
    def get_blob_abs_coords(cls, blobs):
        return cls.get_blob_col(blobs, cls._get_abs_inds())

    @classmethod
    def set_blob_abs_coords(cls, blobs, coords, *args, **kwargs):
        cls.set_blob_col(blobs, cls._get_abs_inds(), coords, *args, **kwargs)
        return blobs

    # -- coordinate shifts ------------------------------------------------------------
    @classmethod
    def shift_blobs(cls, blob, cols, fn: Callable, vals, to_int: bool = False):
        if blob is None:
            return blob
        sub = fn(blob[cols] if blob.ndim == 1 else blob[..., cols], vals)
        if to_int:
            sub = sub.astype(int)
        if blob.ndim == 1:
            blob[cols] = sub
        else:
            blob[..., cols] = sub
        return blob

    @classmethod
    def shift_blob_rel_coords(cls, blob, offset):
        return cls.shift_blobs(blob, cls._get_rel_inds(), np.add, offset)

    @classmethod
    def shift_blob_abs_coords(cls, blob, offset):
        return cls.shift_blobs(blob, cls._get_abs_inds(), np.add, offset)

    @classmethod
    def multiply_blob_rel_coords(cls, blob, factor):
        return cls.shift_blobs(blob, cls._get_rel_inds(), np.multiply, factor, True)

    @classmethod
    def multiply_blob_abs_coords(cls, blob, factor):
        return cls.shift_blobs(blob, cls._get_abs_inds(), np.multiply, factor, True)

    def remove_abs_blob_coords(self, remove_extra: bool = False) -> np.ndarray:
        """Drop the abs-coordinate columns; ``remove_extra`` also drops columns
        not named in ``Cols`` (detector.py:711-731)."""
        inds = Blobs._col_inds.values() if remove_extra else range(self.blobs.shape[1])
        abs_inds = Blobs._get_abs_inds()
        keep = [i for i in inds if i is not None and i not in abs_inds]
        new_cols = [self.cols[i] for i in keep]
        self.blobs = self.blobs[:, keep]
        self.cols = new_cols
        return self.blobs

    @classmethod
    def replace_rel_with_abs_blob_coords(cls, blobs: np.ndarray) -> np.ndarray:
        blobs[:, cls._get_rel_inds()] = blobs[:, cls._get_abs_inds()]
        return blobs

    @classmethod
    def blobs_in_channel(cls, blobs, channel, return_mask: bool = False):
        mask = None
        sel = blobs
        if channel is not None:
            mask = np.isin(cls.get_blobs_channel(blobs), channel)
            sel = blobs[mask]
        return (sel, mask) if return_mask else sel

    @classmethod
    def show_blobs_per_channel(cls, blobs):
        for chl in np.unique(cls.get_blobs_channel(blobs)):
            _logger.info("- blobs in channel %s: %s", int(chl),
                         len(cls.blobs_in_channel(blobs, chl)))

    @classmethod
    def blob_for_db(cls, blob: np.ndarray) -> np.ndarray:
        """``abs_z, abs_y, abs_x, radius, confirmed, truth, channel``."""
        rest = [cls._col_inds[c] for c in (cls.Cols.RADIUS, cls.Cols.CONFIRMED,
                                           cls.Cols.TRUTH, cls.Cols.CHANNEL)]
        return np.array([*blob[cls._get_abs_inds()], *blob[rest]])


def calc_scaling_factor() -> np.ndarray:
    """Pixels per physical unit, ``1 / config.resolutions[0]``.

    Raises:
        AttributeError: if no resolution is set (detector.py:821-823).
    """
    if config.resolutions is None or len(config.resolutions) < 1:
        raise AttributeError("Must load resolutions from file or set a resolution")
    return np.divide(1.0, config.resolutions[0])


def calc_overlap(factor: Optional[int] = None) -> np.ndarray:
    """Chunk overlap in pixels, ``ceil(scaling * factor)`` (detector.py:828-841)."""
    if factor is None:
        factor = OVERLAP_FACTOR
    return np.ceil(np.multiply(calc_scaling_factor(), factor)).astype(int)


def sigma_ladder(settings, scaling_factor: float, image_is_f32: bool = False) -> np.ndarray:
    """The ``blob_log`` scale ladder for a profile (detector.py:903-927).

    scikit-image 0.25.2 (``skimage/feature/blob.py``, ``blob_log``) casts the two
    scalar sigmas to the image's float dtype and builds the ladder as
    ``np.linspace(0, 1, num_sigma)[:, None] * (max_sigma - min_sigma) + min_sigma``,
    not as ``np.linspace(min_sigma, max_sigma, num_sigma)``: the two differ in the
    last bits whenever ``max - min`` is not a power of two (8.9e-16 for 4..10),
    and the ``radius`` column is ``sigma * sqrt(3)`` of exactly these values.  A
    float32 image gets float32-rounded ends and a float32 difference."""
    dt = np.float32 if image_is_f32 else np.float64
    lo = np.asarray(settings["min_sigma_factor"] * scaling_factor, dtype=dt)
    hi = np.asarray(settings["max_sigma_factor"] * scaling_factor, dtype=dt)
    return skimage_ladder(lo, hi, int(settings["num_sigma"]))


def skimage_ladder(lo, hi, num_sigma: int) -> np.ndarray:
    """``scale * (max_sigma - min_sigma) + min_sigma`` with ``scale =
    np.linspace(0, 1, num_sigma)`` (float64) and the difference taken in the dtype
    of ``lo`` / ``hi``, as ``blob_log`` does."""
    lo, hi = np.asarray(lo), np.asarray(hi)
    scale = np.linspace(0, 1, int(num_sigma))
    return (scale * (hi - lo) + lo).astype(np.float64)


def img_as_float_host(arr: np.ndarray) -> np.ndarray:
    """``skimage.util.img_as_float`` on the host for the dtypes the kernels do not read in
    place (bool, signed and 32/64-bit integers, float16): bool -> 0 / 1, unsigned ->
    ``x / max``, signed -> ``(2 x + 1) / (max - min)``, all in float64."""
    dt = arr.dtype
    if dt.kind == "f":
        return arr.astype(np.float32 if dt == np.float16 else dt)
    if dt.kind == "b":
        return arr.astype(np.float64)
    if dt.kind == "u":
        return arr.astype(np.float64) / float(np.iinfo(dt).max)
    if dt.kind == "i":
        info = np.iinfo(dt)
        out = arr.astype(np.float64)
        out *= 2.0
        out += 1.0
        out /= float(info.max) - float(info.min)
        return out
    raise TypeError(f"unsupported image dtype {dt}")


def input_scale(dtype) -> float:
    """``skimage.util.img_as_float`` factor for an input dtype."""
    dtype = np.dtype(dtype)
    if dtype.kind == "u":
        return 1.0 / float(np.iinfo(dtype).max)
    return 1.0


def cands_to_blobs(cands: np.ndarray, sigmas: np.ndarray, shape_yx: Sequence[int],
                   chl: int) -> np.ndarray:
    """Device candidates -> the reference's formatted table, in
    ``peak_local_max`` order (descending response, ties in C order)."""
    Y, X = shape_yx
    lin = ((cands["z"].astype(np.int64) * Y + cands["y"]) * X + cands["x"]) * len(sigmas) \
        + cands["s"]
    order = np.lexsort((lin, -cands["resp"].astype(np.float64)))
    c = cands[order]
    table = np.column_stack([c["z"], c["y"], c["x"], sigmas[c["s"]] * math.sqrt(3)]
                            ).astype(np.float64)
    return Blobs(table).format_blobs(chl)


def _unmixing_of(settings, chl: int):
    """``{channel to subtract: factor}`` for ``chl`` (detector.py:910-921), or {}."""
    spec = getattr(settings, "spectral_unmixing", None)
    if not spec:
        return {}
    out = {}
    for spec_chl, spec_subtr in spec.items():
        if spec_chl == chl:
            out.update(spec_subtr)
    return out


def detection_shape(shape: Sequence[int], channels: Sequence[int]) -> Tuple[int, int, int]:
    """Shape of the volume ``blob_log`` sees for an ROI of ``shape``: the isotropic shape
    when the first channel's profile asks for it (detector.py:893-897)."""
    from . import cv_nd
    isotropic = config.get_roi_profile(channels[0])["isotropic"]
    if isotropic is None:
        return tuple(int(v) for v in shape[:3])
    return cv_nd.isotropic_shape(shape, isotropic)


def preprocessed_roi(roi, multichannel: bool, denoise_max_shape: Sequence[int]):
    """Every channel of an ROI preprocessed block by block (stack_detect.py:122-150) into one
    dense ``(z, y, x[, c])`` float32 CUDA tensor: what the reference hands to ``detect_blobs``
    and to ``colocalize_blobs`` (as float64)."""
    import torch
    from .. import gpu
    shape = tuple(int(v) for v in roi.shape[:3])
    block = tuple(int(v) for v in denoise_max_shape)
    n_chl = int(roi.shape[3]) if multichannel else 1
    out = torch.empty(shape + ((n_chl,) if multichannel else ()), dtype=torch.float32,
                      device=gpu.require_cuda())
    for c in range(n_chl):
        src = gpu.as_source(roi, c if multichannel else None)
        vol = gpu.preprocess_blocks(src, block, plot_3d.preproc_params(config.get_roi_profile(c), c))
        if multichannel:
            out[..., c] = vol[:, :, :shape[2]]
        else:
            out.copy_(vol[:, :, :shape[2]])
    return out


def enqueue_detection(det, roi, channels: Sequence[int], multichannel: bool,
                      denoise_max_shape: Optional[Sequence[int]] = None,
                      as_float64: bool = False):
    """Launch, without waiting, everything ``detect_blobs`` does to an ROI up to and
    including ``blob_log``, one fused chunk launch per channel on ``det`` (a
    ``gpu.ChunkDetector`` that fits ``detection_shape``).  With ``denoise_max_shape``
    the channels are first preprocessed block by block as ``detect_sub_roi`` does
    (stack_detect.py:122-150).  Returns ``[(channel, sigmas, ticket)]``.

    The common case is ONE launch per channel reading the caller's array in place.
    Profiles with ``isotropic`` (detector.py:893-897) or ``spectral_unmixing``
    (:910-921) go through intermediate float volumes instead: preprocess ->
    ``mmb_resize_linear`` -> ``mmb_unmix_subtract`` -> detection.  ``as_float64``: the
    float32 ``roi`` stands for a float64 array of the reference (``preprocessed_roi``), so
    the sigma ladder is built in float64."""
    from .. import gpu
    from . import cv_nd
    scale = calc_scaling_factor()[2]
    isotropic = config.get_roi_profile(channels[0])["isotropic"]
    unmix = {c: _unmixing_of(config.get_roi_profile(c), c) for c in channels}
    if any(unmix.values()) and not multichannel:
        raise IndexError("spectral unmixing needs a multichannel ROI")
    block = tuple(int(v) for v in denoise_max_shape) if denoise_max_shape is not None \
        else (1, 1, 1)
    tickets = []
    if (isinstance(roi, np.ndarray) and denoise_max_shape is None and isotropic is None
            and roi.dtype not in (np.uint8, np.uint16, np.float32, np.float64)):
        # dtypes the kernels do not read in place: scikit-image's own conversion, on the host
        roi = img_as_float_host(roi)
    if isotropic is None and not any(unmix.values()):
        for chl in channels:
            settings = config.get_roi_profile(chl)
            src = gpu.as_source(roi, chl if multichannel else None)
            pre, in_scale = None, 1.0
            f32 = src.dtype == gpu._lib.MMB_F32 and not as_float64
            if denoise_max_shape is not None:
                pre = plot_3d.preproc_params(settings, chl)
                f32 = False      # preprocessing yields float64 in the reference
            else:
                np_dtype = roi.dtype if isinstance(roi, np.ndarray) else None
                in_scale = input_scale(np_dtype) if np_dtype is not None else {
                    gpu._lib.MMB_U8: 1 / 255.0, gpu._lib.MMB_U16: 1 / 65535.0}.get(src.dtype, 1.0)
            sigmas = sigma_ladder(settings, scale, f32)
            if det.free_slots() == 0:
                raise RuntimeError("no free output slot: finish a pending sub-ROI first")
            tickets.append((chl, sigmas, det.enqueue(
                src, sigmas, settings["detection_threshold"], settings["overlap"],
                scale=in_scale, pre=pre, block_shape=block)))
        return tickets

    shape = tuple(int(v) for v in roi.shape[:3])
    iso_shape = shape if isotropic is None else cv_nd.isotropic_shape(shape, isotropic)
    edge = cv_nd.edge_mode_for(roi.shape)
    cache = {}

    def volume(c):
        """(pitched float32 volume in detection shape, raw integer scale or None, f32)"""
        if c not in cache:
            src = gpu.as_source(roi, c if multichannel else None)
            int_scale = {gpu._lib.MMB_U8: 1 / 255.0, gpu._lib.MMB_U16: 1 / 65535.0}.get(src.dtype)
            f32 = src.dtype == gpu._lib.MMB_F32 and not as_float64
            if denoise_max_shape is not None:
                vol = gpu.preprocess_blocks(
                    src, block, plot_3d.preproc_params(config.get_roi_profile(c), c))
                src, int_scale, f32 = gpu.volume_source(vol, shape), None, False
                if isotropic is None:
                    cache[c] = (vol, None, False)
                    return cache[c]
            if isotropic is not None:
                vol = gpu.resize_linear(src, iso_shape, edge)
            else:
                vol = gpu.to_float(src, 1.0)
            cache[c] = (vol, int_scale, f32)
        return cache[c]

    for chl in channels:
        settings = config.get_roi_profile(chl)
        vol, int_scale, f32 = volume(chl)
        in_scale = int_scale if int_scale is not None else 1.0
        if unmix[chl]:
            vol = vol.clone()
            for subt_chl, subt_fac in unmix[chl].items():
                gpu.unmix_subtract(vol, volume(int(subt_chl))[0], iso_shape[2], subt_fac)
            # np.subtract(..., factor * other) is float64: img_as_float leaves it as it is
            in_scale, f32 = 1.0, False
        sigmas = sigma_ladder(settings, scale, f32)
        if det.free_slots() == 0:
            raise RuntimeError("no free output slot: finish a pending sub-ROI first")
        tickets.append((chl, sigmas, det.enqueue(
            gpu.volume_source(vol, iso_shape), sigmas, settings["detection_threshold"],
            settings["overlap"], scale=in_scale)))
    return tickets


def scale_back_isotropic(blobs: np.ndarray, channels: Sequence[int]) -> np.ndarray:
    """Blobs detected on the isotropic ROI back onto the original grid: relative and
    absolute coordinates times ``1 / isotropic_factor`` (detector.py:944-951)."""
    from . import cv_nd
    isotropic = config.get_roi_profile(channels[0])["isotropic"]
    if isotropic is None or blobs is None:
        return blobs
    factor = 1 / cv_nd.calc_isotropic_factor(isotropic)
    blobs = Blobs.multiply_blob_rel_coords(blobs, factor)
    return Blobs.multiply_blob_abs_coords(blobs, factor)


def detect_blobs(roi, channel: Optional[Sequence[int]],
                 exclude_border: Optional[Sequence[Sequence[int]]] = None
                 ) -> Optional[np.ndarray]:
    """Detect blobs in an ROI with the multi-scale LoG detector.

    Args:
        roi: ``(z, y, x)`` or ``(z, y, x, c)`` array (numpy, or a CUDA tensor).
        channel: channels to detect in; None = all.
        exclude_border: optional ``[start, end]`` pairs of ``z, y, x`` margins
            whose blobs are dropped.

    Returns:
        ``(n, 11)`` float64 table in ``Blobs.Cols`` order, or None if nothing
        was found (detector.py:874-957).
    """
    from .. import gpu
    shape = tuple(roi.shape)
    multichannel, channels = plot_3d.setup_channels(roi, channel, 3)
    channels = list(channels)
    det_shape = detection_shape(shape, channels)
    detector = gpu.ChunkDetector(det_shape, n_slots=max(4, len(channels)))
    blobs_all = []
    for chl, sigmas, ticket in enqueue_detection(detector, roi, channels, multichannel):
        cands, _ = detector.collect(ticket)
        if len(cands) < 1:
            _logger.debug("No blobs detected for channel %s", chl)
            continue
        blobs_all.append(cands_to_blobs(cands, sigmas, det_shape[1:3], chl))
    if not blobs_all:
        return None
    blobs_all = scale_back_isotropic(np.vstack(blobs_all), channels)
    if exclude_border is not None:
        blobs_all = get_blobs_interior(blobs_all, shape, *exclude_border)
    return blobs_all


def sort_blobs(blobs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Sort by z, then y, then x; returns ``(sorted copy, order)``."""
    order = np.lexsort((blobs[:, 2], blobs[:, 1], blobs[:, 0]))
    return blobs[order], order


def _find_close_blobs(blobs: np.ndarray, blobs_master: np.ndarray, tol):
    """The box match of ``remove_close_blobs`` on the GPU (``mmb_prune_seams``), in the form
    the caller needs instead of the reference's pair lists (detector.py:1000-1006):
    ``(last, hit)`` - for every master the index of the LAST check blob within ``tol`` on
    all three axes in the reference's (tile, row) order, or -1; for every check blob
    whether any master matched it."""
    import torch
    from .. import gpu
    dev = gpu.require_cuda()
    m = torch.from_numpy(np.ascontiguousarray(blobs_master[:, :3]).astype(np.int32)).to(dev)
    c = torch.from_numpy(np.ascontiguousarray(blobs[:, :3]).astype(np.int32)).to(dev)
    last, hit = gpu.prune_seams(m, c, [int(t) for t in np.broadcast_to(tol, (3,))])
    return last.cpu().numpy(), hit.cpu().numpy().astype(bool)


def remove_close_blobs(blobs: np.ndarray, blobs_master: np.ndarray, tol,
                       chunk_size: int = 1000) -> Tuple[np.ndarray, np.ndarray]:
    """Drop every blob of ``blobs`` that lies within ``tol`` (per-axis, inclusive)
    of a blob in ``blobs_master``, and move each matched master's absolute
    coordinates to the rounded mean with its match (detector.py:1009-1085).

    The box test is done on integer-truncated coordinates, as in the reference
    (which casts to the smallest signed integer dtype).  When several blobs
    match one master the reference's repeated fancy-index assignment leaves the
    LAST match in place, i.e. the matching check blob with the largest index;
    that rule is applied directly.  ``chunk_size`` is accepted for signature
    compatibility; the GPU match needs no tiling on the host.
    """
    if len(blobs) < 1 or len(blobs_master) < 1:
        return blobs, blobs_master
    last, hit = _find_close_blobs(blobs, blobs_master, tol)
    pruned = blobs[~hit]
    sel = last >= 0
    if np.any(sel):
        abs_inds = Blobs(blobs)._get_abs_inds()
        between = np.around((blobs_master[np.ix_(sel, abs_inds)]
                             + blobs[np.ix_(last[sel], abs_inds)]) / 2)
        blobs_master[np.ix_(sel, abs_inds)] = between
    return pruned, blobs_master


def meas_pruning_ratio(num_blobs_orig, num_blobs_after_pruning, num_blobs_next):
    """``(original count, pruned/original, pruned/adjacent)`` or None
    (detector.py:1122-1144)."""
    if num_blobs_next > 0 and num_blobs_orig > 0:
        return (num_blobs_orig, num_blobs_after_pruning / num_blobs_orig,
                num_blobs_after_pruning / num_blobs_next)
    return None


def get_blobs_in_roi(blobs: np.ndarray, offset: Sequence[int], size: Sequence[int],
                     margin: Sequence[int] = (0, 0, 0), reverse: bool = True
                     ) -> Tuple[np.ndarray, np.ndarray]:
    """Blobs inside ``[offset - margin, offset + size + margin)``; ``reverse``
    means the three triples are given as x,y,z (detector.py:1210-1245)."""
    if reverse:
        offset, size, margin = offset[::-1], size[::-1], margin[::-1]
    mask = np.ones(len(blobs), dtype=bool)
    for a in range(3):
        mask &= blobs[:, a] >= offset[a] - margin[a]
        mask &= blobs[:, a] < offset[a] + size[a] + margin[a]
    return blobs[mask], mask


def get_blobs_interior(blobs: np.ndarray, shape: Sequence[int], pad_start: Sequence[int],
                       pad_end: Sequence[int]) -> np.ndarray:
    """Blobs at least ``pad_start`` from the low faces and ``pad_end`` from the
    high faces of a region of ``shape`` (detector.py:1248-1268)."""
    mask = np.ones(len(blobs), dtype=bool)
    for a in range(3):
        mask &= blobs[:, a] >= pad_start[a]
        mask &= blobs[:, a] < shape[a] - pad_end[a]
    return blobs[mask]
