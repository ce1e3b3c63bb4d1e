"""Restatement of the scikit-image 0.25.2 functions on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED for
scikit-image internals: the wheel is pinned by the reference
(``/root/reference/envs/requirements.txt:46``) but is not installable here, so
the control flow below is restated from the published 0.25.2 sources
(``skimage/feature/blob.py``, ``skimage/feature/peak.py``,
``skimage/_shared/coord.py``, ``skimage/filters/_gaussian.py``,
``skimage/morphology/gray.py``).  Every numeric kernel is the same scipy call
scikit-image makes.

Reference call sites restated here:
  ``magmap/cv/detector.py:931-933``   blob_log(roi, min_sigma, max_sigma, num_sigma, threshold, overlap)
  ``magmap/plot/plot_3d.py:157``      filters.gaussian(denoised, 8)
  ``magmap/plot/plot_3d.py:165``      morphology.erosion(denoised, morphology.octahedron(1))
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy import ndimage as ndi
from scipy import spatial

_INT_MAX = {np.dtype(t): float(np.iinfo(t).max) for t in
            (np.uint8, np.uint16, np.uint32, np.int8, np.int16, np.int32)}


def img_as_float(image: np.ndarray) -> np.ndarray:
    """``skimage.util.img_as_float``: floats pass through untouched (float16
    is widened to float32); unsigned ints are divided by their dtype max in
    float64; signed ints map to [-1, 1]."""
    dt = image.dtype
    if dt.kind == "f":
        return image.astype(np.float32) if dt == np.float16 else image
    if dt.kind == "b":
        return image.astype(np.float64)
    if dt.kind == "u":
        return image.astype(np.float64) / _INT_MAX[dt]
    if dt.kind == "i":
        # signed: (2*x + 1) / (max - min), the dtype_range mapping
        info = np.iinfo(dt)
        out = image.astype(np.float64)
        out *= 2.0
        out += 1.0
        out /= float(info.max) - float(info.min)
        return out
    raise ValueError(f"unsupported dtype {dt}")


def sigma_list(min_sigma: float, max_sigma: float, num_sigma: int,
               float_dtype=np.float64) -> np.ndarray:
    """blob_log's linear scale ladder (``log_scale=False``), as 0.25.2 builds it
    (``skimage/feature/blob.py``): the scalar sigmas are cast to the image's float
    dtype (a float32 image gets float32-rounded ends; every integer or float64
    image - all of MagellanMapper's own call paths - float64 ones) and then

        scale = np.linspace(0, 1, num_sigma)[:, None]
        sigma_list = scale * (max_sigma - min_sigma) + min_sigma

    which is NOT bit-equal to ``np.linspace(min_sigma, max_sigma, num_sigma)``
    unless ``max - min`` is a power of two (3..5 is; 4..10 differs by 8.9e-16)."""
    lo = np.asarray(min_sigma, dtype=float_dtype)
    hi = np.asarray(max_sigma, dtype=float_dtype)
    scale = np.linspace(0, 1, int(num_sigma))
    return (scale * (hi - lo) + lo).astype(np.float64)


def log_cube(image: np.ndarray, sigmas: Sequence[float]) -> np.ndarray:
    """Scale-normalised, sign-flipped LoG stack: ``cube[..., i] =
    -gaussian_laplace(image, s_i) * s_i**2`` (blob.py, "computing gaussian
    laplace"; "average s**2 provides scale invariance")."""
    image = img_as_float(image)
    cube = np.empty(image.shape + (len(sigmas),), dtype=image.dtype
                    if image.dtype.kind == "f" else np.float64)
    for i, s in enumerate(sigmas):
        svec = np.full(image.ndim, s, dtype=cube.dtype)
        cube[..., i] = -ndi.gaussian_laplace(image, svec) * np.mean(svec) ** 2
    return cube


def peak_local_max_4d(cube: np.ndarray, threshold_abs: float
                      ) -> Tuple[np.ndarray, np.ndarray]:
    """``peak_local_max(cube, threshold_abs=thr, footprint=ones(3**ndim),
    exclude_border=0)`` as blob_log calls it.

    Returns ``(coords, responses)``: integer (n, ndim) coordinates sorted by
    descending response with a stable sort (ties keep C order), as
    ``_get_high_intensity_peaks`` does.  ``min_distance=1`` with ``p_norm=inf``
    only rejects points at Chebyshev distance < 1 from an accepted point, i.e.
    duplicates, so ``ensure_spacing`` is the identity on distinct integer
    coordinates and is not re-run here.
    """
    if cube.size == 1:
        mask = cube > threshold_abs
    else:
        mx = ndi.maximum_filter(cube, size=3, mode="nearest")
        mask = cube == mx
        if np.all(mask):          # "no peak for a trivial image"
            mask[:] = False
        mask &= cube > threshold_abs
    coords = np.nonzero(mask)
    vals = cube[coords]
    order = np.argsort(-vals, kind="stable")
    return np.transpose(coords)[order], vals[order]


def blob_overlap(b1: np.ndarray, b2: np.ndarray) -> float:
    """Fraction of the smaller sphere's volume inside the other
    (``_blob_overlap`` + ``_compute_sphere_overlap``, ndim == 3,
    ``sigma_dim == 1``).  Inputs are ``[z, y, x, sigma]`` rows."""
    s1, s2 = b1[-1], b2[-1]
    if s1 == 0 and s2 == 0:
        return 0.0
    root = math.sqrt(3)
    if s1 > s2:
        big, r1, r2 = s1, 1.0, s2 / s1
    else:
        big, r1, r2 = s2, s1 / s2, 1.0
    p1 = b1[:3] / (big * root)
    p2 = b2[:3] / (big * root)
    d = np.sqrt(np.sum((p2 - p1) ** 2))
    if d > r1 + r2:
        return 0.0
    if d <= abs(r1 - r2):
        return 1.0
    vol = (math.pi / (12 * d) * (r1 + r2 - d) ** 2
           * (d ** 2 + 2 * d * (r1 + r2) - 3 * (r1 ** 2 + r2 ** 2) + 6 * r1 * r2))
    return vol / (4.0 / 3 * math.pi * min(r1, r2) ** 3)


@dataclass
class PruneTrace:
    """Diagnostics of one ``prune_blobs`` run."""
    pairs: np.ndarray                     # (m, 2) examined pairs, in iteration order
    kill_edges: np.ndarray                # (k, 2) [killer, victim] on ORIGINAL sigmas
    order_dependent: np.ndarray           # indices whose survival depends on pair order
    keep_reference_order: np.ndarray      # bool mask, scikit-image's set-iteration order
    keep_canonical: np.ndarray            # bool mask, order-independent greedy rule


def _kill_edges(blobs: np.ndarray, pairs: np.ndarray, overlap: float) -> np.ndarray:
    edges = []
    for i, j in pairs:
        if blob_overlap(blobs[i], blobs[j]) > overlap:
            # victim = smaller sigma; on equal sigma the lower index (blob1)
            if blobs[i, -1] > blobs[j, -1]:
                edges.append((i, j))
            else:
                edges.append((j, i))
    return np.array(edges, dtype=np.int64).reshape(-1, 2)


def canonical_keep(n: int, edges: np.ndarray) -> np.ndarray:
    """Order-independent resolution of the kill graph: a blob survives iff
    none of its killers survives.  The graph is acyclic (a killer has a larger
    sigma, or equal sigma and a larger index), so this is well defined."""
    killers: List[List[int]] = [[] for _ in range(n)]
    for k, v in edges:
        killers[v].append(k)
    state = np.full(n, -1, dtype=np.int8)           # -1 unknown, 0 dead, 1 alive
    pending = True
    while pending:
        pending = False
        for v in range(n):
            if state[v] != -1:
                continue
            ks = [state[k] for k in killers[v]]
            if any(s == 1 for s in ks):
                state[v] = 0
            elif all(s == 0 for s in ks):
                state[v] = 1
            else:
                pending = True
    return state == 1


def order_dependent_set(n: int, edges: np.ndarray) -> np.ndarray:
    """Blobs whose fate depends on the iteration order of the pair set: those
    with at least one killer, none of which is a root (a blob nobody kills).
    A root is alive whenever its pairs are visited, so its victims always die;
    any other killer may or may not already be zeroed when its pair comes up."""
    has_killer = np.zeros(n, dtype=bool)
    if len(edges):
        has_killer[edges[:, 1]] = True
    root = ~has_killer
    killed_by_root = np.zeros(n, dtype=bool)
    for k, v in edges:
        if root[k]:
            killed_by_root[v] = True
    return np.nonzero(has_killer & ~killed_by_root)[0]


def prune_blobs(blobs: np.ndarray, overlap: float, trace: bool = False,
                pair_order: Optional[np.ndarray] = None):
    """``_prune_blobs(blobs_array, overlap, sigma_dim=1)``.

    Sequential, in the iteration order of the Python ``set`` returned by
    ``cKDTree.query_pairs`` (``pairs = np.array(list(tree.query_pairs(d)))``):
    a pair whose overlap on the CURRENT sigmas exceeds ``overlap`` zeroes the
    smaller-sigma blob (blob *i*, the stronger response, on equal sigma);
    zeroed blobs stay in the loop but can only re-zero themselves.
    """
    blobs = np.array(blobs, dtype=np.float64, copy=True)
    n = len(blobs)
    if n == 0:
        return (blobs, None) if trace else blobs
    orig = blobs.copy()
    sigma = blobs[:, -1].max()
    distance = 2 * sigma * math.sqrt(blobs.shape[1] - 1)
    tree = spatial.cKDTree(blobs[:, :-1])
    if pair_order is None:
        pairs = np.array(list(tree.query_pairs(distance)), dtype=np.int64).reshape(-1, 2)
    else:
        pairs = np.asarray(pair_order, dtype=np.int64).reshape(-1, 2)
    for i, j in pairs:
        b1, b2 = blobs[i], blobs[j]
        if blob_overlap(b1, b2) > overlap:
            if b1[-1] > b2[-1]:
                b2[-1] = 0
            else:
                b1[-1] = 0
    keep = blobs[:, -1] > 0
    out = orig[keep]
    if not trace:
        return out
    edges = _kill_edges(orig, pairs, overlap)
    tr = PruneTrace(pairs=pairs, kill_edges=edges,
                    order_dependent=order_dependent_set(n, edges),
                    keep_reference_order=keep,
                    keep_canonical=canonical_keep(n, edges))
    return out, tr


@dataclass
class BlobLogResult:
    blobs: np.ndarray                     # (n, 4) [z, y, x, sigma], scikit-image output
    sigmas: np.ndarray
    peaks: np.ndarray                     # (m, 4) int [z, y, x, scale index] before pruning
    responses: np.ndarray                 # (m,) cube values at peaks, descending
    trace: Optional[PruneTrace] = None
    cube: Optional[np.ndarray] = field(default=None, repr=False)


def blob_log(image: np.ndarray, min_sigma: float = 1, max_sigma: float = 50,
             num_sigma: int = 10, threshold: float = 0.2, overlap: float = 0.5,
             full: bool = False, keep_cube: bool = False):
    """``skimage.feature.blob_log`` for scalar sigmas, ``log_scale=False``,
    ``threshold_rel=None``, ``exclude_border=False`` (the arguments the
    reference passes, ``magmap/cv/detector.py:931-933``)."""
    image = img_as_float(image)
    sig = sigma_list(min_sigma, max_sigma, num_sigma, image.dtype)
    cube = log_cube(image, sig)
    peaks, resp = peak_local_max_4d(cube, threshold)
    if len(peaks) == 0:
        empty = np.empty((0, image.ndim + 1))
        if full:
            return BlobLogResult(empty, sig, peaks, resp, None, cube if keep_cube else None)
        return empty
    lm = peaks.astype(cube.dtype)
    lm = np.hstack([lm[:, :-1], sig[peaks[:, -1]][:, None].astype(cube.dtype)])
    if full:
        out, tr = prune_blobs(lm, overlap, trace=True)
        return BlobLogResult(out, sig, peaks, resp, tr, cube if keep_cube else None)
    return prune_blobs(lm, overlap)


def filters_gaussian(image: np.ndarray, sigma: float) -> np.ndarray:
    """``skimage.filters.gaussian(image, sigma)`` with its defaults
    ``mode='nearest'``, ``truncate=4.0``, ``preserve_range=False``,
    ``channel_axis=None``: a float image goes straight to
    ``scipy.ndimage.gaussian_filter``."""
    image = img_as_float(image)
    return ndi.gaussian_filter(image, sigma, mode="nearest", cval=0, truncate=4.0)


def octahedron1() -> np.ndarray:
    """``skimage.morphology.octahedron(1)``: centre plus six face neighbours."""
    fp = np.zeros((3, 3, 3), dtype=np.uint8)
    fp[1, 1, :] = 1
    fp[1, :, 1] = 1
    fp[:, 1, 1] = 1
    return fp


def erosion_octahedron1(image: np.ndarray) -> np.ndarray:
    """``skimage.morphology.erosion(image, octahedron(1))``: grey erosion =
    minimum over the footprint; 0.25.x pads with ``mode='reflect'``, for which
    a radius-1 neighbour outside the array mirrors onto the voxel itself (the
    older max-value padding gives the same minimum)."""
    return ndi.grey_erosion(image, footprint=octahedron1(), mode="reflect")


def transform_resize(image: np.ndarray, output_shape: Sequence[int], order=None,
                     mode: str = "reflect", cval: float = 0, clip: bool = True,
                     preserve_range: bool = False, anti_aliasing=None,
                     anti_aliasing_sigma=None) -> np.ndarray:
    """``skimage.transform.resize`` (``skimage/transform/_warps.py``, 0.19 - 0.25) for the
    n-D ``scipy.ndimage.zoom`` route, as ``magmap.cv.cv_nd.rescale_resize`` calls it:

    * ``order=None`` -> 0 for bool images, 1 otherwise; float conversion with
      ``preserve_range`` keeps float32 / float64 and widens everything else to float64;
    * ``anti_aliasing=None`` -> on when any axis shrinks (never for bool, nor for
      integer images with ``order == 0``): ``ndi.gaussian_filter`` with ``sigma =
      max(0, (in / out - 1) / 2)`` per axis in the translated boundary mode;
    * ``np.pad`` mode names translate to ``scipy.ndimage`` ones: 'reflect' -> 'mirror',
      'symmetric' -> 'reflect', 'edge' -> 'nearest';
    * ``ndi.zoom(..., order, mode, cval, grid_mode=True)`` with zoom factors ``out / in``;
    * ``clip`` -> ``np.clip`` to the input's value range.
    """
    image = np.asarray(image)
    output_shape = tuple(int(v) for v in output_shape)
    if len(output_shape) != image.ndim:
        raise ValueError("output_shape must have one entry per image axis")
    input_shape = image.shape
    input_type = image.dtype
    if input_type == np.float16:
        image = image.astype(np.float32)
    if order is None:
        order = 0 if input_type == bool else 1
    if anti_aliasing is None:
        anti_aliasing = (input_type != bool
                         and not (np.issubdtype(input_type, np.integer) and order == 0)
                         and any(o < i for o, i in zip(output_shape, input_shape)))
    factors = np.divide(input_shape, output_shape)
    if order > 0:
        if preserve_range:
            if image.dtype.char not in "df":
                image = image.astype(float)
        else:
            image = img_as_float(image)
    ndi_mode = {"constant": "constant", "edge": "nearest", "symmetric": "reflect",
                "reflect": "mirror", "wrap": "wrap"}[mode]
    filtered = image
    if anti_aliasing:
        if anti_aliasing_sigma is None:
            anti_aliasing_sigma = np.maximum(0, (factors - 1) / 2)
        filtered = ndi.gaussian_filter(image, anti_aliasing_sigma, cval=cval, mode=ndi_mode)
    zoom_factors = [1 / f for f in factors]
    out = ndi.zoom(filtered, zoom_factors, order=order, mode=ndi_mode, cval=cval,
                   grid_mode=True)
    if clip and out.size:
        np.clip(out, np.min(image), np.max(image), out=out)
    return out


def ball(radius: int) -> np.ndarray:
    """``skimage.morphology.ball``: voxels with ``z^2 + y^2 + x^2 <= radius^2``."""
    n = 2 * radius + 1
    zz, yy, xx = np.mgrid[-radius:radius:n * 1j, -radius:radius:n * 1j, -radius:radius:n * 1j]
    return np.array(zz * zz + yy * yy + xx * xx <= radius * radius, dtype=np.uint8)


def dilation(image: np.ndarray, footprint: np.ndarray) -> np.ndarray:
    """``skimage.morphology.dilation`` (grey, 'reflect' borders):
    ``scipy.ndimage.grey_dilation`` with the (point-symmetric) footprint."""
    return ndi.grey_dilation(image, footprint=footprint.astype(bool))
