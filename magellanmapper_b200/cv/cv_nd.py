"""The resizing helpers of ``magmap/cv/cv_nd.py`` that sit on the detection path
(``:1040-1167``): the isotropic factor and ``make_isotropic``, on the GPU through
``mmb_resize_linear``."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np

from ..settings import config


def calc_isotropic_factor(scale: Union[float, Sequence[float]] = 1,
                          res: Optional[Sequence[float]] = None) -> np.ndarray:
    """Resolutions over their minimum, times ``scale`` (cv_nd.py:1040-1068)."""
    if res is None:
        res = config.resolutions[0]
    resize_factor = np.divide(res, np.amin(res))
    resize_factor = resize_factor * scale
    return resize_factor


def isotropic_shape(shape: Sequence[int], scale=1, res=None) -> Tuple[int, int, int]:
    """``(shape[:3] * factor).astype(int)`` (cv_nd.py:1091-1093)."""
    factor = calc_isotropic_factor(scale, res)
    return tuple(int(v) for v in (np.array(shape[:3]) * factor).astype(int))


def edge_mode_for(shape: Sequence[int]) -> bool:
    """'edge' instead of 'reflect' when any axis is one voxel thick (cv_nd.py:1097-1102)."""
    return bool(np.any(np.array(shape) == 1))


def make_isotropic(roi, scale: Union[float, Sequence[float]] = 1,
                   res: Optional[Sequence[float]] = None, **kwargs) -> np.ndarray:
    """Resize an ROI ``(z, y, x[, c])`` to be isotropic (cv_nd.py:1071-1106): linear
    interpolation with 'reflect' boundaries, Gaussian anti-aliasing on shrinking axes,
    value range preserved and the result cast back to the ROI's dtype.  ``kwargs`` are the
    reference's overrides for ``rescale_resize``; the detection path passes none and only
    the defaults are served."""
    from .. import gpu
    import torch
    for key, val in kwargs.items():
        if (key, val) not in (("preserve_range", True), ("order", 1), ("anti_aliasing", None)):
            raise NotImplementedError(
                f"make_isotropic({key}={val!r}): only the defaults of the detection path "
                "(order 1, preserve_range, default anti-aliasing) are accelerated")
    shape = tuple(roi.shape)
    out_shape = isotropic_shape(shape, scale, res)
    edge = edge_mode_for(shape)
    dtype = roi.dtype if isinstance(roi, np.ndarray) else None
    n_chl = shape[3] if len(shape) > 3 else None
    outs = []
    for c in range(n_chl or 1):
        src = gpu.as_source(roi, c if n_chl else None)
        vol = gpu.resize_linear(src, out_shape, edge)
        torch.cuda.synchronize()
        a = vol[:, :, :out_shape[2]].cpu().numpy()
        outs.append(a.astype(dtype) if dtype is not None else a)
    return np.stack(outs, axis=-1) if n_chl else outs[0]
