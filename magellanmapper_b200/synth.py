"""Seeded synthetic nuclei volumes.

Background uint16 ~ N(400, 30^2) clipped at 0; nuclei are isotropic Gaussian
spots with sigma0 ~ U(2.5, 4.5) px and peak amplitude ~ U(0.3, 0.9) * 65535 at
about one nucleus per 6.7 k voxels, centres uniform over the whole volume
(faces and chunk seams included).  Pure numpy, deterministic for a given seed;
``bench.py`` has a device-side generator of the same recipe for volumes that
are too large to build on the host.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

DENSITY = 1.0 / 6700.0


def nuclei_table(shape: Sequence[int], seed: int, density: float = DENSITY,
                 n: Optional[int] = None) -> np.ndarray:
    """(n, 5) float64 rows ``z, y, x, sigma0, amplitude`` (amplitude in raw
    uint16 counts)."""
    rng = np.random.default_rng(seed)
    shape = np.asarray(shape[:3], dtype=np.int64)
    if n is None:
        n = max(1, int(round(float(np.prod(shape)) * density)))
    ctr = rng.uniform(0, 1, size=(n, 3)) * shape
    sig = rng.uniform(2.5, 4.5, size=n)
    amp = rng.uniform(0.3, 0.9, size=n) * 65535.0
    return np.column_stack([ctr, sig, amp])


def render(shape: Sequence[int], table: np.ndarray, seed: int,
           bg_mean: float = 400.0, bg_std: float = 30.0) -> np.ndarray:
    """Render a nuclei table plus Gaussian background noise to uint16."""
    rng = np.random.default_rng(seed + 0x5EED)
    Z, Y, X = (int(s) for s in shape[:3])
    vol = rng.normal(bg_mean, bg_std, size=(Z, Y, X)).astype(np.float32)
    for cz, cy, cx, s0, a in table:
        r = int(np.ceil(4 * s0))
        z0, z1 = max(0, int(cz) - r), min(Z, int(cz) + r + 1)
        y0, y1 = max(0, int(cy) - r), min(Y, int(cy) + r + 1)
        x0, x1 = max(0, int(cx) - r), min(X, int(cx) + r + 1)
        if z0 >= z1 or y0 >= y1 or x0 >= x1:
            continue
        gz = np.exp(-0.5 * ((np.arange(z0, z1) - cz) / s0) ** 2)
        gy = np.exp(-0.5 * ((np.arange(y0, y1) - cy) / s0) ** 2)
        gx = np.exp(-0.5 * ((np.arange(x0, x1) - cx) / s0) ** 2)
        vol[z0:z1, y0:y1, x0:x1] += (a * gz[:, None, None] * gy[None, :, None]
                                     * gx[None, None, :]).astype(np.float32)
    np.clip(vol, 0, 65535, out=vol)
    return np.rint(vol).astype(np.uint16)


def make_volume(shape: Sequence[int], seed: int, density: float = DENSITY,
                n: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """Return ``(uint16 volume, nuclei table)``."""
    tab = nuclei_table(shape, seed, density, n)
    return render(shape, tab, seed), tab


def near_max_of(vol: np.ndarray, pct: float = 99.5) -> float:
    """The reference's ``near_max`` metadata: the max over z-planes of the
    per-plane upper percentile (``magmap/io/importer.py:1415-1468``)."""
    return float(max(np.percentile(p, pct) for p in vol))


def device_volume(shape: Sequence[int], seed: int, offset: Sequence[int] = (0, 0, 0),
                  density: float = DENSITY, device=None, out=None):
    """The same recipe generated ON THE DEVICE as a pure function of (seed, global
    z, y, x) (``mmb_synth_nuclei``, include/mmb200_tools.h): returns the uint16 box of
    ``shape`` whose first voxel sits at ``offset`` of the unbounded volume, as an
    int16-bit CUDA tensor (torch has no full uint16).  Boxes generated separately -
    the z-slabs of the ranks, a halo, the sub-box an oracle spot check recomputes -
    agree bit for bit where they overlap.  Not the numpy generator's values: a
    different random stream of the same distribution."""
    import ctypes as C
    import torch
    from . import _lib, gpu
    dev = device or gpu.require_cuda()
    Z, Y, X = (int(v) for v in shape[:3])
    if out is None:
        out = torch.empty((Z, Y, X), dtype=torch.int16, device=dev)
    _lib.check(_lib.load().mmb_synth_nuclei(
        C.c_void_p(out.data_ptr()), Z, Y, X, int(offset[0]), int(offset[1]), int(offset[2]),
        int(seed) & 0xFFFFFFFFFFFFFFFF, float(density),
        C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out
