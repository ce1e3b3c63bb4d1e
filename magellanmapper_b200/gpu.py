"""Thin torch-tensor wrappers over the C ABI (``include/mmb200.h``).

PyTorch is plumbing only: device allocations, streams, host<->device copies.
Every function here ends in a call into ``libmmb200.so``; none has a CPU path.

Float volumes are ``(Z, Y, pitch)`` float32 CUDA tensors whose rows are padded
to a multiple of 32 elements; the logical width ``X`` travels alongside.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import MmbCand, MmbPreprocParams

_NP2MMB = {np.dtype(np.uint8): _lib.MMB_U8, np.dtype(np.uint16): _lib.MMB_U16,
           np.dtype(np.float32): _lib.MMB_F32, np.dtype(np.float64): _lib.MMB_F64}
_T2MMB = {torch.uint8: _lib.MMB_U8, torch.uint16: _lib.MMB_U16, torch.int16: _lib.MMB_U16,
          torch.float32: _lib.MMB_F32, torch.float64: _lib.MMB_F64}

CAND_DTYPE = np.dtype([("z", "<i4"), ("y", "<i4"), ("x", "<i4"), ("s", "<i4"), ("resp", "<f4")])


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("magellanmapper_b200 needs a CUDA device; there is no CPU fallback")
    _lib.load()
    return torch.device("cuda", torch.cuda.current_device())


def pitch_for(x: int) -> int:
    return (int(x) + 31) // 32 * 32


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def new_volume(Z: int, Y: int, X: int, device=None) -> torch.Tensor:
    return torch.zeros((Z, Y, pitch_for(X)), dtype=torch.float32, device=device or require_cuda())


@dataclass
class Source:
    """A 3-D (z, y, x) view of device memory in its native dtype."""
    tensor: torch.Tensor          # keeps the storage alive
    ptr: int
    dtype: int
    strides: Tuple[int, int, int]
    shape: Tuple[int, int, int]


def as_source(arr, channel: Optional[int] = None, device=None) -> Source:
    """Upload (if needed) and describe one channel of a (z,y,x[,c]) array.

    numpy uint16 has no full torch support, so it is moved as int16 bits; the
    library is told the true dtype.
    """
    device = device or require_cuda()
    if isinstance(arr, np.ndarray):
        dt = _NP2MMB.get(arr.dtype)
        if dt is None:
            # rare dtypes (int32, float16, bool...) go through float32 on the host side of the copy
            arr = arr.astype(np.float32)
            dt = _lib.MMB_F32
        a = np.ascontiguousarray(arr)
        if a.dtype == np.uint16:
            t = torch.from_numpy(a.view(np.int16))
        else:
            t = torch.from_numpy(a)
        t = t.to(device, non_blocking=True)
    elif isinstance(arr, torch.Tensor):
        dt = _T2MMB.get(arr.dtype)
        if dt is None:
            raise TypeError(f"unsupported tensor dtype {arr.dtype}")
        t = arr if arr.is_cuda else arr.to(device, non_blocking=True)
    else:
        raise TypeError(f"expected numpy array or torch tensor, got {type(arr)}")
    if t.dim() == 4:
        c = 0 if channel is None else int(channel)
        view = t[..., c]
    elif t.dim() == 3:
        view = t
    else:
        raise ValueError(f"expected a (z,y,x) or (z,y,x,c) array, got {tuple(t.shape)}")
    return Source(t, view.data_ptr(), dt, tuple(int(s) for s in view.stride()),
                  tuple(int(s) for s in view.shape))


def upload_if_fits(img, fraction: float = 0.4):
    """Move a C-contiguous host array to the device in one copy when it takes
    less than ``fraction`` of the free device memory; otherwise (or for
    non-contiguous views, e.g. slices of a memmap) return it unchanged and let
    each sub-ROI be uploaded on its own."""
    if not isinstance(img, np.ndarray) or not img.flags.c_contiguous:
        return img
    if img.dtype not in _NP2MMB:
        return img
    device = require_cuda()
    free, _ = torch.cuda.mem_get_info(device)
    if img.nbytes > fraction * free:
        return img
    host = img.view(np.int16) if img.dtype == np.uint16 else img
    t = torch.from_numpy(host).to(device, non_blocking=True)
    return t.view(torch.uint16) if img.dtype == np.uint16 else t


class StripFeeder:
    """Streams y-strips ``img[:, y0:y1]`` of a C-contiguous HOST array to the device
    so that the upload of the next strip overlaps the kernels of the current one
    (a chunk grid is walked strip by strip; a strip holds every chunk of one y
    column).  Each strip is one pitched DMA (``mmb_upload_pieces``: one contiguous
    piece per z-plane) on a copy stream into one of two device buffers."""

    def __init__(self, img: np.ndarray, y_ranges: Sequence[Tuple[int, int]], device=None,
                 prefix: Optional[torch.Tensor] = None, suffix: Optional[torch.Tensor] = None):
        """``prefix`` / ``suffix``: device tensors of whole planes that precede /
        follow ``img`` along z (halo planes received from a neighbouring rank);
        every strip is then ``[prefix; img; suffix][:, y0:y1]``."""
        if not (isinstance(img, np.ndarray) and img.flags.c_contiguous and img.ndim in (3, 4)
                and img.dtype in _NP2MMB):
            raise TypeError("StripFeeder needs a C-contiguous (z,y,x[,c]) array of a supported dtype")
        self.lib = _lib.load()
        self.device = device or require_cuda()
        self.img = img
        self.ranges = [(int(a), int(b)) for a, b in y_ranges]
        self.tdtype = torch.int16 if img.dtype == np.uint16 else torch.from_numpy(
            np.zeros(1, img.dtype)).dtype
        rows = max(b - a for a, b in self.ranges)
        self.prefix = prefix if prefix is not None and prefix.shape[0] else None
        self.suffix = suffix if suffix is not None and suffix.shape[0] else None
        self.n_pre = 0 if self.prefix is None else int(self.prefix.shape[0])
        self.n_suf = 0 if self.suffix is None else int(self.suffix.shape[0])
        self.planes = self.n_pre + int(img.shape[0]) + self.n_suf
        self.bufs = [torch.empty((self.planes, rows) + tuple(img.shape[2:]), dtype=self.tdtype,
                                 device=self.device) for _ in range(min(2, len(self.ranges)))]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        # halo planes are produced on the caller's stream (NCCL recv): order the copy
        # stream after them
        self.copy_stream.wait_stream(torch.cuda.current_stream())
        self.uploaded = [None] * len(self.ranges)      # events on the copy stream
        self.released = [None] * len(self.ranges)      # events on the compute stream
        for j in range(len(self.bufs)):                # both buffers are free: fill them
            self._issue(j)

    def _issue(self, j: int) -> None:
        y0, y1 = self.ranges[j]
        img = self.img
        row_bytes = int(np.prod(img.shape[2:])) * img.itemsize
        buf = self.bufs[j % len(self.bufs)]
        with torch.cuda.stream(self.copy_stream):
            if j >= len(self.bufs) and self.released[j - len(self.bufs)] is not None:
                self.copy_stream.wait_event(self.released[j - len(self.bufs)])
            src = img.ctypes.data + y0 * row_bytes
            piece = (y1 - y0) * row_bytes
            view = self._view(buf, y0, y1)
            if img.shape[0]:
                _lib.check(self.lib.mmb_upload_pieces(
                    C.c_void_p(view[self.n_pre:].data_ptr()), C.c_void_p(src), int(img.shape[0]),
                    piece, int(img.shape[1]) * row_bytes,
                    C.c_void_p(self.copy_stream.cuda_stream)))
            if self.prefix is not None:
                view[:self.n_pre].copy_(self.prefix[:, y0:y1].view(self.tdtype)
                                        if self.prefix.dtype != self.tdtype
                                        else self.prefix[:, y0:y1], non_blocking=True)
            if self.suffix is not None:
                view[self.n_pre + img.shape[0]:].copy_(
                    self.suffix[:, y0:y1].view(self.tdtype) if self.suffix.dtype != self.tdtype
                    else self.suffix[:, y0:y1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.uploaded[j] = ev

    def _view(self, buf: torch.Tensor, y0: int, y1: int) -> torch.Tensor:
        # dense strip: the buffer is viewed with this strip's own row count
        inner = tuple(self.img.shape[2:])
        n = self.planes * (y1 - y0) * int(np.prod(inner))
        return buf.view(-1)[:n].view((self.planes, y1 - y0) + inner)

    def strip(self, j: int) -> torch.Tensor:
        """Device view of strip ``j`` (shape ``(Z, y1 - y0, X[, C])``); the current
        stream waits for its upload, and the upload of strip ``j + 1`` is started."""
        torch.cuda.current_stream().wait_event(self.uploaded[j])
        y0, y1 = self.ranges[j]
        return self._view(self.bufs[j % len(self.bufs)], y0, y1)

    def release(self, j: int) -> None:
        """Every kernel reading strip ``j`` has been enqueued on the current stream."""
        ev = torch.cuda.Event()
        ev.record()
        self.released[j] = ev
        nxt = j + len(self.bufs)                       # the strip that reuses this buffer
        if nxt < len(self.ranges) and self.uploaded[nxt] is None:
            self._issue(nxt)


def percentiles(src: Source, q: Sequence[float], per_plane: bool = False) -> np.ndarray:
    """``np.percentile(view, q)`` ('linear', float64) of a uint8 / uint16 view on the
    device: shape ``(len(q),)`` for the whole view, ``(Z, len(q))`` per z-plane."""
    lib = _lib.load()
    Z, Y, X = src.shape
    groups = Z if per_plane else 1
    dev = src.tensor.device
    out = torch.empty((groups, len(q)), dtype=torch.float64, device=dev)
    work = torch.empty(lib.mmb_percentiles_work_bytes(groups), dtype=torch.uint8, device=dev)
    qs = (C.c_double * len(q))(*[float(v) for v in q])
    _lib.check(lib.mmb_percentiles(C.c_void_p(src.ptr), src.dtype, _lib._I64x3(*src.strides), Z, Y, X,
                                   1 if per_plane else 0, qs, len(q), _ptr(out), _ptr(work),
                                   _stream()))
    res = out.cpu().numpy()
    return res if per_plane else res[0]


def to_float(src: Source, scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    Z, Y, X = src.shape
    out = out if out is not None else new_volume(Z, Y, X, src.tensor.device)
    _lib.check(lib.mmb_to_float(C.c_void_p(src.ptr), src.dtype, _lib._I64x3(*src.strides), Z, Y, X,
                                _ptr(out), out.shape[2], float(scale), _stream()))
    return out


def preprocess_blocks(src: Source, block_shape: Sequence[int], params: MmbPreprocParams,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    Z, Y, X = src.shape
    out = out if out is not None else new_volume(Z, Y, X, src.tensor.device)
    bz, by, bx = (int(b) for b in block_shape)
    _lib.check(lib.mmb_preprocess_blocks(
        C.c_void_p(src.ptr), src.dtype, _lib._I64x3(*src.strides), Z, Y, X, bz, by, bx,
        C.byref(params), _ptr(out), out.shape[2], _stream()))
    return out


def volume_source(vol: torch.Tensor, shape: Sequence[int]) -> Source:
    """A pitched float32 device volume (``new_volume``) as a detection input."""
    Z, Y, X = (int(v) for v in shape)
    pitch = int(vol.shape[2])
    return Source(vol, vol.data_ptr(), _lib.MMB_F32, (Y * pitch, pitch, 1), (Z, Y, X))


def resize_linear(src: Source, out_shape: Sequence[int], edge: bool = False) -> torch.Tensor:
    """``skimage.transform.resize(..., preserve_range=True).astype(dtype)`` as
    ``cv_nd.make_isotropic`` uses it (``mmb_resize_linear``): returns the pitched
    float32 volume of ``out_shape``; integer inputs come back truncated to integers.
    ``edge``: 'edge' instead of 'reflect' boundaries."""
    lib = _lib.load()
    Z, Y, X = src.shape
    Zo, Yo, Xo = (int(v) for v in out_shape)
    out = new_volume(Zo, Yo, Xo, src.tensor.device)
    _lib.check(lib.mmb_resize_linear(
        C.c_void_p(src.ptr), src.dtype, _lib._I64x3(*src.strides), Z, Y, X, _ptr(out), Zo, Yo, Xo,
        out.shape[2], 1 if edge else 0, _stream()))
    return out


def unmix_subtract(target: torch.Tensor, other: torch.Tensor, X: int, factor: float) -> None:
    """``target = max(target - factor * other, 0)`` in place on two pitched volumes."""
    Z, Y, pitch = target.shape
    if tuple(other.shape) != tuple(target.shape):
        raise ValueError(f"volumes differ: {tuple(target.shape)} vs {tuple(other.shape)}")
    _lib.check(_lib.load().mmb_unmix_subtract(_ptr(target), _ptr(other), Z, Y, int(X), pitch,
                                              float(factor), _stream()))


def log_pass(in0, in1, X: int, axis: int, mode: int, sigma: float, scale: float = 1.0):
    """One separable sweep (see ``mmb_log_pass``); returns (out0, out1|None)."""
    lib = _lib.load()
    Z, Y, pitch = in0.shape
    out0 = torch.zeros_like(in0)
    out1 = torch.zeros_like(in0) if mode != 2 else None
    _lib.check(lib.mmb_log_pass(_ptr(in0), _ptr(in1), _ptr(out0), _ptr(out1), Z, Y, X, pitch,
                                axis, mode, float(sigma), float(scale), _stream()))
    return out0, out1


def log_scale(vol: torch.Tensor, X: int, sigma: float, out: Optional[torch.Tensor] = None,
              work: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``-gaussian_laplace(vol, sigma) * sigma**2`` with scipy 'reflect' faces."""
    lib = _lib.load()
    Z, Y, pitch = vol.shape
    out = out if out is not None else torch.zeros_like(vol)
    if work is None:
        work = torch.zeros(lib.mmb_log_work_bytes(Z, Y, pitch), dtype=torch.uint8,
                           device=vol.device)
    _lib.check(lib.mmb_log_scale(_ptr(vol), _ptr(out), _ptr(work), Z, Y, X, pitch, float(sigma),
                                 _stream()))
    return out


def new_cand_buffer(capacity: int, device=None) -> torch.Tensor:
    return torch.zeros((capacity, 5), dtype=torch.int32, device=device or require_cuda())


def cands_to_numpy(buf: torch.Tensor, n: int) -> np.ndarray:
    host = buf[:n].cpu().numpy()
    return host.view(CAND_DTYPE).reshape(-1).copy()


def cands_from_numpy(c: np.ndarray, device=None) -> torch.Tensor:
    raw = np.ascontiguousarray(c.astype(CAND_DTYPE)).view(np.int32).reshape(-1, 5)
    return torch.from_numpy(raw.copy()).to(device or require_cuda())


def localmax(prev, cur, nxt, X: int, s: int, thr: float, cand: torch.Tensor,
             counter: torch.Tensor, z_lo: int = 0, z_hi: Optional[int] = None) -> None:
    lib = _lib.load()
    Z, Y, pitch = cur.shape
    z_hi = Z if z_hi is None else z_hi
    _lib.check(lib.mmb_localmax_compact(_ptr(prev), _ptr(cur), _ptr(nxt), Z, Y, X, pitch, int(s),
                                        float(thr), int(z_lo), int(z_hi), _ptr(cand),
                                        cand.shape[0], _ptr(counter), _stream()))


def prune_within(cand: torch.Tensor, n: int, sigmas: Sequence[float], overlap: float, Y: int,
                 X: int, z_sorted: bool = False) -> torch.Tensor:
    """keep flags of ``_prune_blobs``; ``z_sorted`` promises candidates listed by
    ascending z (``mmb_prune_within_zsorted``: same result, windowed pair search)."""
    lib = _lib.load()
    keep = torch.zeros(max(n, 1), dtype=torch.uint8, device=cand.device)
    sig = (C.c_double * len(sigmas))(*[float(s) for s in sigmas])
    fn = lib.mmb_prune_within_zsorted if z_sorted else lib.mmb_prune_within
    _lib.check(fn(_ptr(cand), int(n), sig, len(sigmas), float(overlap), int(Y), int(X),
                  _ptr(keep), _stream()))
    return keep[:n]


def prune_seams(master_zyx: torch.Tensor, check_zyx: torch.Tensor, tol: Sequence[int]):
    """Box match of int32 (n,3) coordinate tensors; returns (master_last, check_hit)."""
    lib = _lib.load()
    nm, nc = master_zyx.shape[0], check_zyx.shape[0]
    dev = master_zyx.device
    master_last = torch.full((max(nm, 1),), -1, dtype=torch.int32, device=dev)
    check_hit = torch.zeros(max(nc, 1), dtype=torch.uint8, device=dev)
    _lib.check(lib.mmb_prune_seams(_ptr(master_zyx), nm, _ptr(check_zyx), nc,
                                   _lib._I32x3(*[int(t) for t in tol]), _ptr(master_last),
                                   _ptr(check_hit), _stream()))
    return master_last[:nm], check_hit[:nc]


#: running totals over every chunk collected in this process (``bench.py`` reports them):
#: local maxima, survivors of the overlap pruning, and the size of the set whose survival
#: in scikit-image depends on its pair iteration order
STATS = {"peaks": 0, "survivors": 0, "order_dependent": 0}
#: developer aid: when a list, every collected chunk appends its status counters
#: (local maxima, survivors, kill edges, order-dependent set) - tools/side_stream_check.py
CHUNK_LOG = None


class _Slot:
    """Output buffers of one in-flight chunk: device candidates and status, and
    their pinned host mirrors."""

    def __init__(self, capacity: int, device):
        self.capacity = capacity
        self.cand = new_cand_buffer(capacity, device)
        self.status = torch.zeros(4, dtype=torch.int32, device=device)
        self.status_host = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.cand_host: Optional[torch.Tensor] = None      # pinned, grown on demand
        self.busy = False

    def host_rows(self, n: int) -> torch.Tensor:
        if self.cand_host is None or self.cand_host.shape[0] < n:
            rows = max(n, min(self.capacity, 65536))
            self.cand_host = torch.empty((rows, 5), dtype=torch.int32).pin_memory()
        return self.cand_host[:n]


@dataclass
class Ticket:
    """Handle of a chunk enqueued with ``ChunkDetector.enqueue``."""
    slot: _Slot
    event: "torch.cuda.Event"
    args: tuple                  # everything needed to redo the chunk after an overflow
    stream: Optional["torch.cuda.Stream"] = None     # the stream the chunk was enqueued on


class ChunkDetector:
    """Reusable workspace around ``mmb_detect_chunk_enqueue`` for chunks up to a
    maximum shape (the workspace is the dominant allocation: eight float
    volumes).

    ``enqueue`` launches a whole chunk without any host synchronisation and
    ``collect`` fetches its survivors later, so the host-side table assembly of
    chunk *i* overlaps the kernels of chunk *i + 1* (the workspace is shared;
    stream order serialises the chunks, only the small output slots rotate).
    ``detect`` is the synchronous pair of the two."""

    def __init__(self, max_shape: Sequence[int], capacity: Optional[int] = None, device=None,
                 n_slots: int = 4):
        self.lib = _lib.load()
        self.device = device or require_cuda()
        self.max_shape = tuple(int(s) for s in max_shape)
        Z, Y, X = self.max_shape
        nvox = Z * Y * X
        self.capacity = int(capacity) if capacity else max(4096, nvox // 256)
        self.n_slots = int(n_slots)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._alloc()

    def _alloc(self):
        Z, Y, X = self.max_shape
        nbytes = self.lib.mmb_detect_work_bytes(Z, Y, pitch_for(X), self.capacity)
        self.work = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        # tickets in flight keep their old slots alive by reference
        self.slots = [_Slot(self.capacity, self.device) for _ in range(self.n_slots)]

    def free_slots(self) -> int:
        return sum(not s.busy for s in self.slots)

    def enqueue(self, src: Source, sigmas: Sequence[float], threshold: float, overlap: float,
                scale: float = 1.0, pre: Optional[MmbPreprocParams] = None,
                block_shape: Sequence[int] = (25, 25, 25), z_lo: int = 0,
                z_hi: Optional[int] = None) -> Ticket:
        """Launch one chunk asynchronously on the current stream."""
        Z, Y, X = src.shape
        if Z > self.max_shape[0] or Y > self.max_shape[1] or X > self.max_shape[2]:
            raise ValueError(f"chunk {src.shape} exceeds workspace {self.max_shape}")
        slot = next((s for s in self.slots if not s.busy), None)
        if slot is None:
            raise RuntimeError("all output slots are in flight: collect() a ticket first")
        z_hi = Z if z_hi is None else z_hi
        sig = (C.c_double * len(sigmas))(*[float(s) for s in sigmas])
        bz, by, bx = (int(b) for b in block_shape)
        _lib.check(self.lib.mmb_detect_chunk_enqueue(
            C.c_void_p(src.ptr), src.dtype, _lib._I64x3(*src.strides), Z, Y, X, pitch_for(X),
            float(scale), C.byref(pre) if pre is not None else None, bz, by, bx, sig,
            len(sigmas), float(threshold), float(overlap), int(z_lo), int(z_hi),
            _ptr(self.work), _ptr(slot.cand), slot.capacity, _ptr(slot.status), _stream()))
        slot.status_host.copy_(slot.status, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        slot.busy = True
        return Ticket(slot, ev, (src, sigmas, threshold, overlap, scale, pre, block_shape, z_lo,
                                 z_hi), torch.cuda.current_stream())

    def collect(self, ticket: Ticket) -> Tuple[np.ndarray, int]:
        """Wait for a chunk; returns ``(survivors as CAND_DTYPE records, number of
        local maxima)``.  A chunk whose candidates or kill edges overflowed the
        buffers (counted exactly on the device) is redone once with larger ones."""
        slot = ticket.slot
        ticket.event.synchronize()
        n_peaks, n_out, n_edges, n_od = (int(v) for v in slot.status_host)
        edge_cap = self.lib.mmb_detect_edge_capacity(slot.capacity)
        if n_peaks > slot.capacity or n_edges > edge_cap:
            slot.busy = False
            need = max(n_peaks, (n_edges - 4096) // 4 + 1)
            self.capacity = max(self.capacity, int(need * 1.25) + 1024)
            self._alloc()
            return self.detect(*ticket.args)
        STATS["peaks"] += n_peaks
        STATS["survivors"] += n_out
        STATS["order_dependent"] += n_od
        if n_out == 0:
            slot.busy = False
            return np.zeros(0, dtype=CAND_DTYPE), n_peaks
        host = slot.host_rows(n_out)
        with torch.cuda.stream(self.copy_stream):
            host.copy_(slot.cand[:n_out], non_blocking=True)
        self.copy_stream.synchronize()
        out = host.numpy().view(CAND_DTYPE).reshape(-1).copy()
        slot.busy = False
        return out, n_peaks

    def collect_device(self, ticket: Ticket) -> Tuple[torch.Tensor, int]:
        """Like ``collect`` but the survivors stay on the device: returns ``((n, 5)
        int32 tensor of candidate records, number of local maxima)``.  Only the
        three status counters cross to the host."""
        slot = ticket.slot
        ticket.event.synchronize()
        n_peaks, n_out, n_edges, n_od = (int(v) for v in slot.status_host)
        edge_cap = self.lib.mmb_detect_edge_capacity(slot.capacity)
        if n_peaks > slot.capacity or n_edges > edge_cap:
            slot.busy = False
            need = max(n_peaks, (n_edges - 4096) // 4 + 1)
            self.capacity = max(self.capacity, int(need * 1.25) + 1024)
            self._alloc()
            with torch.cuda.stream(ticket.stream or torch.cuda.current_stream()):
                redo = self.enqueue(*ticket.args)
            return self.collect_device(redo)
        STATS["peaks"] += n_peaks
        STATS["survivors"] += n_out
        STATS["order_dependent"] += n_od
        if CHUNK_LOG is not None:
            CHUNK_LOG.append((tuple(ticket.args[0].shape), n_peaks, n_out, n_edges, n_od))
        # copy on the stream the chunk ran on: the slot may then be reused by a later
        # chunk of that stream right away (consumers on another stream must order
        # themselves after it)
        with torch.cuda.stream(ticket.stream or torch.cuda.current_stream()):
            out = slot.cand[:n_out].clone()
        slot.busy = False
        return out, n_peaks

    def detect(self, src: Source, sigmas: Sequence[float], threshold: float, overlap: float,
               scale: float = 1.0, pre: Optional[MmbPreprocParams] = None,
               block_shape: Sequence[int] = (25, 25, 25), z_lo: int = 0,
               z_hi: Optional[int] = None) -> Tuple[np.ndarray, int]:
        """Returns ``(survivors as CAND_DTYPE records, number of local maxima)``."""
        return self.collect(self.enqueue(src, sigmas, threshold, overlap, scale, pre,
                                         block_shape, z_lo, z_hi))


def whole_roi_preprocess(src: Source, params: MmbPreprocParams) -> np.ndarray:
    """``saturate_roi`` / ``denoise_roi`` with the whole ROI as one block (the
    GUI path, magmap/gui/visualizer.py:2742-2743).  ROIs up to 32 voxels a side run
    in the one-CTA kernel, larger ones in the whole-volume kernels of
    ``preprocess_large.cu`` (any size that fits the device)."""
    Z, Y, X = src.shape
    out = preprocess_blocks(src, (Z, Y, X), params)
    torch.cuda.synchronize()
    return out[:, :, :X].cpu().numpy()
