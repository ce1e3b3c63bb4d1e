"""ctypes binding of ``libmmb200.so`` (``include/mmb200.h``).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C
magellanmapper_b200/csrc``).  There is no fallback: if it is missing or a
symbol is absent, ``load()`` raises and every compute entry point of this
package fails with it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMB200_LIB") or os.path.join(_HERE, "libmmb200.so")   # env override: developer A/B builds

MMB_U8, MMB_U16, MMB_F32, MMB_F64 = 0, 1, 2, 3
MMB_OK, MMB_ERR_INVALID, MMB_ERR_CUDA, MMB_ERR_OVERFLOW, MMB_ERR_UNSUPPORTED = 0, -1, -2, -3, -4


class MmbCand(C.Structure):
    _fields_ = [("z", C.c_int32), ("y", C.c_int32), ("x", C.c_int32),
                ("s", C.c_int32), ("resp", C.c_float)]


class MmbPreprocParams(C.Structure):
    _fields_ = [("clip_vmin", C.c_double), ("clip_vmax", C.c_double),
                ("max_thresh", C.c_double), ("clip_min", C.c_double),
                ("clip_max", C.c_double), ("unsharp_strength", C.c_double),
                ("erosion_threshold", C.c_double)]


class MmbRow(C.Structure):
    _fields_ = [("z", C.c_int32), ("y", C.c_int32), ("x", C.c_int32), ("s", C.c_int32),
                ("resp", C.c_float), ("chunk", C.c_int32), ("channel", C.c_int32),
                ("reserved", C.c_int32)]


class MmbStackGeom(C.Structure):
    _fields_ = [("grid", C.c_int32 * 3), ("overlap", C.c_int32 * 3), ("tol", C.c_int32 * 3),
                ("pad", C.c_int32 * 3), ("start", C.POINTER(C.c_int32) * 3),
                ("size", C.POINTER(C.c_int32) * 3), ("n_channels", C.c_int32),
                ("num_sigma", C.c_int32), ("sigmas", C.POINTER(C.c_double)),
                ("channel_ids", C.POINTER(C.c_int32))]


class MmbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mmb200 error {code}: {msg}")
        self.code = code


class MmbOverflow(MmbError):
    pass


_I64x3 = C.c_int64 * 3
_I32x3 = C.c_int32 * 3
_vp = C.c_void_p

#: every exported symbol of include/mmb200.h with (restype, argtypes)
SIGNATURES = {
    "mmb_version": (C.c_int, []),
    "mmb_last_error": (C.c_char_p, []),
    "mmb_launch_count": (C.c_int64, []),
    "mmb_profile_enable": (C.c_int, [C.c_int]),
    "mmb_profile_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_double)]),
    "mmb_upload_pieces": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, C.c_int64, _vp]),
    "mmb_percentiles_work_bytes": (C.c_int64, [C.c_int]),
    "mmb_percentiles": (C.c_int, [_vp, C.c_int, _I64x3, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_double), C.c_int, _vp, _vp, _vp]),
    "mmb_to_float": (C.c_int, [_vp, C.c_int, _I64x3, C.c_int, C.c_int, C.c_int, _vp,
                               C.c_int64, C.c_double, _vp]),
    "mmb_preprocess_blocks": (C.c_int, [_vp, C.c_int, _I64x3, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int,
                                        C.POINTER(MmbPreprocParams), _vp, C.c_int64, _vp]),
    "mmb_resize_linear": (C.c_int, [_vp, C.c_int, _I64x3, C.c_int, C.c_int, C.c_int, _vp, C.c_int,
                                    C.c_int, C.c_int, C.c_int64, C.c_int, _vp]),
    "mmb_unmix_subtract": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64,
                                     C.c_double, _vp]),
    "mmb_coloc_work_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "mmb_coloc_sums": (C.c_int, [_vp, C.c_int, C.c_int64 * 4, C.c_int, C.c_int, C.c_int, C.c_int,
                                 _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "mmb_log_work_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int64]),
    "mmb_log_scale": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64,
                                C.c_double, _vp]),
    "mmb_log_pass": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64,
                               C.c_int, C.c_int, C.c_double, C.c_double, _vp]),
    "mmb_localmax_compact": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64,
                                       C.c_int, C.c_float, C.c_int, C.c_int, _vp, C.c_int,
                                       _vp, _vp]),
    "mmb_prune_within": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double,
                                   C.c_int, C.c_int, _vp, _vp]),
    "mmb_prune_within_zsorted": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double), C.c_int,
                                           C.c_double, C.c_int, C.c_int, _vp, _vp]),
    "mmb_prune_seams": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _I32x3, _vp, _vp, _vp]),
    "mmb_rows_from_cands": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp]),
    "mmb_cands_append": (C.c_int, [_vp, C.c_int, _vp, _I32x3, _I32x3, _I32x3, _vp, C.c_int, _vp,
                                   _vp]),
    "mmb_stack_tables_work_bytes": (C.c_int64, [C.c_int]),
    "mmb_stack_tables": (C.c_int, [_vp, C.c_int, C.POINTER(MmbStackGeom), C.c_int, _vp, _vp, _vp,
                                   _vp, _vp]),
    "mmb_detect_work_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int64, C.c_int]),
    "mmb_detect_edge_capacity": (C.c_int, [C.c_int]),
    "mmb_detect_chunk_enqueue": (C.c_int, [_vp, C.c_int, _I64x3, C.c_int, C.c_int, C.c_int,
                                           C.c_int64, C.c_double, C.POINTER(MmbPreprocParams),
                                           C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                           C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                           _vp, _vp, C.c_int, _vp, _vp]),
    "mmb_detect_chunk": (C.c_int, [_vp, C.c_int, _I64x3, C.c_int, C.c_int, C.c_int, C.c_int64,
                                   C.c_double, C.POINTER(MmbPreprocParams), C.c_int, C.c_int,
                                   C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double,
                                   C.c_double, C.c_int, C.c_int, _vp, _vp, C.c_int,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int), _vp]),
}

#: bench / test utilities of include/mmb200_tools.h (not the reference-facing boundary)
TOOLS_SIGNATURES = {
    "mmb_debug_smem_poison": (C.c_int, [C.c_int]),
    "mmb_log_xy_fused": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64,
                                   C.c_double, _vp]),
    "mmb_synth_nuclei": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                   C.c_int64, C.c_uint64, C.c_double, _vp]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the library and bind every symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SIGNATURES.items()) + list(TOOLS_SIGNATURES.items()):
        fn = getattr(lib, name)          # AttributeError if the symbol is absent
        fn.restype = res
        fn.argtypes = args
    if lib.mmb_version() != 1:
        raise ImportError(f"libmmb200 version {lib.mmb_version()} != 1")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == MMB_OK:
        return
    msg = load().mmb_last_error().decode("utf-8", "replace")
    if rc == MMB_ERR_OVERFLOW:
        raise MmbOverflow(rc, msg)
    if rc == MMB_ERR_INVALID:
        raise ValueError(f"mmb200: {msg}")
    if rc == MMB_ERR_UNSUPPORTED:
        raise NotImplementedError(f"mmb200: {msg}")
    raise MmbError(rc, msg)
