"""The few ``magmap/io/libmag.py`` helpers the detection path uses."""
from __future__ import annotations

import os
import shutil
from typing import Callable, Optional, Sequence

import numpy as np

from ..settings import config


def is_seq(val) -> bool:
    """Non-string sequence or array with at least one dimension (libmag.py:1176)."""
    return isinstance(val, (list, tuple)) or np.ndim(val) != 0


def printv(*args):
    if config.verbose:
        print(*args)


def combine_arrs(arrs, filter_none: bool = True, fn: Optional[Callable] = None, **kwargs):
    """Concatenate (or ``fn``) arrays after dropping ``None`` entries; a single
    survivor is returned as is, nothing gives ``None`` (libmag.py:196-227)."""
    if arrs is None:
        return None
    if filter_none:
        arrs = [a for a in arrs if a is not None]
    if len(arrs) == 0:
        return None
    if len(arrs) == 1:
        return arrs[0]
    return (fn or np.concatenate)(arrs, **kwargs)


_SIGNED = (np.int8, np.int16, np.int32, np.int64)
_UNSIGNED = (np.uint8, np.uint16, np.uint32, np.uint64)
_FLOATS = (np.float16, np.float32, np.float64)


def dtype_within_range(min_val, max_val, integer=None, signed=None):
    """Smallest dtype whose range contains ``[min_val, max_val]`` (libmag.py:1116)."""
    if signed is None:
        signed = min_val < 0
    if integer is None:
        integer = float(max_val).is_integer()
    if integer:
        cands, info = (_SIGNED if signed else _UNSIGNED), np.iinfo
    else:
        cands, info = _FLOATS, np.finfo
    for dt in cands:
        if info(dt).min <= min_val and info(dt).max >= max_val:
            return dt
    raise TypeError(f"no dtype (integer={integer}, signed={signed}) holds "
                    f"{min_val}..{max_val}")


def splitext(path: str):
    return os.path.splitext(path)


def insert_before_ext(path: str, insert: str, sep: str = "") -> str:
    root, ext = os.path.splitext(path)
    return f"{root}{sep}{insert}{ext}"


def combine_paths(base_path: Optional[str], suffix: str, sep: str = "_",
                  ext: Optional[str] = None) -> str:
    """``base`` without its extension + sep + ``suffix`` (libmag.py:331-369)."""
    if not base_path:
        return suffix
    if not os.path.basename(base_path):
        path = os.path.join(base_path, suffix)
    else:
        path = os.path.splitext(base_path)[0] + sep + suffix
    if ext:
        path = f"{os.path.splitext(path)[0]}.{ext}"
    return path


def backup_file(path: str, modifier: str = "") -> None:
    """Move an existing file to ``name(modifier)(i).ext`` with the first free
    ``i`` (libmag.py:969-1015)."""
    if not os.path.exists(path):
        return
    i = 1
    while True:
        cand = insert_before_ext(path, f"{modifier}({i})")
        if not os.path.exists(cand):
            shutil.move(path, cand)
            return
        i += 1
