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


def printv(*s):
    """Print only when ``config.verbose`` (libmag.py:488-496)."""
    if config.verbose:
        print(*s)


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


#: extensions with more than one period (libmag.py:43)
_EXTENSIONS_MULTIPLE = (".tar", ".nii")
#: files that travel together: a backup of the first also backs up the second (libmag.py:30-33)
_FILE_TYPE_GROUPS = {"obj": "mtl", "mhd": "raw"}


def splitext(path):
    """``os.path.splitext`` that keeps multi-period extensions such as ``.nii.gz`` whole:
    the split is at the last occurrence of the first listed extension start that the path
    contains (libmag.py:272-293)."""
    for start in _EXTENSIONS_MULTIPLE:
        at = path.rfind(start)
        if at >= 0:
            return path[:at], path[at:]
    return os.path.splitext(path)


def insert_before_ext(name, insert: str, sep: str = "") -> str:
    """``name`` with ``sep + insert`` spliced in front of its last extension; appended
    when the base name has no period (libmag.py:247-269)."""
    name = str(name)
    if "." not in os.path.basename(name):
        return f"{name}{sep}{insert}"
    stem, _, ext = name.rpartition(".")
    return f"{stem}{sep}{insert}.{ext}"


def combine_paths(base_path: Optional[str], suffix: str, sep: str = "_",
                  ext: Optional[str] = None, check_dir: bool = False,
                  keep_ext: bool = False) -> str:
    """``base_path`` (without its extension unless ``keep_ext``) + ``sep`` + ``suffix``; a
    directory - a trailing separator, or an existing directory when ``check_dir`` - is
    joined with ``suffix`` instead; ``ext`` replaces the extension of the result; no base
    path gives ``suffix`` (libmag.py:331-369)."""
    if not base_path:
        return suffix
    is_dir = not os.path.basename(base_path) or (check_dir and os.path.isdir(base_path))
    if is_dir:
        out = os.path.join(base_path, suffix)
    else:
        stem = base_path if keep_ext else splitext(base_path)[0]
        out = f"{stem}{sep}{suffix}"
    return f"{splitext(out)[0]}.{ext}" if ext else out


def backup_file(path, modifier: str = "", i: Optional[int] = None) -> None:
    """Move an existing file out of the way: to ``name<modifier>.ext`` when a modifier is
    given and that name is free, else to ``name<modifier>(n).ext`` with the first free
    ``n >= 1``.  A file that travels with it (``.mtl`` of an ``.obj``, ``.raw`` of an
    ``.mhd``) is moved the same way starting from the same ``n``, which is what ``i`` is
    for (libmag.py:969-1015)."""
    n = int(i) if i else 0
    if n == 0 and not os.path.exists(path):
        return
    while True:
        if n == 0 and modifier == "":
            n = 1
        target = insert_before_ext(path, modifier if n == 0 else f"{modifier}({n})")
        if not os.path.exists(target):
            break
        n += 1
    shutil.move(path, target)
    stem, ext = os.path.splitext(path)
    partner = _FILE_TYPE_GROUPS.get(ext[1:])
    if partner:
        backup_file(f"{stem}.{partner}", modifier, n)
