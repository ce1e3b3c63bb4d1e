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
    """``os.path.splitext`` that keeps multi-period extensions such as ``.nii.gz`` whole
    (libmag.py:272-293)."""
    i = -1
    for ext in _EXTENSIONS_MULTIPLE:
        i = path.rfind(ext)
        if i != -1:
            break
    return os.path.splitext(path) if i == -1 else (path[:i], path[i:])


def insert_before_ext(name, insert: str, sep: str = "") -> str:
    """Splice ``insert`` in front of the extension of ``name`` (libmag.py:247-269)."""
    name = str(name)
    if os.path.basename(name).find(".") == -1:
        return name + sep + insert
    return "{0}{2}{3}.{1}".format(*name.rsplit(".", 1), sep, insert)


def combine_paths(base_path: Optional[str], suffix: str, sep: str = "_",
                  ext: Optional[str] = None, check_dir: bool = False,
                  keep_ext: bool = False) -> str:
    """``base_path`` (without its extension unless ``keep_ext``) + ``sep`` + ``suffix``; a
    directory (trailing separator, or an existing one with ``check_dir``) is joined instead;
    ``ext`` replaces the extension of the result (libmag.py:331-369)."""
    if not base_path:
        return suffix
    if not os.path.basename(base_path) or check_dir and os.path.isdir(base_path):
        path = os.path.join(base_path, suffix)
    else:
        path = base_path if keep_ext else splitext(base_path)[0]
        path = path + sep + suffix
    if ext:
        path = f"{splitext(path)[0]}.{ext}"
    return path


def backup_file(path, modifier: str = "", i: Optional[int] = None) -> None:
    """Move an existing file to ``name[modifier](i).ext`` with the first free ``i`` - to
    ``name[modifier].ext`` itself when a modifier is given and that name is free - and do
    the same, with the same index, for a file that travels with it (libmag.py:969-1015)."""
    if not i:
        if not os.path.exists(path):
            return
        i = 0
    while True:
        if i == 0 and modifier != "":
            backup_path = insert_before_ext(path, modifier)
        else:
            if i == 0:
                i = 1
            backup_path = insert_before_ext(path, "{}({})".format(modifier, i))
        if not os.path.exists(backup_path):
            shutil.move(path, backup_path)
            root, ext = os.path.splitext(path)
            associated = _FILE_TYPE_GROUPS.get(ext[1:])
            if associated:
                backup_file("{}.{}".format(root, associated), modifier, i)
            break
        i += 1
