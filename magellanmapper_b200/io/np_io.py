"""``Image5d`` container and archive reader (``magmap/io/np_io.py:33-71,159-177``)."""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

import numpy as np


class Image5d:
    """Main image holder: ``img`` is ``t, z, y, x[, c]``."""

    def __init__(self, img=None, path_img: Optional[str] = None,
                 path_meta: Optional[str] = None, img_io=None):
        self.img = img
        self.path_img = path_img
        self.path_meta = path_meta
        self.img_io = img_io
        self.subimg_offset: Optional[Sequence[int]] = None
        self.subimg_size: Optional[Sequence[int]] = None
        self.meta: Optional[Dict[Any, Any]] = None
        self.rgb = False
        self.is_roi = False
        self.shapes = None


def read_np_archive(archive) -> Dict[str, Any]:
    out = {}
    for key in archive.keys():
        try:
            out[key] = archive[key]
        except ValueError:
            print(f"unable to load {key} from archive, will ignore")
    return out
