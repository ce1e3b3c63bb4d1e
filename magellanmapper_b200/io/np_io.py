"""``Image5d`` container, archive reader and the ``.npy`` image feed of
``magmap/io/np_io.py`` (``:33-71, 159-177, 193-592, 787-862``): an imported image is an
``<base>_image5d.npy`` file opened as a read-only memory map plus a ``<base>_meta.yml`` file
with resolutions and the per-channel ``near_min`` / ``near_max`` the preprocessing needs.
The stack detector streams such a map to the GPU strip by strip (``gpu.StripFeeder``); it is
never read into host memory whole."""
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


def get_num_channels(img=None, is_3d: bool = False) -> int:
    """Channel count of a ``t,z,y,x[,c]`` image, or of a ``z,y,x[,c]`` one with ``is_3d``
    (np_io.py:610-625)."""
    axis = 3 if is_3d else 4
    if img is None or img.ndim <= axis:
        return 1
    return int(img.shape[axis])


def read_np_archive(archive) -> Dict[str, Any]:
    out = {}
    for key in archive.keys():
        try:
            out[key] = archive[key]
        except ValueError:
            print(f"unable to load {key} from archive, will ignore")
    return out


def setup_images(path: str, series=None, offset=None, size=None, proc_type=None,
                 allow_import: bool = True, fallback_main_img: bool = True, bg_atlas=None,
                 labels_ref_path=None) -> Image5d:
    """Open the imported image of ``path`` for detection (the ``.npy`` route of
    np_io.py:193-592): memory-map ``<base>_image5d.npy`` read-only, load
    ``<base>_meta.yml`` into ``config`` (resolutions, ``near_min`` / ``near_max``,
    magnification, zoom) and set ``config.filename``.  With ``offset`` / ``size`` (z, y, x)
    the returned image is that view of the map, as the saved sub-image route (:283-296)
    would hand over.  The remaining parameters are the reference's; only their defaults are
    served (no processing-type specific loading, no import of other formats, no atlas or
    label images - all outside the accelerated path).

    Raises:
        FileNotFoundError: if the image file does not exist (other formats - TIFF, CZI
            import - are outside the accelerated path).
        NotImplementedError: for ``proc_type``, ``bg_atlas`` or ``labels_ref_path``.
    """
    import os
    from . import importer
    from ..settings import config
    for name, val in (("proc_type", proc_type), ("bg_atlas", bg_atlas),
                      ("labels_ref_path", labels_ref_path)):
        if val is not None:
            raise NotImplementedError(f"setup_images({name}=...) is outside the detection path")
    base = path
    for suffix in ("_" + config.SUFFIX_IMAGE5D, "_" + config.SUFFIX_META):
        if base.endswith(suffix):
            base = base[:-len(suffix)]
    filename_image5d, filename_meta = importer.make_filenames(base, series)
    if not os.path.exists(filename_image5d):
        filename_image5d, filename_meta = importer.make_filenames(base, series, keep_ext=True)
    if not os.path.exists(filename_image5d):
        raise FileNotFoundError(f"no imported image {filename_image5d}")
    img5d = Image5d(np.load(filename_image5d, mmap_mode="r"), filename_image5d, filename_meta)
    importer.load_metadata(filename_meta, img5d=img5d)
    if offset is not None and size is not None:
        z, y, x = (int(v) for v in offset)
        sz, sy, sx = (int(v) for v in size)
        img5d.img = img5d.img[:, z:z + sz, y:y + sy, x:x + sx]
        img5d.subimg_offset, img5d.subimg_size = offset, size
    config.filename = path
    return img5d


def write_npy(image5d, md: Dict[Any, Any], path: str, find_near_bounds: bool = True) -> None:
    """Save a ``t, z, y, x[, c]`` image as the reference's imported-image pair
    (np_io.py:787-862): ``near_min`` / ``near_max`` per channel from the per-plane 0.5 /
    99.5 percentiles (on the GPU, ``importer.calc_near_bounds``), the metadata file, and the
    voxels plane by plane through a memory map.  ``md``: ``config.MetaKeys.RESOLUTIONS`` /
    ``MAGNIFICATION`` / ``ZOOM`` (or the lower-case names as strings).  An existing image file is left alone."""
    import os
    from enum import Enum
    from . import importer
    # the reference keys the dictionary with `config.MetaKeys` members; plain lower-case
    # names are accepted as well
    md = {(k.name.lower() if isinstance(k, Enum) else k): v for k, v in md.items()}
    from . import libmag
    filename_image5d, filename_meta = importer.make_filenames(libmag.splitext(str(path))[0],
                                                              keep_ext=True)
    if os.path.exists(filename_image5d):
        print(f"File {filename_image5d} already exists, skipping saving image5d")
        return
    if find_near_bounds:
        near_mins, near_maxs = importer.calc_near_bounds(image5d[0], dim_channel=3)
    else:
        info = np.iinfo(image5d.dtype) if np.issubdtype(image5d.dtype, np.integer) \
            else np.finfo(image5d.dtype)
        near_mins, near_maxs = [info.min], [info.max]
    importer.save_image_info(
        filename_meta, [os.path.basename(path)], [tuple(image5d.shape)], md.get("resolutions"),
        md.get("magnification"), md.get("zoom"), near_mins, near_maxs)
    out = np.lib.format.open_memmap(filename_image5d, mode="w+", dtype=image5d.dtype,
                                    shape=tuple(int(v) for v in image5d.shape))
    for t in range(image5d.shape[0]):
        for z in range(image5d.shape[1]):
            out[t, z] = image5d[t, z]
        out.flush()
    del out
