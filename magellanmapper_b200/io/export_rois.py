"""Blob export sinks of ``magmap/io/export_rois.py`` that consume the detector's table."""
from __future__ import annotations

import os

import numpy as np


def blobs_to_csv(blobs: np.ndarray, path: str) -> str:
    """Write ``z, y, x, r`` of every blob to ``<path without extension>_blobs.csv.gz``
    (export_rois.py:278-289: ``np.savetxt`` with a ``z,y,x,r`` header, GZIP by suffix).
    Returns the path written."""
    path_out = "{}_blobs.csv.gz".format(os.path.splitext(path)[0])
    np.savetxt(path_out, np.asarray(blobs)[:, :4], delimiter=",", header="z,y,x,r")
    return path_out
