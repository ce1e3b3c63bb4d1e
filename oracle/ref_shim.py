"""Import the UNMODIFIED reference from /root/reference in this container.

TEST INFRASTRUCTURE ONLY; used by ``oracle/make_golden.py`` to generate the
vectors under ``tests/golden/``.  Never imported on the GPU box
(``/root/reference`` does not exist there).

The reference needs wheels this image lacks (scikit-image, appdirs, tifffile,
matplotlib, SimpleITK, ...).  Those module names are satisfied by mocks so the
reference's own Python (chunking, detector glue, stack_detect, plot_3d) runs
as written; the handful of scikit-image FUNCTIONS the hot path calls are wired
to ``oracle.skimage_restated`` (the mock cannot compute).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import sys
from unittest import mock

REFERENCE_ROOT = "/root/reference"
MOCK_TOP = {"skimage", "appdirs", "tifffile", "matplotlib", "SimpleITK",
            "mpl_toolkits", "javabridge", "bioformats", "traits", "traitsui",
            "pyface", "mayavi", "tvtk", "vtk", "brainglobe_atlasapi",
            "bg_atlasapi", "boto3", "keras", "tensorflow"}


class _MockLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _MissingFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] not in MOCK_TOP:
            return None
        return importlib.machinery.ModuleSpec(name, _MockLoader(), is_package=True)


_loaded = None


def load_reference():
    """Return a namespace of reference modules with skimage calls wired."""
    global _loaded
    if _loaded is not None:
        return _loaded
    sys.meta_path.append(_MissingFinder())
    sys.path.insert(0, REFERENCE_ROOT)
    from magmap.cv import chunking, cv_nd, detector, stack_detect      # noqa
    from magmap.plot import plot_3d                              # noqa
    from magmap.settings import config, roi_prof                 # noqa
    from magmap.io import np_io                                  # noqa
    from oracle import skimage_restated as ski

    # detector.py:25,931 - blob_log
    detector.blob_log = ski.blob_log
    # plot_3d.py:157,165 - filters.gaussian, morphology.erosion/octahedron
    plot_3d.filters.gaussian = ski.filters_gaussian
    plot_3d.morphology.octahedron = lambda r: ski.octahedron1()
    plot_3d.morphology.erosion = lambda img, fp: ski.erosion_octahedron1(img)

    # colocalizer.py:373,395 - morphology.ball, morphology.dilation
    from magmap.cv import colocalizer
    colocalizer.morphology.ball = ski.ball
    colocalizer.morphology.dilation = lambda img, footprint=None, selem=None: ski.dilation(
        img, footprint if footprint is not None else selem)

    # cv_nd.py:1147 - transform.resize (make_isotropic)
    cv_nd.transform.resize = ski.transform_resize

    class NS:
        pass
    ns = NS()
    ns.chunking, ns.detector, ns.stack_detect = chunking, detector, stack_detect
    ns.cv_nd = cv_nd
    ns.colocalizer = colocalizer
    ns.plot_3d, ns.config, ns.roi_prof, ns.np_io = plot_3d, config, roi_prof, np_io
    _loaded = ns
    return ns


def set_profile(ns, resolution, near_max=-1.0, **mods):
    """Set the config globals the path reads (config.py:211,246,883-902)."""
    prof = ns.roi_prof.ROIProfile()
    prof.add_profiles("/root/reference/profiles/roi_blobs.yaml")
    for k, v in mods.items():
        prof[k] = v
    ns.config.roi_profile = prof
    ns.config.roi_profiles = [prof]
    ns.config.resolutions = [list(resolution)]
    ns.config.near_max = [near_max]
    ns.config.channel = None
    return prof
