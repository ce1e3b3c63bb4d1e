"""The reference's OWN unit tests for this path (``magmap/tests/test_chunking.py``,
``magmap/tests/test_detector.py``), unmodified, executed against the drop-in: the names
``magmap.cv.chunking``, ``magmap.cv.detector`` and ``magmap.settings.config`` are bound to
the mirror's modules the way ``INTEGRATION.md`` (B) tells a maintainer to alias them, then the
reference's test files are loaded from ``/root/reference`` and run.  Build container only
(the reference does not travel); ``tests/test_host_mirror.py`` restates the same cases for
everywhere else."""
import importlib.util
import os
import sys
import types
import unittest

import pytest

REF_TESTS = "/root/reference/magmap/tests"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_TESTS),
                                reason="the unmodified reference is only in the build container")


@pytest.mark.parametrize("name", ["test_chunking", "test_detector"])
def test_reference_unit_tests_pass_on_the_mirror(name):
    from magellanmapper_b200.cv import chunking, detector
    from magellanmapper_b200.settings import config

    saved = {k: sys.modules.get(k) for k in (
        "magmap", "magmap.cv", "magmap.settings", "magmap.cv.chunking", "magmap.cv.detector",
        "magmap.settings.config")}
    try:
        # packages that hold nothing but the aliased modules: no reference code is imported
        for pkg in ("magmap", "magmap.cv", "magmap.settings"):
            mod = types.ModuleType(pkg)
            mod.__path__ = []
            sys.modules[pkg] = mod
        for full, mod in (("magmap.cv.chunking", chunking), ("magmap.cv.detector", detector),
                          ("magmap.settings.config", config)):
            sys.modules[full] = mod
            parent, leaf = full.rsplit(".", 1)
            setattr(sys.modules[parent], leaf, mod)
        spec = importlib.util.spec_from_file_location(f"_ref_{name}", os.path.join(REF_TESTS, name + ".py"))
        ref_tests = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_tests)
        suite = unittest.defaultTestLoader.loadTestsFromModule(ref_tests)
        assert suite.countTestCases() >= 1
        with open(os.devnull, "w") as sink:
            result = unittest.TextTestRunner(stream=sink, verbosity=0).run(suite)
        assert result.wasSuccessful(), result.failures + result.errors
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _run_selected(ref_file, aliases, keep):
    """Load one of the reference's test files with ``aliases`` bound and run the test
    methods whose names ``keep`` accepts."""
    saved = {}
    pkgs = {"magmap"} | {full.rsplit(".", 1)[0] for full in aliases}
    try:
        for pkg in sorted(pkgs):
            saved[pkg] = sys.modules.get(pkg)
            mod = types.ModuleType(pkg)
            mod.__path__ = []
            sys.modules[pkg] = mod
        for full, mod in aliases.items():
            saved[full] = sys.modules.get(full)
            sys.modules[full] = mod
            parent, leaf = full.rsplit(".", 1)
            setattr(sys.modules[parent], leaf, mod)
        spec = importlib.util.spec_from_file_location(
            "_ref_" + os.path.basename(ref_file)[:-3], ref_file)
        ref_tests = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_tests)
        suite = unittest.TestSuite()
        for case in unittest.defaultTestLoader.loadTestsFromModule(ref_tests):
            for t in case:
                if keep(t.id().rsplit(".", 1)[1]):
                    suite.addTest(t)
        n = suite.countTestCases()
        with open(os.devnull, "w") as sink:
            result = unittest.TextTestRunner(stream=sink, verbosity=0).run(suite)
        assert result.wasSuccessful(), result.failures + result.errors
        return n
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_reference_np_io_and_libmag_unit_tests_on_the_mirror():
    """``test_np_io.py`` whole, and of ``test_libmag.py`` the cases for the helpers this path
    uses and the mirror therefore carries (path splitting and suffixing for the archive and
    image-feed file names)."""
    from magellanmapper_b200.io import libmag, np_io
    assert _run_selected(os.path.join(REF_TESTS, "test_np_io.py"),
                         {"magmap.io.np_io": np_io}, lambda name: True) == 1
    mirrored = {"test_insert_before_ext", "test_splitext", "test_combine_paths"}
    assert _run_selected(os.path.join(REF_TESTS, "test_libmag.py"),
                         {"magmap.io.libmag": libmag}, mirrored.__contains__) == 3
