"""Property tests of the chunk geometry on random, ragged shapes: the host mirror against the
oracle, and - in the build container, where ``/root/reference`` exists - against the unmodified
reference itself (``magmap/cv/chunking.py:214-256``, ``stack_detect.py:282-335``)."""
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from magellanmapper_b200.cv import chunking, stack_detect
from magellanmapper_b200.settings import config, roi_prof
from oracle import magmap_restated as mm

HAVE_REFERENCE = os.path.isdir("/root/reference/magmap")

axis = st.integers(1, 40)
shapes = st.tuples(axis, axis, axis)
pixels = st.tuples(st.integers(1, 20), st.integers(1, 20), st.integers(1, 20))
overlaps = st.one_of(st.none(), st.tuples(st.integers(0, 4), st.integers(0, 4), st.integers(0, 4)))


def _equal_grids(a, b):
    (sl_a, off_a), (sl_b, off_b) = a, b
    assert sl_a.shape == sl_b.shape
    for c in np.ndindex(*sl_a.shape):
        assert tuple(sl_a[c]) == tuple(sl_b[c]), c
    np.testing.assert_array_equal(off_a, off_b)


@settings(max_examples=150, deadline=None)
@given(shapes, pixels, overlaps)
def test_stack_splitter_mirror_equals_oracle_and_covers_the_stack(shape, max_pixels, overlap):
    got = chunking.stack_splitter(shape, max_pixels, overlap)
    _equal_grids(got, mm.stack_splitter(shape, max_pixels, overlap))
    # every voxel belongs to a chunk, cores tile the stack without gaps, and a chunk
    # reaches past its core by the overlap unless the stack ends first
    sl = got[0]
    seen = np.zeros(shape, dtype=np.uint8)
    for c in np.ndindex(*sl.shape):
        seen[sl[c]] = 1
        for a in range(3):
            assert sl[c][a].start == c[a] * max_pixels[a]
            want = sl[c][a].start + max_pixels[a] + (0 if overlap is None else overlap[a])
            assert sl[c][a].stop == min(want, shape[a])
    assert seen.all()


@settings(max_examples=60, deadline=None)
@given(shapes, pixels, st.tuples(st.integers(0, 3), st.integers(0, 3), st.integers(0, 3)))
def test_split_and_remerge_is_the_identity(shape, max_pixels, overlap):
    roi = np.arange(int(np.prod(shape))).reshape(shape)
    sl, _ = chunking.stack_splitter(shape, max_pixels, overlap)
    subs = np.empty(sl.shape, dtype=object)
    for c in np.ndindex(*sl.shape):
        subs[c] = roi[sl[c]]
    np.testing.assert_array_equal(
        chunking.merge_split_stack(subs, max_pixels, np.array(overlap)), roi)


@pytest.mark.skipif(not HAVE_REFERENCE, reason="the unmodified reference is only in the build container")
def test_random_geometry_against_the_unmodified_reference():
    """Live run of the reference's ``stack_splitter`` and ``setup_blocks`` on 60 random
    geometries (the committed vectors hold a handful of fixed ones)."""
    from oracle import ref_shim
    ns = ref_shim.load_reference()
    rng = np.random.default_rng(17)
    for _ in range(60):
        shape = tuple(int(v) for v in rng.integers(1, 60, 3))
        max_pixels = tuple(int(v) for v in rng.integers(1, 25, 3))
        overlap = None if rng.random() < 0.3 else tuple(int(v) for v in rng.integers(0, 5, 3))
        _equal_grids(chunking.stack_splitter(shape, max_pixels, overlap),
                     ns.chunking.stack_splitter(shape, max_pixels, overlap))
    for _ in range(25):
        shape = tuple(int(v) for v in rng.integers(20, 700, 3))
        res = [float(v) for v in rng.choice([0.5, 0.913, 1.0, 2.5, 5.0, 6.6], 3)]
        mods = {"segment_size": int(rng.integers(20, 500)),
                "denoise_size": int(rng.choice([10, 25, 2000])),
                "prune_tol_factor": tuple(float(v) for v in rng.choice([0.9, 1.0, 1.5], 3)),
                "exclude_border": None if rng.random() < 0.5 else (1, 0, 0)}
        ref_prof = ref_shim.set_profile(ns, res, **mods)
        want = ns.stack_detect.setup_blocks(ref_prof, shape)
        prof = roi_prof.ROIProfile()
        prof.add_profiles("roi_blobs.yaml")
        for k, v in mods.items():
            prof[k] = v
        config.roi_profile, config.roi_profiles = prof, [prof]
        config.resolutions = [list(res)]
        got = stack_detect.setup_blocks(prof, shape)
        _equal_grids((got.sub_roi_slices, got.sub_rois_offsets),
                     (want.sub_roi_slices, want.sub_rois_offsets))
        for name in ("denoise_max_shape", "exclude_border", "tol", "overlap_base",
                     "overlap", "overlap_padding", "max_pixels"):
            a, b = getattr(got, name), getattr(want, name)
            if b is None:
                assert a is None, name
            else:
                np.testing.assert_array_equal(np.asarray(a), np.asarray(b), err_msg=name)
