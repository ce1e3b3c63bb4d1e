"""Randomised cases of the oracle against the UNMODIFIED reference, run live through
``oracle/ref_shim.py``.  Only in the build container: ``/root/reference`` does not travel, so
these tests skip on the GPU box, where the committed vectors of ``tests/golden/`` stand in.
They widen the pinning of ``oracle/magmap_restated.py`` from the fixed golden cases to fresh
geometries, resolutions, profile settings and blob tables on every seed below."""
import os
import tempfile

import numpy as np
import pytest

from magellanmapper_b200 import synth
from oracle import magmap_restated as mm

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/magmap"),
                                reason="the unmodified reference is only in the build container")


@pytest.fixture(scope="module")
def ns():
    from oracle import ref_shim
    ref = ref_shim.load_reference()
    ref.config.verbose = False
    return ref


def _table(rng, n, span, channel=0):
    t = np.full((n, 11), -1.0)
    t[:, :3] = rng.integers(0, span, (n, 3))
    t[:, 3] = rng.uniform(4, 9, n)
    t[:, 6] = channel
    t[:, 7:10] = t[:, :3]
    return t


@pytest.mark.parametrize("seed", range(6))
def test_remove_close_blobs_random(ns, seed):
    rng = np.random.default_rng(100 + seed)
    n_master, n_check = (int(v) for v in rng.integers(0, 1300, 2))
    n_master = max(n_master, 1)
    span = int(rng.integers(8, 200))
    tol = rng.integers(0, 6, 3)
    master, check = _table(rng, n_master, span), _table(rng, n_check, span)
    want = ns.detector.remove_close_blobs(check.copy(), master.copy(), tol)
    got = mm.remove_close_blobs(check.copy(), master.copy(), tol)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("seed", range(4))
def test_prune_blobs_mp_random(ns, seed):
    """Random grids (incl. thin trailing chunks), anisotropic resolutions and tolerance
    factors; duplicates planted across the seams with jitter around the tolerance."""
    from oracle import ref_shim
    rng = np.random.default_rng(200 + seed)
    shape = tuple(int(v) for v in rng.integers(40, 150, 3))
    res = tuple(float(v) for v in rng.choice([0.7, 1.0, 2.0, 5.0], 3))
    mods = {"segment_size": int(rng.integers(25, 70)),
            "prune_tol_factor": tuple(float(v) for v in rng.choice([0.9, 1.0, 1.6], 3))}
    prof = ref_shim.set_profile(ns, res, **mods)
    b = ns.stack_detect.setup_blocks(prof, shape)
    ob = mm.setup_blocks(mm.Profile(**mods), shape, res)
    seg = np.zeros(b.sub_roi_slices.shape, dtype=object)
    base = rng.integers(0, shape, (int(rng.integers(200, 1500)), 3)).astype(float)
    for c in np.ndindex(*seg.shape):
        sl = b.sub_roi_slices[c]
        inside = np.all([(base[:, a] >= sl[a].start) & (base[:, a] < sl[a].stop)
                         for a in range(3)], axis=0)
        pts = base[inside] + rng.integers(-2, 3, (int(inside.sum()), 3))
        pts = np.clip(pts, [s.start for s in sl], [s.stop - 1 for s in sl])
        if len(pts) == 0:
            seg[c] = None
            continue
        t = _table(rng, len(pts), 2)
        t[:, :3] = pts - [s.start for s in sl]          # chunk-relative, as detect_sub_roi
        t[:, 7:10] = pts
        seg[c] = t
    roi = np.zeros(shape, dtype=np.uint8)
    copy = np.zeros(seg.shape, dtype=object)
    for c in np.ndindex(*seg.shape):
        copy[c] = None if seg[c] is None else seg[c].copy()
    np.testing.assert_array_equal(mm.merge_blobs(copy), ns.chunking.merge_blobs(seg))
    want, _ = ns.stack_detect.StackPruner.prune_blobs_mp(
        roi, seg, b.overlap, b.tol, b.sub_roi_slices, b.sub_rois_offsets, [0], b.overlap_padding)
    got = mm.prune_blobs_mp(shape, copy, ob.overlap, ob.tol, ob.sub_roi_slices,
                            ob.sub_rois_offsets, (0,), ob.overlap_padding)
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("seed", range(3))
def test_detect_blobs_random_profiles(ns, seed):
    """``detect_blobs`` with random sigma ladders, overlap, threshold factor and border on
    raw and preprocessed ROIs of random shape and resolution."""
    from oracle import ref_shim
    rng = np.random.default_rng(300 + seed)
    shape = tuple(int(v) for v in (rng.integers(12, 30), rng.integers(40, 90), rng.integers(40, 90)))
    res = tuple(float(v) for v in rng.choice([0.8, 1.0, 1.3, 2.0], 3))
    vol, _ = synth.make_volume(shape, seed=400 + seed, density=1 / 1500.0)
    near_max = synth.near_max_of(vol)
    lo = float(rng.choice([2, 3, 4]))
    mods = {"min_sigma_factor": lo, "max_sigma_factor": lo + float(rng.choice([1, 2, 4])),
            "num_sigma": int(rng.choice([3, 5, 10])), "overlap": float(rng.choice([0.3, 0.55, 0.8])),
            "detection_threshold": float(rng.choice([0.05, 0.1, 0.2]))}
    ref_shim.set_profile(ns, res, near_max=near_max, **mods)
    prof = mm.Profile(**mods)
    np.testing.assert_array_equal(mm.detect_blobs(vol, prof, res), ns.detector.detect_blobs(vol, [0]))
    pre_ref = ns.plot_3d.denoise_roi(ns.plot_3d.saturate_roi(vol))
    pre = mm.denoise_roi(mm.saturate_roi(vol, prof, near_max), prof)
    np.testing.assert_array_equal(pre, pre_ref)
    border = rng.integers(0, 5, (2, 3))
    np.testing.assert_array_equal(mm.detect_blobs(pre, prof, res, 0, border),
                                  ns.detector.detect_blobs(pre_ref, [0], border))


@pytest.mark.parametrize("seed", range(2))
def test_detect_blobs_blocks_random_stack(ns, seed):
    """The whole stack driver (fork pool, merge, seam pruning, final layout) on a random
    multi-chunk volume with ragged trailing chunks and an anisotropic resolution."""
    from oracle import ref_shim
    rng = np.random.default_rng(500 + seed)
    shape = tuple(int(v) for v in (rng.integers(30, 70), rng.integers(70, 140), rng.integers(70, 140)))
    res = (float(rng.choice([1.0, 2.0])), 1.0, 1.0)
    vol, _ = synth.make_volume(shape, seed=600 + seed, density=1 / 2500.0)
    near_max = synth.near_max_of(vol)
    mods = {"segment_size": int(rng.integers(35, 60))}
    if seed % 2:
        mods["exclude_border"] = (1, 2, 0)
    ref_shim.set_profile(ns, res, near_max=near_max, **mods)
    with tempfile.TemporaryDirectory() as td:
        ns.config.filename = os.path.join(td, "synth")
        cwd = os.getcwd()
        os.chdir(td)
        try:
            _, _, blobs = ns.stack_detect.detect_blobs_blocks(
                ns.config.filename, ns.np_io.Image5d(vol[None]), None, None, [0], False, False,
                True)
        finally:
            os.chdir(cwd)
    got = mm.detect_blobs_blocks(vol, mm.Profile(**mods), res, near_max)
    assert blobs.blobs is not None and len(blobs.blobs) > 10
    np.testing.assert_array_equal(got, blobs.blobs)


@pytest.mark.parametrize("seed", range(4))
def test_make_isotropic_and_isotropic_detection_random(ns, seed):
    """``make_isotropic`` (up- and down-scaling axes mixed, integer and float ROIs, one and
    two channels) and ``detect_blobs`` with ``isotropic`` incl. the back-scaling of the blobs."""
    from oracle import ref_shim
    rng = np.random.default_rng(700 + seed)
    shape = tuple(int(v) for v in (rng.integers(8, 24), rng.integers(30, 60), rng.integers(30, 60)))
    res = tuple(float(v) for v in rng.choice([0.6, 1.0, 2.5, 4.0], 3))
    iso = tuple(float(v) for v in rng.choice([0.5, 0.8, 1.0, 1.3], 3))
    vol, _ = synth.make_volume(shape, seed=800 + seed, density=1 / 1200.0)
    near_max = synth.near_max_of(vol)
    ref_shim.set_profile(ns, res, near_max=near_max, isotropic=iso)
    np.testing.assert_array_equal(mm.make_isotropic(vol, iso, res), ns.cv_nd.make_isotropic(vol, iso))
    as_float = vol.astype(np.float64) / 65535
    np.testing.assert_array_equal(mm.make_isotropic(as_float, iso, res),
                                  ns.cv_nd.make_isotropic(as_float, iso))
    two = np.stack([vol, vol[::-1]], axis=-1)
    np.testing.assert_array_equal(mm.make_isotropic(two, iso, res), ns.cv_nd.make_isotropic(two, iso))
    prof = mm.Profile(isotropic=iso)
    np.testing.assert_array_equal(mm.detect_blobs(vol, prof, res), ns.detector.detect_blobs(vol, [0]))


@pytest.mark.parametrize("seed", range(3))
def test_colocalize_blobs_random(ns, seed):
    rng = np.random.default_rng(900 + seed)
    shape = tuple(int(v) for v in (rng.integers(10, 26), rng.integers(30, 50), rng.integers(30, 50)))
    n_chl = int(rng.choice([2, 3]))
    vols = [synth.make_volume(shape, seed=950 + 10 * seed + c, density=1 / 900.0)[0]
            for c in range(n_chl)]
    roi = np.stack(vols, axis=-1)
    n = int(rng.integers(20, 150))
    blobs = np.full((n, 11), -1.0)
    blobs[:, :3] = rng.integers(-2, np.add(shape, 2), (n, 3))     # some outside the ROI
    blobs[:, 3] = 5.0
    blobs[:, 6] = rng.integers(0, n_chl, n)
    blobs[: n // 8, :3] = blobs[n // 8: 2 * (n // 8), :3]         # shared voxels
    for thresh in (None, 5, 60):
        want = ns.colocalizer.colocalize_blobs(roi, blobs, thresh)
        np.testing.assert_array_equal(mm.colocalize_blobs(roi, blobs, thresh), want)


@pytest.mark.parametrize("seed", range(3))
def test_near_bounds_random(ns, seed):
    from magmap.io import importer
    rng = np.random.default_rng(1000 + seed)
    multichannel = bool(seed % 2)
    shape = (int(rng.integers(2, 7)), int(rng.integers(20, 70)), int(rng.integers(20, 70)))
    if multichannel:
        shape += (2,)
    hi = int(rng.choice([255, 4000, 65535]))
    vol = rng.integers(0, hi, shape).astype(np.uint8 if hi == 255 else np.uint16)
    lows, highs = [], []
    for z in range(shape[0]):
        lo, hi_ = importer.calc_intensity_bounds(vol[z], dim_channel=2)
        lows.append(lo)
        highs.append(hi_)
    want_min, want_max = importer.calc_near_intensity_bounds([], [], lows, highs)
    got_min, got_max = mm.calc_near_bounds(vol, multichannel)
    np.testing.assert_array_equal(np.ravel(got_min), np.ravel(want_min))
    np.testing.assert_array_equal(np.ravel(got_max), np.ravel(want_max))
