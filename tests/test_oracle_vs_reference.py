"""Pin the oracle against outputs of the UNMODIFIED reference.

The vectors under tests/golden/ were produced by ``python -m oracle.make_golden``
importing /root/reference in the build container (oracle/ref_shim.py).  These
tests never touch /root/reference, so they also run on the GPU box.
"""
import os

import numpy as np
import pytest

from oracle import magmap_restated as mm
from oracle import skimage_restated as ski


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _slices_arr(sl):
    arr = np.zeros(sl.shape + (3, 2), dtype=np.int64)
    for c in np.ndindex(*sl.shape):
        arr[c] = [[s.start, s.stop] for s in sl[c]]
    return arr


def test_chunk_geometry(golden_dir):
    g = _load(golden_dir, "chunk_geometry.npz")
    for i in range(int(g["n"])):
        ov = g[f"c{i}_overlap"]
        ov = None if ov[0] < 0 else ov
        sl, off = mm.stack_splitter(g[f"c{i}_shape"], g[f"c{i}_max_pixels"], ov)
        np.testing.assert_array_equal(_slices_arr(sl), g[f"c{i}_slices"])
        np.testing.assert_array_equal(off, g[f"c{i}_offsets"])
    np.testing.assert_array_equal(mm.calc_overlap([6.6, 1.1, 1.1], 2), g["overlap2_res661111"])
    np.testing.assert_array_equal(mm.calc_overlap([6.6, 1.1, 1.1]), g["overlap_default_res661111"])


def test_setup_blocks(golden_dir):
    g = _load(golden_dir, "setup_blocks.npz")
    for i in range(int(g["n"])):
        mods = {}
        for k in g[f"b{i}_mods_keys"]:
            v = g[f"b{i}_mod_{k}"]
            mods[str(k)] = float(v) if v.ndim == 0 else tuple(v.tolist())
        if "exclude_border" in mods:
            mods["exclude_border"] = tuple(int(v) for v in mods["exclude_border"])
        prof = mm.Profile(**mods)
        b = mm.setup_blocks(prof, g[f"b{i}_shape"], g[f"b{i}_res"])
        np.testing.assert_array_equal(_slices_arr(b.sub_roi_slices), g[f"b{i}_slices"])
        np.testing.assert_array_equal(b.sub_rois_offsets, g[f"b{i}_offsets"])
        for key in ("denoise_max_shape", "tol", "overlap_base", "overlap",
                    "overlap_padding", "max_pixels"):
            np.testing.assert_array_equal(getattr(b, key), g[f"b{i}_{key}"], err_msg=key)


def test_blob_layout(golden_dir):
    g = _load(golden_dir, "blob_layout.npz")
    full = mm.format_blobs(g["b4"], 2)
    np.testing.assert_array_equal(full, g["full"])
    assert tuple(g["cols"]) == mm.COLS
    inter = mm.get_blobs_interior(full, (100, 100, 100), (10, 5, 0), (20, 0, 30))
    np.testing.assert_array_equal(inter, g["interior"])


def test_remove_close_blobs(golden_dir):
    g = _load(golden_dir, "remove_close.npz")
    for i in range(int(g["n"])):
        pruned, master = mm.remove_close_blobs(
            g[f"r{i}_check"].copy(), g[f"r{i}_master"].copy(), g[f"r{i}_tol"])
        np.testing.assert_array_equal(pruned, g[f"r{i}_pruned"])
        np.testing.assert_array_equal(master, g[f"r{i}_master_out"])


def test_prune_blobs_mp(golden_dir):
    g = _load(golden_dir, "prune_mp.npz")
    shape = tuple(g["shape"])
    b = mm.setup_blocks(mm.Profile(segment_size=50), shape, (1, 1, 1))
    seg = np.zeros(tuple(g["grid"]), dtype=object)
    for c in np.ndindex(*seg.shape):
        t = g["seg_%d_%d_%d" % c]
        seg[c] = t if len(t) else None
    np.testing.assert_array_equal(mm.merge_blobs(seg), g["merged"])
    out = mm.prune_blobs_mp(shape, seg, b.overlap, b.tol, b.sub_roi_slices,
                            b.sub_rois_offsets, (0,), b.overlap_padding)
    np.testing.assert_array_equal(out, g["pruned"])


def test_preprocess_blocks(golden_dir):
    g = _load(golden_dir, "preprocess_blocks.npz")
    prof = mm.Profile()
    nm = float(g["near_max"])
    for name in g["names"]:
        blk = g[f"{name}_in"]
        sat = mm.saturate_roi(blk, prof, nm)
        np.testing.assert_array_equal(sat, g[f"{name}_sat"], err_msg=str(name))
        assert sat.dtype == g[f"{name}_sat"].dtype
        den = mm.denoise_roi(sat, prof)
        np.testing.assert_array_equal(den, g[f"{name}_out"], err_msg=str(name))


def test_detect_blobs_small(golden_dir):
    g = _load(golden_dir, "detect_small.npz")
    prof = mm.Profile()
    raw = mm.detect_blobs(g["vol"], prof, (1, 1, 1))
    np.testing.assert_array_equal(raw, g["raw"])
    pre = mm.denoise_roi(mm.saturate_roi(g["vol"], prof, float(g["near_max"])), prof)
    np.testing.assert_array_equal(pre, g["pre"])
    gui = mm.detect_blobs(pre, prof, (1, 1, 1))
    np.testing.assert_array_equal(gui, g["gui"])
    excl = mm.detect_blobs(pre, prof, (1, 1, 1), 0, np.array([[3, 4, 5], [2, 0, 6]]))
    np.testing.assert_array_equal(excl, g["excl"])
    assert len(raw) > 5 and len(gui) > 5


@pytest.mark.parametrize("tag,mods", [("plain", {}), ("excl", {"exclude_border": (2, 1, 1)})])
def test_detect_blobs_blocks_small(golden_dir, tag, mods):
    g = _load(golden_dir, "stack_small.npz")
    prof = mm.Profile(segment_size=50, **mods)
    out = mm.detect_blobs_blocks(g["vol"], prof, (1, 1, 1), float(g["near_max"]))
    np.testing.assert_array_equal(out, g[f"{tag}_blobs"])
    assert tuple(g[f"{tag}_cols"]) == ("z", "y", "x", "radius", "confirmed", "truth",
                                         "channel", "region")


def test_prune_order_dependence_is_characterised(golden_dir):
    """Any difference between scikit-image's set-order pruning and the
    order-independent rule is confined to the order-dependent set."""
    g = _load(golden_dir, "detect_small.npz")
    res = ski.blob_log(g["pre"], 3, 5, 10, 0.1, 0.5, full=True)
    tr = res.trace
    diff = np.nonzero(tr.keep_reference_order != tr.keep_canonical)[0]
    assert set(diff.tolist()) <= set(tr.order_dependent.tolist())
    # shuffling the pair order never changes blobs outside that set
    rng = np.random.default_rng(0)
    lm = np.hstack([res.peaks[:, :3].astype(float), res.sigmas[res.peaks[:, 3]][:, None]])
    base = tr.keep_reference_order
    for _ in range(5):
        perm = rng.permutation(len(tr.pairs))
        _, tr2 = ski.prune_blobs(lm, 0.5, trace=True, pair_order=tr.pairs[perm])
        d = np.nonzero(tr2.keep_reference_order != base)[0]
        assert set(d.tolist()) <= set(tr.order_dependent.tolist())


def test_near_bounds_oracle_vs_reference(golden_dir):
    """oracle.calc_intensity_bounds / calc_near_bounds against the unmodified
    reference's importer functions (tests/golden/near_bounds.npz)."""
    from oracle import magmap_restated as mm
    g = np.load(os.path.join(golden_dir, "near_bounds.npz"))
    for name in ("u16", "u16_narrow", "u8", "u16_2c"):
        vol = g[f"{name}_vol"]
        multichannel = bool(g[f"{name}_multichannel"])
        near_mins, near_maxs = mm.calc_near_bounds(vol, multichannel)
        np.testing.assert_array_equal(np.ravel(near_mins), np.ravel(g[f"{name}_near_mins"]))
        np.testing.assert_array_equal(np.ravel(near_maxs), np.ravel(g[f"{name}_near_maxs"]))
        lo, hi = mm.calc_intensity_bounds(vol, channel_axis=3 if multichannel else None)
        np.testing.assert_array_equal(np.array([lo, hi]), g[f"{name}_whole"])


def test_isotropic_and_unmixing(golden_dir):
    """``make_isotropic`` (up-scaling, and shrinking with the anti-aliasing Gaussian),
    ``detect_blobs`` with ``isotropic`` and with spectral unmixing, and the stack driver
    with ``isotropic`` + ``exclude_border``, against the unmodified reference."""
    g = _load(golden_dir, "iso_unmix.npz")
    vol, pre, nm = g["vol"], g["pre"], float(g["near_max"])
    np.testing.assert_array_equal(mm.make_isotropic(vol, (0.96, 1, 1), (3, 1, 1)),
                                  g["iso_up_resized"])
    np.testing.assert_array_equal(mm.make_isotropic(pre, (0.96, 1, 1), (3, 1, 1)),
                                  g["iso_up_pre_resized"])
    np.testing.assert_array_equal(mm.make_isotropic(vol, (1.5, 0.6, 0.75), (1, 1, 1)),
                                  g["iso_mixed_resized"])
    up = mm.Profile(isotropic=(0.96, 1, 1))
    np.testing.assert_array_equal(mm.detect_blobs(vol, up, (3, 1, 1)), g["iso_up_raw"])
    np.testing.assert_array_equal(
        mm.detect_blobs(pre, up, (3, 1, 1), 0, np.array([[1, 2, 0], [0, 3, 2]])), g["iso_up_pre"])
    mixed = mm.Profile(isotropic=(1.5, 0.6, 0.75))
    np.testing.assert_array_equal(mm.detect_blobs(pre, mixed, (1, 1, 1)), g["iso_mixed_pre"])
    p0 = mm.Profile(spectral_unmixing={0: {1: 0.4}})
    np.testing.assert_array_equal(
        mm.detect_blobs(g["pre2"], [p0, mm.Profile()], (1, 1, 1), [0, 1]), g["unmix_pre"])
    prof = mm.Profile(isotropic=(0.96, 1, 1), segment_size=40, exclude_border=(1, 0, 0))
    final = mm.detect_blobs_blocks(g["svol"], prof, (2.5, 1, 1), float(g["snm"]))
    np.testing.assert_array_equal(final, g["stack_iso_blobs"])


def test_colocalize_blobs(golden_dir):
    """``colocalize_blobs`` (ball(2) label dilation, "min" and percentile thresholds, blobs
    sharing a voxel, on the ROI faces and outside the ROI) against the unmodified reference,
    on raw uint16 and on preprocessed float64 intensities."""
    g = _load(golden_dir, "coloc.npz")
    for key, roi, thresh in (("min_raw", g["roi"], None), ("p5_raw", g["roi"], 5),
                             ("min_pre", g["pre"], None), ("p30_pre", g["pre"], 30)):
        got = mm.colocalize_blobs(roi, g["blobs"], thresh)
        assert got.dtype == np.uint8
        np.testing.assert_array_equal(got, g[key], err_msg=key)
    assert mm.colocalize_blobs(g["roi"][..., 0], g["blobs"]) is None
    assert mm.colocalize_blobs(g["roi"], None) is None
