"""CPU tests of the host-side mirror of the reference interface: the
reference's own unit tests for this path re-expressed against the drop-in
(magmap/tests/test_chunking.py, magmap/tests/test_detector.py) and the
reference-generated geometry vectors."""
import os

import numpy as np
import pytest

from magellanmapper_b200.cv import chunking, detector, stack_detect
from magellanmapper_b200.settings import config, roi_prof


def _roi_blobs(**mods):
    prof = roi_prof.ROIProfile()
    prof.add_profiles("roi_blobs.yaml")
    for k, v in mods.items():
        prof[k] = v
    config.roi_profile = prof
    config.roi_profiles = [prof]
    return prof


def _split_remerge(roi, max_pixels, overlap):
    sl, off = chunking.stack_splitter(roi.shape, max_pixels, overlap)
    subs = np.empty(sl.shape, dtype=object)
    for c in np.ndindex(*sl.shape):
        subs[c] = roi[sl[c]]
    return chunking.merge_split_stack(subs, max_pixels, overlap)


def test_stack_splitter_roundtrip():
    """magmap/tests/test_chunking.py:47-66"""
    roi = np.arange(5 * 4 * 4).reshape((5, 4, 4))
    max_pixels = [1, 3, 3]
    for ov in ((0, 1, 1), (0, 1, 2), (1, 1, 2)):
        np.testing.assert_array_equal(roi, _split_remerge(roi, max_pixels, np.array(ov)))
    config.resolutions = [[6.6, 1.1, 1.1]]
    np.testing.assert_array_equal(roi, _split_remerge(roi, max_pixels, detector.calc_overlap(2)))


def test_merge_split_stack2_roundtrip():
    roi = np.arange(7 * 9 * 11).reshape((7, 9, 11))
    # the reference only ever calls these with overlap=None (stack_detect.py:124-149);
    # like it, the write cursor advances by the first chunk's full shape
    for mp_, ov in (((3, 4, 5), None), ((2, 9, 4), None), ((7, 9, 11), (2, 2, 2))):
        sl, _ = chunking.stack_splitter(roi.shape, mp_, ov)
        subs = np.empty(sl.shape, dtype=object)
        for c in np.ndindex(*sl.shape):
            subs[c] = roi[sl[c]]
        total = chunking.get_split_stack_total_shape(subs, ov)
        np.testing.assert_array_equal(total, roi.shape)
        out = np.zeros(tuple(total), dtype=roi.dtype)
        chunking.merge_split_stack2(subs, ov, 0, out)
        np.testing.assert_array_equal(out, roi)


def test_chunk_geometry_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "chunk_geometry.npz"))
    for i in range(int(g["n"])):
        ov = g[f"c{i}_overlap"]
        ov = None if ov[0] < 0 else ov
        sl, off = chunking.stack_splitter(g[f"c{i}_shape"], g[f"c{i}_max_pixels"], ov)
        arr = np.zeros(sl.shape + (3, 2), dtype=np.int64)
        for c in np.ndindex(*sl.shape):
            arr[c] = [[s.start, s.stop] for s in sl[c]]
        np.testing.assert_array_equal(arr, g[f"c{i}_slices"])
        np.testing.assert_array_equal(off, g[f"c{i}_offsets"])
        assert off.dtype == g[f"c{i}_offsets"].dtype


def test_setup_blocks_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "setup_blocks.npz"))
    for i in range(int(g["n"])):
        mods = {}
        for k in g[f"b{i}_mods_keys"]:
            v = g[f"b{i}_mod_{k}"]
            mods[str(k)] = float(v) if v.ndim == 0 else tuple(v.tolist())
        if "exclude_border" in mods:
            mods["exclude_border"] = tuple(int(v) for v in mods["exclude_border"])
        prof = _roi_blobs(**mods)
        config.resolutions = [g[f"b{i}_res"].tolist()]
        b = stack_detect.setup_blocks(prof, tuple(g[f"b{i}_shape"]))
        arr = np.zeros(b.sub_roi_slices.shape + (3, 2), dtype=np.int64)
        for c in np.ndindex(*b.sub_roi_slices.shape):
            arr[c] = [[s.start, s.stop] for s in b.sub_roi_slices[c]]
        np.testing.assert_array_equal(arr, g[f"b{i}_slices"])
        np.testing.assert_array_equal(b.sub_rois_offsets, g[f"b{i}_offsets"])
        for key in ("denoise_max_shape", "tol", "overlap_base", "overlap", "overlap_padding",
                    "max_pixels"):
            np.testing.assert_array_equal(getattr(b, key), g[f"b{i}_{key}"], err_msg=key)


def test_blobs_schema():
    """magmap/tests/test_detector.py:13-71"""
    rng = np.random.default_rng(0)
    blobs = rng.random(20).reshape((5, 4))
    blobs[:, :3] = np.multiply(blobs[:, :3], 100).astype(int)
    blobs[:, 3] = blobs[:, 3] * 10
    bl = detector.Blobs(blobs)
    assert bl.cols == [c.value for c in bl.Cols][:4]
    assert bl._col_inds[bl.Cols.RADIUS] == 3
    assert bl._col_inds[bl.Cols.ABS_X] is None
    bl.format_blobs()
    assert bl.cols == [c.value for c in bl.Cols]
    assert bl._col_inds[bl.Cols.RADIUS] == 3
    assert bl._col_inds[bl.Cols.ABS_X] == 9

    np.testing.assert_array_equal(bl.get_blob_confirmed(bl.blobs), bl.blobs[:, 4])
    bl.set_blob_confirmed(bl.blobs, 1)
    assert np.all(bl.get_blob_confirmed(bl.blobs) == 1)
    np.testing.assert_array_equal(bl.get_blob_truth(bl.blobs), bl.blobs[:, 5])
    bl.set_blob_truth(bl.blobs, 2)
    assert np.all(bl.get_blob_truth(bl.blobs) == 2)
    np.testing.assert_array_equal(bl.get_blobs_channel(bl.blobs), bl.blobs[:, 6])
    bl.set_blob_channel(bl.blobs, 3)
    assert np.all(bl.get_blobs_channel(bl.blobs) == 3)
    np.testing.assert_array_equal(bl.get_blob_abs_coords(bl.blobs), bl.blobs[:, 7:10])
    bl.set_blob_abs_coords(bl.blobs, (1, 2, 3))
    assert all(np.all(bl.get_blob_abs_coords(bl.blobs) == (1, 2, 3), axis=1))


def test_blob_layout_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "blob_layout.npz"))
    bl = detector.Blobs(g["b4"].copy())
    full = bl.format_blobs(2)
    np.testing.assert_array_equal(full, g["full"])
    assert list(g["cols"]) == bl.cols
    inter = detector.get_blobs_interior(full, (100, 100, 100), (10, 5, 0), (20, 0, 30))
    np.testing.assert_array_equal(inter, g["interior"])


def test_merge_blobs_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "prune_mp.npz"))
    seg = np.zeros(tuple(g["grid"]), dtype=object)
    for c in np.ndindex(*seg.shape):
        t = g["seg_%d_%d_%d" % c]
        seg[c] = t if len(t) else None
    np.testing.assert_array_equal(chunking.merge_blobs(seg), g["merged"])
    empty = np.zeros((2, 1, 1), dtype=object)
    empty[:] = None
    assert chunking.merge_blobs(empty) is None


def test_archive_roundtrip(tmp_path):
    rng = np.random.default_rng(1)
    bl = detector.Blobs(rng.random((6, 4)))
    bl.format_blobs(0)
    bl.replace_rel_with_abs_blob_coords(bl.blobs)
    bl.remove_abs_blob_coords(True)
    assert bl.cols == ["z", "y", "x", "radius", "confirmed", "truth", "channel", "region"]
    bl.path = str(tmp_path / "x_blobs.npz")
    bl.resolutions = [[1.0, 1.0, 1.0]]
    bl.basename = "x"
    bl.roi_offset, bl.roi_size = (0, 0, 0), (5, 6, 7)
    arc = bl.save_archive()
    assert set(arc) == {"ver", "segments", "resolutions", "basename", "offset", "roi_size",
                        "colocs", "columns"}
    bl.save_archive()                      # second save backs the first one up
    assert os.path.exists(str(tmp_path / "x_blobs(1).npz"))
    back = detector.Blobs().load_blobs(bl.path)
    np.testing.assert_array_equal(back.blobs, bl.blobs)
    assert back.cols == bl.cols and back.ver == 5


def test_profile_layers():
    prof = roi_prof.ROIProfile()
    assert prof["segment_size"] == 500 and prof["denoise_size"] == 25
    prof.add_profiles("lightsheet,4xnuc")
    assert prof["settings_name"] == "default,lightsheet,4xnuc"
    assert prof["max_sigma_factor"] == 4 and prof["exclude_border"] == (1, 0, 0)
    a, b = roi_prof.ROIProfile(), roi_prof.ROIProfile()
    assert roi_prof.ROIProfile.is_identical_settings([a, b], roi_prof.ROIProfile.BLOCK_SIZES)
    b["segment_size"] = 100
    assert not roi_prof.ROIProfile.is_identical_settings([a, b], roi_prof.ROIProfile.BLOCK_SIZES)


def test_errors_match_reference():
    config.resolutions = None
    with pytest.raises(AttributeError):
        detector.calc_scaling_factor()
    from magellanmapper_b200.io import np_io
    with pytest.raises(IOError):
        stack_detect.detect_blobs_stack("x", None)
    with pytest.raises(ValueError):
        stack_detect.detect_blobs_blocks("x", np_io.Image5d(None))


def _np_find_close(blobs, master, tol):
    """numpy stand-in for the GPU box match (same contract as detector._find_close_blobs)."""
    c = np.asarray(blobs)[:, :3].astype(np.int32)
    m = np.asarray(master)[:, :3].astype(np.int32)
    close = np.all(np.abs(m[:, None, :] - c[None, :, :]) <= np.asarray(tol)[None, None, :], axis=2)
    hit = close.any(axis=0)
    last = np.where(close.any(axis=1), close.shape[1] - 1 - np.argmax(close[:, ::-1], axis=1), -1)
    return last.astype(np.int64), hit


def _prune_seam_by_seam(seg_rois, blocks, channels):
    """prune_blobs_mp spelled out with StackPruner.prune_overlap on the wide table,
    seam by seam, the way the reference structures it (stack_detect.py:680-861)."""
    merged = chunking.merge_blobs(seg_rois)
    overlap, tol, pad = blocks.overlap, blocks.tol, blocks.overlap_padding
    slices, offsets = blocks.sub_roi_slices, blocks.sub_rois_offsets
    last = tuple(np.subtract(slices.shape, 1))
    per_channel, ratios_all = [], []
    for chl in channels:
        blobs = detector.Blobs.blobs_in_channel(merged, chl)
        for axis in range(3):
            n_sec = offsets.shape[axis]
            if n_sec <= 1:
                continue
            pos = blobs[:, axis]
            keep_idx, work = [], []
            for j in range(n_sec):
                coord = [0, 0, 0]
                coord[axis] = j
                coord = tuple(coord)
                start = offsets[coord][axis]
                size = slices[coord][axis].stop - slices[coord][axis].start
                end = start + size
                shift = overlap[axis] + pad[axis]
                blobs_ol, n_next = None, None
                if j < n_sec - 1:
                    lo, hi = end - shift, end + pad[axis]
                    blobs_ol = blobs[(pos >= lo) & (pos < hi)]
                    nlo = end + tol[axis]
                    nhi = nlo + overlap[axis] + 2 * pad[axis]
                    total = offsets[last][axis] + size
                    if nlo < total and nhi < total:
                        n_next = int(np.count_nonzero((pos >= nlo) & (pos < nhi)))
                    upper = lo
                else:
                    upper = end
                lower = start + (shift if j > 0 else 0)
                keep_idx.append(np.flatnonzero((pos < upper) & (pos >= lower)))
                work.append((blobs_ol, axis, tol, n_next))
            parts = []
            for j, w in enumerate(work):
                res, ratios = stack_detect.StackPruner.prune_overlap(j, w)
                if res is not None:
                    parts.append(res)
                if ratios:
                    ratios_all.append(ratios)
            blobs = np.concatenate([blobs[np.concatenate(keep_idx)]] + parts)
        per_channel.append(blobs)
    return np.vstack(per_channel)[:, :-3], ratios_all


def test_prune_blobs_mp_index_form_equals_seam_by_seam(monkeypatch):
    """The index-array implementation of prune_blobs_mp against prune_overlap applied
    seam by seam on the wide table: same rows, same order, same averaged positions,
    same ratios - on a 3x3x2 chunk grid crowded with near-duplicates at the seams."""
    monkeypatch.setattr(detector, "_find_close_blobs", _np_find_close)
    _roi_blobs(segment_size=40)
    config.resolutions = [[1.0, 1.0, 1.0]]
    shape = (100, 110, 70)
    blocks = stack_detect.setup_blocks(config.roi_profile, shape)
    rng = np.random.default_rng(11)
    seg = np.empty(blocks.sub_roi_slices.shape, dtype=object)
    for c in np.ndindex(*seg.shape):
        sl = blocks.sub_roi_slices[c]
        size = [s.stop - s.start for s in sl]
        n = int(rng.integers(0, 60))
        if n == 0:
            seg[c] = None
            continue
        zyx = np.column_stack([rng.integers(0, s, n) for s in size]).astype(float)
        t = detector.Blobs(np.column_stack([zyx, np.full(n, 6.0)])).format_blobs(int(n % 2))
        detector.Blobs.shift_blob_rel_coords(t, blocks.sub_rois_offsets[c])
        detector.Blobs.shift_blob_abs_coords(t, blocks.sub_rois_offsets[c])
        seg[c] = t
    for channels in ([0], [0, 1]):
        want, want_ratios = _prune_seam_by_seam(seg, blocks, channels)
        got, df = stack_detect.StackPruner.prune_blobs_mp(
            None, seg, blocks.overlap, blocks.tol, blocks.sub_roi_slices,
            blocks.sub_rois_offsets, channels, blocks.overlap_padding)
        np.testing.assert_array_equal(got, want)
        assert len(df) == len(want_ratios)
        if len(df):
            np.testing.assert_allclose(df.to_numpy(), np.array(want_ratios))
    assert len(want) < sum(len(seg[c]) for c in np.ndindex(*seg.shape) if seg[c] is not None)


def test_calc_near_intensity_bounds_list_semantics(golden_dir):
    """importer.calc_near_intensity_bounds (host part, importer.py:1447-1468): one
    channel APPENDS the extremes to the given lists, several channels REPLACE them by
    arrays - checked against the reference's outputs for the golden per-plane bounds."""
    from magellanmapper_b200.io import importer
    g = np.load(os.path.join(golden_dir, "near_bounds.npz"))
    lows = [list(r) for r in g["u16_plane_lows"]]
    highs = [list(r) for r in g["u16_plane_highs"]]
    mins, maxs = importer.calc_near_intensity_bounds([1.0], [2.0], lows, highs)
    assert mins[0] == 1.0 and maxs[0] == 2.0 and len(mins) == 2
    assert mins[1] == g["u16_near_mins"][0] and maxs[1] == g["u16_near_maxs"][0]
    lows = [list(r) for r in g["u16_2c_plane_lows"]]
    highs = [list(r) for r in g["u16_2c_plane_highs"]]
    mins, maxs = importer.calc_near_intensity_bounds([9.0], [9.0], lows, highs)
    np.testing.assert_array_equal(mins, g["u16_2c_near_mins"])
    np.testing.assert_array_equal(maxs, g["u16_2c_near_maxs"])
    assert importer.calc_near_intensity_bounds([3.0], [4.0], [], []) == ([3.0], [4.0])


def test_csv_and_sqlite_sinks_match_reference(golden_dir, tmp_path):
    """``export_rois.blobs_to_csv`` and the blob tables of ``sqlite`` against files written
    by the unmodified reference: CSV text, schema, stored rows, replace-on-duplicate,
    delete, ROI select-or-insert."""
    import gzip
    from magellanmapper_b200.io import export_rois, sqlite
    g = np.load(os.path.join(golden_dir, "sinks.npz"))
    blobs = g["blobs"]
    out = export_rois.blobs_to_csv(blobs, str(tmp_path / "img.npy"))
    assert out.endswith("img_blobs.csv.gz")
    with gzip.open(out, "rb") as f:
        assert f.read() == g["csv"].tobytes()
    conn, cur = sqlite.create_db(str(tmp_path / "magmap.db"))
    exp_id = sqlite.insert_experiment(conn, cur, "synth")
    roi_id, _ = sqlite.select_or_insert_roi(conn, cur, exp_id, None, (5, 6, 7), (64, 64, 40))
    again, msg = sqlite.select_or_insert_roi(conn, cur, exp_id, 0, (5, 6, 7), (64, 64, 40))
    assert again == roi_id and msg.startswith("Found")
    sqlite.insert_blobs(conn, cur, roi_id, blobs[:, :7])
    assert sqlite.delete_blobs(conn, cur, roi_id, blobs[7:9]) == int(g["n_deleted"])
    cur.execute("SELECT {} FROM blobs ORDER BY id".format(sqlite._COLS_BLOBS))
    rows = np.array([list(r) for r in cur.fetchall()], dtype=np.float64)
    np.testing.assert_array_equal(rows, g["rows"])
    np.testing.assert_array_equal(sqlite.select_blobs_confirmed(cur, 1), g["confirmed1"])
    cur.execute("SELECT name, sql FROM sqlite_master WHERE type = 'table' AND name NOT LIKE "
                "'sqlite_%' ORDER BY name")
    schema = ["{}|{}".format(r[0], " ".join(r[1].split())) for r in cur.fetchall()]
    norm = lambda t: str(t).replace(", ", ",")            # same SQL up to blanks after commas
    assert [norm(t) for t in schema] == [norm(t) for t in g["schema"]]
    cur.execute("SELECT experiment_id, series, offset_x, offset_y, offset_z, size_x, size_y, "
                "size_z FROM rois")
    np.testing.assert_array_equal(np.array([list(r) for r in cur.fetchall()]), g["rois"])
    got, ids = sqlite.select_blobs_by_roi(cur, roi_id)
    assert got.shape == (len(rows), 7) and len(ids) == len(rows)
    conn.close()


def test_setup_images_opens_reference_written_files(golden_dir):
    """The ``.npy`` feed: an image + metadata pair written by the unmodified reference's
    ``np_io.write_npy`` is memory-mapped and its metadata lands in ``config``."""
    from magellanmapper_b200.io import np_io as nio
    from magellanmapper_b200.settings import config
    base = os.path.join(golden_dir, "feed", "sample")
    img5d = nio.setup_images(base)
    assert isinstance(img5d.img, np.memmap) and not img5d.img.flags.writeable
    assert img5d.img.shape == (1, 6, 40, 36, 2) and img5d.img.dtype == np.uint16
    np.testing.assert_allclose(config.resolutions, [[5.0, 1.1, 1.1]])
    assert config.magnification == 20.0 and config.zoom == 1.0
    assert len(config.near_max) == 2 and len(config.near_min) == 2
    assert img5d.meta["ver"] == 15 and img5d.shapes == [[1, 6, 40, 36, 2]]
    sub = nio.setup_images(base + "_image5d.npy", offset=(1, 4, 5), size=(3, 20, 10))
    assert sub.img.shape == (1, 3, 20, 10, 2)
    np.testing.assert_array_equal(sub.img[0], img5d.img[0, 1:4, 4:24, 5:15])
    with pytest.raises(FileNotFoundError):
        nio.setup_images(os.path.join(golden_dir, "feed", "absent"))


def test_sigma_ladder_is_scikit_images_own_form():
    """``blob_log`` builds its ladder as ``linspace(0, 1, n) * (max - min) + min``, which is
    not ``linspace(min, max, n)`` in the last bits unless ``max - min`` is a power of two:
    product, oracle and the radius column (sigma * sqrt(3)) follow scikit-image's form."""
    from oracle import skimage_restated as ski
    from magellanmapper_b200.cv import detector
    from magellanmapper_b200.settings import roi_prof
    for lo, hi, n in ((3, 5, 10), (4, 10, 10), (1, 50, 7), (10, 14, 10)):
        want = np.linspace(0, 1, n) * (np.float64(hi) - np.float64(lo)) + np.float64(lo)
        np.testing.assert_array_equal(ski.sigma_list(lo, hi, n), want)
        prof = roi_prof.ROIProfile()
        prof["min_sigma_factor"], prof["max_sigma_factor"], prof["num_sigma"] = lo, hi, n
        np.testing.assert_array_equal(detector.sigma_ladder(prof, 1.0), want)
    # the case the plain linspace gets wrong (config 4's second channel)
    plain = np.linspace(4, 10, 10)
    assert np.max(np.abs(plain - ski.sigma_list(4, 10, 10))) > 0
    # a float32 image rounds the ends and their difference to float32 first
    f32 = detector.skimage_ladder(np.float32(3.3), np.float32(5.7), 10)
    want32 = np.linspace(0, 1, 10) * (np.float32(5.7) - np.float32(3.3)) + np.float32(3.3)
    np.testing.assert_array_equal(f32, want32.astype(np.float64))
