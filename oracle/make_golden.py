"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference here.

TEST INFRASTRUCTURE ONLY.  Run in the build container (which has
``/root/reference``):  ``python -m oracle.make_golden``.  The GPU box never runs
this; it only reads the committed vectors.

Each vector records inputs and the reference's own outputs for one piece of
the path.  scikit-image functions inside are ``oracle.skimage_restated`` (see
``oracle/ref_shim.py``), so what these vectors pin is the reference's Python:
chunk geometry, block setup, blob table layout, per-chunk orchestration, seam
pruning, saturate/denoise glue.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim                     # noqa: E402
from magellanmapper_b200 import synth           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _slices_to_arr(sl):
    arr = np.zeros(sl.shape + (3, 2), dtype=np.int64)
    for c in np.ndindex(*sl.shape):
        arr[c] = [[s.start, s.stop] for s in sl[c]]
    return arr


def gen_chunk_geometry(ns):
    cases = [
        ((5, 4, 4), (1, 3, 3), (0, 1, 1)),        # magmap/tests/test_chunking.py
        ((5, 4, 4), (1, 3, 3), (0, 1, 2)),
        ((5, 4, 4), (1, 3, 3), (1, 1, 2)),
        ((5, 4, 4), (1, 3, 3), (1, 3, 3)),        # calc_overlap(2) at res 6.6,1.1,1.1 -> ceil(2/res)
        ((512, 2048, 2048), (500, 500, 500), (5, 5, 5)),   # BASELINE config 2
        ((1024, 4096, 4096), (100, 500, 500), (1, 5, 5)),  # config 5
        ((60, 128, 128), (50, 50, 50), None),
        ((50, 500, 500), (25, 25, 25), None),
        ((7, 7, 7), (10, 10, 10), (2, 2, 2)),
    ]
    out = {}
    for i, (shape, mp_, ov) in enumerate(cases):
        sl, off = ns.chunking.stack_splitter(shape, mp_, None if ov is None else np.array(ov))
        out[f"c{i}_shape"] = np.array(shape)
        out[f"c{i}_max_pixels"] = np.array(mp_)
        out[f"c{i}_overlap"] = np.array(ov if ov is not None else (-1, -1, -1))
        out[f"c{i}_slices"] = _slices_to_arr(sl)
        out[f"c{i}_offsets"] = off
    out["n"] = np.array(len(cases))
    # calc_overlap / calc_scaling_factor
    ns.config.resolutions = [[6.6, 1.1, 1.1]]
    out["overlap2_res661111"] = ns.detector.calc_overlap(2)
    out["overlap_default_res661111"] = ns.detector.calc_overlap()
    np.savez_compressed(os.path.join(OUT, "chunk_geometry.npz"), **out)


def gen_setup_blocks(ns):
    out = {}
    cases = [
        ((512, 2048, 2048), (1, 1, 1), {}),
        ((1024, 4096, 4096), (5, 1, 1), {}),
        ((51, 200, 200), (6.6, 1.1, 1.1), {"exclude_border": (1, 0, 0), "segment_size": 150,
                                            "prune_tol_factor": (1, 0.9, 0.9)}),
        ((60, 128, 128), (1, 1, 1), {"segment_size": 50}),
        ((60, 128, 128), (1, 1, 1), {"segment_size": 50, "exclude_border": (4, 3, 0)}),
    ]
    for i, (shape, res, mods) in enumerate(cases):
        prof = ref_shim.set_profile(ns, res, **mods)
        b = ns.stack_detect.setup_blocks(prof, shape)
        out[f"b{i}_shape"] = np.array(shape)
        out[f"b{i}_res"] = np.array(res, dtype=float)
        out[f"b{i}_mods_keys"] = np.array(list(mods.keys()), dtype=str)
        for k, v in mods.items():
            out[f"b{i}_mod_{k}"] = np.array(v, dtype=float)
        out[f"b{i}_slices"] = _slices_to_arr(b.sub_roi_slices)
        out[f"b{i}_offsets"] = b.sub_rois_offsets
        out[f"b{i}_denoise_max_shape"] = np.array(b.denoise_max_shape)
        out[f"b{i}_tol"] = np.array(b.tol)
        out[f"b{i}_overlap_base"] = np.array(b.overlap_base)
        out[f"b{i}_overlap"] = np.array(b.overlap)
        out[f"b{i}_overlap_padding"] = np.array(b.overlap_padding)
        out[f"b{i}_max_pixels"] = np.array(b.max_pixels)
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "setup_blocks.npz"), **out)


def gen_blob_layout(ns):
    rng = np.random.default_rng(11)
    b4 = np.column_stack([rng.integers(0, 100, (6, 3)).astype(float),
                          rng.uniform(3, 9, 6)])
    bl = ns.detector.Blobs(b4.copy())
    full = bl.format_blobs(2)
    cols = np.array(bl.cols, dtype=str)
    interior = ns.detector.get_blobs_interior(full, (100, 100, 100), (10, 5, 0), (20, 0, 30))
    np.savez_compressed(os.path.join(OUT, "blob_layout.npz"), b4=b4, full=full,
                        cols=cols, interior=interior)


def _rand_table(rng, n, lo, hi, chunk_tag):
    """(n, 14) table like merge_blobs output: 11 cols + chunk tag."""
    t = np.full((n, 14), -1.0)
    t[:, 0:3] = rng.integers(lo, hi, (n, 3))
    t[:, 3] = rng.uniform(5, 9, n)
    t[:, 6] = 0
    t[:, 7:10] = t[:, 0:3]
    t[:, 11:14] = chunk_tag
    return t


def gen_remove_close(ns):
    rng = np.random.default_rng(5)
    out = {}
    cases = [(40, 60, (2, 2, 2), 30), (1500, 1200, (1, 3, 3), 60), (5, 0, (1, 1, 1), 10),
             (300, 300, (5, 5, 5), 400)]
    for i, (nm, nc, tol, span) in enumerate(cases):
        master = _rand_table(rng, nm, 0, span, (0, 0, 0))
        check = _rand_table(rng, nc, 0, span, (1, 0, 0))
        pruned, master_out = ns.detector.remove_close_blobs(
            check.copy(), master.copy(), np.array(tol))
        out[f"r{i}_master"], out[f"r{i}_check"], out[f"r{i}_tol"] = master, check, np.array(tol)
        out[f"r{i}_pruned"], out[f"r{i}_master_out"] = pruned, master_out
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "remove_close.npz"), **out)


def gen_prune_mp(ns):
    """prune_blobs_mp on hand-built per-chunk tables with planted duplicates
    across all three seam axes."""
    rng = np.random.default_rng(9)
    shape = (60, 128, 128)
    prof = ref_shim.set_profile(ns, (1, 1, 1), segment_size=50)
    b = ns.stack_detect.setup_blocks(prof, shape)
    seg = np.zeros(b.sub_roi_slices.shape, dtype=object)
    base = rng.integers(0, (60, 128, 128), (700, 3)).astype(float)
    for c in np.ndindex(*seg.shape):
        sl = b.sub_roi_slices[c]
        inside = np.all([(base[:, a] >= sl[a].start) & (base[:, a] < sl[a].stop)
                         for a in range(3)], axis=0)
        pts = base[inside]
        # jitter so that duplicates in overlaps are near- but not always exact matches
        pts = pts + rng.integers(-2, 3, pts.shape)
        pts = np.clip(pts, [s.start for s in sl], [s.stop - 1 for s in sl])
        if len(pts) == 0:
            seg[c] = None
            continue
        t = np.full((len(pts), 11), -1.0)
        t[:, 0:3] = pts
        t[:, 3] = rng.uniform(5, 9, len(pts))
        t[:, 6] = 0
        t[:, 7:10] = pts
        seg[c] = t
    roi = np.zeros(shape, dtype=np.uint8)
    merged = ns.chunking.merge_blobs(seg)
    pruned, _ = ns.stack_detect.StackPruner.prune_blobs_mp(
        roi, seg, b.overlap, b.tol, b.sub_roi_slices, b.sub_rois_offsets, [0],
        b.overlap_padding)
    out = {"shape": np.array(shape), "merged": merged, "pruned": pruned,
           "grid": np.array(seg.shape)}
    for c in np.ndindex(*seg.shape):
        key = "seg_%d_%d_%d" % c
        out[key] = seg[c] if seg[c] is not None else np.zeros((0, 11))
    np.savez_compressed(os.path.join(OUT, "prune_mp.npz"), **out)


def gen_preprocess(ns):
    """saturate_roi + denoise_roi on single blocks (reference glue around
    np.percentile / gaussian / erosion)."""
    vol, _ = synth.make_volume((40, 80, 80), seed=21, density=1 / 1500.0)
    near_max = synth.near_max_of(vol)
    ref_shim.set_profile(ns, (1, 1, 1), near_max=near_max)
    rng = np.random.default_rng(3)
    blocks = {
        "dense": vol[5:30, 10:35, 20:45],
        "thin": vol[0:5, 0:25, 0:25],
        "ragged": vol[28:40, 55:80, 57:80],
        "constant": np.full((25, 25, 25), 417, dtype=np.uint16),
        "zeros": np.zeros((6, 7, 8), dtype=np.uint16),
        "bright": (rng.uniform(20000, 60000, (25, 25, 25))).astype(np.uint16),
        "two_level": np.where(rng.uniform(size=(25, 25, 25)) < 0.03, 5000, 300).astype(np.uint16),
        "single_voxel": vol[3:4, 3:4, 3:4],
    }
    out = {"near_max": np.array(near_max), "names": np.array(list(blocks), dtype=str)}
    for name, blk in blocks.items():
        sat = ns.plot_3d.saturate_roi(blk)
        den = ns.plot_3d.denoise_roi(sat)
        out[f"{name}_in"] = blk
        out[f"{name}_sat"] = sat
        out[f"{name}_out"] = den
    np.savez_compressed(os.path.join(OUT, "preprocess_blocks.npz"), **out)


def gen_detect_small(ns):
    """detector.detect_blobs on a raw uint16 ROI and on a preprocessed one,
    plus an exclude_border call."""
    vol, tab = synth.make_volume((40, 64, 64), seed=31, density=1 / 3000.0)
    near_max = synth.near_max_of(vol)
    ref_shim.set_profile(ns, (1, 1, 1), near_max=near_max)
    raw = ns.detector.detect_blobs(vol, [0])
    pre = ns.plot_3d.denoise_roi(ns.plot_3d.saturate_roi(vol))
    gui = ns.detector.detect_blobs(pre, [0])
    excl = ns.detector.detect_blobs(pre, [0], np.array([[3, 4, 5], [2, 0, 6]]))
    np.savez_compressed(os.path.join(OUT, "detect_small.npz"), vol=vol, table=tab,
                        near_max=np.array(near_max), raw=raw, pre=pre.astype(np.float64),
                        gui=gui, excl=excl)


def gen_stack_small(ns):
    """stack_detect.detect_blobs_blocks end to end (fork pool, seam pruning) on a
    multi-chunk volume: segment_size=50 -> 2x3x3 chunks of <=55 px, 25^3
    preprocessing blocks."""
    shape = (60, 128, 128)
    vol, tab = synth.make_volume(shape, seed=41, density=1 / 2500.0)
    near_max = synth.near_max_of(vol)
    out = {"vol": vol, "table": tab, "near_max": np.array(near_max)}
    for tag, mods in (("plain", {"segment_size": 50}),
                      ("excl", {"segment_size": 50, "exclude_border": (2, 1, 1)})):
        ref_shim.set_profile(ns, (1, 1, 1), near_max=near_max, **mods)
        img5d = ns.np_io.Image5d(vol[None])
        with tempfile.TemporaryDirectory() as td:
            ns.config.filename = os.path.join(td, "synth")
            cwd = os.getcwd()
            os.chdir(td)
            try:
                _, _, blobs = ns.stack_detect.detect_blobs_blocks(
                    ns.config.filename, img5d, None, None, [0], False, False, True)
            finally:
                os.chdir(cwd)
        out[f"{tag}_blobs"] = blobs.blobs
        out[f"{tag}_cols"] = np.array(blobs.cols, dtype=str)
    np.savez_compressed(os.path.join(OUT, "stack_small.npz"), **out)


def gen_iso_unmix(ns):
    """detector.detect_blobs with the ``isotropic`` resize (up- and down-scaling, raw
    uint16 and preprocessed float64 input), with spectral unmixing on a two-channel ROI,
    and detect_blobs_blocks with ``isotropic`` + ``exclude_border`` (the lightsheet
    profile's combination) - all through the UNMODIFIED reference."""
    out = {}
    vol, _ = synth.make_volume((20, 56, 48), seed=61, density=1 / 1500.0)
    near_max = synth.near_max_of(vol)
    out["vol"] = vol
    out["near_max"] = np.array(near_max)
    # up-scaling z: 3 um planes, 1 um pixels, 0.96 of isotropic -> 57 planes
    ref_shim.set_profile(ns, (3, 1, 1), near_max=near_max, isotropic=(0.96, 1, 1))
    out["iso_up_resized"] = ns.cv_nd.make_isotropic(vol, (0.96, 1, 1))
    out["iso_up_raw"] = ns.detector.detect_blobs(vol, [0])
    pre = ns.plot_3d.denoise_roi(ns.plot_3d.saturate_roi(vol))
    out["pre"] = pre.astype(np.float64)
    out["iso_up_pre_resized"] = ns.cv_nd.make_isotropic(pre, (0.96, 1, 1))
    out["iso_up_pre"] = ns.detector.detect_blobs(pre, [0], np.array([[1, 2, 0], [0, 3, 2]]))
    # shrinking y and x (anti-aliasing Gaussian), growing z
    ref_shim.set_profile(ns, (1, 1, 1), near_max=near_max, isotropic=(1.5, 0.6, 0.75))
    out["iso_mixed_resized"] = ns.cv_nd.make_isotropic(vol, (1.5, 0.6, 0.75))
    out["iso_mixed_pre_resized"] = ns.cv_nd.make_isotropic(pre, (1.5, 0.6, 0.75))
    out["iso_mixed_pre"] = ns.detector.detect_blobs(pre, [0])
    # spectral unmixing: channel 0 minus 0.4 x channel 1, channel 1 as it is
    v1, _ = synth.make_volume(vol.shape, seed=62, density=1 / 1500.0)
    two = np.stack([vol, v1], axis=-1)
    out["two"] = two
    p0 = ref_shim.set_profile(ns, (1, 1, 1), near_max=near_max)
    p1 = ns.roi_prof.ROIProfile()
    p1.add_profiles("/root/reference/profiles/roi_blobs.yaml")
    p0.spectral_unmixing = {0: {1: 0.4}}
    ns.config.roi_profiles = [p0, p1]
    ns.config.near_max = [near_max, synth.near_max_of(v1)]
    out["near_max1"] = np.array(ns.config.near_max[1])
    pre2 = ns.plot_3d.denoise_roi(ns.plot_3d.saturate_roi(two))
    out["pre2"] = pre2.astype(np.float64)
    out["unmix_pre"] = ns.detector.detect_blobs(pre2, None)
    # stack: isotropic + exclude_border, anisotropic chunk geometry
    shape = (24, 100, 90)
    svol, _ = synth.make_volume(shape, seed=63, density=1 / 2000.0)
    snm = synth.near_max_of(svol)
    out["svol"] = svol
    out["snm"] = np.array(snm)
    ref_shim.set_profile(ns, (2.5, 1, 1), near_max=snm, isotropic=(0.96, 1, 1),
                         segment_size=40, exclude_border=(1, 0, 0))
    img5d = ns.np_io.Image5d(svol[None])
    with tempfile.TemporaryDirectory() as td:
        ns.config.filename = os.path.join(td, "synth")
        cwd = os.getcwd()
        os.chdir(td)
        try:
            _, _, blobs = ns.stack_detect.detect_blobs_blocks(
                ns.config.filename, img5d, None, None, [0], False, False, True)
        finally:
            os.chdir(cwd)
    out["stack_iso_blobs"] = blobs.blobs
    np.savez_compressed(os.path.join(OUT, "iso_unmix.npz"), **out)


def gen_sinks(ns):
    """The reference's CSV and SQLite sinks fed with a detector table: the CSV text, the
    database schema and the stored blob rows (incl. the replace-on-duplicate rule)."""
    import gzip
    import importlib
    import sqlite3
    export_rois = importlib.import_module("magmap.io.export_rois")
    ref_sqlite = importlib.import_module("magmap.io.sqlite")
    g = np.load(os.path.join(OUT, "detect_small.npz"))
    blobs = g["raw"].copy()
    blobs[::3, 4] = 1            # some confirmed
    blobs[1::4, 5] = 0
    dup = np.vstack([blobs, blobs[:5] * [1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1]])   # same key, new radius
    out = {"blobs": dup}
    with tempfile.TemporaryDirectory() as td:
        export_rois.blobs_to_csv(dup, os.path.join(td, "img.npy"))
        with gzip.open(os.path.join(td, "img_blobs.csv.gz"), "rb") as f:
            out["csv"] = np.frombuffer(f.read(), dtype=np.uint8)
        conn, cur = ref_sqlite._create_db(os.path.join(td, "magmap.db"))
        exp_id = ref_sqlite.insert_experiment(conn, cur, "synth", None)
        roi_id, _ = ref_sqlite.select_or_insert_roi(conn, cur, exp_id, None, (5, 6, 7), (64, 64, 40))
        roi_again, _ = ref_sqlite.select_or_insert_roi(conn, cur, exp_id, 0, (5, 6, 7), (64, 64, 40))
        assert roi_id == roi_again
        ref_sqlite.insert_blobs(conn, cur, roi_id, dup[:, :7])
        n_del = ref_sqlite.delete_blobs(conn, cur, roi_id, dup[7:9])
        cur.execute("SELECT {} FROM blobs ORDER BY id".format(ref_sqlite._COLS_BLOBS))
        out["rows"] = np.array([list(r) for r in cur.fetchall()], dtype=np.float64)
        out["n_deleted"] = np.array(n_del)
        out["confirmed1"] = ref_sqlite.select_blobs_confirmed(cur, 1)
        cur.execute("SELECT name, sql FROM sqlite_master WHERE type = 'table' AND name NOT LIKE "
                    "'sqlite_%' ORDER BY name")
        out["schema"] = np.array(["{}|{}".format(r[0], " ".join(r[1].split()))
                                  for r in cur.fetchall()], dtype=str)
        cur.execute("SELECT experiment_id, series, offset_x, offset_y, offset_z, size_x, size_y, "
                    "size_z FROM rois")
        out["rois"] = np.array([list(r) for r in cur.fetchall()], dtype=np.int64)
        conn.close()
    np.savez_compressed(os.path.join(OUT, "sinks.npz"), **out)


def gen_coloc(ns):
    """colocalizer.colocalize_blobs of the unmodified reference on a two-channel ROI: raw
    uint16 and preprocessed float64 intensities, "min" and percentile thresholds, blobs that
    share a voxel, touch the ROI faces or lie outside the ROI."""
    rng = np.random.default_rng(91)
    shape = (24, 48, 44)
    v0, t0 = synth.make_volume(shape, seed=92, density=1 / 900.0)
    v1, t1 = synth.make_volume(shape, seed=93, density=1 / 900.0)
    # half of channel 1's nuclei sit on channel 0's: real co-localisation
    v1 = np.maximum(v1, (v0 * 0.8).astype(np.uint16) * (rng.random(shape) < 0.5))
    roi = np.stack([v0, v1], axis=-1)
    rows = []
    for chl, tab in ((0, t0), (1, t1)):
        for z, y, x, s, a in tab:
            rows.append([int(z), int(y), int(x), 5.0, -1, -1, chl])
    rows += [rows[0][:6] + [0], rows[3][:3] + [5.0, -1, -1, 1],          # shared voxels
             [0, 0, 0, 5.0, -1, -1, 0], [23, 47, 43, 5.0, -1, -1, 1],    # corners
             [1, 46, 2, 5.0, -1, -1, 0], [30, 10, 10, 5.0, -1, -1, 0],   # outside in z
             [5, -1, 7, 5.0, -1, -1, 1]]
    blobs = np.array(rows, dtype=np.float64)
    blobs = np.hstack([blobs, blobs[:, :3], np.full((len(blobs), 1), -1.0)])
    out = {"roi": roi, "blobs": blobs}
    ns.config.verbose = False
    out["min_raw"] = ns.colocalizer.colocalize_blobs(roi, blobs)
    out["p5_raw"] = ns.colocalizer.colocalize_blobs(roi, blobs, 5)
    ref_shim.set_profile(ns, (1, 1, 1), near_max=synth.near_max_of(v0))
    ns.config.near_max = [synth.near_max_of(v0), synth.near_max_of(v1)]
    pre = ns.plot_3d.denoise_roi(ns.plot_3d.saturate_roi(roi))
    out["pre"] = pre.astype(np.float64)
    out["min_pre"] = ns.colocalizer.colocalize_blobs(pre, blobs)
    out["p30_pre"] = ns.colocalizer.colocalize_blobs(pre, blobs, 30)
    np.savez_compressed(os.path.join(OUT, "coloc.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_shim.load_reference()
    gen_chunk_geometry(ns)
    gen_setup_blocks(ns)
    gen_blob_layout(ns)
    gen_remove_close(ns)
    gen_prune_mp(ns)
    gen_preprocess(ns)
    gen_detect_small(ns)
    gen_stack_small(ns)
    gen_iso_unmix(ns)
    gen_sinks(ns)
    gen_coloc(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


def make_near_bounds():
    """importer.calc_intensity_bounds / calc_near_intensity_bounds of the UNMODIFIED
    reference on small uint16 / uint8 volumes (one and two channels)."""
    ref_shim.load_reference()
    from magmap.io import importer
    rng = np.random.default_rng(21)
    out = {}
    vols = {
        "u16": rng.integers(0, 65535, (6, 41, 37)).astype(np.uint16),
        "u16_narrow": (400 + 30 * rng.standard_normal((5, 64, 50))).clip(0, 65535).astype(np.uint16),
        "u8": rng.integers(0, 255, (4, 30, 33)).astype(np.uint8),
        "u16_2c": rng.integers(0, 4000, (5, 28, 31, 2)).astype(np.uint16),
    }
    for name, vol in vols.items():
        multichannel = vol.ndim == 4
        lows, highs = [], []
        for z in range(vol.shape[0]):
            lo, hi = importer.calc_intensity_bounds(vol[z], dim_channel=2)
            lows.append(lo)
            highs.append(hi)
        near_mins, near_maxs = importer.calc_near_intensity_bounds([], [], lows, highs)
        whole_lo, whole_hi = importer.calc_intensity_bounds(vol[None], dim_channel=4)
        out[f"{name}_vol"] = vol
        out[f"{name}_plane_lows"] = np.array(lows, dtype=np.float64)
        out[f"{name}_plane_highs"] = np.array(highs, dtype=np.float64)
        out[f"{name}_near_mins"] = np.array(near_mins, dtype=np.float64)
        out[f"{name}_near_maxs"] = np.array(near_maxs, dtype=np.float64)
        out[f"{name}_whole"] = np.array([whole_lo, whole_hi], dtype=np.float64)
        out[f"{name}_multichannel"] = np.array(multichannel)
    np.savez_compressed(os.path.join(OUT, "near_bounds.npz"), **out)
    print("near_bounds.npz:", {k: v.shape for k, v in out.items() if not k.endswith("_vol")})


if __name__ == "__main__":
    import sys as _sys
    if len(_sys.argv) > 1 and _sys.argv[1] == "near_bounds":
        make_near_bounds()          # only the vectors added after the first batch
    else:
        main()
        make_near_bounds()
