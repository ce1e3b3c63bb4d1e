"""The host-only helpers of the mirror (`Blobs` accessors and coordinate shifts, overlap and
scaling factors, ROI selections, pruning ratios, sorting) against the UNMODIFIED reference on
random tables - live, build container only (``/root/reference`` does not travel)."""
import os

import numpy as np
import pytest

from magellanmapper_b200.cv import detector
from magellanmapper_b200.settings import config

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/magmap"),
                                reason="the unmodified reference is only in the build container")


@pytest.fixture(scope="module")
def ns():
    from oracle import ref_shim
    ref = ref_shim.load_reference()
    ref.config.verbose = False
    return ref


def _tables(rng, n):
    t = np.full((n, 4), 0.0)
    t[:, :3] = rng.integers(0, 300, (n, 3))
    t[:, 3] = rng.uniform(3, 9, n)
    return t


@pytest.mark.parametrize("seed", range(4))
def test_blobs_accessors_and_shifts(ns, seed):
    rng = np.random.default_rng(seed)
    raw = _tables(rng, int(rng.integers(1, 200)))
    ours, theirs = detector.Blobs(raw.copy()), ns.detector.Blobs(raw.copy())
    assert ours.cols == theirs.cols
    np.testing.assert_array_equal(ours.format_blobs(2), theirs.format_blobs(2))
    assert ours.cols == theirs.cols
    a, b = ours.blobs, theirs.blobs
    off = rng.integers(-20, 20, 3)
    fac = rng.uniform(0.3, 3, 3)
    for name, arg in (("shift_blob_rel_coords", off), ("shift_blob_abs_coords", off),
                      ("multiply_blob_rel_coords", fac), ("multiply_blob_abs_coords", fac)):
        np.testing.assert_array_equal(getattr(detector.Blobs, name)(a, arg),
                                      getattr(ns.detector.Blobs, name)(b, arg), err_msg=name)
    for name in ("get_blob_confirmed", "get_blob_truth", "get_blobs_channel",
                 "get_blob_abs_coords"):
        np.testing.assert_array_equal(getattr(detector.Blobs, name)(a),
                                      getattr(ns.detector.Blobs, name)(b), err_msg=name)
    detector.Blobs.set_blob_truth(a, 1)
    ns.detector.Blobs.set_blob_truth(b, 1)
    detector.Blobs.set_blob_abs_coords(a, (4, 5, 6))
    ns.detector.Blobs.set_blob_abs_coords(b, (4, 5, 6))
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(detector.Blobs.blob_for_db(a[0]), ns.detector.Blobs.blob_for_db(b[0]))
    for chl in (0, 2, [1, 2], None):
        np.testing.assert_array_equal(detector.Blobs.blobs_in_channel(a, chl),
                                      ns.detector.Blobs.blobs_in_channel(b, chl))
    np.testing.assert_array_equal(detector.Blobs.replace_rel_with_abs_blob_coords(a.copy()),
                                  ns.detector.Blobs.replace_rel_with_abs_blob_coords(b.copy()))
    np.testing.assert_array_equal(ours.remove_abs_blob_coords(True),
                                  theirs.remove_abs_blob_coords(True))
    assert ours.cols == theirs.cols


@pytest.mark.parametrize("seed", range(4))
def test_roi_selection_sorting_ratios_and_factors(ns, seed):
    rng = np.random.default_rng(50 + seed)
    blobs = np.full((int(rng.integers(5, 400)), 11), -1.0)
    blobs[:, :3] = rng.integers(0, 120, (len(blobs), 3))
    blobs[:, 3] = rng.uniform(3, 9, len(blobs))
    offset, size = rng.integers(0, 60, 3), rng.integers(1, 70, 3)
    margin = rng.integers(0, 4, 3)
    for kw in ({}, {"margin": margin}, {"reverse": False}, {"margin": margin, "reverse": False}):
        got = detector.get_blobs_in_roi(blobs, offset, size, **kw)
        want = ns.detector.get_blobs_in_roi(blobs, offset, size, **kw)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
    pad_a, pad_b = rng.integers(0, 10, 3), rng.integers(0, 10, 3)
    np.testing.assert_array_equal(detector.get_blobs_interior(blobs, (120, 120, 120), pad_a, pad_b),
                                  ns.detector.get_blobs_interior(blobs, (120, 120, 120), pad_a, pad_b))
    got, want = detector.sort_blobs(blobs), ns.detector.sort_blobs(blobs)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])
    for args in ((0, 0, 0), (10, 7, 12), (10, 10, 0), (5, 0, 9), (123, 100, 140)):
        assert detector.meas_pruning_ratio(*args) == ns.detector.meas_pruning_ratio(*args)
    res = [[float(v) for v in rng.choice([0.5, 0.913, 1.0, 2.5, 6.6], 3)]]
    config.resolutions = res
    ns.config.resolutions = res
    np.testing.assert_array_equal(detector.calc_scaling_factor(), ns.detector.calc_scaling_factor())
    for factor in (None, 2, 7):
        np.testing.assert_array_equal(detector.calc_overlap(factor), ns.detector.calc_overlap(factor))


def _plain(v):
    """Profile values as comparable plain objects (enums by name, dicts recursively)."""
    if isinstance(v, dict):
        return {str(getattr(k, "name", k)): _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple, np.ndarray)):
        return [_plain(x) for x in v]
    if hasattr(v, "name") and hasattr(v, "value") and not isinstance(v, (int, float)):
        return v.name
    return v


def test_every_roi_profile_modifier_layers_like_the_reference(ns):
    """``ROIProfile`` defaults, every named modifier on its own and in the combinations the
    reference's documentation uses, plus the YAML profile: equal on every key the mirror
    carries (the keys this path reads)."""
    from magellanmapper_b200.settings import roi_prof
    ref_default = ns.roi_prof.ROIProfile()
    names = sorted(ref_default.profiles)
    assert names, "the reference lists no modifiers"
    ours_default = roi_prof.ROIProfile()
    assert sorted(ours_default.profiles) == names
    combos = [n for n in names] + ["lightsheet,4xnuc", "lightsheet,cleared,zebrafish",
                                   "2p20x,lowres", "roi_blobs.yaml", "lightsheet,roi_blobs.yaml"]
    cwd = os.getcwd()
    os.chdir("/root/reference")          # the reference resolves YAML files under ./profiles
    try:
        for combo in [""] + combos:
            theirs = ns.roi_prof.ROIProfile()
            ours = roi_prof.ROIProfile()
            if combo:
                theirs.add_profiles(combo)
                ours.add_profiles(combo)
            assert ours[ours.NAME_KEY] == theirs[theirs.NAME_KEY], combo
            for key in ours:
                assert key in theirs, (combo, key)
                assert _plain(ours[key]) == _plain(theirs[key]), (combo, key)
    finally:
        os.chdir(cwd)


@pytest.mark.parametrize("seed", range(3))
def test_csv_and_sqlite_sinks_random(ns, seed, tmp_path):
    """Random blob tables (duplicates of the (roi, z, y, x) key, confirmed / truth flags,
    several channels, several ROIs) through both implementations of the CSV and SQLite
    sinks: the same CSV bytes and the same stored rows after inserts, replaces and deletes."""
    import gzip
    import importlib
    from magellanmapper_b200.io import export_rois, sqlite
    ref_export = importlib.import_module("magmap.io.export_rois")
    ref_sqlite = importlib.import_module("magmap.io.sqlite")
    rng = np.random.default_rng(70 + seed)
    n = int(rng.integers(5, 400))
    blobs = np.full((n, 11), -1.0)
    blobs[:, :3] = rng.integers(0, 40, (n, 3))
    blobs[:, 3] = np.round(rng.uniform(3, 9, n), 3)
    blobs[:, 4] = rng.choice([-1, 0, 1], n)
    blobs[:, 5] = rng.choice([-1, 0, 1, 2], n)
    blobs[:, 6] = rng.integers(0, 3, n)
    blobs[:, 7:10] = blobs[:, :3] + 100
    dirs = {}
    for tag in ("ours", "theirs"):
        dirs[tag] = tmp_path / tag
        dirs[tag].mkdir()
    export_rois.blobs_to_csv(blobs, str(dirs["ours"] / "img.npy"))
    ref_export.blobs_to_csv(blobs, str(dirs["theirs"] / "img.npy"))
    texts = []
    for tag in ("ours", "theirs"):
        with gzip.open(dirs[tag] / "img_blobs.csv.gz", "rb") as f:
            texts.append(f.read())
    assert texts[0] == texts[1]

    stored = []
    for tag, mod, create in (("ours", sqlite, sqlite.create_db), ("theirs", ref_sqlite, ref_sqlite._create_db)):
        conn, cur = create(str(dirs[tag] / "magmap.db"))
        exp = mod.insert_experiment(conn, cur, "synth", None)
        rois = [mod.select_or_insert_roi(conn, cur, exp, s, off, (30, 30, 30))[0]
                for s, off in ((None, (1, 2, 3)), (0, (1, 2, 3)), (1, (4, 5, 6)))]
        assert rois[0] == rois[1] != rois[2]
        mod.insert_blobs(conn, cur, rois[0], blobs[:, :7])
        mod.insert_blobs(conn, cur, rois[2], blobs[: n // 2, :7])
        again = blobs[: n // 3, :7].copy()
        again[:, 3] *= 2                                        # same key: replaced
        mod.insert_blobs(conn, cur, rois[0], again)
        n_del = mod.delete_blobs(conn, cur, rois[0], blobs[n // 4: n // 4 + 3])
        cur.execute("SELECT {} FROM blobs ORDER BY roi_id, z, y, x, channel, radius".format(
            mod._COLS_BLOBS))
        rows = np.array([list(r) for r in cur.fetchall()], dtype=np.float64)
        conf = mod.select_blobs_confirmed(cur, 1)
        # the reference's `ClrDB.select_blobs_by_roi` (sqlite.py:823-836) spelled out
        cur.execute("SELECT {}, id FROM blobs WHERE roi_id = ?".format(mod._COLS_BLOBS), (rois[2],))
        got, ids = mod._parse_blobs(cur.fetchall())
        if mod is sqlite:
            same, same_ids = sqlite.select_blobs_by_roi(cur, rois[2])
            np.testing.assert_array_equal(same, got)
            assert same_ids == ids
        stored.append((rows, n_del, np.asarray(conf), np.asarray(got), len(ids)))
        conn.close()
    for a, b in zip(stored[0], stored[1]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("seed", range(3))
def test_write_npy_random_without_near_bounds(ns, seed, tmp_path):
    """``np_io.write_npy`` keyed with ``config.MetaKeys`` like a reference caller keys it
    (``find_near_bounds=False``: no GPU needed): the same ``.npy`` bytes and metadata file as
    the reference's, and ``setup_images`` / ``load_metadata`` read back what it reads."""
    import yaml
    from magellanmapper_b200.io import np_io
    rng = np.random.default_rng(90 + seed)
    shape = (1, int(rng.integers(2, 9)), int(rng.integers(5, 40)), int(rng.integers(5, 40)))
    if seed % 2:
        shape += (int(rng.integers(2, 4)),)
    dtype = [np.uint8, np.uint16, np.float32][seed % 3]
    img = (rng.random(shape) * 200).astype(dtype)
    res = [[float(v) for v in rng.choice([0.5, 1.1, 5.0], 3)]]
    mag, zoom = float(rng.choice([5.0, 20.0])), float(rng.choice([0.8, 1.0]))
    files = {}
    for tag, mod, keys in (("ours", np_io, config.MetaKeys), ("theirs", ns.np_io, ns.config.MetaKeys)):
        d = tmp_path / tag
        d.mkdir()
        md = {keys.RESOLUTIONS: res, keys.MAGNIFICATION: mag, keys.ZOOM: zoom}
        mod.write_npy(img, md, str(d / "sample.czi"), False)
        with open(d / "sample_image5d.npy", "rb") as f:
            raw = f.read()
        with open(d / "sample_meta.yml") as f:
            meta = yaml.safe_load(f)
        files[tag] = (raw, meta)
    assert files["ours"][0] == files["theirs"][0]
    assert files["ours"][1] == files["theirs"][1]
    # the mirror reads the reference's pair
    config.filename = None
    img5d = np_io.setup_images(str(tmp_path / "theirs" / "sample"))
    np.testing.assert_array_equal(np.asarray(img5d.img), img)
    assert [list(r) for r in config.resolutions] == res
    assert config.magnification == mag and config.zoom == zoom


MIRRORED_MODULES = [
    ("cv.detector", "magmap.cv.detector"), ("cv.stack_detect", "magmap.cv.stack_detect"),
    ("cv.chunking", "magmap.cv.chunking"), ("plot.plot_3d", "magmap.plot.plot_3d"),
    ("cv.cv_nd", "magmap.cv.cv_nd"), ("cv.colocalizer", "magmap.cv.colocalizer"),
    ("io.np_io", "magmap.io.np_io"), ("io.importer", "magmap.io.importer"),
    ("io.sqlite", "magmap.io.sqlite"), ("io.export_rois", "magmap.io.export_rois"),
    ("io.libmag", "magmap.io.libmag"), ("settings.config", "magmap.settings.config"),
    ("settings.roi_prof", "magmap.settings.roi_prof"),
    ("settings.profiles", "magmap.settings.profiles"),
]


def test_every_mirrored_callable_has_the_reference_signature(ns):
    """Every function, class and method the mirror defines under a name the reference module
    also has takes the reference's parameters: same names in the same order, defaults where
    the reference has defaults, the same kind of method.  (Names only the mirror has are its
    own additions and are not checked.)"""
    import importlib
    import inspect

    def params(fn):
        try:
            sig = inspect.signature(fn)
        except (TypeError, ValueError):
            return None
        return [(p.name, p.kind, p.default is not inspect.Parameter.empty)
                for p in sig.parameters.values()]

    def compare(label, ours, theirs, problems):
        a, b = params(ours), params(theirs)
        if a is None or b is None:
            return
        var = (inspect.Parameter.VAR_POSITIONAL, inspect.Parameter.VAR_KEYWORD)
        names_a = [p[0] for p in a if p[1] not in var]
        names_b = [p[0] for p in b if p[1] not in var]
        if names_a[:len(names_b)] != names_b:
            problems.append(f"{label}: {names_a} vs reference {names_b}")
            return
        for pa, pb in zip([p for p in a if p[1] not in var], [p for p in b if p[1] not in var]):
            if pa[2] != pb[2]:
                problems.append(f"{label}: default of `{pa[0]}`")
        for extra in [p for p in a if p[1] not in var][len(names_b):]:
            if not extra[2]:
                problems.append(f"{label}: extra parameter `{extra[0]}` without a default")
        if {p[1] for p in b if p[1] in var} - {p[1] for p in a if p[1] in var}:
            problems.append(f"{label}: reference takes *args / **kwargs")

    problems, checked = [], 0
    for ours_name, ref_name in MIRRORED_MODULES:
        ours_mod = importlib.import_module("magellanmapper_b200." + ours_name)
        ref_mod = importlib.import_module(ref_name)
        for name, obj in vars(ours_mod).items():
            if name.startswith("__") or getattr(obj, "__module__", None) != ours_mod.__name__:
                continue
            if not hasattr(ref_mod, name):
                continue
            ref_obj = getattr(ref_mod, name)
            if inspect.isclass(obj):
                for mname, member in vars(obj).items():
                    if mname.startswith("__") and mname != "__init__":
                        continue
                    static = inspect.getattr_static(ref_obj, mname, None)
                    if static is None or isinstance(member, property):
                        continue
                    fn = member.__func__ if isinstance(member, (classmethod, staticmethod)) else member
                    ref_fn = static.__func__ if isinstance(static, (classmethod, staticmethod)) else static
                    if not callable(fn) or not callable(ref_fn):
                        continue
                    if type(member).__name__ != type(static).__name__:
                        problems.append(f"{ours_name}.{name}.{mname}: {type(member).__name__} vs "
                                        f"reference {type(static).__name__}")
                    compare(f"{ours_name}.{name}.{mname}", fn, ref_fn, problems)
                    checked += 1
            elif callable(obj):
                compare(f"{ours_name}.{name}", obj, ref_obj, problems)
                checked += 1
    assert checked > 80
    assert not problems, "\n".join(problems)


def test_config_defaults_and_enums_equal_the_reference(ns):
    """Every global the mirror's ``settings.config`` shares with the reference's starts with
    the reference's value; the enums of the path list the same members."""
    import enum
    import inspect
    from magellanmapper_b200.cv import stack_detect as sd
    from magellanmapper_b200.settings import config as ours
    theirs = ns.config
    own = {"annotations", "gpu_device", "logger"}
    # values the tests of this session may have assigned: compared on a fresh import instead
    import importlib
    fresh = importlib.util.module_from_spec(importlib.util.find_spec(ours.__name__))
    fresh.__spec__.loader.exec_module(fresh)
    ref_fresh = importlib.util.module_from_spec(importlib.util.find_spec(theirs.__name__))
    ref_fresh.__spec__.loader.exec_module(ref_fresh)
    for name, val in vars(fresh).items():
        if name.startswith("_") or name in own or inspect.ismodule(val):
            continue
        if callable(val) and not inspect.isclass(val):
            continue
        assert hasattr(ref_fresh, name), f"config.{name} is not a reference global"
        ref_val = getattr(ref_fresh, name)
        if inspect.isclass(val):
            if issubclass(val, enum.Enum):
                assert [m.name for m in val] == [m.name for m in ref_val], name
            continue
        if isinstance(val, dict):
            assert {str(k): v for k, v in val.items()} == {str(k): v for k, v in ref_val.items()}, name
        else:
            assert val == ref_val, f"config.{name}: {val!r} vs reference {ref_val!r}"
    for a, b in ((detector.Blobs.Keys, ns.detector.Blobs.Keys), (detector.Blobs.Cols, ns.detector.Blobs.Cols),
                 (sd.StackTimes, ns.stack_detect.StackTimes)):
        assert [(m.name, m.value) for m in a] == [(m.name, m.value) for m in b]
    assert detector.Blobs.BLOBS_NP_VER == ns.detector.Blobs.BLOBS_NP_VER


def test_error_behaviour_equals_the_reference(ns, tmp_path):
    """Bad calls into the host-only part of the path end the same way in both: the same
    exception type, or both succeed with the same kind of result."""
    import importlib
    from magellanmapper_b200.cv import chunking, stack_detect as sd
    from magellanmapper_b200.io import importer, libmag, np_io
    from magellanmapper_b200.settings import roi_prof
    ref_importer = importlib.import_module("magmap.io.importer")
    ref_libmag = importlib.import_module("magmap.io.libmag")
    tmp = str(tmp_path)

    def outcome(fn):
        try:
            return "ok", type(fn()).__name__
        except Exception as e:                                   # noqa: BLE001
            return (type(e).__name__,)

    empty_grid = np.zeros((1, 1, 1), dtype=object)
    cases = [
        (lambda: detector.Blobs().load_blobs(tmp + "/nope.npz"),
         lambda: ns.detector.Blobs().load_blobs(tmp + "/nope.npz")),
        (lambda: chunking.stack_splitter((4, 4), (2, 2)), lambda: ns.chunking.stack_splitter((4, 4), (2, 2))),
        (lambda: chunking.merge_blobs(empty_grid), lambda: ns.chunking.merge_blobs(empty_grid)),
        (lambda: detector.Blobs(None).format_blobs(), lambda: ns.detector.Blobs(None).format_blobs()),
        (lambda: detector.get_blobs_in_roi(None, (0, 0, 0), (1, 1, 1)),
         lambda: ns.detector.get_blobs_in_roi(None, (0, 0, 0), (1, 1, 1))),
        (lambda: detector.remove_close_blobs(None, np.zeros((2, 4)), (1, 1, 1)),
         lambda: ns.detector.remove_close_blobs(None, np.zeros((2, 4)), (1, 1, 1))),
        (lambda: detector.remove_close_blobs(np.zeros((2, 4)), None, (1, 1, 1)),
         lambda: ns.detector.remove_close_blobs(np.zeros((2, 4)), None, (1, 1, 1))),
        (lambda: detector.sort_blobs(np.zeros((0, 4))), lambda: ns.detector.sort_blobs(np.zeros((0, 4)))),
        (lambda: detector.meas_pruning_ratio(None, 1, 1), lambda: ns.detector.meas_pruning_ratio(None, 1, 1)),
        (lambda: importer.load_metadata(tmp + "/no_meta.yml"),
         lambda: ref_importer.load_metadata(tmp + "/no_meta.yml")),
        (lambda: libmag.combine_paths(None, "x"), lambda: ref_libmag.combine_paths(None, "x")),
        (lambda: roi_prof.ROIProfile().add_profiles("nonesuch"),
         lambda: ns.roi_prof.ROIProfile().add_profiles("nonesuch")),
        (lambda: np_io.get_num_channels(None), lambda: ns.np_io.get_num_channels(None)),
        (lambda: sd.detect_blobs_stack("x", None), lambda: ns.stack_detect.detect_blobs_stack("x", None)),
        (lambda: sd.detect_blobs_blocks("x", np_io.Image5d(None)),
         lambda: ns.stack_detect.detect_blobs_blocks("x", ns.np_io.Image5d(None))),
    ]
    config.resolutions = ns.config.resolutions = [[1.0, 1.0, 1.0]]
    for i, (ours, theirs) in enumerate(cases):
        assert outcome(ours) == outcome(theirs), i
    config.resolutions = ns.config.resolutions = None
    for ours, theirs in (
            (detector.calc_overlap, ns.detector.calc_overlap),
            (lambda: sd.setup_blocks(roi_prof.ROIProfile(), (10, 10, 10)),
             lambda: ns.stack_detect.setup_blocks(ns.roi_prof.ROIProfile(), (10, 10, 10)))):
        assert outcome(ours) == outcome(theirs) == ("AttributeError",)
    config.resolutions = ns.config.resolutions = [[1.0, 1.0, 1.0]]


def _np_find_close(blobs, master, tol):
    """numpy stand-in for the GPU box match (the contract of detector._find_close_blobs)."""
    c = np.asarray(blobs)[:, :3].astype(np.int32)
    m = np.asarray(master)[:, :3].astype(np.int32)
    close = np.all(np.abs(m[:, None, :] - c[None, :, :]) <= np.asarray(tol)[None, None, :], axis=2)
    hit = close.any(axis=0)
    last = np.where(close.any(axis=1), close.shape[1] - 1 - np.argmax(close[:, ::-1], axis=1), -1)
    return last.astype(np.int64), hit


@pytest.mark.parametrize("seed", range(4))
def test_host_route_merge_and_prune_random_grids(ns, seed, monkeypatch):
    """The mirror's own ``merge_blobs`` + index-form ``StackPruner.prune_blobs_mp`` +
    ``remove_close_blobs`` bookkeeping (the GPU box match replaced by numpy) against the
    unmodified reference on random, ragged, anisotropic grids with one or two channels:
    the same table row for row and the same pruning-ratio frame."""
    from oracle import ref_shim
    from magellanmapper_b200.cv import chunking, stack_detect as sd
    from magellanmapper_b200.settings import roi_prof
    monkeypatch.setattr(detector, "_find_close_blobs", _np_find_close)
    # the column map is class-level state that every `Blobs` constructor rewrites (as in the
    # reference, detector.py:116, 142-162); detection always leaves the full set behind
    detector.Blobs().cols = [c.value for c in detector.Blobs.Cols]
    ns.detector.Blobs().cols = [c.value for c in ns.detector.Blobs.Cols]
    rng = np.random.default_rng(300 + seed)
    shape = tuple(int(v) for v in rng.integers(40, 150, 3))
    res = [float(v) for v in rng.choice([0.7, 1.0, 2.0, 5.0], 3)]
    mods = {"segment_size": int(rng.integers(25, 70)),
            "prune_tol_factor": tuple(float(v) for v in rng.choice([0.9, 1.0, 1.6], 3))}
    ref_prof = ref_shim.set_profile(ns, res, **mods)
    rb = ns.stack_detect.setup_blocks(ref_prof, shape)
    prof = roi_prof.ROIProfile()
    prof.add_profiles("roi_blobs.yaml")
    for k, v in mods.items():
        prof[k] = v
    config.roi_profile, config.roi_profiles, config.resolutions = prof, [prof], [res]
    ob = sd.setup_blocks(prof, shape)
    n_chl = 1 + seed % 2
    base = rng.integers(0, shape, (int(rng.integers(300, 1500)), 3)).astype(float)
    seg_a = np.zeros(rb.sub_roi_slices.shape, dtype=object)
    seg_b = np.zeros(rb.sub_roi_slices.shape, dtype=object)
    for c in np.ndindex(*seg_a.shape):
        sl = rb.sub_roi_slices[c]
        inside = np.all([(base[:, a] >= sl[a].start) & (base[:, a] < sl[a].stop)
                         for a in range(3)], axis=0)
        pts = base[inside] + rng.integers(-2, 3, (int(inside.sum()), 3))
        pts = np.clip(pts, [s.start for s in sl], [s.stop - 1 for s in sl])
        if len(pts) == 0:
            seg_a[c] = seg_b[c] = None
            continue
        t = np.full((len(pts), 11), -1.0)
        t[:, :3] = pts
        t[:, 3] = rng.uniform(4, 9, len(pts))
        t[:, 6] = rng.integers(0, n_chl, len(pts))
        t[:, 7:10] = pts
        seg_a[c], seg_b[c] = t.copy(), t.copy()
    np.testing.assert_array_equal(chunking.merge_blobs(seg_a), ns.chunking.merge_blobs(seg_b))
    channels = list(range(n_chl))
    want, want_df = ns.stack_detect.StackPruner.prune_blobs_mp(
        np.zeros(shape, np.uint8), seg_b, rb.overlap, rb.tol, rb.sub_roi_slices,
        rb.sub_rois_offsets, channels, rb.overlap_padding)
    got, got_df = sd.StackPruner.prune_blobs_mp(
        None, seg_a, ob.overlap, ob.tol, ob.sub_roi_slices, ob.sub_rois_offsets, channels,
        ob.overlap_padding)
    np.testing.assert_array_equal(got, want)
    assert list(got_df.columns) == list(want_df.columns)
    np.testing.assert_allclose(got_df.to_numpy(dtype=float), want_df.to_numpy(dtype=float))


def test_path_helpers_and_backup_rule_random(ns, tmp_path):
    """``libmag.splitext`` / ``insert_before_ext`` / ``combine_paths`` on a few hundred
    generated paths, and ``backup_file`` replayed in two directories (plain, with a modifier,
    repeated, with a companion file): the same strings and the same directory listings."""
    import importlib
    import itertools
    from magellanmapper_b200.io import libmag
    ref = importlib.import_module("magmap.io.libmag")
    stems = ["item", "foo/bar/item", "foo.d/bar", "a.b/c.d/e", "x/", "", "img.nii", "vol.nii.gz",
             "arch.tar.gz", "deep/er.tar/thing", ".hidden", "dir.with.dots/plain"]
    exts = ["", ".py", ".npz", ".file.ext", ".nii.gz", ".tar"]
    for stem, ext in itertools.product(stems, exts):
        path = stem + ext
        assert libmag.splitext(path) == ref.splitext(path), path
        for insert, sep in (("totest", "_"), ("(1)", ""), ("a.b", "-")):
            assert libmag.insert_before_ext(path, insert, sep) == ref.insert_before_ext(path, insert, sep)
        for suffix, sep, new_ext, keep in (("file.py", "_", None, False), ("blobs.npz", "_", None, True),
                                           ("file", "-", "ext", False), ("image5d.npy", "_", "npz", True)):
            for base in (path, None):
                assert (libmag.combine_paths(base, suffix, sep, new_ext, False, keep)
                        == ref.combine_paths(base, suffix, sep, new_ext, False, keep)), (base, suffix)
    existing = str(tmp_path)
    assert libmag.combine_paths(existing, "x.npz", check_dir=True) == \
        ref.combine_paths(existing, "x.npz", check_dir=True)

    listings = []
    for tag, mod in (("ours", libmag), ("theirs", ref)):
        d = tmp_path / tag
        d.mkdir()

        def touch(name):
            (d / name).write_text(name)
        for _ in range(3):                                   # plain: (1), (2), (3)
            touch("s_blobs.npz")
            mod.backup_file(str(d / "s_blobs.npz"))
        for _ in range(3):                                   # modifier: _old, _old(1), _old(2)
            touch("t.csv")
            mod.backup_file(str(d / "t.csv"), "_old")
        for _ in range(2):                                   # companion file moves along
            touch("mesh.obj")
            touch("mesh.mtl")
            mod.backup_file(str(d / "mesh.obj"))
        mod.backup_file(str(d / "absent.npz"))               # nothing to do
        listings.append(sorted((p.name, p.read_text()) for p in d.iterdir()))
    assert listings[0] == listings[1]
    assert ("s_blobs(3).npz", "s_blobs.npz") in listings[0] and ("t_old(2).csv", "t.csv") in listings[0]
