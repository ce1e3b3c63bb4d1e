"""GPU tests of the reference-facing Python surface (detector.detect_blobs,
stack_detect.detect_blobs_blocks / detect_blobs_stack, StackPruner) against the
reference-generated vectors and the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import magmap_restated as mm                      # noqa: E402
from magellanmapper_b200 import synth                         # noqa: E402
from magellanmapper_b200.cv import detector, stack_detect     # noqa: E402
from magellanmapper_b200.io import np_io                      # noqa: E402
from magellanmapper_b200.settings import config, roi_prof     # noqa: E402


@pytest.fixture(autouse=True)
def _cuda():
    from magellanmapper_b200 import gpu
    gpu.require_cuda()
    yield
    stack_detect.StackDetector.release_workspace()


def _setup(resolution=(1, 1, 1), near_max=-1.0, **mods):
    prof = roi_prof.ROIProfile()
    prof.add_profiles("roi_blobs.yaml")
    for k, v in mods.items():
        prof[k] = v
    config.roi_profile = prof
    config.roi_profiles = [prof]
    config.resolutions = [list(resolution)]
    config.near_max = [near_max]
    config.channel = None
    return prof


def _rows(t):
    """order-independent view of a blob table: integer z,y,x + radius"""
    return sorted((int(r[0]), int(r[1]), int(r[2]), round(float(r[3]), 9)) for r in t)


def test_detect_blobs_vs_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "detect_small.npz"))
    _setup(near_max=float(g["near_max"]))
    raw = detector.detect_blobs(g["vol"], [0])
    assert raw.shape[1] == 11 and raw.dtype == np.float64
    assert _rows(raw) == _rows(g["raw"])
    # identical row ORDER too (descending response) and identical non-coordinate columns
    np.testing.assert_array_equal(raw[:, 4:7], g["raw"][:, 4:7])
    np.testing.assert_array_equal(raw[:, 7:10], raw[:, 0:3])
    assert np.mean(np.all(raw[:, :4] == g["raw"][:, :4], axis=1)) > 0.95
    gui = detector.detect_blobs(g["pre"], [0])
    assert _rows(gui) == _rows(g["gui"])
    excl = detector.detect_blobs(g["pre"], [0], np.array([[3, 4, 5], [2, 0, 6]]))
    assert _rows(excl) == _rows(g["excl"])


def test_config1_full_size_raw_and_whole_roi_preprocessed():
    """BASELINE config 1 at its named size, 50x500x500 uint16, ``roi_blobs`` profile:
    ``detect_blobs`` on the raw ROI, and again after ``plot_3d.saturate_roi`` +
    ``denoise_roi`` with the WHOLE ROI as one block (the GUI leg,
    magmap/gui/visualizer.py:2742-2743) - preprocessing within 2e-5 of the oracle, blob
    sets equal except candidates within 1e-4 of the threshold."""
    from oracle import skimage_restated as ski
    from magellanmapper_b200.plot import plot_3d
    shape = (50, 500, 500)
    vol, _ = synth.make_volume(shape, seed=101)
    nm = synth.near_max_of(vol)
    _setup(near_max=nm)
    prof = mm.Profile()

    def compare(got, img, label):
        res = ski.blob_log(img, 3, 5, 10, 0.1, 0.5, full=True)
        want = {(int(z), int(y), int(x), round(float(s) * np.sqrt(3), 9))
                for z, y, x, s in res.blobs}
        have = set(_rows(got))
        near = {tuple(int(v) for v in p[:3]) for p, r in zip(res.peaks, res.responses)
                if abs(r - 0.1) < 1e-4}
        od = set()
        if res.trace is not None:
            od = {tuple(int(v) for v in res.peaks[i][:3]) for i in res.trace.order_dependent}
        diff = {d for d in have ^ want if d[:3] not in near and d[:3] not in od}
        tp = len(have & want)
        f1 = 2 * tp / max(len(have) + len(want), 1)
        print(f"config 1 {label}: gpu={len(have)} oracle={len(want)} F1={f1:.6f} "
              f"near-thr={len(near)} order-dependent={len(od)} unexplained={len(diff)}")
        assert len(want) > 500 and f1 > 0.999
        # what remains must be explained by float32-vs-float64 near-ties only
        assert len(diff) <= max(2, len(want) // 2000), sorted(diff)[:10]

    compare(detector.detect_blobs(vol, [0]), vol, "raw")

    sat = plot_3d.saturate_roi(vol, channel=[0])
    want_sat = mm.saturate_roi(vol, prof, nm)
    assert sat.shape == shape and sat.dtype == np.float64
    assert np.max(np.abs(sat - want_sat)) < 2e-5
    den = plot_3d.denoise_roi(sat, channel=[0])
    want_den = mm.denoise_roi(want_sat, prof)
    err = np.max(np.abs(den - want_den))
    print(f"config 1 whole-ROI saturate+denoise: max abs err {err:.3e}")
    assert err < 2e-5
    # both halves fused in one call equal the two-step route (same kernels, one pass)
    compare(detector.detect_blobs(den, [0]), want_den, "whole-ROI preprocessed")


def test_detect_blobs_none_when_empty():
    _setup()
    assert detector.detect_blobs(np.full((10, 30, 30), 7, dtype=np.uint16), [0]) is None


def test_detect_blobs_two_channels_one_call():
    """channel-last input, per-channel profiles (BASELINE config 4 shape of call)"""
    v0, _ = synth.make_volume((30, 70, 80), seed=51, density=1 / 2500.0)
    v1, _ = synth.make_volume((30, 70, 80), seed=52, density=1 / 2500.0)
    roi = np.stack([v0, v1], axis=-1)
    p0 = _setup()
    p1 = roi_prof.ROIProfile()
    p1.add_profiles("roi_blobs.yaml")
    p1["min_sigma_factor"], p1["max_sigma_factor"] = 4, 10
    config.roi_profiles = [p0, p1]
    both = detector.detect_blobs(roi, None)
    want0 = mm.detect_blobs(v0, mm.Profile(), (1, 1, 1), 0)
    want1 = mm.detect_blobs(v1, mm.Profile(min_sigma_factor=4, max_sigma_factor=10), (1, 1, 1), 1)
    assert _rows(both[both[:, 6] == 0]) == _rows(want0)
    assert _rows(both[both[:, 6] == 1]) == _rows(want1)


def test_remove_close_blobs_vs_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "remove_close.npz"))
    for i in range(int(g["n"])):
        pruned, master = detector.remove_close_blobs(
            g[f"r{i}_check"].copy(), g[f"r{i}_master"].copy(), g[f"r{i}_tol"])
        np.testing.assert_array_equal(pruned, g[f"r{i}_pruned"])
        np.testing.assert_array_equal(master, g[f"r{i}_master_out"])


def test_prune_blobs_mp_vs_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "prune_mp.npz"))
    prof = _setup(segment_size=50)
    shape = tuple(g["shape"])
    b = stack_detect.setup_blocks(prof, shape)
    seg = np.zeros(tuple(g["grid"]), dtype=object)
    for c in np.ndindex(*seg.shape):
        t = g["seg_%d_%d_%d" % c]
        seg[c] = t.copy() if len(t) else None
    out, df = stack_detect.StackPruner.prune_blobs_mp(
        np.zeros(shape, np.uint8), seg, b.overlap, b.tol, b.sub_roi_slices,
        b.sub_rois_offsets, [0], b.overlap_padding)
    np.testing.assert_array_equal(out, g["pruned"])
    assert list(df.columns) == ["blobs", "ratio_pruning", "ratio_adjacent"]


@pytest.mark.parametrize("tag,mods", [("plain", {}), ("excl", {"exclude_border": (2, 1, 1)})])
def test_detect_blobs_blocks_vs_reference_vectors(golden_dir, tmp_path, tag, mods):
    """The reference's own detect_blobs_blocks output (fork pool, seam pruning)
    on a 2x3x3-chunk volume must be reproduced row for row."""
    g = np.load(os.path.join(golden_dir, "stack_small.npz"))
    _setup(near_max=float(g["near_max"]), segment_size=50, **mods)
    config.filename = str(tmp_path / "synth")
    img5d = np_io.Image5d(g["vol"][None])
    os.chdir(tmp_path)
    _, _, blobs = stack_detect.detect_blobs_blocks(
        config.filename, img5d, None, None, [0], False, True, True)
    want = g[f"{tag}_blobs"]
    assert blobs.cols == list(g[f"{tag}_cols"])
    assert blobs.blobs.shape == want.shape
    np.testing.assert_array_equal(blobs.blobs, want)
    assert os.path.exists(tmp_path / "stack_detection_times.csv")


def test_detect_blobs_stack_saves_archive(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "stack_small.npz"))
    _setup(near_max=float(g["near_max"]), segment_size=50)
    config.filename = str(tmp_path / "vol")
    os.chdir(tmp_path)
    img5d = np_io.Image5d(g["vol"][None])
    img5d.is_roi = True
    _, _, blobs = stack_detect.detect_blobs_stack(config.filename, img5d)
    np.testing.assert_array_equal(blobs.blobs, g["plain_blobs"])
    back = detector.Blobs().load_blobs(str(tmp_path / "vol_blobs.npz"))
    np.testing.assert_array_equal(back.blobs, g["plain_blobs"])
    assert back.cols == ["z", "y", "x", "radius", "confirmed", "truth", "channel", "region"]


def test_stack_anisotropic_geometry_vs_oracle(tmp_path):
    """config 5 geometry (resolution 5x1x1: 5x25x25 preprocessing blocks,
    100x500x500-voxel chunks scaled down via segment_size) against the oracle."""
    shape = (40, 120, 110)
    vol, _ = synth.make_volume(shape, seed=61, density=1 / 2500.0)
    nm = synth.near_max_of(vol)
    res = (5, 1, 1)
    _setup(res, nm, segment_size=60)
    config.filename = str(tmp_path / "aniso")
    os.chdir(tmp_path)
    want = mm.detect_blobs_blocks(vol, mm.Profile(segment_size=60), res, nm)
    _, _, blobs = stack_detect.detect_blobs_blocks(
        config.filename, np_io.Image5d(vol[None]), None, None, [0], False, False, True)
    np.testing.assert_array_equal(blobs.blobs, want)
    assert len(want) > 20


def test_stack_accepts_device_tensor(golden_dir, tmp_path):
    """A CUDA tensor image is consumed in place (no host round trip)."""
    g = np.load(os.path.join(golden_dir, "stack_small.npz"))
    _setup(near_max=float(g["near_max"]), segment_size=50)
    config.filename = str(tmp_path / "dev")
    os.chdir(tmp_path)
    t = torch.from_numpy(g["vol"].view(np.int16)).cuda()[None]
    _, _, blobs = stack_detect.detect_blobs_blocks(
        config.filename, np_io.Image5d(t), None, None, [0], False, False, True)
    np.testing.assert_array_equal(blobs.blobs, g["plain_blobs"])


def test_device_tables_equal_host_tables(golden_dir, tmp_path, monkeypatch):
    """The device-resident route (tables merged and seam-pruned in HBM) and the
    host route through the reference's structure give the same table, row for
    row, and the same pruning-ratio frame - one channel and two channels."""
    g = np.load(os.path.join(golden_dir, "stack_small.npz"))
    os.chdir(tmp_path)
    vol = g["vol"]
    two = np.stack([vol, vol[::-1, :, ::-1]], axis=-1)
    for img, chls in ((vol, [0]), (two, [0, 1])):
        outs = []
        for flag in (True, False):
            monkeypatch.setattr(stack_detect, "DEVICE_TABLES", flag)
            _setup(near_max=float(g["near_max"]), segment_size=50)
            config.near_max = [float(g["near_max"])] * len(chls)
            config.filename = str(tmp_path / f"dt{int(flag)}")
            _, _, blobs = stack_detect.detect_blobs_blocks(
                config.filename, np_io.Image5d(img[None]), None, None, chls, False, True, True)
            ratios = np.loadtxt(tmp_path / "blob_ratios.csv", delimiter=",", skiprows=1)
            outs.append((blobs.blobs, ratios))
        np.testing.assert_array_equal(outs[0][0], outs[1][0])
        np.testing.assert_array_equal(outs[0][1], outs[1][1])
        assert len(outs[0][0]) > 100
    np.testing.assert_array_equal(outs[0][0][outs[0][0][:, 6] == 0][:5],
                                  outs[1][0][outs[1][0][:, 6] == 0][:5])


def test_strip_feeder_delivers_every_strip():
    """mmb_upload_pieces through gpu.StripFeeder: every y-strip of a (z, y, x) and
    of a channel-last (z, y, x, c) host array arrives intact, with more strips
    than device buffers (buffer reuse gated on release events)."""
    from magellanmapper_b200 import gpu
    rng = np.random.default_rng(3)
    for shape in ((7, 40, 33), (5, 31, 18, 2)):
        img = rng.integers(0, 65535, size=shape, dtype=np.uint16)
        ranges = [(0, 13), (8, 21), (16, 29), (24, shape[1])]
        feeder = gpu.StripFeeder(img, ranges)
        for j, (y0, y1) in enumerate(ranges):
            strip = feeder.strip(j)
            got = strip.cpu().numpy().view(np.uint16)
            np.testing.assert_array_equal(got, img[:, y0:y1])
            feeder.release(j)
    with pytest.raises(TypeError):
        gpu.StripFeeder(np.asfortranarray(np.zeros((4, 5, 6), np.uint16)), [(0, 5)])


def test_config2_full_size_resident_equals_streamed(tmp_path):
    """BASELINE config 2 at full size (512x2048x2048 uint16, 50 chunks): the stack
    resident in HBM and the same stack streamed from host memory strip by strip
    (different chunk order, different buffers) give the same table bit for bit; every
    blob lies inside the volume, on integer coordinates, with a radius of the ladder."""
    import bench
    dev = torch.device("cuda", 0)
    shape = (512, 2048, 2048)
    vol = synth.device_volume(shape, 1, device=dev)
    nm = bench.near_max_device(vol)
    bench.setup_config(nm, str(tmp_path / "c2"))
    os.chdir(tmp_path)
    _, _, res = stack_detect.detect_blobs_blocks(
        str(tmp_path / "c2"), np_io.Image5d(vol[None]), None, None, [0], False, False, True)
    host = vol.cpu().numpy().view(np.uint16)
    del vol
    torch.cuda.empty_cache()
    _, _, hst = stack_detect.detect_blobs_blocks(
        str(tmp_path / "c2h"), np_io.Image5d(host[None]), None, None, [0], False, False, True)
    np.testing.assert_array_equal(res.blobs, hst.blobs)
    b = res.blobs
    assert b.shape[1] == 8 and len(b) > 200_000
    assert np.all(b[:, :3] >= 0) and np.all(b[:, :3] < np.array(shape))
    # integer coordinates (seam averages are rounded), radii from the sigma ladder only
    assert np.all(b[:, :3] == np.round(b[:, :3]))
    radii = np.unique(np.round(b[:, 3], 9))
    assert len(radii) <= 10 and radii.min() >= 3 * np.sqrt(3) - 1e-6 and radii.max() <= 5 * np.sqrt(3) + 1e-6


def test_near_bounds_on_device_vs_reference_vectors(golden_dir):
    """io.importer on the GPU (mmb_percentiles) against the unmodified reference's
    importer.calc_intensity_bounds / calc_near_intensity_bounds: bit-equal float64 for
    uint16, a narrow-histogram uint16, uint8 and a two-channel channel-last volume."""
    from magellanmapper_b200.io import importer
    g = np.load(os.path.join(golden_dir, "near_bounds.npz"))
    for name in ("u16", "u16_narrow", "u8", "u16_2c"):
        vol = g[f"{name}_vol"]
        lows, highs = importer.calc_plane_bounds(vol)
        np.testing.assert_array_equal(np.array(lows), g[f"{name}_plane_lows"])
        np.testing.assert_array_equal(np.array(highs), g[f"{name}_plane_highs"])
        near_mins, near_maxs = importer.calc_near_bounds(vol)
        np.testing.assert_array_equal(np.ravel(near_mins), np.ravel(g[f"{name}_near_mins"]))
        np.testing.assert_array_equal(np.ravel(near_maxs), np.ravel(g[f"{name}_near_maxs"]))
        lo, hi = importer.calc_intensity_bounds(vol[None], dim_channel=4)
        np.testing.assert_array_equal(np.array([lo, hi]), g[f"{name}_whole"])
        # device-resident input is read in place
        t = torch.from_numpy(vol.view(np.int16) if vol.dtype == np.uint16 else vol).cuda()
        m2, x2 = importer.calc_near_bounds(t)
        np.testing.assert_array_equal(np.ravel(m2), np.ravel(near_mins))
    with pytest.raises(NotImplementedError):
        importer.calc_near_bounds(np.zeros((2, 8, 8), np.float32))


def test_near_max_of_config2_plane_sized_input():
    """A 2048x2048 plane stack: per-plane percentiles equal numpy's on planes with
    4.2 M samples (counts beyond 16 bits, ranks that fall between duplicates)."""
    from magellanmapper_b200.io import importer
    rng = np.random.default_rng(8)
    vol = (400 + 30 * rng.standard_normal((3, 2048, 2048))).clip(0, 65535).astype(np.uint16)
    vol[1, :64] = 60000
    lows, highs = importer.calc_plane_bounds(vol)
    for z in range(3):
        lo, hi = np.percentile(vol[z], (0.5, 99.5))
        assert lows[z][0] == lo and highs[z][0] == hi


def _setup_two(near_maxs, res=(1, 1, 1)):
    p0 = _setup(res, near_maxs[0])
    p1 = roi_prof.ROIProfile()
    p1.add_profiles("roi_blobs.yaml")
    config.roi_profiles = [p0, p1]
    config.near_max = list(near_maxs)
    return p0, p1


def test_make_isotropic_vs_reference_vectors(golden_dir):
    """cv_nd.make_isotropic: integer ROIs bit-exact (float64 interpolation, truncating
    cast), float ROIs within float32 rounding; growing and shrinking axes."""
    from magellanmapper_b200.cv import cv_nd
    g = np.load(os.path.join(golden_dir, "iso_unmix.npz"))
    _setup((3, 1, 1))
    got = cv_nd.make_isotropic(g["vol"], (0.96, 1, 1))
    assert got.dtype == np.uint16
    np.testing.assert_array_equal(got, g["iso_up_resized"])
    got = cv_nd.make_isotropic(g["pre"], (0.96, 1, 1))
    assert np.max(np.abs(got - g["iso_up_pre_resized"])) < 1e-6
    _setup((1, 1, 1))
    np.testing.assert_array_equal(cv_nd.make_isotropic(g["vol"], (1.5, 0.6, 0.75)),
                                  g["iso_mixed_resized"])
    got = cv_nd.make_isotropic(g["pre"], (1.5, 0.6, 0.75))
    assert np.max(np.abs(got - g["iso_mixed_pre_resized"])) < 1e-6
    # a one-voxel-thick ROI switches to 'edge' boundaries
    thin = g["vol"][:1]
    _setup((3, 1, 1))
    np.testing.assert_array_equal(cv_nd.make_isotropic(thin, 1),
                                  mm.make_isotropic(thin, 1, (3, 1, 1)))


def test_detect_blobs_isotropic_vs_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "iso_unmix.npz"))
    _setup((3, 1, 1), float(g["near_max"]), isotropic=(0.96, 1, 1))
    raw = detector.detect_blobs(g["vol"], [0])
    np.testing.assert_array_equal(raw, g["iso_up_raw"])
    pre = detector.detect_blobs(g["pre"], [0], np.array([[1, 2, 0], [0, 3, 2]]))
    np.testing.assert_array_equal(pre, g["iso_up_pre"])
    _setup((1, 1, 1), float(g["near_max"]), isotropic=(1.5, 0.6, 0.75))
    np.testing.assert_array_equal(detector.detect_blobs(g["pre"], [0]), g["iso_mixed_pre"])


def test_detect_blobs_spectral_unmixing_vs_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "iso_unmix.npz"))
    p0, _ = _setup_two([float(g["near_max"]), float(g["near_max1"])])
    p0.spectral_unmixing = {0: {1: 0.4}}
    got = detector.detect_blobs(g["pre2"], None)
    assert _rows(got) == _rows(g["unmix_pre"])
    np.testing.assert_array_equal(got[:, 4:7], g["unmix_pre"][:, 4:7])
    # the same on the raw two-channel ROI (float64 after the subtraction: no 1/65535)
    want = mm.detect_blobs(g["two"], [mm.Profile(spectral_unmixing={0: {1: 0.4}}), mm.Profile()],
                           (1, 1, 1), [0, 1])
    raw = detector.detect_blobs(g["two"], None)
    assert len(want) > 20
    assert len(set(_rows(raw)) ^ set(_rows(want))) <= max(2, len(want) // 100)


def test_stack_isotropic_exclude_border_vs_reference_vectors(golden_dir, tmp_path):
    """The lightsheet profile's combination through detect_blobs_blocks (host route)."""
    g = np.load(os.path.join(golden_dir, "iso_unmix.npz"))
    _setup((2.5, 1, 1), float(g["snm"]), isotropic=(0.96, 1, 1), segment_size=40,
           exclude_border=(1, 0, 0))
    os.chdir(tmp_path)
    _, _, blobs = stack_detect.detect_blobs_blocks(
        str(tmp_path / "iso"), np_io.Image5d(g["svol"][None]), None, None, [0], False, False,
        True)
    np.testing.assert_array_equal(blobs.blobs, g["stack_iso_blobs"])


def test_stack_unmixing_device_route_equals_host_route_and_oracle(tmp_path, monkeypatch):
    """Two channels, channel 0 unmixed by channel 1, through detect_blobs_blocks: the
    device-table route equals the host route row for row, and one chunk's detection
    (both channels preprocessed block by block, then unmixed) equals the oracle's except
    candidates within 1e-4 of the threshold."""
    from oracle import skimage_restated as ski
    v0, _ = synth.make_volume((40, 90, 80), seed=71, density=1 / 2500.0)
    v1, _ = synth.make_volume((40, 90, 80), seed=72, density=1 / 2500.0)
    two = np.stack([v0, v1], axis=-1)
    nms = [synth.near_max_of(v0), synth.near_max_of(v1)]
    p0, p1 = _setup_two(nms)
    for p in (p0, p1):
        p["segment_size"] = 35
    p0.spectral_unmixing = {0: {1: 0.5}}
    os.chdir(tmp_path)

    def run():
        _, _, b = stack_detect.detect_blobs_blocks(
            str(tmp_path / "um"), np_io.Image5d(two[None]), None, None, [0, 1], False, False,
            True)
        return b.blobs
    dev = run()
    monkeypatch.setattr(stack_detect, "DEVICE_TABLES", False)
    host = run()
    assert len(dev) > 50
    np.testing.assert_array_equal(dev, host)

    # one chunk against the oracle
    prof0 = mm.Profile(segment_size=35, spectral_unmixing={0: {1: 0.5}})
    prof1 = mm.Profile(segment_size=35)
    blocks = mm.setup_blocks(prof0, two.shape[:3], (1, 1, 1))
    coord = (0, 1, 1)
    sub = two[blocks.sub_roi_slices[coord]]
    pre = np.stack([mm.preprocess_blocks(sub[..., k], blocks.denoise_max_shape, pr, nm)
                    for k, (pr, nm) in enumerate(zip((prof0, prof1), nms))], axis=-1)
    unmixed = np.subtract(pre[..., 0], 0.5 * pre[..., 1])
    unmixed[unmixed < 0] = 0
    res = ski.blob_log(unmixed, 3, 5, 10, 0.1, 0.5, full=True)
    want = {(int(z), int(y), int(x), round(float(sg) * np.sqrt(3), 9)) for z, y, x, sg in res.blobs}
    _, seg = stack_detect.StackDetector.detect_sub_roi(
        coord, np.zeros(3), np.subtract(blocks.sub_roi_slices.shape, 1),
        blocks.denoise_max_shape, None, None, sub, [0])
    got = set(_rows(seg))
    near = {tuple(int(v) for v in p[:3]) for p, r in zip(res.peaks, res.responses)
            if abs(r - 0.1) < 1e-4}
    od = set()
    if res.trace is not None:
        od = {tuple(int(v) for v in res.peaks[i][:3]) for i in res.trace.order_dependent}
    diff = {d for d in got ^ want if d[:3] not in near and d[:3] not in od}
    print(f"unmixed chunk: gpu={len(got)} oracle={len(want)} near={len(near)} od={len(od)} "
          f"unexplained={len(diff)}")
    assert len(want) > 5 and not diff, sorted(diff)


def test_write_npy_equals_reference_files_and_feeds_the_detector(golden_dir, tmp_path):
    """``np_io.write_npy`` (near bounds computed on the GPU) reproduces the image and
    metadata files the unmodified reference wrote for the same array; ``setup_images`` of
    the result drives ``detect_blobs_stack`` straight from the memory map."""
    import yaml
    from magellanmapper_b200.io import np_io as nio
    ref_img = np.load(os.path.join(golden_dir, "feed", "sample_image5d.npy"))
    with open(os.path.join(golden_dir, "feed", "sample_meta.yml")) as f:
        ref_md = yaml.safe_load(f)
    nio.write_npy(ref_img, {"resolutions": [[5.0, 1.1, 1.1]], "magnification": 20.0, "zoom": 1.0},
                  str(tmp_path / "sample.czi"))
    with open(tmp_path / "sample_image5d.npy", "rb") as a, \
            open(os.path.join(golden_dir, "feed", "sample_image5d.npy"), "rb") as b:
        assert a.read() == b.read()
    with open(tmp_path / "sample_meta.yml") as f:
        assert yaml.safe_load(f) == ref_md
    _setup()
    img5d = nio.setup_images(str(tmp_path / "sample"))
    config.channel = None
    os.chdir(tmp_path)
    _, _, blobs = stack_detect.detect_blobs_stack(str(tmp_path / "sample"), img5d)
    _, _, want = stack_detect.detect_blobs_stack(str(tmp_path / "sample2"),
                                                 np_io.Image5d(np.array(img5d.img)))
    assert blobs.blobs is not None and len(blobs.blobs) > 5
    np.testing.assert_array_equal(blobs.blobs, want.blobs)


def test_colocalize_blobs_vs_reference_vectors(golden_dir):
    """colocalizer.colocalize_blobs against the unmodified reference: raw uint16 and
    preprocessed float64 ROIs, "min" and percentile thresholds, blobs sharing a voxel,
    on the ROI faces and outside the ROI."""
    from magellanmapper_b200.cv import colocalizer
    g = np.load(os.path.join(golden_dir, "coloc.npz"))
    blobs = g["blobs"]
    for key, roi, thresh in (("min_raw", g["roi"], None), ("p5_raw", g["roi"], 5),
                             ("min_pre", g["pre"], None), ("p30_pre", g["pre"], 30)):
        got = colocalizer.colocalize_blobs(roi, blobs, thresh)
        assert got.dtype == np.uint8 and got.shape == g[key].shape
        np.testing.assert_array_equal(got, g[key], err_msg=key)
    assert colocalizer.colocalize_blobs(g["roi"][..., 0], blobs) is None
    assert colocalizer.colocalize_blobs(g["roi"], None) is None
    # exact integer sums: means of the raw ROI equal numpy's bit for bit
    in_roi = detector.get_blobs_in_roi(blobs, (0, 0, 0), g["roi"].shape[:3], reverse=False)[0]
    means, counts = colocalizer.blob_surround_means(g["roi"], in_roi)
    b = int(np.argmax(counts))
    z, y, x = in_roi[b, :3].astype(int)
    assert 1 <= counts[b] <= 33


def test_stack_colocalization_two_channels(tmp_path):
    """detect_blobs_blocks(coloc=True): same blobs as without co-localisation, one flag
    per channel in the reference's (column-10) slice, and per chunk the flags the oracle
    computes from the oracle-preprocessed sub-ROI."""
    v0, _ = synth.make_volume((40, 90, 80), seed=71, density=1 / 2500.0)
    v1, _ = synth.make_volume((40, 90, 80), seed=72, density=1 / 2500.0)
    v1 = np.maximum(v1, (v0 * 0.7).astype(np.uint16))
    two = np.stack([v0, v1], axis=-1)
    nms = [synth.near_max_of(v0), synth.near_max_of(v1)]
    p0, p1 = _setup_two(nms)
    for p in (p0, p1):
        p["segment_size"] = 35
    os.chdir(tmp_path)

    def run(coloc):
        _, _, b = stack_detect.detect_blobs_blocks(
            str(tmp_path / "co"), np_io.Image5d(two[None]), None, None, [0, 1], False, False,
            True, coloc)
        return b
    plain, co = run(False), run(True)
    np.testing.assert_array_equal(co.blobs, plain.blobs)
    assert plain.colocalizations is None
    assert co.colocalizations.shape == (len(co.blobs), 2) and co.colocalizations.dtype == np.uint8
    # one chunk against the oracle
    prof = mm.Profile(segment_size=35)
    blocks = mm.setup_blocks(prof, two.shape[:3], (1, 1, 1))
    coord = (0, 1, 1)
    sub = two[blocks.sub_roi_slices[coord]]
    pre = np.stack([mm.preprocess_blocks(sub[..., k], blocks.denoise_max_shape, prof, nms[k])
                    for k in range(2)], axis=-1)
    _, seg = stack_detect.StackDetector.detect_sub_roi(
        coord, np.zeros(3), np.subtract(blocks.sub_roi_slices.shape, 1),
        blocks.denoise_max_shape, None, None, sub, [0, 1], coloc=True)
    assert seg.shape[1] == 13 and len(seg) > 10
    want = mm.colocalize_blobs(pre, seg[:, :11])
    np.testing.assert_array_equal(seg[:, 11:13].astype(np.uint8), want)
    assert want.sum() > len(seg)            # every blob at least in its own channel


def test_detect_blobs_other_integer_dtypes_follow_img_as_float():
    """int16 / uint32 / int32 ROIs are converted as scikit-image's ``img_as_float`` does
    (signed: ``(2 x + 1) / (max - min)``), not read as raw counts."""
    vol, _ = synth.make_volume((24, 70, 64), seed=33, density=1 / 1500.0)
    _setup()
    for dt, conv in ((np.int16, lambda v: (v // 2).astype(np.int16)),
                     (np.uint32, lambda v: v.astype(np.uint32) * 65537),
                     (np.int32, lambda v: (v.astype(np.int64) * 32768).astype(np.int32))):
        roi = conv(vol)
        want = mm.detect_blobs(roi, mm.Profile(), (1, 1, 1), 0)
        got = detector.detect_blobs(roi, [0])
        assert want is not None and len(want) > 10
        assert len(set(_rows(got)) ^ set(_rows(want))) <= max(1, len(want) // 200), dt


def test_streamed_stack_with_overflowing_candidate_buffers(tmp_path):
    """A host stack streamed as five y-strips through a workspace whose candidate
    buffers overflow on the first chunks (capacity 8): every overflowing chunk is redone
    from its strip, which may only be handed back to the feeder once all of its chunks
    are collected - same table as the resident run with ample buffers."""
    from magellanmapper_b200 import gpu
    shape = (40, 170, 64)
    vol, _ = synth.make_volume(shape, seed=77, density=1 / 1500.0)
    _setup(near_max=synth.near_max_of(vol), segment_size=35)
    os.chdir(tmp_path)

    def run(img):
        _, _, b = stack_detect.detect_blobs_blocks(
            str(tmp_path / "ov"), np_io.Image5d(img[None]), None, None, [0], False, False, True)
        return b.blobs
    want = run(torch.from_numpy(vol.view(np.int16)).cuda())
    stack_detect.StackDetector.release_workspace()
    settings = config.get_roi_profile(0)
    blocks = stack_detect.setup_blocks(settings, shape)
    assert blocks.sub_roi_slices.shape[1] >= 4          # at least four strips, two buffers
    largest = [max(s[a].stop - s[a].start for s in blocks.sub_roi_slices.flat) for a in range(3)]
    small = gpu.ChunkDetector(tuple(largest), capacity=8)
    stack_detect.StackDetector._gpu_detector = small
    got = run(vol)
    assert small.capacity > 8                           # buffers were regrown on overflow
    assert len(want) > 100
    np.testing.assert_array_equal(got, want)
