"""GPU parity tests: every C-ABI entry point against the CPU oracle.

Run on the B200 box: ``python -m pytest tests -m gpu``.  All calls go through
``libmmb200.so`` (ctypes); nothing here reads /root/reference.

Tolerances: filtered volumes within 1e-4 relative of the
float64 oracle (fp32 arithmetic; measured ~1e-6); integer outputs (peak
coordinates, scale index, prune decisions, seam matches) bit-exact, except
candidates whose response lies within 1e-4 of the detection threshold or whose
float64 neighbourhood has a near-tie, which the tests list explicitly.
"""
import os

import numpy as np
import pytest
from scipy import ndimage as ndi

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import magmap_restated as mm           # noqa: E402
from oracle import skimage_restated as ski         # noqa: E402
from magellanmapper_b200 import synth              # noqa: E402

REL_TOL = 1e-4


@pytest.fixture(scope="module")
def gpu():
    from magellanmapper_b200 import gpu as g
    g.require_cuda()
    return g


def _rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def _vol_to_dev(gpu, a):
    """float array (Z,Y,X) -> padded device volume"""
    Z, Y, X = a.shape
    v = gpu.new_volume(Z, Y, X)
    v[:, :, :X] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    return v


def _kernels(sigma):
    """scipy's own weights for order 0 and 2"""
    from scipy.ndimage._filters import _gaussian_kernel1d
    r = int(4.0 * sigma + 0.5)
    return _gaussian_kernel1d(sigma, 0, r)[::-1], _gaussian_kernel1d(sigma, 2, r)[::-1], r


# ------------------------------------------------------------------ to_float

@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.float32, np.float64])
def test_to_float(gpu, dtype):
    rng = np.random.default_rng(0)
    a = (rng.uniform(0, 250, (7, 33, 45))).astype(dtype)
    scale = 1 / 65535.0 if dtype == np.uint16 else (1 / 255.0 if dtype == np.uint8 else 1.0)
    out = gpu.to_float(gpu.as_source(a), scale)
    torch.cuda.synchronize()
    got = out[:, :, :45].cpu().numpy()
    want = ski.img_as_float(a) if dtype in (np.uint8, np.uint16) else a
    assert _rel_err(got, want.astype(np.float64)) < 1e-6


def test_to_float_channel_last(gpu):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 65535, (5, 20, 30, 2)).astype(np.uint16)
    for c in range(2):
        out = gpu.to_float(gpu.as_source(a, channel=c), 1 / 65535.0)
        torch.cuda.synchronize()
        got = out[:, :, :30].cpu().numpy()
        assert _rel_err(got, a[..., c] / 65535.0) < 1e-6


# ------------------------------------------------------------ single sweeps

@pytest.mark.parametrize("shape", [(40, 70, 90), (12, 48, 505), (3, 5, 7), (50, 130, 64),
                                   (24, 37, 300), (20, 33, 777), (70, 300, 330)])
@pytest.mark.parametrize("sigma", [1.0, 3.0, 3.6666666666666665, 5.0, 8.0])
def test_log_pass_each_axis(gpu, shape, sigma):
    rng = np.random.default_rng(int(sigma * 10) + shape[0])
    a = rng.uniform(0, 1, shape)
    b = rng.uniform(-1, 1, shape)
    g, h, r = _kernels(sigma)
    da, db = _vol_to_dev(gpu, a), _vol_to_dev(gpu, b)
    X = shape[2]
    for axis in (0, 1, 2):
        ga = ndi.correlate1d(a, g, axis, mode="reflect")
        ha = ndi.correlate1d(a, h, axis, mode="reflect")
        gb = ndi.correlate1d(b, g, axis, mode="reflect")
        # mode 0: (g*a, h*a)
        o0, o1 = gpu.log_pass(da, None, X, axis, 0, sigma)
        torch.cuda.synchronize()
        assert _rel_err(o0[:, :, :X].cpu().numpy(), ga) < REL_TOL
        assert _rel_err(o1[:, :, :X].cpu().numpy(), ha) < REL_TOL
        # mode 1: (g*a, h*a + g*b); mode 2: scale*(h*a + g*b)
        o0, o1 = gpu.log_pass(da, db, X, axis, 1, sigma)
        torch.cuda.synchronize()
        assert _rel_err(o0[:, :, :X].cpu().numpy(), ga) < REL_TOL
        assert _rel_err(o1[:, :, :X].cpu().numpy(), ha + gb) < REL_TOL
        o0, _ = gpu.log_pass(da, db, X, axis, 2, sigma, scale=-sigma * sigma)
        torch.cuda.synchronize()
        assert _rel_err(o0[:, :, :X].cpu().numpy(), -sigma * sigma * (ha + gb)) < REL_TOL


@pytest.mark.parametrize("sigma", [3.0, 4.111111111111111, 5.0])
def test_x_sweep_long_launch(gpu, sigma):
    """Launches with >= 16 tiles per SM take the warp-specialised x sweep
    (conv_x_ws_kernel: TMA producer warp, 16 consumer warps, store warp); both x
    faces, a nearly empty last tile (X = 520 = 2 * 256 + 8) and a row tail."""
    shape = (61, 630, 520)
    rng = np.random.default_rng(int(sigma * 7))
    a = rng.uniform(0, 1, shape).astype(np.float32).astype(np.float64)
    g, h, r = _kernels(sigma)
    da = _vol_to_dev(gpu, a)
    o0, o1 = gpu.log_pass(da, None, shape[2], 2, 0, sigma)
    torch.cuda.synchronize()
    assert _rel_err(o0[:, :, :shape[2]].cpu().numpy(), ndi.correlate1d(a, g, 2, mode="reflect")) < REL_TOL
    assert _rel_err(o1[:, :, :shape[2]].cpu().numpy(), ndi.correlate1d(a, h, 2, mode="reflect")) < REL_TOL


def test_log_pass_large_radius_fallback(gpu):
    """sigma = 20 -> radius 80 > 64: generic kernel"""
    rng = np.random.default_rng(5)
    a = rng.uniform(0, 1, (20, 30, 50))
    g, h, r = _kernels(20.0)
    da = _vol_to_dev(gpu, a)
    for axis in (0, 1, 2):
        o0, o1 = gpu.log_pass(da, None, 50, axis, 0, 20.0)
        torch.cuda.synchronize()
        assert _rel_err(o0[:, :, :50].cpu().numpy(), ndi.correlate1d(a, g, axis, mode="reflect")) < REL_TOL
        assert _rel_err(o1[:, :, :50].cpu().numpy(), ndi.correlate1d(a, h, axis, mode="reflect")) < REL_TOL


# ---------------------------------------------------------------- LoG scale

@pytest.mark.parametrize("shape,seed", [((50, 120, 140), 0), ((12, 48, 130), 1), ((30, 505, 48), 2)])
def test_log_scale_matches_scipy(gpu, shape, seed):
    vol, _ = synth.make_volume(shape, seed=seed, density=1 / 3000.0)
    img = ski.img_as_float(vol)
    src = gpu.as_source(vol)
    f = gpu.to_float(src, 1 / 65535.0)
    worst = 0.0
    for sigma in ski.sigma_list(3, 5, 10)[[0, 3, 9]]:
        out = gpu.log_scale(f, shape[2], sigma)
        torch.cuda.synchronize()
        want = -ndi.gaussian_laplace(img, sigma) * sigma ** 2
        worst = max(worst, _rel_err(out[:, :, :shape[2]].cpu().numpy(), want))
    print(f"log_scale max rel err {worst:.3e}")
    assert worst < REL_TOL


@pytest.mark.parametrize("shape", [(9, 70, 140), (3, 64, 33), (5, 505, 505), (12, 200, 48),
                                   (2, 97, 300), (40, 66, 129), (7, 48, 505), (3, 33, 70),
                                   (4, 32, 32)])
@pytest.mark.parametrize("sigma", [1.9, 3.0, 4.0, 4.4, 4.6, 5.0])
def test_fused_xy_sweep_equals_separate_sweeps(gpu, shape, sigma):
    """The fused x -> y sweep (log_xy.cu; radii <= 20, x and y extent >= 32) inside
    ``mmb_log_scale`` against the three separate sweeps of ``mmb_log_pass``: same
    accumulation order per output, so the LoG volumes are equal BIT FOR BIT; and against
    scipy within 1e-4.  Shapes cover ragged last tiles in x and y, pitch > X padding, planes
    too few to fill the GPU (the y axis is then cut into segments), radii of every ring
    geometry (RP = 8, 16, 24)."""
    rng = np.random.default_rng(int(sigma * 10) + shape[1])
    a = rng.random(shape, dtype=np.float32)
    a[rng.random(shape) < 0.01] += 3.0
    f = _vol_to_dev(gpu, a)
    X = shape[2]
    fused = gpu.log_scale(f, X, sigma)
    A, B = gpu.log_pass(f, None, X, 2, 0, sigma)
    C, D = gpu.log_pass(A, B, X, 1, 1, sigma)
    sep, _ = gpu.log_pass(C, D, X, 0, 2, sigma, scale=-(sigma * sigma))
    torch.cuda.synchronize()
    got, want = fused[:, :, :X].cpu().numpy(), sep[:, :, :X].cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_array_equal(got, want)
    ref = -ndi.gaussian_laplace(a.astype(np.float64), sigma) * sigma ** 2
    assert _rel_err(got, ref) < REL_TOL


# ---------------------------------------------------------------- local max

def _gpu_cube(gpu, f, X, sigmas):
    outs = [gpu.log_scale(f, X, s) for s in sigmas]
    torch.cuda.synchronize()
    return outs


def test_localmax_exact_on_same_cube(gpu):
    """Given the SAME float32 cube, the peak set must equal peak_local_max's."""
    shape = (40, 90, 100)
    vol, _ = synth.make_volume(shape, seed=3, density=1 / 2500.0)
    f = gpu.to_float(gpu.as_source(vol), 1 / 65535.0)
    sigmas = ski.sigma_list(3, 5, 10)
    cube_d = _gpu_cube(gpu, f, shape[2], sigmas)
    cube = np.stack([c[:, :, :shape[2]].cpu().numpy() for c in cube_d], axis=-1)
    want, want_resp = ski.peak_local_max_4d(cube, np.float32(0.1))
    cand = gpu.new_cand_buffer(100000)
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    for i in range(len(sigmas)):
        gpu.localmax(cube_d[i - 1] if i > 0 else None, cube_d[i],
                     cube_d[i + 1] if i + 1 < len(sigmas) else None, shape[2], i, 0.1,
                     cand, counter)
    torch.cuda.synchronize()
    n = int(counter.item())
    got = gpu.cands_to_numpy(cand, n)
    got_set = {(int(c["z"]), int(c["y"]), int(c["x"]), int(c["s"])) for c in got}
    want_set = {tuple(int(v) for v in row) for row in want}
    assert len(got_set) == n, "duplicates emitted"
    assert got_set == want_set
    assert n > 20
    # responses are the cube values
    lut = {tuple(int(v) for v in row): r for row, r in zip(want, want_resp)}
    for c in got:
        assert lut[(int(c["z"]), int(c["y"]), int(c["x"]), int(c["s"]))] == c["resp"]


def test_localmax_plateau_and_borders(gpu):
    """Ties count for every voxel of a plateau; faces and scale ends clamp."""
    Z, Y, X = 6, 7, 9
    rng = np.random.default_rng(2)
    cube = rng.uniform(0, 0.05, (Z, Y, X, 3)).astype(np.float32)
    cube[0, 0, 0, 0] = 0.9                       # corner, first scale
    cube[5, 6, 8, 2] = 0.8                       # opposite corner, last scale
    cube[2, 3, 4, 1] = cube[2, 3, 5, 1] = 0.7    # two-voxel plateau
    cube[4, 1, 1, 1] = 0.1                       # equal to threshold: not a peak (strict >)
    want, _ = ski.peak_local_max_4d(cube, np.float32(0.1))
    vols = [_vol_to_dev(gpu, cube[..., i]) for i in range(3)]
    cand = gpu.new_cand_buffer(64)
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    for i in range(3):
        gpu.localmax(vols[i - 1] if i > 0 else None, vols[i], vols[i + 1] if i < 2 else None,
                     X, i, 0.1, cand, counter)
    torch.cuda.synchronize()
    got = gpu.cands_to_numpy(cand, int(counter.item()))
    got_set = {(int(c["z"]), int(c["y"]), int(c["x"]), int(c["s"])) for c in got}
    assert got_set == {tuple(int(v) for v in r) for r in want}
    assert (2, 3, 4, 1) in got_set and (2, 3, 5, 1) in got_set and (4, 1, 1, 1) not in got_set


def test_localmax_overflow_is_counted(gpu):
    cube = np.zeros((4, 8, 64), dtype=np.float32)
    cube[::2, ::2, ::2] = 1.0                    # many isolated peaks
    v = _vol_to_dev(gpu, cube)
    cand = gpu.new_cand_buffer(10)
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    gpu.localmax(None, v, None, 64, 0, 0.5, cand, counter)
    torch.cuda.synchronize()
    assert int(counter.item()) == int((cube > 0.5).sum()) > 10


# -------------------------------------------------------------------- prune

def _cands_from_peaks(peaks, resp):
    from magellanmapper_b200.gpu import CAND_DTYPE
    c = np.zeros(len(peaks), dtype=CAND_DTYPE)
    c["z"], c["y"], c["x"], c["s"] = peaks[:, 0], peaks[:, 1], peaks[:, 2], peaks[:, 3]
    c["resp"] = resp
    return c


def test_prune_within_matches_oracle(gpu):
    """Clustered candidates (many overlaps, chains, equal-sigma ties)."""
    rng = np.random.default_rng(7)
    sigmas = ski.sigma_list(3, 5, 10)
    n = 3000
    centres = rng.integers(0, 120, (60, 3))
    pts = centres[rng.integers(0, 60, n)] + rng.integers(-9, 10, (n, 3))
    pts = np.clip(pts, 0, 127)
    pts = np.unique(pts, axis=0)
    n = len(pts)
    s = rng.integers(0, 10, n)
    resp = rng.uniform(0.1, 1, n).astype(np.float32)
    resp[::7] = resp[0]                                  # response ties
    order = np.lexsort((s, pts[:, 2], pts[:, 1], pts[:, 0]))   # C order of (z,y,x,s)
    pts, s, resp = pts[order], s[order], resp[order]
    o2 = np.argsort(-resp, kind="stable")                # peak_local_max order
    pts, s, resp = pts[o2], s[o2], resp[o2]
    lm = np.hstack([pts.astype(float), sigmas[s][:, None]])
    _, tr = ski.prune_blobs(lm, 0.5, trace=True)
    cand = _cands_from_peaks(np.column_stack([pts, s]), resp)
    # feed the GPU a SHUFFLED list: the result must not depend on input order
    perm = rng.permutation(n)
    keep = gpu.prune_within(gpu.cands_from_numpy(cand[perm]), n, sigmas, 0.5, 128, 128)
    torch.cuda.synchronize()
    got = np.zeros(n, dtype=bool)
    got[perm] = keep.cpu().numpy().astype(bool)
    assert len(tr.kill_edges) > 100
    np.testing.assert_array_equal(got, tr.keep_canonical)
    # and against scikit-image's own iteration order outside the order-dependent set
    stable = np.ones(n, dtype=bool)
    stable[tr.order_dependent] = False
    np.testing.assert_array_equal(got[stable], tr.keep_reference_order[stable])
    print(f"prune: n={n} edges={len(tr.kill_edges)} order-dependent={len(tr.order_dependent)} "
          f"differ={(got != tr.keep_reference_order).sum()}")


def test_prune_within_empty_and_single(gpu):
    sigmas = ski.sigma_list(3, 5, 10)
    from magellanmapper_b200.gpu import CAND_DTYPE
    one = np.zeros(1, dtype=CAND_DTYPE)
    keep = gpu.prune_within(gpu.cands_from_numpy(one), 1, sigmas, 0.5, 10, 10)
    assert keep.cpu().numpy().tolist() == [1]
    keep = gpu.prune_within(gpu.new_cand_buffer(4), 0, sigmas, 0.5, 10, 10)
    assert keep.numel() == 0


def test_prune_seams_matches_reference_vectors(gpu, golden_dir):
    g = np.load(os.path.join(golden_dir, "remove_close.npz"))
    for i in range(int(g["n"])):
        master, check, tol = g[f"r{i}_master"], g[f"r{i}_check"], g[f"r{i}_tol"]
        m = torch.from_numpy(master[:, :3].astype(np.int32)).cuda().contiguous()
        c = torch.from_numpy(check[:, :3].astype(np.int32)).cuda().contiguous()
        last, hit = gpu.prune_seams(m, c, tol)
        torch.cuda.synchronize()
        last, hit = last.cpu().numpy(), hit.cpu().numpy().astype(bool)
        pruned = check[~hit] if len(check) else check
        np.testing.assert_array_equal(pruned, g[f"r{i}_pruned"])
        mo = master.copy()
        sel = last >= 0
        mo[sel, 7:10] = np.around((master[sel, 7:10] + check[last[sel], 7:10]) / 2)
        np.testing.assert_array_equal(mo, g[f"r{i}_master_out"])


# --------------------------------------------------------------- preprocess

def _params(gpu, prof, near_max):
    from magellanmapper_b200._lib import MmbPreprocParams
    return MmbPreprocParams(prof.clip_vmin, prof.clip_vmax, near_max * prof.max_thresh_factor,
                            prof.clip_min, prof.clip_max, prof.unsharp_strength or 0.0,
                            prof.erosion_threshold or 0.0)


def test_preprocess_blocks_golden(gpu, golden_dir):
    """Single blocks from the reference-generated vectors (dense, thin, ragged,
    constant, zeros, bright/eroded, two-level, single voxel)."""
    g = np.load(os.path.join(golden_dir, "preprocess_blocks.npz"))
    prof = mm.Profile()
    p = _params(gpu, prof, float(g["near_max"]))
    for name in g["names"]:
        blk = g[f"{name}_in"]
        out = gpu.preprocess_blocks(gpu.as_source(blk), (25, 25, 25), p)
        torch.cuda.synchronize()
        got = out[:, :, :blk.shape[2]].cpu().numpy()
        want = g[f"{name}_out"].astype(np.float64)
        err = np.max(np.abs(got - want))
        print(f"preprocess {name}: max abs err {err:.3e}")
        assert err < 2e-5, name


@pytest.mark.parametrize("shape", [(50, 75, 100), (55, 48, 30), (12, 26, 51)])
def test_preprocess_blocks_volume(gpu, shape):
    vol, _ = synth.make_volume(shape, seed=13, density=1 / 1500.0)
    nm = synth.near_max_of(vol)
    prof = mm.Profile()
    want = mm.preprocess_blocks(vol, (25, 25, 25), prof, nm)
    out = gpu.preprocess_blocks(gpu.as_source(vol), (25, 25, 25), _params(gpu, prof, nm))
    torch.cuda.synchronize()
    got = out[:, :, :shape[2]].cpu().numpy()
    err = np.max(np.abs(got - want))
    print(f"preprocess volume {shape}: max abs err {err:.3e}")
    assert err < 2e-5


def test_preprocess_anisotropic_blocks(gpu):
    """config 5 geometry: 5x25x25 blocks"""
    vol, _ = synth.make_volume((23, 60, 60), seed=17, density=1 / 1500.0)
    nm = synth.near_max_of(vol)
    prof = mm.Profile()
    want = mm.preprocess_blocks(vol, (5, 25, 25), prof, nm)
    out = gpu.preprocess_blocks(gpu.as_source(vol), (5, 25, 25), _params(gpu, prof, nm))
    torch.cuda.synchronize()
    assert np.max(np.abs(out[:, :, :60].cpu().numpy() - want)) < 2e-5



@pytest.mark.parametrize("shape,block,dtype", [
    ((40, 90, 70), (40, 45, 35), np.uint16),      # every side above the one-CTA limit
    ((37, 50, 64), (33, 50, 64), np.uint16),      # ragged trailing layer (4 planes)
    ((20, 66, 41), (25, 33, 64), np.uint8),       # mixed: z and x clipped to the volume
    ((34, 40, 40), (34, 40, 40), np.float32),     # one block, float keys
    ((12, 70, 35), (2000, 2000, 2000), np.float64),   # `lowres`: block = chunk
    ((30, 50, 26), (25, 25, 25), np.float64),         # float64 keys at the default block size
])
def test_preprocess_large_blocks(gpu, shape, block, dtype):
    """Blocks above 32 voxels a side (global-memory path, preprocess_large.cu): radix-select
    percentiles, per-block 65-tap 'nearest' blur, block-mean erosion switch."""
    vol, _ = synth.make_volume(shape, seed=23, density=1 / 1200.0)
    nm = synth.near_max_of(vol)
    if dtype == np.uint8:
        vol = (vol >> 8).astype(np.uint8)
        nm = nm / 256.0
    elif dtype != np.uint16:
        vol = (vol / 65535.0).astype(dtype)
        nm = nm / 65535.0
    prof = mm.Profile()
    want = mm.preprocess_blocks(vol, block, prof, nm)
    out = gpu.preprocess_blocks(gpu.as_source(vol), block, _params(gpu, prof, nm))
    torch.cuda.synchronize()
    got = out[:, :, :shape[2]].cpu().numpy()
    err = np.max(np.abs(got - want))
    print(f"large-block preprocess {shape} / {block} {np.dtype(dtype).name}: max abs err {err:.3e}")
    assert err < 2e-5


def test_preprocess_large_eroded_and_degenerate(gpu):
    """A bright block (mean above erosion_threshold -> octahedron erosion inside the block
    only) next to a constant block (vmin == vmax -> passed through unstretched)."""
    rng = np.random.default_rng(5)
    vol = np.zeros((40, 40, 80), dtype=np.uint16)
    vol[:, :, :40] = rng.integers(20000, 60000, size=(40, 40, 40))
    vol[:, :, 40:] = 1234
    prof = mm.Profile()
    want = mm.preprocess_blocks(vol, (40, 40, 40), prof, 30000.0)
    out = gpu.preprocess_blocks(gpu.as_source(vol), (40, 40, 40), _params(gpu, prof, 30000.0))
    torch.cuda.synchronize()
    got = out[:, :, :80].cpu().numpy()
    assert np.mean(mm.saturate_roi(vol[:, :, :40], prof, 30000.0)) > prof.erosion_threshold
    scale = np.maximum(np.abs(want), 1.0)
    assert np.max(np.abs(got - want) / scale) < 2e-5

# ------------------------------------------------------------- fused driver

def _blob_sets(cands, sigmas):
    return {(int(c["z"]), int(c["y"]), int(c["x"]), int(c["s"])) for c in cands}


def _compare_detection(got, res, thr, label):
    """got: CAND records; res: oracle BlobLogResult.  Exact match required
    except candidates within 1e-4 of the threshold (north_star)."""
    sig = res.sigmas
    want = {(int(z), int(y), int(x), int(np.argmin(np.abs(sig - s)))) for z, y, x, s in res.blobs}
    got_set = _blob_sets(got, sig)
    near = {tuple(int(v) for v in p) for p, r in zip(res.peaks, res.responses)
            if abs(r - thr) < 1e-4}
    near |= {k for c, k in zip(got, [tuple(int(c[f]) for f in "zyxs") for c in got])
             if abs(float(c["resp"]) - thr) < 1e-4}
    od = set()
    if res.trace is not None:
        od = {tuple(int(v) for v in res.peaks[i]) for i in res.trace.order_dependent}
    diff = (got_set ^ want) - near - od
    tp = len(got_set & want)
    f1 = 2 * tp / max(len(got_set) + len(want), 1)
    print(f"{label}: gpu={len(got_set)} oracle={len(want)} F1={f1:.6f} near-thr={len(near)} "
          f"order-dependent={len(od)} unexplained={len(diff)}")
    assert not diff, f"{label}: unexplained differences {sorted(diff)[:10]}"
    assert f1 > 0.999 or len(want) < 50
    return f1


def test_detect_chunk_raw_uint16(gpu):
    """BASELINE config 1 shape of call at reduced size: detect_blobs on raw uint16."""
    shape = (50, 200, 200)
    vol, _ = synth.make_volume(shape, seed=0)
    res = ski.blob_log(vol, 3, 5, 10, 0.1, 0.5, full=True)
    det = gpu.ChunkDetector(shape)
    got, npk = det.detect(gpu.as_source(vol), res.sigmas, 0.1, 0.5, scale=1 / 65535.0)
    assert npk == len(res.peaks) or abs(npk - len(res.peaks)) <= 2
    _compare_detection(got, res, 0.1, "raw 50x200x200")


def test_detect_chunk_preprocessed(gpu):
    """StackDetector.detect_sub_roi arithmetic: 25^3 block preprocessing + detection."""
    shape = (55, 130, 105)
    vol, _ = synth.make_volume(shape, seed=4, density=1 / 3000.0)
    nm = synth.near_max_of(vol)
    prof = mm.Profile()
    pre = mm.preprocess_blocks(vol, (25, 25, 25), prof, nm)
    res = ski.blob_log(pre, 3, 5, 10, 0.1, 0.5, full=True)
    det = gpu.ChunkDetector(shape)
    got, _ = det.detect(gpu.as_source(vol), res.sigmas, 0.1, 0.5, pre=_params(gpu, prof, nm),
                        block_shape=(25, 25, 25))
    _compare_detection(got, res, 0.1, "preprocessed 55x130x105")


def test_detect_chunk_whole_chunk_block(gpu):
    """`lowres`-style profile inside the fused chunk driver: the preprocessing block is the
    whole chunk, so the large-block path runs in the sweep buffers the driver lends it."""
    shape = (45, 100, 90)
    vol, _ = synth.make_volume(shape, seed=8, density=1 / 3000.0)
    nm = synth.near_max_of(vol)
    prof = mm.Profile()
    pre = mm.preprocess_blocks(vol, (2000, 2000, 2000), prof, nm)
    res = ski.blob_log(pre, 3, 5, 10, 0.1, 0.5, full=True)
    det = gpu.ChunkDetector(shape)
    got, _ = det.detect(gpu.as_source(vol), res.sigmas, 0.1, 0.5, pre=_params(gpu, prof, nm),
                        block_shape=(2000, 2000, 2000))
    _compare_detection(got, res, 0.1, "whole-chunk block 45x100x90")


def test_detect_chunk_overflow_regrows(gpu):
    shape = (30, 64, 64)
    vol, _ = synth.make_volume(shape, seed=6, density=1 / 1500.0)
    det = gpu.ChunkDetector(shape, capacity=4)
    got, npk = det.detect(gpu.as_source(vol), ski.sigma_list(3, 5, 10), 0.1, 0.5,
                          scale=1 / 65535.0)
    assert npk > 4 and det.capacity >= npk and len(got) > 0


def test_detect_chunk_empty(gpu):
    vol = np.full((20, 40, 40), 300, dtype=np.uint16)
    det = gpu.ChunkDetector(vol.shape)
    got, npk = det.detect(gpu.as_source(vol), ski.sigma_list(3, 5, 10), 0.1, 0.5,
                          scale=1 / 65535.0)
    assert npk == 0 and len(got) == 0
