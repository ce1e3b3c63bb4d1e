"""Multi-GPU host logic: slab / chunk-row geometry, the halo exchange and the
variable-length gathers (world size 2 on gloo, CPU tensors), and - on a GPU -
the two shardings against the single-GPU results they must reproduce."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist                               # noqa: E402
import torch.multiprocessing as mp                             # noqa: E402

from magellanmapper_b200 import multi_gpu as mg               # noqa: E402
from magellanmapper_b200.cv import chunking, stack_detect     # noqa: E402
from magellanmapper_b200.settings import config, roi_prof     # noqa: E402


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


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ---- geometry ---------------------------------------------------------------------

def test_slab_bounds_partition_and_alignment():
    for Z, world, align in ((1024, 2, 25), (4096, 8, 25), (100, 8, 25), (2048, 8, 1), (7, 3, 5)):
        b = mg.slab_bounds(Z, world, align)
        assert len(b) == world and b[0][0] == 0 and b[-1][1] == Z
        for (a0, a1), (b0, b1) in zip(b[:-1], b[1:]):
            assert a1 == b0 and a0 <= a1
            assert a1 % align == 0
    assert mg.slab_bounds(1024, 2, 25) == [(0, 500), (500, 1024)]


def test_chunk_rows_balanced_and_wanted_ranges():
    """config-2 stacks piled up as one 4-rank volume: 512-plane slabs, 500-plane
    chunk pitch with 5 planes of overlap (chunking.stack_splitter)."""
    _setup()
    held = mg.slab_bounds(2048, 4)
    blocks = stack_detect.setup_blocks(config.roi_profile, (2048, 2048, 2048))
    rows, z_bounds, wanted = mg.chunk_row_plan((2048, 2048, 2048), blocks, held)
    assert z_bounds == [(0, 505), (500, 1005), (1000, 1505), (1500, 2005), (2000, 2048)]
    assert rows == [[0], [1], [2], [3, 4]]
    assert wanted == [(0, 505), (500, 1005), (1000, 1505), (1500, 2048)]
    plan = mg.transfer_plan(held, wanted)
    # halo planes only travel between slab neighbours
    assert all(abs(src - dst) == 1 for src, dst, _, _ in plan)
    assert (0, 1, 500, 512) in plan and (1, 2, 1000, 1024) in plan and (2, 3, 1500, 1536) in plan
    # the first-plane rule (kept for comparison) piles two rows on rank 0
    assert mg.assign_chunk_rows([b[0] for b in z_bounds], held) == [[0, 1], [2], [3], [4]]
    # two slabs of config 2: 505 | 505 + 24 planes instead of 1010 | 24
    b2 = [(0, 505), (500, 1005), (1000, 1024)]
    assert mg.assign_chunk_rows_balanced(b2, 2) == [[0], [1, 2]]
    # more ranks than rows: trailing ranks idle, every row assigned once
    r3 = mg.assign_chunk_rows_balanced(b2, 5)
    assert sorted(k for r in r3 for k in r) == [0, 1, 2] and max(len(r) for r in r3) == 1


def test_loan_units_even_out_the_odd_row():
    """N slabs of 512 planes make ceil(N*512/500) chunk rows; lending (row, y-column)
    units brings every rank within 1 % of the mean load, each unit keeps exactly one
    worker, and the box plan covers every plane of a lent unit exactly once."""
    for n in (2, 4, 8):
        Z = 512 * n
        zb = [(500 * k, min(500 * k + 505, Z)) for k in range(-(-Z // 500))]
        yb = [(500 * j, min(500 * j + 505, 2048)) for j in range(5)]
        held = mg.slab_bounds(Z, n)
        rows = mg.assign_chunk_rows_balanced(zb, n)
        loans = mg.loan_units(rows, zb, yb)
        assert len({(k, j) for k, j, _, _ in loans}) == len(loans)
        w = lambda k, j: (zb[k][1] - zb[k][0]) * (yb[j][1] - yb[j][0])
        load = [sum(w(k, j) for k in rows[r] for j in range(5)) for r in range(n)]
        before = max(load) / (sum(load) / n)
        for k, j, owner, worker in loans:
            assert k in rows[owner] and owner != worker
            load[owner] -= w(k, j)
            load[worker] += w(k, j)
        assert max(load) / (sum(load) / n) < 1.01 < before
        plan = mg.box_transfer_plan(loans, zb, yb, held)
        for k, j, _, worker in loans:
            pieces = sorted((z0, z1) for s, d, kk, jj, z0, z1 in plan if (kk, jj, d) == (k, j, worker))
            assert pieces[0][0] == zb[k][0] and pieces[-1][1] == zb[k][1]
            assert all(a[1] == b[0] for a, b in zip(pieces[:-1], pieces[1:]))
    assert mg.loan_units([[0, 1, 2]], [(0, 505), (500, 1005), (1000, 1024)], [(0, 505)]) == []


def test_seamless_plan_halo_is_whole_block_layers():
    own, ext = mg.seamless_plan(1024, 2, 25, 21)
    assert own == [(0, 500), (500, 1024)] and ext == [(0, 525), (475, 1024)]
    own, ext = mg.seamless_plan(2048, 8, 25, 21)
    for (a, b), (ea, eb) in zip(own, ext):
        assert ea % 25 == 0 and (eb % 25 == 0 or eb == 2048)
        assert (a - ea >= 21 or ea == 0) and (eb - b >= 21 or eb == 2048)


def test_pack_unpack_tables_roundtrip():
    rng = np.random.default_rng(5)
    grid = (2, 3, 2)
    seg = np.empty(grid, dtype=object)
    for c in np.ndindex(*grid):
        n = int(rng.integers(0, 4))
        seg[c] = rng.random((n, 11)) if n else None
    packed = chunking.merge_blobs(seg)
    half = len(packed) // 2
    # a cut in the middle of one chunk's rows (as if two ranks held the halves)
    back = mg.unpack_tables([packed[:0], packed], grid)
    for c in np.ndindex(*grid):
        if seg[c] is None:
            assert back[c] is None
        else:
            np.testing.assert_array_equal(back[c], seg[c])
    assert half >= 0


# ---- collectives on gloo, world size 2 -----------------------------------------------

def _gloo_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        vol = torch.arange(40 * 3 * 4, dtype=torch.int16).reshape(40, 3, 4)
        held = [(0, 17), (17, 40)]
        wanted = [(0, 25), (10, 40)]
        ext = mg.exchange_planes(vol[held[rank][0]:held[rank][1]].clone(), held, wanted)
        assert torch.equal(ext, vol[wanted[rank][0]:wanted[rank][1]])
        # a rank that wants nothing still serves its planes
        wanted2 = [(0, 40), (20, 20)]
        ext2 = mg.exchange_planes(vol[held[rank][0]:held[rank][1]].clone(), held, wanted2)
        assert torch.equal(ext2, vol[wanted2[rank][0]:wanted2[rank][1]])
        # host-resident slab: only the halo planes are exchanged
        host = vol.numpy().view(np.uint16)
        pre, own, suf = mg.exchange_edges(host[held[rank][0]:held[rank][1]], held, wanted,
                                          device=torch.device("cpu"))
        w0, w1 = wanted[rank]
        whole = np.concatenate([pre.numpy().view(np.uint16), own, suf.numpy().view(np.uint16)])
        np.testing.assert_array_equal(whole, host[w0:w1])
        assert own.base is not None and len(own) == min(w1, held[rank][1]) - max(w0, held[rank][0])
        # lent units travel as dense sub-boxes (rows of chunk row k restricted to the y
        # range of chunk column j), pieces from every slab that holds part of the row
        zb = [(0, 22), (20, 40)]
        yb = [(0, 2), (1, 3)]
        loans = [(1, 1, 1, 0), (0, 0, 0, 1)]        # (k, j, owner, worker)
        boxes = mg._exchange_boxes(vol[held[rank][0]:held[rank][1]].clone(), held, loans, zb, yb,
                                   None, torch.device("cpu"))
        mine = {(k, j) for k, j, _, w in loans if w == rank}
        assert set(boxes) == mine
        for (k, j), box in boxes.items():
            assert torch.equal(box, vol[zb[k][0]:zb[k][1], yb[j][0]:yb[j][1]])
        # variable-length gathers (float64 tables and int32 candidate records)
        mine = np.full((3 + 2 * rank, 14), float(rank)) + np.arange(14)
        parts = mg.gather_rows(mine if rank == 0 else mine, 14)
        cand = (np.arange(5 * (rank + 1), dtype=np.int32) + 100 * rank).reshape(-1, 5)
        cparts = mg.gather_rows(cand if rank == 1 else None, 5, dtype=np.int32)
        if rank == 0:
            assert [p.shape for p in parts] == [(3, 14), (5, 14)]
            np.testing.assert_array_equal(parts[1], np.full((5, 14), 1.0) + np.arange(14))
            assert cparts[0].shape == (0, 5) and cparts[1].dtype == np.int32
            np.testing.assert_array_equal(cparts[1], cand * 0 + (np.arange(10) + 100).reshape(2, 5))
        else:
            assert parts is None and cparts is None
        assert mg._agree_max(7 + rank) == 8
        # device-table gather: one batch of point-to-point transfers, int32 mmb_row records
        rows = (torch.arange(8 * (2 + rank), dtype=torch.int32) + 1000 * rank).reshape(-1, 8)
        tparts = mg.gather_tensor_rows(rows if rank == 1 else None, 8, dtype=torch.int32,
                                       device=torch.device("cpu"))
        if rank == 0:
            assert tparts[0].shape == (0, 8) and tparts[1].shape == (3, 8)
            assert torch.equal(tparts[1], (torch.arange(24, dtype=torch.int32) + 1000).reshape(3, 8))
        else:
            assert tparts is None
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def test_exchange_and_gather_gloo_world2(tmp_path):
    port = _free_port()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


# ---- GPU: both shardings against the single-GPU results ----------------------------------

@pytest.mark.gpu
def test_slab_driver_world1_equals_reference_vectors(golden_dir, tmp_path):
    from magellanmapper_b200 import gpu
    gpu.require_cuda()
    g = np.load(os.path.join(golden_dir, "stack_small.npz"))
    _setup(near_max=float(g["near_max"]), segment_size=50)
    config.filename = str(tmp_path / "slab")
    os.chdir(tmp_path)
    vol = torch.from_numpy(g["vol"].view(np.int16)).cuda()
    _, _, blobs = mg.detect_blobs_blocks_slabs(config.filename, vol, [(0, vol.shape[0])],
                                               vol.shape)
    np.testing.assert_array_equal(blobs.blobs, g["plain_blobs"])
    # host slab: streamed through gpu.StripFeeder
    _, _, blobs = mg.detect_blobs_blocks_slabs(config.filename, np.ascontiguousarray(g["vol"]),
                                               [(0, vol.shape[0])], vol.shape)
    np.testing.assert_array_equal(blobs.blobs, g["plain_blobs"])
    stack_detect.StackDetector.release_workspace()


@pytest.mark.gpu
def test_slab_driver_chunk_rows_on_shifted_slabs(golden_dir, tmp_path):
    """Each 'rank' of a two-slab split run in turn on one GPU (planes handed over
    by slicing instead of send/recv): the union of the per-rank chunk tables,
    seam-pruned, equals the reference's output."""
    from magellanmapper_b200 import gpu
    from magellanmapper_b200.cv import detector
    gpu.require_cuda()
    g = np.load(os.path.join(golden_dir, "stack_small.npz"))
    _setup(near_max=float(g["near_max"]), segment_size=50)
    vol = torch.from_numpy(g["vol"].view(np.int16)).cuda()
    Z = vol.shape[0]
    held = mg.slab_bounds(Z, 2)
    blocks = stack_detect.setup_blocks(config.roi_profile, tuple(vol.shape))
    rows, z_bounds, wanted = mg.chunk_row_plan(vol.shape, blocks, held)
    assert sorted(rows[0] + rows[1]) == list(range(blocks.sub_roi_slices.shape[0]))
    grid = blocks.sub_roi_slices.shape
    parts = []
    for rank in range(2):
        w0, w1 = wanted[rank]
        ext = vol[w0:w1]
        local = np.empty(grid, dtype=object)
        coords = []
        for c in np.ndindex(*grid):
            sz, sy, sx = blocks.sub_roi_slices[c]
            if c[0] in rows[rank]:
                local[c] = (slice(sz.start - w0, sz.stop - w0), sy, sx)
                coords.append(c)
            else:
                local[c] = (slice(0, 0),) * 3
        seg = stack_detect.StackDetector.detect_blobs_sub_rois(
            None, ext, local, blocks.sub_rois_offsets, blocks.denoise_max_shape,
            blocks.exclude_border, False, [0], coords=coords)
        parts.append(mg.pack_tables(seg))
    seg_all = mg.unpack_tables(parts, grid)
    table, _ = stack_detect.StackPruner.prune_blobs_mp(
        None, seg_all, blocks.overlap, blocks.tol, blocks.sub_roi_slices,
        blocks.sub_rois_offsets, [0], blocks.overlap_padding)
    blobs = detector.Blobs(table)
    blobs.replace_rel_with_abs_blob_coords(table)
    blobs.blobs = table
    np.testing.assert_array_equal(blobs.remove_abs_blob_coords(True), g["plain_blobs"])
    stack_detect.StackDetector.release_workspace()


@pytest.mark.gpu
@pytest.mark.parametrize("tile_yx", [None, (50, 75)])
def test_seamless_slabs_equal_one_chunk(tile_yx):
    """Three z-slabs with block-aligned halos (and optionally y/x tiles inside each
    slab), candidates pooled and pruned once == the whole volume detected as a
    single chunk: identical blob table, row for row."""
    from magellanmapper_b200 import gpu, synth
    from magellanmapper_b200.cv import detector
    gpu.require_cuda()
    shape = (160, 140, 150)
    vol, _ = synth.make_volume(shape, seed=77, density=1 / 2500.0)
    nm = synth.near_max_of(vol)
    _setup(near_max=nm)
    settings, pre, sigmas, halo, bd = mg._seamless_setup(shape, 0)
    assert halo == 21 and bd == (25, 25, 25)
    dev = torch.from_numpy(vol.view(np.int16)).cuda()
    det = gpu.ChunkDetector(shape)
    whole, _ = det.detect(gpu.as_source(dev), sigmas, settings["detection_threshold"],
                          settings["overlap"], pre=pre, block_shape=bd)
    want = detector.cands_to_blobs(whole, sigmas, shape[1:], 0)
    own, ext = mg.seamless_plan(shape[0], 3, bd[0], halo)
    cands = [mg.seamless_candidates(dev[e0:e1], (e0, e1), o, shape, 0, tile_yx)
             for o, (e0, e1) in zip(own, ext)]
    assert all(c.is_cuda and c.dtype == torch.int32 for c in cands)
    got = mg.prune_global(torch.cat(cands), sigmas, settings["overlap"], shape, 0)
    assert len(want) > 50
    np.testing.assert_array_equal(got, want)


def _nccl_worker(rank, world, port, out_dir, golden_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from magellanmapper_b200 import gpu, synth
        from magellanmapper_b200.cv import detector
        g = np.load(os.path.join(golden_dir, "stack_small.npz"))
        _setup(near_max=float(g["near_max"]), segment_size=50)
        os.chdir(out_dir)
        vol = torch.from_numpy(g["vol"].view(np.int16))
        held = mg.slab_bounds(vol.shape[0], world)
        slab = vol[held[rank][0]:held[rank][1]].cuda()
        _, _, blobs = mg.detect_blobs_blocks_slabs(os.path.join(out_dir, "nccl"), slab, held,
                                                   vol.shape)
        # the same from a HOST slab (streamed strip by strip, halo planes exchanged first)
        host_slab = np.ascontiguousarray(g["vol"][held[rank][0]:held[rank][1]])
        _, _, blobs_h = mg.detect_blobs_blocks_slabs(os.path.join(out_dir, "nccl_h"), host_slab,
                                                     held, vol.shape)
        if rank == 0:
            np.testing.assert_array_equal(blobs.blobs, g["plain_blobs"])
            np.testing.assert_array_equal(blobs_h.blobs, g["plain_blobs"])
        else:
            assert blobs is None and blobs_h is None
        stack_detect.StackDetector.release_workspace()
        # a taller stack with the default 500-voxel chunks scaled down so that units are
        # lent between ranks (several chunk rows, y columns and a thin trailing row)
        shape = (150, 120, 110)
        big = synth.device_volume(shape, 91, device=torch.device("cuda", rank))
        nm = float(np.percentile(big.cpu().numpy().view(np.uint16), 99.5))
        _setup(near_max=nm, segment_size=40)
        held = mg.slab_bounds(shape[0], world)
        _, _, b_multi = mg.detect_blobs_blocks_slabs(
            os.path.join(out_dir, "lend"), big[held[rank][0]:held[rank][1]].contiguous(), held,
            shape)
        if rank == 0:
            from magellanmapper_b200.io import np_io
            _, _, b_one = stack_detect.detect_blobs_blocks(
                os.path.join(out_dir, "one"), np_io.Image5d(big[None]), None, None, [0], False,
                False, True)
            assert len(b_one.blobs) > 100
            np.testing.assert_array_equal(b_multi.blobs, b_one.blobs)
        stack_detect.StackDetector.release_workspace()
        # seamless z-slabs over NCCL == the whole volume as one chunk on one GPU
        _setup(near_max=nm)
        table = mg.detect_seamless(big[held[rank][0]:held[rank][1]].contiguous(), held, shape, 0,
                                   tile_yx=(50, 75) if world > 2 else None)
        if rank == 0:
            settings, pre, sigmas, halo, bd = mg._seamless_setup(shape, 0)
            det = gpu.ChunkDetector(shape)
            whole, _ = det.detect(gpu.as_source(big), sigmas, settings["detection_threshold"],
                                  settings["overlap"], pre=pre, block_shape=bd)
            want = detector.cands_to_blobs(whole, sigmas, shape[1:], 0)
            assert len(want) > 50
            np.testing.assert_array_equal(table, want)
        else:
            assert table is None
        open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_slab_drivers_nccl(golden_dir, tmp_path, world):
    """Both shardings on N real ranks over NCCL (halo exchange, lent units, batched
    row gather, rank-0 table kernels) against the single-GPU results; runs whenever the
    box has N GPUs."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    mp.spawn(_nccl_worker, args=(world, port, str(tmp_path), golden_dir), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
