#!/usr/bin/env python
"""Benchmark of the blob-detection hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config 2|3] [--mode chunks|seamless]

One step = one pass of ``detect_blobs_blocks`` (chunked preprocessing, 10-scale
LoG, 4-D local maxima, overlap pruning, seam pruning) over one synthetic
cleared-tissue stack.

* ``--config 2`` (default; BASELINE ``configs[1]``): 512x2048x2048 uint16, ``roi_blobs``
  profile, 1 um isotropic, 50 chunks of 500^3 (+5 overlap).  N > 1 runs one such stack
  per GPU as the z-slabs of ONE N*512-plane volume (weak scaling,
  ``multi_gpu.detect_blobs_blocks_slabs``).
* ``--config 3`` (BASELINE ``configs[2]``): ONE 2048x8192x8192 uint16 volume sharded as
  N z-slabs (``--shape`` shrinks it for fewer GPUs), chunk-faithful
  (``--mode chunks``) or as one seamless chunk (``--mode seamless``,
  ``multi_gpu.detect_seamless``).

The volume is a pure function of (seed, z, y, x) generated on the device
(``synth.device_volume``), so every rank's slab, the single-GPU re-run and the boxes the
oracle recomputes are the same voxels.

Prints ONE JSON line (rank 0).  ``value`` = GVoxel/s with the stack resident in
HBM; ``e2e`` = the same through the public API from pinned HOST memory,
host->device copy and result read-back inside the timed region; ``config.parity`` =
the blob table of the TIMED stack against the CPU oracle on sub-boxes (F1, unexplained
differences) and, for N > 1, against the single-GPU table of the same volume.
``--impl reference`` times the CPU oracle (the reference algorithm restated on
scipy, multiprocessing pool over chunks) on a bounded sample instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPES = {2: (512, 2048, 2048), 3: (2048, 8192, 8192)}
RESOLUTION = (1.0, 1.0, 1.0)
SEED = 1
METRIC = "blob_detection_throughput"
UNIT = "GVoxel/s"
CORE = 96          # edge of the boxes the oracle recomputes for ``config.parity``


def near_max_device(vol_i16):
    """The importer's ``near_max`` metadata: max over z-planes of the per-plane 99.5th
    percentile (magmap/io/importer.py:571-583), computed on the device by
    ``mmb_percentiles`` (every plane, exact)."""
    from magellanmapper_b200.io import importer
    _, near_maxs = importer.calc_near_bounds(vol_i16)
    return float(np.ravel(near_maxs)[0])


# ----------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------
# profile / config plumbing
# ----------------------------------------------------------------------------

def setup_config(near_max, filename, resolution=RESOLUTION, channel_mods=None):
    """``roi_blobs`` for every channel (``channel_mods``: one dict of profile overrides per
    channel), resolution and the importer's ``near_max`` metadata."""
    from magellanmapper_b200.settings import config, roi_prof
    near_maxs = list(near_max) if isinstance(near_max, (list, tuple)) else [near_max]
    profs = []
    for c in range(len(near_maxs)):
        prof = roi_prof.ROIProfile()
        prof.add_profiles("roi_blobs.yaml")
        for k, v in ((channel_mods or [{}] * len(near_maxs))[c]).items():
            prof[k] = v
        profs.append(prof)
    config.roi_profile = profs[0]
    config.roi_profiles = profs
    config.resolutions = [list(resolution)]
    config.near_max = near_maxs
    config.channel = None
    config.filename = filename
    return profs[0]


def peaks_json():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


KINDS = ["to_float", "preprocess", "log_x", "log_y", "log_z", "localmax", "prune_edges",
         "prune_resolve", "compact", "seam_match", "log_xy"]
#: algorithmic bytes per voxel and launch: float32 volumes
#: read + written, uint16 input for the preprocessing
ALG_BYTES = {"preprocess": 6.0, "log_x": 12.0, "log_y": 16.0, "log_z": 12.0, "localmax": 4.0,
             "log_xy": 12.0}


def collect_profile(lib):
    import ctypes as C
    ms = (C.c_double * 16)(); cnt = (C.c_int64 * 16)(); units = (C.c_double * 16)()
    lib.mmb_profile_collect(ms, cnt, units)
    lib.mmb_profile_enable(0)
    return ms, cnt, units


def roofline_block(ms, cnt, units, sm_mhz=None):
    """Roofline of the dominant kernel (the y sweep: 8 B in + 8 B out per voxel) from the
    CUDA events the library records around every launch on the launching stream."""
    peaks, peak_src = peaks_json()
    kinds, alg_bytes = KINDS, ALG_BYTES
    per_kind = {}
    total_ms = sum(ms)
    for i, k in enumerate(kinds):
        if cnt[i]:
            per_kind[k] = {"ms": ms[i], "launches": int(cnt[i]), "share": ms[i] / total_ms,
                           "gbps": (alg_bytes[k] * units[i] / (ms[i] * 1e-3) / 1e9
                                    if k in alg_bytes else None)}
    dom = max((k for k in per_kind if k in alg_bytes), key=lambda k: per_kind[k]["ms"])
    # DRAM traffic of the dominant kernel from the committed ncu capture (bytes per
    # voxel measured on one 505^3 chunk), scaled to this run's average launch size
    traffic, traffic_src = None, None
    for name in ("r02_traffic.json", "r01_traffic_v10.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(tpath):
            continue
        with open(tpath) as f:
            tj = json.load(f)
        if dom in tj["kernels"]:
            vox_per_launch = units[kinds.index(dom)] / per_kind[dom]["launches"]
            traffic = tj["kernels"][dom]["dram_bytes_per_voxel"] * vox_per_launch
            traffic_src = (f"dram__bytes_read.sum + dram__bytes_write.sum per voxel of a 505^3 "
                           f"chunk from profiles/{name} x this run's voxels per launch (the "
                           f"same file holds a 12x505x505 thin chunk, whose output stays in L2)")
            break
    # the fused x -> y sweep is bound by the FP32 pipe, not by HBM: report its pipe fraction
    # next to the HBM figure the contract asks for, and the HBM-bound sweep (z) as well
    fp32 = None
    if "log_xy" in per_kind:
        clock = (sm_mhz or 1965.0) * 1e6
        lane_fma = 165.0          # (10 r + 5) per voxel, mean over the sigma 3..5 ladder (r = 12..20)
        k = per_kind["log_xy"]
        rate = lane_fma * units[kinds.index("log_xy")] / (k["ms"] * 1e-3)
        fp32 = {"kernel": "log_xy", "lane_fma_per_voxel": lane_fma,
                "achieved_tflops": 2 * rate / 1e12,
                "peak_tflops": 2 * 148 * 128 * clock / 1e12,
                "frac": rate / (148 * 128 * clock),
                "note": "direct-form arithmetic (10 r + 5 FMAs per voxel) over the pipe peak of "
                        "128 lane-FMAs per clock per SM at the sampled SM clock; the x taps "
                        "run in the symmetric-pair form, so fewer pipe slots are issued"}
    hbm_kernel = None
    if "log_z" in per_kind:
        hbm_kernel = {"kernel": "log_z", "achieved": per_kind["log_z"]["gbps"],
                      "frac": per_kind["log_z"]["gbps"] / peaks["hbm_gbs"]}
    return {"bound": "hbm", "kernel": dom, "achieved": per_kind[dom]["gbps"],
            "fp32_pipe": fp32, "largest_hbm_bound_kernel": hbm_kernel,
            "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": per_kind[dom]["gbps"] / peaks["hbm_gbs"], "traffic": traffic,
            "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": alg_bytes[dom] * units[kinds.index(dom)]
            / per_kind[dom]["launches"],
            "peak_source": f"{peak_src} copy bandwidth (MEASURED_PEAKS.json hbm_gbs)",
            "avg_launch_ms": per_kind[dom]["ms"] / per_kind[dom]["launches"],
            "algorithmic_bytes_per_voxel": alg_bytes[dom],
            "units": "true voxels (X, not the padded row pitch) of rank 0's launches",
            "fp32_colimit": "one LoG scale costs (14 r + 7) fp32 FMAs per voxel (r = 12..20) as "
                            "packed FFMA2; at the 128 lane-FMA/clk/SM pipe peak that alone takes "
                            "as long as moving the 40 B/voxel of three separate sweeps at the "
                            "measured HBM peak, so x and y run fused (12 B/voxel, FP32-bound: see "
                            "fp32_pipe) and z stays the HBM-side sweep (DESIGN.md section 4)",
            "per_kernel": per_kind}


# ----------------------------------------------------------------------------
# CPU baseline (oracle = the reference algorithm restated on scipy)
# ----------------------------------------------------------------------------

def cpu_sample_shape(cores, shape, seconds=20.0):
    """The first P planes of the stack (up to 2048 x 2048 of each), P sized for about
    ``seconds`` (20) s of CPU work at ~0.7 MVoxel/s/core (measured on the GPU box's host).  The reference's own chunking
    (segment_size 500, overlap 5) then cuts it into up to 25 chunks of P x 505 x 505."""
    y, x = min(shape[1], 2048), min(shape[2], 2048)
    target = 0.7e6 * cores * seconds
    p = int(max(16, min(shape[0], round(target / (y * x)))))
    return (p, y, x)


def run_cpu_baseline(sample, near_max, cores, prof=None, resolution=RESOLUTION):
    from oracle import magmap_restated as mm
    prof = prof or mm.Profile()               # roi_blobs: 500-voxel chunks, 5 overlap
    t0 = time.perf_counter()
    blobs = mm.detect_blobs_blocks(sample, prof, resolution, near_max, processes=cores)
    dt = time.perf_counter() - t0
    blocks = mm.setup_blocks(prof, sample.shape, resolution)
    return sample.size / dt / 1e9, dt, 0 if blobs is None else len(blobs), \
        int(np.prod(blocks.sub_roi_slices.shape))


# ----------------------------------------------------------------------------
# parity of the timed stack against the oracle (outside every timed region)
# ----------------------------------------------------------------------------

def _fetcher(device, seed=SEED):
    """Boxes of the timed volume regenerated on the device (a pure function of the
    global coordinates) and copied to the host."""
    from magellanmapper_b200 import synth

    def fetch(z0, z1, y0, y1, x0, x1):
        box = synth.device_volume((z1 - z0, y1 - y0, x1 - x0), seed, offset=(z0, y0, x0),
                                  device=device)
        return box.cpu().numpy().view(np.uint16)
    return fetch


def parity_single_gpu(vol, shape, near_max, final_blobs, device, cores):
    """Config 2 on one GPU: oracle recomputation of five cores (volume corner, chunk
    interior, a chunk's high faces next to three seams, two boxes in the thin trailing
    chunks) plus the oracle's seam pruning of the GPU's per-chunk tables at full size
    against the GPU's final table (``oracle.subbox_check.check_stack``)."""
    from oracle import magmap_restated as mm
    from oracle import subbox_check as sb
    from magellanmapper_b200.cv import stack_detect
    from magellanmapper_b200.settings import config
    settings = config.get_roi_profile(0)
    blocks = stack_detect.setup_blocks(settings, shape)
    seg_rois = stack_detect.StackDetector.detect_blobs_sub_rois(
        None, vol, blocks.sub_roi_slices, blocks.sub_rois_offsets, blocks.denoise_max_shape,
        blocks.exclude_border, False, [0])
    out = sb.check_stack(_fetcher(device), shape, mm.Profile(), RESOLUTION, near_max, seg_rois,
                         final_blobs, processes=min(cores, 5), core_size=CORE)
    out.pop("box_results", None)
    return out


def parity_cores_global(final_rows, gshape, near_max, device, cores, seamless, prof=None,
                        resolution=RESOLUTION, seed=SEED, size=CORE, max_boxes=5):
    """Configs sharded over ranks: oracle recomputation of cores of the WHOLE volume
    against rank 0's final table.  Seamless: the volume is one chunk, any box will do.
    Chunk-faithful: cores at least 12 voxels from every inner chunk face, where the final
    table is the chunk's own table (seam pruning only touches the overlap zones)."""
    from oracle import magmap_restated as mm
    from oracle import subbox_check as sb
    RESOLUTION = resolution
    prof = prof or mm.Profile()
    fetch = _fetcher(device, seed)
    blocks = mm.setup_blocks(prof, gshape, RESOLUTION)
    grid = blocks.sub_roi_slices.shape
    jobs = []
    if seamless:
        Z, Y, X = gshape
        places = [((0, size), (0, size), (0, size)),
                  tuple((n // 2 - size // 2, n // 2 - size // 2 + size) for n in gshape),
                  ((Z - size, Z), (Y // 3, Y // 3 + size), (X - size, X)),
                  ((Z // 4, Z // 4 + size), (Y - size, Y), (X // 5, X // 5 + size))]
        for core in places:
            core = tuple((max(0, a), min(n, b)) for (a, b), n in zip(core, gshape))
            region = sb.plan_region(core, gshape, prof, RESOLUTION, blocks.denoise_max_shape)
            raw = fetch(*[v for r in region for v in r])
            rows = _rows_in(final_rows, region, (0, 0, 0))
            jobs.append((raw, gshape, core, rows, prof, RESOLUTION, near_max,
                         blocks.denoise_max_shape))
    else:
        last = tuple(g - 1 for g in grid)
        picks = [((0, 0, 0), ("lo", "lo", "lo")), ((0, 0, 0), ("mid", "mid", "mid")),
                 (tuple(min(1, g - 1) for g in grid), ("mid", "lo+", "mid")),
                 (last, ("hi", "hi", "hi")),
                 ((grid[0] // 2, last[1], grid[2] // 2), ("mid", "hi", "mid"))]
        margin = 12
        for coord, place in picks[:max_boxes]:
            sl = blocks.sub_roi_slices[coord]
            o = [s.start for s in sl]
            cs = tuple(s.stop - s.start for s in sl)
            core = []
            for ax, (n, where) in enumerate(zip(cs, place)):
                inner_lo = coord[ax] > 0                 # low face is a seam
                inner_hi = coord[ax] < grid[ax] - 1      # high face is a seam
                a_min = margin if inner_lo else 0
                b_max = n - margin if inner_hi else n
                w = min(size, max(1, b_max - a_min))
                lo = {"lo": a_min, "lo+": a_min, "mid": max(a_min, (n - w) // 2),
                      "hi": b_max - w}[where]
                core.append((lo, lo + w))
            region = sb.plan_region(core, cs, prof, RESOLUTION, blocks.denoise_max_shape)
            raw = fetch(o[0] + region[0][0], o[0] + region[0][1], o[1] + region[1][0],
                        o[1] + region[1][1], o[2] + region[2][0], o[2] + region[2][1])
            rows = _rows_in(final_rows, [(a + oo, b + oo) for (a, b), oo in zip(region, o)], o)
            jobs.append((raw, cs, tuple(core), rows, prof, RESOLUTION, near_max,
                         blocks.denoise_max_shape))
    results = sb.check_cores(jobs, processes=min(cores, len(jobs)))
    return sb.summarize(results)


def _rows_in(final_rows, box, origin):
    """Rows ``z, y, x, radius`` of a final table inside the absolute ``box``, shifted to
    be relative to ``origin``."""
    if final_rows is None or len(final_rows) == 0:
        return None
    t = np.asarray(final_rows)
    m = np.ones(len(t), dtype=bool)
    for ax, (a, b) in enumerate(box):
        m &= (t[:, ax] >= a) & (t[:, ax] < b)
    rows = np.array(t[m][:, :4], dtype=np.float64)
    rows[:, :3] -= np.asarray(origin, dtype=np.float64)
    return rows



# ----------------------------------------------------------------------------
# BASELINE configs 1, 4 and 5 (one GPU each)
# ----------------------------------------------------------------------------

NAMED = {
    1: dict(shape=(50, 500, 500), resolution=(1.0, 1.0, 1.0), mods=[{}],
            what="BASELINE config 1: synthetic 4x nuclei ROI uint16 zyx {shape}, single "
                 "channel, roi_blobs profile, detector.detect_blobs (img_as_float, 10-scale "
                 "LoG sigma 3..5, 4-D local maxima, overlap pruning) on the raw ROI"),
    4: dict(shape=(1024, 4096, 4096), resolution=(1.0, 1.0, 1.0),
            mods=[{}, {"min_sigma_factor": 4, "max_sigma_factor": 10}],
            what="BASELINE config 4: two-channel synthetic volume {shape} uint16 "
                 "(channel-last), per-channel profiles (sigma 3..5 and 4..10, num_sigma 10), "
                 "both channels detected in ONE detect_blobs_stack pass; GVoxel counts every "
                 "channel's voxels"),
    5: dict(shape=(1024, 4096, 4096), resolution=(5.0, 1.0, 1.0), mods=[{}],
            what="BASELINE config 5: scale-anisotropic volume {shape} uint16 at 5x1x1 um "
                 "(chunks 101x505x505, preprocessing blocks 5x25x25, overlap 1x5x5), "
                 "detect_blobs_stack with saturate + denoise preprocessing fused"),
}


def main_named(args, shape, cores):
    """Configs 1, 4, 5 on one GPU, same JSON contract as the default run."""
    import torch
    from magellanmapper_b200 import gpu, _lib, synth
    from magellanmapper_b200.cv import detector, stack_detect
    from magellanmapper_b200.io import np_io
    from oracle import magmap_restated as mm
    cfg = NAMED[args.config]
    res = cfg["resolution"]
    n_chl = len(cfg["mods"])
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    lib = _lib.load()
    gpu.require_cuda()
    tmp = tempfile.mkdtemp(prefix="mmb_bench_")
    os.chdir(tmp)
    seeds = [SEED + 100 * c for c in range(n_chl)]
    if n_chl == 1:
        vol = synth.device_volume(shape, seeds[0], device=device)
        near_maxs = [near_max_device(vol)]
    else:
        vol = torch.empty(tuple(shape) + (n_chl,), dtype=torch.int16, device=device)
        near_maxs = [0.0] * n_chl
        for c in range(n_chl):
            for z0 in range(0, shape[0], 64):
                z1 = min(shape[0], z0 + 64)
                part = synth.device_volume((z1 - z0, shape[1], shape[2]), seeds[c],
                                           offset=(z0, 0, 0), device=device)
                near_maxs[c] = max(near_maxs[c], near_max_device(part))
                vol[z0:z1, :, :, c] = part
            del part
    setup_config(near_maxs, os.path.join(tmp, "bench"), res, cfg["mods"])
    nvox = float(np.prod(shape)) * n_chl
    chls = list(range(n_chl))

    def run(img, base):
        if args.config == 1:
            return detector.detect_blobs(img, chls)
        if isinstance(img, np.ndarray):
            i5 = np_io.Image5d(img[None])
            i5.is_roi = True
            _, _, b = stack_detect.detect_blobs_stack(os.path.join(tmp, base), i5)
        else:
            _, _, b = stack_detect.detect_blobs_blocks(
                os.path.join(tmp, base), np_io.Image5d(img[None]), None, None, chls, False,
                False, True)
        return None if b is None else b.blobs

    # the named configurations are measured on one stream (their committed lines under
    # profiles/ were), so that every event duration is of a kernel alone on the device
    stack_detect.THIN_CHUNK_FRACTION = 0.0
    lib.mmb_profile_enable(1)
    for _ in range(args.warmup):
        final = run(vol, "res")
    lib.mmb_profile_enable(0)

    parity = None
    if not args.skip_parity:
        t_par = time.perf_counter()
        parts = []
        for c in chls:
            mods = {k: v for k, v in cfg["mods"][c].items()}
            prof = mm.Profile(**mods)
            rows = None if final is None else final[final[:, 6] == c]
            wide = prof.max_sigma_factor > 6          # wide ladders need wide margins
            if args.config == 1:
                from oracle import subbox_check as sb
                cores_ = [((0, shape[0]), (0, CORE), (0, CORE)),
                          ((0, shape[0]), (shape[1] // 2, shape[1] // 2 + CORE),
                           (shape[2] - CORE, shape[2]))]
                fetch = _fetcher(device, seeds[c])
                jobs = []
                for core in cores_:
                    region = sb.plan_region(core, shape, prof, res, None)
                    # the ROI is one chunk: table rows are already chunk-relative
                    jobs.append((fetch(*[v for r in region for v in r]), shape, core,
                                 None if rows is None else rows[:, :4], prof, res, near_maxs[c],
                                 None))
                parts.append(sb.summarize(sb.check_cores(jobs, processes=min(cores, 2))))
            else:
                parts.append(parity_cores_global(
                    rows, shape, near_maxs[c], device, cores, False, prof, res, seeds[c],
                    48 if wide else CORE, 3 if wide else 5))
        parity = {k: (sum(p[k] for p in parts) if k != "f1" else None) for k in parts[0]}
        tp2 = sum(p["f1"] * (p["gpu_blobs"] + p["oracle_blobs"]) for p in parts)
        parity["f1"] = tp2 / max(parity["gpu_blobs"] + parity["oracle_blobs"], 1)
        parity["seconds"] = round(time.perf_counter() - t_par, 1)

    sampler = ClockSampler(0)
    sampler.start()
    launches0 = lib.mmb_launch_count()
    lib.mmb_profile_enable(1)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        final = run(vol, "res")
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t_step = max(wall, ev0.elapsed_time(ev1) / 1e3)
    launches = lib.mmb_launch_count() - launches0
    ms, cnt, units = collect_profile(lib)
    clocks = sampler.stop()
    value = nvox * args.steps / t_step / 1e9

    e2e = None
    nbytes = int(nvox) * 2
    if not args.skip_e2e:
        import psutil
        if nbytes * 1.5 < psutil.virtual_memory().available:
            host = torch.empty(vol.shape, dtype=torch.int16).pin_memory()
            host.copy_(vol)
            torch.cuda.synchronize()
            host_np = host.numpy().view(np.uint16)
            b = run(host_np, "e2e")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                b = run(host_np, "e2e")
            torch.cuda.synchronize()
            t_e2e = time.perf_counter() - t0
            nb = 0 if b is None else int(b.shape[0])
            e2e = {"value": nvox * args.steps / t_e2e / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nb * b.shape[1] * 8
                   if nb else 0,
                   "table_equals_resident": bool(b is not None and final is not None
                                                 and np.array_equal(b, final))}
            del host, host_np

    roof = roofline_block(ms, cnt, units, clocks.get("sm_mhz"))
    cpu = None
    if not args.skip_cpu:
        if args.config == 1:
            roi = vol.cpu().numpy().view(np.uint16)
            t0 = time.perf_counter()
            tab = mm.detect_blobs(roi, mm.Profile(), res)
            dt = time.perf_counter() - t0
            cpu = {"value": roi.size / dt / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"the whole timed ROI, oracle detect_blobs in one process (the "
                             f"reference's detect_blobs is single-process), {dt:.1f} s, "
                             f"{0 if tab is None else len(tab)} blobs"}
        else:
            sshape = cpu_sample_shape(cores, shape)
            if n_chl > 1:
                sshape = (max(8, sshape[0] // (2 * n_chl)),) + sshape[1:]   # wide ladders cost more
            tot_t, tot_v, notes = 0.0, 0.0, []
            for c in chls:
                sub = vol[:sshape[0], :sshape[1], :sshape[2]]
                sub = (sub[..., c] if n_chl > 1 else sub).contiguous().cpu().numpy().view(np.uint16)
                gv, dt, nb, nchunks = run_cpu_baseline(sub, near_maxs[c], cores,
                                                       mm.Profile(**cfg["mods"][c]), res)
                tot_t += dt
                tot_v += sub.size
                notes.append(f"channel {c}: {nchunks} chunks, {dt:.1f} s, {nb} blobs")
            cpu = {"value": tot_v / tot_t / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {sshape[0]} planes x {sshape[1]} x {sshape[2]} of the timed "
                             f"volume, the reference's own chunking, fork pool of {cores} "
                             f"processes; " + "; ".join(notes)}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": cfg["what"].format(shape="x".join(str(v) for v in shape)),
                   "l2": "inputs larger than L2" if nvox > 1e8 else
                         "25 MB input: L2 flushed by the 51 MB float volumes each scale writes",
                   "blobs_per_step": 0 if final is None else int(len(final)), "parity": parity},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(out))


# ----------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--mode", default="chunks", choices=["chunks", "seamless"])
    ap.add_argument("--shape", type=str, default=None,
                    help="z,y,x override: per GPU for config 2, the whole volume for config 3")
    ap.add_argument("--tile", type=str, default="2050,2050",
                    help="seamless mode: y,x (or z,y,x) voxels of owned volume per tile")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-parity", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    shape = tuple(int(v) for v in args.shape.split(",")) if args.shape else (
        SHAPES.get(args.config) or NAMED[args.config]["shape"])
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        run_reference(args, rank, world, shape, cores)
        return

    if args.config in NAMED:
        if world > 1:
            if rank == 0:
                raise SystemExit("configs 1, 4 and 5 are single-GPU workloads")
            return
        main_named(args, shape, cores)
        return

    import torch
    import torch.distributed as dist
    from magellanmapper_b200 import gpu, _lib, multi_gpu, synth
    from magellanmapper_b200.cv import stack_detect
    from magellanmapper_b200.io import np_io

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    gpu.require_cuda()

    tmp = tempfile.mkdtemp(prefix="mmb_bench_")
    os.chdir(tmp)

    # ---- input: this rank's z-slab of the global volume --------------------------------
    seamless = args.mode == "seamless"
    if args.config == 2:
        gshape = (shape[0] * world, shape[1], shape[2])        # weak: one stack per GPU
    else:
        gshape = shape                                         # one volume, N slabs
    held = multi_gpu.slab_bounds(gshape[0], world)
    z0, z1 = held[rank]
    ext_out = None
    if seamless:
        # room for the halo around the slab, generated in place (a second copy of a
        # whole-brain slab would not fit next to the first)
        settings_probe = setup_config(1.0, os.path.join(tmp, "probe"))
        _, _, _, halo, bd = multi_gpu._seamless_setup(gshape, 0)
        own, ext_ranges = multi_gpu.seamless_plan(gshape[0], world, bd[0], halo)
        held = own
        z0, z1 = held[rank]
        e0, e1 = ext_ranges[rank]
        ext_out = torch.empty((e1 - e0, gshape[1], gshape[2]), dtype=torch.int16, device=device)
        vol = ext_out[z0 - e0:z1 - e0]
        synth.device_volume(vol.shape, SEED, offset=(z0, 0, 0), device=device, out=vol)
    else:
        vol = synth.device_volume((z1 - z0, gshape[1], gshape[2]), SEED, offset=(z0, 0, 0),
                                  device=device)
    near_max = near_max_device(vol)
    if world > 1:
        t = torch.tensor([near_max], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        near_max = float(t.item())
    setup_config(near_max, os.path.join(tmp, f"bench_r{rank}"))
    nvox_global = float(np.prod(gshape))
    tile = tuple(int(v) for v in args.tile.split(","))

    def step_resident():
        if seamless:
            table = multi_gpu.detect_seamless(vol, held, gshape, 0, tile_yx=tile,
                                              ext_out=ext_out)
            return table
        if world == 1:
            _, _, blobs = stack_detect.detect_blobs_blocks(
                os.path.join(tmp, f"bench_r{rank}"), np_io.Image5d(vol[None]), None, None, [0],
                False, False, True)
        else:
            _, _, blobs = multi_gpu.detect_blobs_blocks_slabs(
                os.path.join(tmp, f"bench_r{rank}"), vol, held, gshape, [0])
        return blobs

    def table_of(res):
        if res is None:
            return None
        return res if isinstance(res, np.ndarray) else res.blobs

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    lib.mmb_profile_enable(1)          # warm the library's event pool as well
    for _ in range(args.warmup):
        res = step_resident()
    lib.mmb_profile_enable(0)

    # ---- parity of this very stack (before timing; nothing of it is timed) -------------
    parity = None
    if not args.skip_parity:
        final = table_of(res)
        t_par = time.perf_counter()
        if world > 1 and args.config == 2 and not seamless:
            barrier()
            if rank == 0:
                whole = synth.device_volume(gshape, SEED, device=device)
                _, _, one = stack_detect.detect_blobs_blocks(
                    os.path.join(tmp, "single"), np_io.Image5d(whole[None]), None, None, [0],
                    False, False, True)
                del whole
                same = (one.blobs is not None and final is not None
                        and one.blobs.shape == final.shape
                        and bool(np.array_equal(one.blobs, final)))
                parity = {"multi_gpu_table_equals_single_gpu_table": same,
                          "rows": 0 if final is None else int(len(final)),
                          "rows_single_gpu": 0 if one.blobs is None else int(len(one.blobs))}
                if not same:
                    raise SystemExit(f"parity: the {world}-rank table differs from the "
                                     f"single-GPU table of the same volume: {parity}")
                stack_detect.StackDetector.release_workspace()
            barrier()
        if rank == 0:
            if world == 1 and not seamless:
                p = parity_single_gpu(vol, gshape, near_max, final, device, cores)
            else:
                p = parity_cores_global(final, gshape, near_max, device, cores, seamless)
            parity = dict(parity or {}, **p)
            parity["seconds"] = round(time.perf_counter() - t_par, 1)
        barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for k in gpu.STATS:
        gpu.STATS[k] = 0
    launches0 = lib.mmb_launch_count()
    lib.mmb_profile_enable(1)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        res = step_resident()
    ev1.record()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    launches = lib.mmb_launch_count() - launches0
    final = table_of(res)
    n_blobs = 0 if final is None else len(final)
    stage_times = None if res is None or not getattr(res, "times", None) else {
        k.value: round(float(v[0]), 4) for k, v in res.times.items()}
    if world > 1 and multi_gpu.LAST_STAGE_S:
        stage_times = {k: round(v, 4) for k, v in multi_gpu.LAST_STAGE_S.items()}
    ms, cnt, units = collect_profile(lib)
    clocks = sampler.stop() if rank == 0 else None

    # The resident single-GPU call runs the thin trailing chunks on a side stream, under the
    # kernels of the full chunks.  Events around a launch that shares the SMs with another
    # launch time the pair, not the kernel, so the roofline's durations come from the same
    # number of further steps of the same workload with that side stream off (one stream,
    # every kernel alone on the device); the overlapped region's figures are kept beside them.
    overlapped = None
    side_on = world == 1 and not seamless and stack_detect.THIN_CHUNK_FRACTION > 0
    if side_on:
        overlapped = (ms, cnt, units)
        frac_side = stack_detect.THIN_CHUNK_FRACTION
        stack_detect.THIN_CHUNK_FRACTION = 0.0
        try:
            step_resident()
            torch.cuda.synchronize()
            lib.mmb_profile_enable(1)
            t1 = time.perf_counter()
            for _ in range(args.steps):
                step_resident()
            torch.cuda.synchronize()
            one_stream_ms = (time.perf_counter() - t1) / args.steps * 1e3
            ms, cnt, units = collect_profile(lib)
        finally:
            stack_detect.THIN_CHUNK_FRACTION = frac_side

    # a step ends with host-side table assembly, so the step time is the larger of
    # the device span and the wall clock between the two synchronisations
    t_step = max(wall, dev_ms / 1e3)
    if world > 1:
        tt = torch.tensor([t_step], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_step = float(tt.item())
    value = nvox_global * args.steps / t_step / 1e9

    # ---- e2e: pinned host stack -> public API -> blob table on the host ----
    e2e = None
    slab_bytes = int(np.prod(vol.shape)) * 2
    do_e2e = not args.skip_e2e and not seamless
    if do_e2e:
        import psutil
        # every rank pins its slab; leave the host comfortable head-room
        if slab_bytes * world * 1.5 > psutil.virtual_memory().available:
            do_e2e = False
    if world > 1:
        tt = torch.tensor([int(do_e2e)], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        do_e2e = bool(tt.item())
    if do_e2e:
        host = torch.empty(vol.shape, dtype=torch.int16).pin_memory()
        host.copy_(vol)
        torch.cuda.synchronize()
        host_np = host.numpy().view(np.uint16)
        img5d_host = np_io.Image5d(host_np[None])
        img5d_host.is_roi = True

        def step_e2e():
            if world == 1:
                _, _, b = stack_detect.detect_blobs_stack(os.path.join(tmp, f"e2e_r{rank}"),
                                                          img5d_host)
            else:
                _, _, b = multi_gpu.detect_blobs_blocks_slabs(
                    os.path.join(tmp, f"e2e_r{rank}"), host_np, held, gshape, [0])
                if b is not None:
                    b.save_archive()
            return b

        b = step_e2e()                                    # warm the pinned path once
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            b = step_e2e()
        barrier()
        t_e2e = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([t_e2e], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_e2e = float(tt.item())
        nb_e2e = 0 if b is None or b.blobs is None else int(b.blobs.shape[0])
        if world > 1:
            tt = torch.tensor([nb_e2e], device=device, dtype=torch.int64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            nb_e2e = int(tt.item())
        d2h = nb_e2e * 8 * 8          # final table: 8 float64 columns per blob
        e2e = {"value": nvox_global * args.steps / t_e2e / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(nvox_global * 2), "d2h_bytes_per_step": d2h}
        if final is not None and b is not None and b.blobs is not None:
            e2e["table_equals_resident"] = bool(np.array_equal(b.blobs, final))
        del host, host_np, img5d_host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    roof = roofline_block(ms, cnt, units, clocks and clocks.get("sm_mhz"))
    if overlapped is not None:
        o = roofline_block(*overlapped, clocks and clocks.get("sm_mhz"))
        roof["timed_on"] = (
            f"{args.steps} further steps of the same resident workload on one stream "
            f"({one_stream_ms:.1f} ms per step); the region `value` is timed on runs the "
            "thin chunks on a side stream, where the events around a launch also span the "
            "launch it shares the SMs with (overlapped_region)")
        roof["overlapped_region"] = {
            "kernel": o["kernel"], "achieved": o["achieved"], "frac": o["frac"],
            "avg_launch_ms": o["avg_launch_ms"],
            "sum_of_event_ms_per_step": sum(v["ms"] for v in o["per_kernel"].values()) / args.steps,
            "ms_per_step": t_step / args.steps * 1e3}

    cpu = None
    if not args.skip_cpu:
        sshape = cpu_sample_shape(cores, vol.shape)
        sample = vol[:sshape[0], :sshape[1], :sshape[2]].cpu().numpy().view(np.uint16)
        gv, dt, nb, nchunks = run_cpu_baseline(np.ascontiguousarray(sample), near_max, cores)
        cpu = {"value": gv, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {sshape[0]} planes x {sshape[1]} x {sshape[2]} of the timed "
                         f"stack, the reference's own chunking (segment_size 500, overlap 5: "
                         f"{nchunks} chunks), fork pool of {cores} processes, {dt:.1f} s, "
                         f"{nb} blobs"}

    if args.config == 2:
        workload = (f"BASELINE config 2: synthetic cleared-tissue stack "
                    f"{shape[0]}x{shape[1]}x{shape[2]} uint16 per GPU, roi_blobs "
                    f"profile, detect_blobs_blocks (25^3 block preprocessing, 10-scale "
                    f"LoG sigma 3..5, 4-D local maxima, overlap + seam pruning), "
                    f"chunk-faithful, 500^3-voxel chunks with 5-voxel overlap")
    else:
        workload = (f"BASELINE config 3: synthetic whole-brain volume "
                    f"{gshape[0]}x{gshape[1]}x{gshape[2]} uint16 as {world} z-slabs, roi_blobs "
                    f"profile, " + ("ONE seamless chunk (halo exchange, local maxima per slab "
                                    f"tile of {tile}, one global overlap pruning on rank 0)"
                                    if seamless else
                                    "chunk-faithful detect_blobs_blocks (500^3-voxel chunks, "
                                    "5-voxel overlap, seam pruning)"))
    multi = None
    if world > 1:
        multi = (f"{world} z-slabs of one {gshape[0]}x{gshape[1]}x{gshape[2]} volume, "
                 + ("halo planes by NCCL send/recv, candidates gathered to rank 0 and pruned "
                    "there once" if seamless else
                    "chunk rows dealt in balanced runs (single row x y-column units lent "
                    "between ranks), halo planes and lent sub-boxes by NCCL send/recv, device "
                    "tables gathered to rank 0 and seam-pruned there"))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_step / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.config == 2 else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload,
                   "l2": "inputs larger than L2 (every sweep streams >= 1 GB per launch)",
                   "blobs_per_step": n_blobs, "host_stage_s_last_step": stage_times,
                   "multi_gpu": multi, "parity": parity,
                   "prune_within_chunks_rank0_per_step": {
                       k: v // max(args.steps, 1) for k, v in gpu.STATS.items()},
                   "prune_note": "order_dependent = local maxima whose survival in scikit-image's "
                                 "_prune_blobs depends on the iteration order of its pair set "
                                 "(kill chains); the GPU resolves them order-independently "
                                 "(DESIGN.md section 4)"},
        "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args, rank, world, shape, cores):
    """Reference arm: the reference's CPU algorithm (oracle port; scikit-image is
    not installable here) with a fork pool over all host cores, on a bounded
    sample of the same workload, cut by the reference's own chunking."""
    if rank != 0:
        return
    from magellanmapper_b200 import synth
    # many steps of this arm run back to back: about 8 s of CPU work each
    sshape = cpu_sample_shape(cores, shape, seconds=8.0)
    sample, _ = synth.make_volume(sshape, SEED)
    near_max = synth.near_max_of(sample)
    for _ in range(min(args.warmup, 1)):
        run_cpu_baseline(sample, near_max, cores)
    t_tot, nb, nchunks = 0.0, 0, 0
    for _ in range(args.steps):
        gv, dt, nb, nchunks = run_cpu_baseline(sample, near_max, cores)
        t_tot += dt
    value = sample.size * args.steps / t_tot / 1e9
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": t_tot / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if args.config == 2 else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"bounded sample of BASELINE config {args.config}: the first "
                               f"{sshape[0]} planes x {sshape[1]} x {sshape[2]} of a stack of the "
                               f"same recipe (numpy generator) and profile, the reference's own "
                               f"chunking (segment_size 500, overlap 5: {nchunks} chunks)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sshape[0]}x{sshape[1]}x{sshape[2]}, fork pool of {cores}, "
                                   f"{nb} blobs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
