"""Per-kernel micro-benchmark on one 505^3 chunk (developer tool, not bench.py).

    python tools/kbench.py [--shape 505,505,505] [--reps 5]

Times every kernel of the chunk pipeline alone with CUDA events on the launching
stream (3 warm-up launches; the working set is far larger than L2) and prints
algorithmic GB/s, % of the measured HBM peak and % of the FP32 FFMA issue peak.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from magellanmapper_b200 import gpu, _lib, synth     # noqa: E402
import bench                                           # noqa: E402


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="505,505,505")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ncu", action="store_true",
                    help="launch every kernel once (r = 16 only), no warm-up: for ncu captures")
    args = ap.parse_args()
    if args.ncu:
        global timeit

        def timeit(fn, reps):          # noqa: F811
            fn()
            torch.cuda.synchronize()
            return 1.0
    Z, Y, X = (int(v) for v in args.shape.split(","))
    dev = torch.device("cuda", 0)
    gpu.require_cuda()
    lib = _lib.load()
    peaks, _ = bench.peaks_json()
    hbm = peaks["hbm_gbs"]
    sm_clock = 1.965e9
    fma_peak = 148 * 128 * sm_clock            # FFMA lanes / s

    vol = synth.device_volume((Z, Y, X), 1, device=dev)
    nm = bench.near_max_device(vol)
    src = gpu.as_source(vol)
    from magellanmapper_b200._lib import MmbPreprocParams
    pre = MmbPreprocParams(5, 99.5, nm * 0.5, 0.2, 1.0, 0.3, 0.2)
    nvox = Z * Y * X
    F = gpu.new_volume(Z, Y, X)
    pitch = F.shape[2]
    nv_p = Z * Y * pitch

    rows = []

    def report(name, ms, bytes_per_voxel, fma_per_voxel=None, n=nvox):
        gbs = bytes_per_voxel * n / (ms * 1e-3) / 1e9
        r = {"kernel": name, "ms": round(ms, 4), "GB/s": round(gbs, 1),
             "hbm_frac": round(gbs / hbm, 3)}
        if fma_per_voxel:
            r["fma_frac"] = round(fma_per_voxel * n / (ms * 1e-3) / fma_peak, 3)
        rows.append(r)
        print(json.dumps(r))

    ms = timeit(lambda: gpu.percentiles(src, (0.5, 99.5), per_plane=True), args.reps)
    report("percentiles per plane (0.5, 99.5)", ms, 4.0)
    ms = timeit(lambda: gpu.preprocess_blocks(src, (25, 25, 25), pre, out=F), args.reps)
    report("preprocess_25^3", ms, 6.0)
    ms = timeit(lambda: gpu.to_float(src, 1 / 65535.0, out=torch.empty_like(F)), args.reps)
    report("to_float", ms, 6.0)

    A, B = torch.zeros_like(F), torch.zeros_like(F)
    C, D = torch.zeros_like(F), torch.zeros_like(F)
    O = torch.zeros_like(F)
    import ctypes as Ct

    def lp(i0, i1, o0, o1, axis, mode, sigma):
        rc = lib.mmb_log_pass(gpu._ptr(i0), gpu._ptr(i1), gpu._ptr(o0), gpu._ptr(o1), Z, Y, X,
                              pitch, axis, mode, float(sigma), -sigma * sigma, gpu._stream())
        assert rc == 0, lib.mmb_last_error()

    os.environ.setdefault("MMB_XY_RMAX", "20")       # time the fused sweep at every radius
    for sigma in ((4.111111111111111,) if args.ncu else (3.0, 4.111111111111111, 4.5, 5.0)):
        r = int(4 * sigma + 0.5)
        ms = timeit(lambda: lp(F, None, A, B, 2, 0, sigma), args.reps)
        report(f"log_x r={r}", ms, 12.0, 4 * r + 2)
        ms = timeit(lambda: lp(A, B, C, D, 1, 1, sigma), args.reps)
        report(f"log_y r={r}", ms, 16.0, 6 * r + 3, nv_p)
        ms = timeit(lambda: lp(C, D, O, None, 0, 2, sigma), args.reps)
        report(f"log_z r={r}", ms, 12.0, 4 * r + 2, nv_p)

        def fused():
            rc = lib.mmb_log_xy_fused(gpu._ptr(F), gpu._ptr(C), gpu._ptr(D), Z, Y, X, pitch,
                                      float(sigma), gpu._stream())
            assert rc == 0, lib.mmb_last_error()
        ms = timeit(fused, args.reps)
        report(f"log_xy fused r={r}", ms, 12.0, 10 * r + 5, nv_p)

    # local maxima on a real 3-scale neighbourhood
    sig = np.linspace(3, 5, 10)
    work = torch.zeros(lib.mmb_log_work_bytes(Z, Y, pitch), dtype=torch.uint8, device=dev)
    cube = [gpu.log_scale(F, X, s, work=work) for s in sig[3:6]]
    cand = gpu.new_cand_buffer(4_000_000)
    counter = torch.zeros(1, dtype=torch.int32, device=dev)

    def lm():
        counter.zero_()
        gpu.localmax(cube[0], cube[1], cube[2], X, 4, 0.1, cand, counter)
    ms = timeit(lm, args.reps)
    hot = float((cube[1][:, :, :X] > 0.1).float().mean())
    report(f"localmax (hot fraction {hot:.3f}, peaks {int(counter.item())})", ms, 4.0)

    if args.ncu:
        return
    # whole chunk through the fused driver
    det = gpu.ChunkDetector((Z, Y, X))
    def chunk():
        det.detect(src, sig, 0.1, 0.5, pre=pre, block_shape=(25, 25, 25))
    ms = timeit(chunk, max(2, args.reps // 2))
    report("detect_chunk (10 scales, end to end)", ms, 126.0)
    print(json.dumps({"GVoxel/s chunk": round(nvox / (ms * 1e-3) / 1e9, 3)}))


if __name__ == "__main__":
    main()
