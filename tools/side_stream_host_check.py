"""Developer tool: the thin-chunk side stream with a HOST image streamed strip by strip
(stack_detect.THIN_SIDE_FOR_HOST_IMAGES) on config 2: the resident one-stream table is the
reference; runs from pinned host memory with the side stream off and on are compared with it
row for row and timed."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from magellanmapper_b200 import synth
from magellanmapper_b200.cv import stack_detect
from magellanmapper_b200.io import np_io

dev = torch.device("cuda", 0)
tmp = tempfile.mkdtemp()
os.chdir(tmp)
vol = synth.device_volume((512, 2048, 2048), 1, device=dev)
nm = bench.near_max_device(vol)
bench.setup_config(nm, tmp + "/c2")


def run(img):
    _, _, b = stack_detect.detect_blobs_blocks(tmp + "/c2", np_io.Image5d(img[None]), None, None,
                                               [0], False, False, True)
    return b.blobs


stack_detect.THIN_CHUNK_FRACTION = 0.0
ref = run(vol)
print("ref rows", ref.shape, flush=True)
stack_detect.THIN_CHUNK_FRACTION = 0.3
host = torch.empty(vol.shape, dtype=torch.int16).pin_memory()
host.copy_(vol)
torch.cuda.synchronize()
host_np = host.numpy().view(np.uint16)
for flag, runs in ((False, 3), (True, int(os.environ.get("RUNS", "8")))):
    stack_detect.THIN_SIDE_FOR_HOST_IMAGES = flag
    run(host_np)
    bad, ts = 0, []
    for i in range(runs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        v = run(host_np)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
        d = int(np.count_nonzero(np.any(v != ref, axis=1))) if v.shape == ref.shape else -1
        bad += d != 0
    print(f"side stream for host images {flag}: {runs} runs, {bad} differ from the resident "
          f"table; ms per stack min {min(ts):.1f} median {sorted(ts)[len(ts) // 2]:.1f}", flush=True)
