"""Developer tool: do the chunk driver's results depend on WHAT runs next to it?
Config 2 resident on the main stream, one-stream route (no thin-chunk side stream), while a
second stream is kept busy with unrelated library kernels (device copies and elementwise
math on its own buffers).  If the table still changes from run to run, concurrency alone is
enough - a race inside one of the driver's kernels, or an ordering assumption that only holds
on an otherwise idle GPU; if it never changes, the differences of the side-stream route
(tools/side_stream_check.py) need the thin chunks' own kernels."""
import os
import sys
import tempfile
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from magellanmapper_b200 import gpu, synth
from magellanmapper_b200.cv import stack_detect
from magellanmapper_b200.io import np_io

dev = torch.device("cuda", 0)
tmp = tempfile.mkdtemp()
os.chdir(tmp)
vol = synth.device_volume((512, 2048, 2048), 1, device=dev)
bench.setup_config(bench.near_max_device(vol), tmp + "/c2")


def run():
    gpu.CHUNK_LOG = []
    _, _, b = stack_detect.detect_blobs_blocks(tmp + "/c2", np_io.Image5d(vol[None]), None, None,
                                               [0], False, False, True)
    return b.blobs, sorted(gpu.CHUNK_LOG)


ref, ref_log = run()
print("ref rows", ref.shape, flush=True)
side = torch.cuda.Stream(device=dev)
with torch.cuda.stream(side):
    a = torch.rand(16 << 20, device=dev)
    b = torch.empty_like(a)
stop = threading.Event()


def noise():
    with torch.cuda.stream(side):
        while not stop.is_set():
            for _ in range(8):
                b.copy_(a)
                torch.sin(a, out=b)
                a.add_(b, alpha=1e-3)
            side.synchronize()


bad = 0
for i in range(int(os.environ.get("PROBE_RUNS", "8"))):
    stop.clear()
    t = threading.Thread(target=noise)
    t.start()
    v, log = run()
    stop.set()
    t.join()
    d = int(np.count_nonzero(np.any(v != ref, axis=1))) if v.shape == ref.shape else -1
    bad += d != 0
    import collections
    shapes = collections.Counter(x[0] for x, y in zip(ref_log, log) if x != y)
    print(i, d, dict(shapes), flush=True)
print("BAD RUNS", bad)
