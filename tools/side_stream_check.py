"""Developer tool: determinism of the side-stream route for thin chunks
(stack_detect.THIN_CHUNK_FRACTION > 0) on config 2 with a device-resident stack: the
one-stream table is the reference, twelve side-stream runs are compared with it row for
row.  Last result (B200, round 1): 8 of 12 runs differ (a few rows, twice the row count)."""
import sys, os, tempfile
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from magellanmapper_b200 import synth
from magellanmapper_b200.cv import stack_detect
from magellanmapper_b200.io import np_io
from magellanmapper_b200.settings import config
dev = torch.device("cuda", 0)
tmp = tempfile.mkdtemp(); os.chdir(tmp)
# sanity of this binary against the reference-generated vectors
g = np.load("/root/repo/tests/golden/stack_small.npz")
bench.setup_config(float(g["near_max"]), tmp + "/g")
config.roi_profile["segment_size"] = 50
_, _, b = stack_detect.detect_blobs_blocks(tmp + "/g", np_io.Image5d(g["vol"][None]), None, None, [0], False, False, True)
print("golden equal:", np.array_equal(b.blobs, g["plain_blobs"]))
stack_detect.StackDetector.release_workspace()
vol = synth.device_volume((512, 2048, 2048), 1, device=dev)
nm = bench.near_max_device(vol)
bench.setup_config(nm, tmp + "/c2")
def run(img):
    _, _, b = stack_detect.detect_blobs_blocks(tmp + "/c2", np_io.Image5d(img[None]), None, None, [0], False, False, True)
    return b.blobs
from magellanmapper_b200 import gpu
gpu.CHUNK_LOG = []
ref = run(vol)
ref_log = list(gpu.CHUNK_LOG)
print("ref rows", ref.shape)
stack_detect.THIN_CHUNK_FRACTION = 0.3
bad = 0
for i in range(12):
    gpu.CHUNK_LOG = []
    v = run(vol)
    d = int(np.count_nonzero(np.any(v != ref, axis=1))) if v.shape == ref.shape else -1
    bad += d != 0
    print(i, d, end="; ")
    if d != 0:
        # which chunks differ, and already in the number of local maxima (before pruning)?
        a = sorted(ref_log)
        b = sorted(gpu.CHUNK_LOG)
        diff = [(x, y) for x, y in zip(a, b) if x != y]
        print("\n   chunks whose counters differ (shape, peaks, survivors, edges, od):", diff[:4])
print("\nBAD RUNS", bad)
