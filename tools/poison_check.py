"""Developer tool: does any kernel of the chunk driver consume workspace bytes it did not
write?  Config 2 with a device-resident stack on ONE stream; the chunk workspace is
filled with a byte pattern before every chunk (0xFF = NaN floats / -1 ints, 0x00, 0x7F =
3.4e38 floats) and the final table is compared with the unpoisoned run row for row.  A
difference names a read of uninitialised memory - the prime suspect for the run-to-run
differences of the side-stream route (tools/side_stream_check.py), where a thin chunk
meets a workspace with a different history."""
import os
import sys
import tempfile

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
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "512,2048,2048").split(","))
vol = synth.device_volume(shape, 1, device=dev)
nm = bench.near_max_device(vol)
bench.setup_config(nm, tmp + "/c2")


def run():
    _, _, b = stack_detect.detect_blobs_blocks(tmp + "/c2", np_io.Image5d(vol[None]), None, None,
                                               [0], False, False, True)
    return b.blobs


ref = run()
print("ref rows", ref.shape, flush=True)
orig = gpu.ChunkDetector.enqueue
pattern = [None]


def poisoned(self, *a, **k):
    if pattern[0] is not None:
        self.work.fill_(pattern[0])
        for s in self.slots:
            if not s.busy:
                s.cand.view(torch.uint8).fill_(pattern[0])
    return orig(self, *a, **k)


gpu.ChunkDetector.enqueue = poisoned
for pat in (0xFF, 0x00, 0x7F, 0xFF):
    pattern[0] = pat
    v = run()
    if v.shape == ref.shape:
        d = int(np.count_nonzero(np.any(v != ref, axis=1)))
    else:
        d = -1
    print(f"pattern 0x{pat:02X}: rows {v.shape[0]} differing {d}", flush=True)
    if d:
        a = {tuple(r) for r in ref[:, :4]}
        b = {tuple(r) for r in v[:, :4]}
        print("  only in ref:", sorted(a - b)[:6])
        print("  only in poisoned:", sorted(b - a)[:6])

# shared memory: every launch followed by a NaN fill of the shared memory of every SM
gpu.ChunkDetector.enqueue = orig
from magellanmapper_b200 import _lib
_lib.load().mmb_debug_smem_poison(1)
v = run()
_lib.load().mmb_debug_smem_poison(0)
d = int(np.count_nonzero(np.any(v != ref, axis=1))) if v.shape == ref.shape else -1
print(f"shared-memory poison: rows {v.shape[0]} differing {d}", flush=True)
if d:
    a = {tuple(r) for r in ref[:, :4]}
    b = {tuple(r) for r in v[:, :4]}
    print("  only in ref:", sorted(a - b)[:6])
    print("  only in poisoned:", sorted(b - a)[:6])
