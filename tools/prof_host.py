import sys, os, cProfile, pstats, time, tempfile
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import bench
from magellanmapper_b200 import synth
from magellanmapper_b200.cv import stack_detect
from magellanmapper_b200.io import np_io
dev = torch.device("cuda", 0)
vol = synth.device_volume(bench.SHAPES[2], 1, device=dev)
nm = bench.near_max_device(vol)
tmp = tempfile.mkdtemp(); os.chdir(tmp)
bench.setup_config(nm, os.path.join(tmp, "x"))
img = np_io.Image5d(vol[None])
def step():
    return stack_detect.detect_blobs_blocks(os.path.join(tmp, "x"), img, None, None, [0], False, False, True)
step(); torch.cuda.synchronize()
t=time.perf_counter(); step(); torch.cuda.synchronize(); print("step", time.perf_counter()-t)
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
