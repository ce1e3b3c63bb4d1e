"""Developer tool: one small chunk through the fused driver, for
`compute-sanitizer --tool racecheck|initcheck python tools/racecheck_small.py`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magellanmapper_b200 import gpu, synth
from magellanmapper_b200._lib import MmbPreprocParams

shape = (30, 60, 70)
vol, _ = synth.make_volume(shape, seed=5, density=1 / 2500.0)
nm = synth.near_max_of(vol)
pre = MmbPreprocParams(5, 99.5, nm * 0.5, 0.2, 1.0, 0.3, 0.2)
det = gpu.ChunkDetector(shape)
got, n = det.detect(gpu.as_source(vol), np.linspace(3, 5, 4), 0.1, 0.5, pre=pre, block_shape=(25, 25, 25))
print("blobs", len(got), "peaks", n)
