import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from magellanmapper_b200 import synth
from magellanmapper_b200 import gpu
from magellanmapper_b200.cv import stack_detect, device_tables
from magellanmapper_b200.io import np_io
from magellanmapper_b200.settings import config
dev = torch.device("cuda", 0)
vol = synth.device_volume((512, 2048, 2048), 1, device=dev)
nm = bench.near_max_device(vol)
bench.setup_config(nm, "/tmp/x")
settings = config.get_roi_profile(0)
blocks = stack_detect.setup_blocks(settings, vol.shape)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tables = stack_detect.StackDetector.detect_blobs_sub_rois_device(vol, blocks.sub_roi_slices, blocks.sub_rois_offsets, blocks.denoise_max_shape, [0])
    rows = tables.rows()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    out, df = device_tables.prune_rows(rows, tables.ladders(), blocks.overlap, blocks.tol, blocks.sub_roi_slices, [0], blocks.overlap_padding, final_layout=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"detect {1e3*(t1-t0):.1f} ms  prune {1e3*(t2-t1):.1f} ms  rows {rows.shape[0]} -> {out.shape[0]}")
