"""Developer tool: device table -> host, single copy vs the pipelined route of
device_tables._to_host, for an eight-GPU-sized final table (2.07 M rows x 8 float64)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from magellanmapper_b200.cv import device_tables as dt
t = torch.rand((2_066_965, 8), dtype=torch.float64, device="cuda")
for name, chunk in (("pipelined", 8 << 20), ("single copy", 1 << 40)):
    dt._CHUNK_BYTES = chunk
    for _ in range(2):
        a = dt._to_host(t)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        a = dt._to_host(t)
    dtm = (time.perf_counter() - t0) / 5
    print(f"{name}: {dtm * 1e3:.1f} ms for {t.numel() * 8 / 1e6:.0f} MB, equal {bool((torch.from_numpy(a) == t.cpu()).all())}")
