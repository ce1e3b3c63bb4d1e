"""Developer tool (no GPU): numpy.savez against io/npz_writer.savez for blob archives of the
sizes config 2 produces on 1 and 8 GPUs and a quarter of config 3's, each written to a NEW
file (five repeats, minimum and median)."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from magellanmapper_b200.io import npz_writer

tmp = tempfile.mkdtemp()
rng = np.random.default_rng(0)
print(f"host: {os.cpu_count()} cores, files under {tmp}")
for rows in (263022, 2066966, 4120000):
    arc = {"ver": 5, "segments": rng.random((rows, 8)), "resolutions": np.array([[1.0, 1.0, 1.0]]),
           "basename": "sample", "offset": None, "roi_size": None, "colocs": None,
           "cols": ["z", "y", "x", "radius", "confirmed", "truth", "channel", "region"]}
    res = {}
    for name, fn in (("numpy.savez", lambda p: np.savez(open(p, "wb"), **arc)),
                     ("npz_writer.savez", lambda p: npz_writer.savez(p, arc, add_suffix=False))):
        ts = []
        for i in range(5):
            p = os.path.join(tmp, f"{name}_{rows}_{i}.npz")
            t0 = time.perf_counter()
            fn(p)
            ts.append((time.perf_counter() - t0) * 1e3)
            os.remove(p)
        res[name] = (min(ts), sorted(ts)[2])
    mb = rows * 64 / 1e6
    print(f"{rows:8d} rows ({mb:7.1f} MB): " + "; ".join(
        f"{k} min {v[0]:.1f} median {v[1]:.1f} ms" for k, v in res.items()))
