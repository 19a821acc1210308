"""Time the device neighbourhood builds (kNN and grid) on the BASELINE point sets. Usage: python tools/knn_time.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ransac_b200 import GpuContext, capi, generator as gen   # noqa: E402

ctx = GpuContext(0)
for cfg, k in ((2, 5), (4, 5), (5, 5), (5, 8)):
    pts, _, _ = gen.make(cfg)
    est = {2: capi.EST_HOMOGRAPHY, 4: capi.EST_ESSENTIAL, 5: capi.EST_HOMOGRAPHY}[cfg]
    ctx.set_points(est, pts)
    ctx.build_neighbors_knn(0, k)
    t = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.build_neighbors_knn(0, k)
        t.append((time.perf_counter() - t0) * 1e3)
    line = f"C{cfg} n={len(pts)} k={k}: kNN build {min(t):.2f} ms"
    if est == capi.EST_HOMOGRAPHY:
        ctx.set_neighbors_grid(0, 50)
        t0 = time.perf_counter()
        ctx.set_neighbors_grid(0, 50)
        line += f"; grid build (cell 50) {(time.perf_counter() - t0) * 1e3:.2f} ms"
    print(line, flush=True)
