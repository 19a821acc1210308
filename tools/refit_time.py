#!/usr/bin/env python
"""Time of Quality::getInliers and of the final refit loop (ransac.cpp:157-207) on the 1M-point problem (C5), next to the fit."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ransac_b200 import GpuContext, capi  # noqa: E402
from ransac_b200 import generator as gen  # noqa: E402

pts, H, mask = gen.make(5)
ctx = GpuContext(0)
ctx.set_points(capi.EST_HOMOGRAPHY, pts)
ctx.set_neighbors_grid(0, 50)
r = ctx.fit(2.0, 0.95, 10000, seed=1, round_size=5000, sampler=capi.SAMPLER_NAPSAC, neighbors=capi.NEIGH_GRID)[0]


def best(fn, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), out


t_inl, ids = best(lambda: ctx.get_inliers(r["model"], 2.0))
t_ref, ref = best(lambda: ctx.refit(r["model"], r["inliers"], 2.0))
t_gt, ids_gt = best(lambda: ctx.get_inliers(np.asarray(H, np.float32), 2.0))
print(f"C5 1M points: fit inliers {r['inliers']}; get_inliers {t_inl:.3f} ms ({len(ids)} ids, ascending {bool(np.all(np.diff(ids) > 0))}); "
      f"refit loop {t_ref:.3f} ms -> {ref['inliers']} inliers; get_inliers of the ground-truth model {t_gt:.3f} ms ({len(ids_gt)} ids)")
