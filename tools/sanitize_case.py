"""Smallest case that touches every kernel family once (for compute-sanitizer --tool memcheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ransac_b200 import GpuContext, capi, generator as gen
ctx = GpuContext(0)
pts = gen.homography(n=777, seed=1)[0]
ctx.set_points(capi.EST_HOMOGRAPHY, pts)
r = ctx.fit(2.0, 0.95, 300, seed=1, round_size=64)[0]
ctx.score(np.stack([r["model"]] * 5), 2.0); ctx.errors(r["model"]); ctx.get_inliers(r["model"], 2.0)
ctx.refit(r["model"], r["inliers"], 2.0)
ctx.fit(2.0, 0.95, 200, seed=1, round_size=64, lo=1)
ctx.set_neighbors_grid(0, 100)
ctx.fit(2.0, 0.95, 200, seed=1, round_size=64, sampler=capi.SAMPLER_NAPSAC, neighbors=capi.NEIGH_GRID)
ctx.set_sprt_pool(0, np.random.default_rng(0).permutation(777).astype(np.int32))
ctx.fit(2.0, 0.95, 200, seed=1, round_size=64, sprt=True)
ctx.set_points(capi.EST_HOMOGRAPHY, np.concatenate([pts, pts[:301]]), [777, 301])
ctx.fit(2.0, 0.95, 200, seed=2, round_size=64)
f = gen.fundamental(n=501, seed=2)[0]
ctx.set_points(capi.EST_FUNDAMENTAL, f)
ctx.fit(2.0, 0.95, 200, seed=1, round_size=64, sampler=capi.SAMPLER_PROSAC)
e = gen.essential(n=401, inlier_ratio=0.5, seed=3)[0]
ctx.set_points(capi.EST_ESSENTIAL, e)
ctx.fit(2.5e-3, 0.95, 128, seed=1, round_size=64)
l = gen.line2d(n=333, seed=4)[0]
ctx.set_points(capi.EST_LINE2D, l)
r = ctx.fit(8.0, 0.99, 100, seed=1, round_size=32)[0]
ctx.refit(r["model"], r["inliers"], 8.0)
ctx.close()
print("sanitize case done")
