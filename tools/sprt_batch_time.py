#!/usr/bin/env python
"""ms per fit of a batch of SPRT problems: all problems in flight (fit_sprt_batch) vs one problem at a time (USAC_GPU_SPRT_BATCH=0).
usage: sprt_batch_time.py [problems=256] [config=3] [n=4000]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402  (pool shuffle only)
from ransac_b200 import GpuContext  # noqa: E402
from ransac_b200 import generator as gen  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
est = {2: O.EST_HOMOGRAPHY, 3: O.EST_FUNDAMENTAL, 4: O.EST_ESSENTIAL}[cfg]
thr, conf = gen.CONFIGS[cfg]["threshold"], gen.CONFIGS[cfg]["confidence"]
sets = [gen.make(cfg, seed_offset=i, n=n)[0] for i in range(B)]
ctx = GpuContext(0)
ctx.set_points(est, np.concatenate(sets), [n] * B)
pool = O.sprt_pool(1, n)
for p in range(B):
    ctx.set_sprt_pool(p, pool)
ts = []
for rep in range(4):
    t0 = time.perf_counter()
    r = ctx.fit(thr, conf, 10000, seed=1, round_size=256, sprt=True)
    ts.append((time.perf_counter() - t0) * 1e3)
print(f"config {cfg}, {B} problems of {n} points, SPRT, batch={os.environ.get('USAC_GPU_SPRT_BATCH', '1')}: {min(ts[1:]):.2f} ms per batch, "
      f"{min(ts[1:]) / B:.4f} ms per fit; mean iterations {np.mean([x['iterations'] for x in r]):.0f}, mean inliers {np.mean([x['inliers'] for x in r]):.0f}")
