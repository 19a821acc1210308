#!/usr/bin/env python
"""Scoring-kernel microbenchmark without a profiler: one round (max_iterations = K) over B C2 problems, so that the fit
launches score_kernel exactly once with B x K models x 4000 points; CUDA-event time of that launch from usac_gpu_last_timing."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ransac_b200 import GpuContext, capi  # noqa: E402
from ransac_b200 import generator as gen  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
N = 4000
pts = np.concatenate([gen.homography(n=N, inlier_ratio=0.0, seed=1000 + i)[0] for i in range(B)])
ctx = GpuContext(0)
ctx.set_points(capi.EST_HOMOGRAPHY, pts, [N] * B)
info = ctx.device_info()
peak = 2.0 * 128 * info["sm_count"] * info["sm_clock_khz"] * 1e3 / 1e12
for K in (64, 128, 256, 512):
    best = 1e9
    for rep in range(6):
        res = ctx.fit_records(2.0, 0.95, K, seed=rep + 1, round_size=K)
        t = ctx.last_timing()
        assert t["score_launches"] == 1
        if rep:
            best = min(best, t["score_ms"])
    ev = float(res["evals"].sum())
    print(f"B={B} K={K}: {best:.3f} ms  {ev / best / 1e9:.1f} G evals/ms-scaled  {42 * ev / (best * 1e-3) / 1e12:.1f} TFLOP/s  frac {42 * ev / (best * 1e-3) / 1e12 / peak:.3f}")
