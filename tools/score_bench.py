#!/usr/bin/env python
"""Scoring-kernel microbenchmark without a profiler: one round (outlier-only data, so every fit stops after max_iterations = K
samples) over B problems of N points: the fit launches score_kernel exactly once; CUDA-event time of that launch from
usac_gpu_last_timing.   usage: score_bench.py [B] [homography|fundamental|essential|line2d]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ransac_b200 import GpuContext, capi  # noqa: E402
from ransac_b200 import generator as gen  # noqa: E402

# environment: SCORE_BENCH_INLIERS (inlier ratio of the epipolar data, default 0.01 = structureless), SCORE_BENCH_K (comma list of round sizes)
INL = float(os.environ.get("SCORE_BENCH_INLIERS", "0.01"))
KS = tuple(int(v) for v in os.environ.get("SCORE_BENCH_K", "128,256,512").split(","))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
kind = sys.argv[2] if len(sys.argv) > 2 else "homography"
N = 4000
FLOPS = {"homography": 42, "fundamental": 33, "essential": 44, "line2d": 4}[kind]
EST = {"homography": capi.EST_HOMOGRAPHY, "fundamental": capi.EST_FUNDAMENTAL, "essential": capi.EST_ESSENTIAL, "line2d": capi.EST_LINE2D}[kind]
make = {"homography": lambda s: gen.homography(n=N, inlier_ratio=0.0, seed=s)[0], "fundamental": lambda s: gen.fundamental(n=N, inlier_ratio=INL, seed=s)[0],
        "essential": lambda s: gen.essential(n=N, inlier_ratio=INL, seed=s)[0], "line2d": lambda s: gen.line2d(n=N, inlier_ratio=0.0, seed=s)[0]}[kind]
thr = {"homography": 2.0, "fundamental": 2.0, "essential": 2.5e-3, "line2d": 8.0}[kind]
pts = np.concatenate([make(1000 + i) for i in range(B)])
ctx = GpuContext(0)
ctx.set_points(EST, pts, [N] * B)
info = ctx.device_info()
peak = 2.0 * 128 * info["sm_count"] * info["sm_clock_khz"] * 1e3 / 1e12
for K in KS:
    best, ev = 1e9, 0.0
    for rep in range(5):
        res = ctx.fit_records(thr, 0.95, K, seed=rep + 1, round_size=K)
        t = ctx.last_timing()
        per = t["score_ms"] / t["score_launches"]
        if rep and t["score_launches"] == 1 and per < best:
            best, ev = per, float(res["evals"].sum())
    if ev:
        print(f"{kind} B={B} K={K}: {best:.3f} ms  {ev / (best * 1e-3) / 1e12:.3f} T evals/s  {FLOPS * ev / (best * 1e-3) / 1e12:.1f} TFLOP/s  frac {FLOPS * ev / (best * 1e-3) / 1e12 / peak:.3f}")
    else:
        print(f"{kind} B={B} K={K}: more than one scoring launch per fit (termination table allowed more iterations)")

# ---- mode 2: the plain Quality call, M models x one large point set (one scoring launch, every lane busy) -----------------
if len(sys.argv) > 3 and sys.argv[3] == "quality":
    Nbig, M = 400000, 4096
    big = {"homography": lambda: gen.homography(n=Nbig, inlier_ratio=0.3, seed=5), "fundamental": lambda: gen.fundamental(n=Nbig, inlier_ratio=0.3, seed=5),
           "essential": lambda: gen.essential(n=Nbig, inlier_ratio=0.3, seed=5), "line2d": lambda: gen.line2d(n=Nbig, inlier_ratio=0.5, seed=5)}[kind]()
    pts2, gt = big[0], big[1]
    g = np.random.default_rng(0)
    w = 3 if kind == "line2d" else 9
    models = (gt.reshape(1, w).astype(np.float64) * (1 + 1e-3 * g.normal(size=(M, w)))).astype(np.float32)
    ctx.set_points(EST, pts2)
    best = 1e9
    for rep in range(5):
        ctx.score(models, thr)
        t = ctx.last_timing()
        if rep:
            best = min(best, t["score_ms"])
    ev = float(M) * Nbig
    print(f"{kind} quality call M={M} N={Nbig}: {best:.3f} ms  {ev / (best * 1e-3) / 1e12:.3f} T evals/s  {FLOPS * ev / (best * 1e-3) / 1e12:.1f} TFLOP/s  frac {FLOPS * ev / (best * 1e-3) / 1e12 / peak:.3f}")
