#!/usr/bin/env python
"""ms per robust fit of the five BASELINE.json configurations: the GPU path (median of 7 fits after 2 warm-ups, host wall clock
around usac_gpu_fit, points resident) next to the CPU oracle (one fit, one thread, same sample stream / batched semantics),
with the parity of the two results. Output goes to profiles/."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from ransac_b200 import GpuContext, capi  # noqa: E402
from ransac_b200 import generator as gen  # noqa: E402

CASES = [
    ("C1 line2d N=1000 50% uniform", 1, {}, dict(est=O.EST_LINE2D, thr=8.0, conf=0.99), {}),
    ("C2 homography N=4000 30% uniform", 2, {}, dict(est=O.EST_HOMOGRAPHY, thr=2.0, conf=0.95), {}),
    ("C3 fundamental N=10000 25% PROSAC+SPRT", 3, {}, dict(est=O.EST_FUNDAMENTAL, thr=2.0, conf=0.95),
     dict(sampler="prosac", sprt=True, K=512)),
    ("C4 essential N=20000 20% uniform+SPRT (no LO)", 4, {}, dict(est=O.EST_ESSENTIAL, thr=2.5e-3, conf=0.95), dict(sprt=True, K=512)),
    ("C4 essential N=20000 20% uniform+SPRT+LO", 4, {}, dict(est=O.EST_ESSENTIAL, thr=2.5e-3, conf=0.95), dict(sprt=True, K=512, lo=1)),
    ("C2 homography N=4000 30% uniform+LO", 2, {}, dict(est=O.EST_HOMOGRAPHY, thr=2.0, conf=0.95), dict(lo=1)),
    ("C5 homography N=1M 10% NAPSAC grid", 5, {}, dict(est=O.EST_HOMOGRAPHY, thr=2.0, conf=0.95), dict(sampler="napsac", K=2048, oracle_iters=64)),
]
ONLY = os.environ.get("CONFIG_TIMES_ONLY")            # substring filter on the case name, e.g. "+LO"
NO_CPU = os.environ.get("CONFIG_TIMES_NO_CPU") == "1"  # skip the oracle leg (profiling runs)
REPS = int(os.environ.get("CONFIG_TIMES_REPS", "9"))
ctx = GpuContext(0)
print(f"{'config':48s} {'GPU ms/fit':>10s} {'CPU ms/fit':>11s} {'speed-up':>8s} {'iters':>6s} {'inliers':>8s}  parity")
for name, cfg, genkw, p, mode in CASES:
    if ONLY and ONLY not in name:
        continue
    pts = gen.make(cfg, **genkw)[0]
    ctx.set_points(p["est"], pts)
    kw = dict(threshold=p["thr"], confidence=p["conf"], max_iterations=10000, seed=1)
    okw = dict(threshold=p["thr"], confidence=p["conf"], max_iterations=10000, seed=1, rng=O.RNG_PHILOX)
    if mode.get("sampler") == "prosac":
        kw["sampler"] = capi.SAMPLER_PROSAC; okw["sampler"] = O.SAMPLER_PROSAC
    if mode.get("sampler") == "napsac":
        ctx.set_neighbors_grid(0, 50)
        kw.update(sampler=capi.SAMPLER_NAPSAC, neighbors=capi.NEIGH_GRID); okw.update(sampler=O.SAMPLER_NAPSAC, neighbors=O.NEIGH_GRID, cell_size=50)
    if mode.get("sprt"):
        ctx.set_sprt_pool(0, O.sprt_pool(1, len(pts)))
        kw["sprt"] = True; okw["sprt"] = True
    if mode.get("lo"):
        kw["lo"] = mode["lo"]; okw["lo"] = mode["lo"]
    if "K" in mode:
        kw["round_size"] = mode["K"]
        if mode.get("sprt") or mode.get("sampler") == "prosac":
            okw["batch"] = mode["K"]
    times = []
    for rep in range(REPS):
        t0 = time.perf_counter()
        r = ctx.fit(**kw)[0]
        times.append((time.perf_counter() - t0) * 1e3)
    gpu_ms = float(np.median(times[2:] if len(times) > 2 else times))
    if NO_CPU:
        print(f"{name:48s} {gpu_ms:10.3f} {'-':>11s} {'-':>8s} {r['iterations']:6d} {r['inliers']:8d}  (not compared)")
        continue
    scale = 1.0
    if "oracle_iters" in mode:                      # the full 1e10-evaluation fit takes ~1 min on one core: time a prefix and scale
        okw["max_iterations"] = mode["oracle_iters"]
        scale = 10000 / mode["oracle_iters"]
    t0 = time.perf_counter()
    ref = O.ransac(pts, p["est"], **okw)
    cpu_ms = (time.perf_counter() - t0) * 1e3 * scale
    if scale == 1.0:
        same = all(r[k] == ref[k] for k in ("inliers", "iterations", "best_hyp")) and np.array_equal(r["model"].view(np.uint32), np.asarray(ref["model"], np.float32).view(np.uint32))
        par = "identical" if same else "DIFFERENT"
    else:
        par = f"(CPU: first {mode['oracle_iters']} hypotheses, scaled)"
    print(f"{name:48s} {gpu_ms:10.3f} {cpu_ms:11.1f} {cpu_ms / gpu_ms:8.1f} {r['iterations']:6d} {r['inliers']:8d}  {par}")
