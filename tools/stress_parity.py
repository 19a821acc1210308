#!/usr/bin/env python
"""Randomised parity sweep against the CPU oracle (test infrastructure, like tests/): estimators x samplers x SPRT / LO x odd
problem sizes x seeds x round sizes, every fit compared field by field (inliers, iterations, winning sample, model bits, and
under SPRT samples drawn / evaluations). Also get_inliers and the refit loop at sizes on both sides of the multi-launch switch.
usage: stress_parity.py [cases=120] [seed=0]   -> prints the number of cases and every mismatch; exit code 1 on any mismatch."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402
from ransac_b200 import GpuContext, capi  # noqa: E402
from ransac_b200 import generator as gen  # noqa: E402

CASES = int(sys.argv[1]) if len(sys.argv) > 1 else 120
g = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
EST = {1: O.EST_LINE2D, 2: O.EST_HOMOGRAPHY, 3: O.EST_FUNDAMENTAL, 4: O.EST_ESSENTIAL}


def bits(a):
    a = np.ascontiguousarray(a, dtype=np.float32).copy()
    a[np.isnan(a)] = np.float32(np.nan)
    return a.view(np.uint32)


ONLY = set(int(x) for x in os.environ.get("STRESS_ONLY_CASES", "").split(",") if x)   # re-run single cases of a sweep (same random stream)
ctx = GpuContext(0)
bad, t0 = [], time.time()
for case in range(CASES):
    cfg = int(g.choice([1, 2, 2, 3, 3, 4]))
    est = EST[cfg]
    n = int(g.integers(200, 6000)) | 1
    ratio = float(g.choice([0.15, 0.3, 0.5, 0.7]))
    pts = gen.make(cfg, seed_offset=1000 + case, n=n, inlier_ratio=ratio)[0]
    if len(pts) != n:
        n = len(pts)
    thr, conf = gen.CONFIGS[cfg]["threshold"], gen.CONFIGS[cfg]["confidence"]
    seed = int(g.integers(1, 1000))
    K = int(g.choice([1, 7, 32, 100, 256, 512]))
    max_it = int(g.choice([50, 300, 1000]))
    sprt = bool(g.random() < 0.5)
    lo = int(g.choice([0, 0, 1, 2])) if cfg != 1 else 0
    sampler = str(g.choice(["uniform", "uniform", "prosac", "napsac"])) if cfg != 1 else "uniform"
    kw = dict(threshold=thr, confidence=conf, max_iterations=max_it, seed=seed, round_size=K)
    okw = dict(threshold=thr, confidence=conf, max_iterations=max_it, seed=seed, rng=O.RNG_PHILOX)
    ctx.set_points(est, pts)
    if sampler == "prosac":
        kw["sampler"] = capi.SAMPLER_PROSAC; okw["sampler"] = O.SAMPLER_PROSAC
    elif sampler == "napsac":
        if g.random() < 0.5:
            ctx.set_neighbors_grid(0, 50)
            kw.update(sampler=capi.SAMPLER_NAPSAC, neighbors=capi.NEIGH_GRID); okw.update(sampler=O.SAMPLER_NAPSAC, neighbors=O.NEIGH_GRID, cell_size=50)
        else:
            ctx.build_neighbors_knn(0, 7)
            kw.update(sampler=capi.SAMPLER_NAPSAC, neighbors=capi.NEIGH_KNN); okw.update(sampler=O.SAMPLER_NAPSAC, neighbors=O.NEIGH_KNN, knn_table=O.knn_build(pts, 7))
    if sprt:
        ctx.set_sprt_pool(0, O.sprt_pool(seed, n))
        kw["sprt"] = True; okw["sprt"] = True
    if lo:
        kw["lo"] = lo; okw["lo"] = lo
    if sprt or sampler == "prosac" or lo:
        okw["batch"] = min(K, max_it)          # usac_gpu_fit never draws more than max_iterations samples per round
    tag = f"case {case}: cfg {cfg} n {n} ratio {ratio} {sampler} sprt {sprt} lo {lo} K {K} max_it {max_it} seed {seed}"
    if ONLY and case not in ONLY:
        continue
    try:
        r = ctx.fit(**kw)[0]
        ref = O.ransac(pts, est, **okw)
        keys = ["inliers", "iterations", "best_hyp", "best_model_idx"] + (["samples_drawn", "evals"] if sprt else [])
        diff = [k for k in keys if r[k] != ref[k]]
        if not np.array_equal(bits(r["model"]), bits(ref["model"])):
            diff.append("model")
        if diff:
            bad.append((tag, diff, {k: (r[k], ref[k]) for k in diff if k != "model"}))
            if os.environ.get("STRESS_DEBUG"):
                kw2 = {k: v for k, v in kw.items() if k not in ("lo", "sprt")}
                okw2 = {k: v for k, v in okw.items() if k not in ("lo", "sprt", "batch")}
                r2, ref2 = ctx.fit(**kw2)[0], O.ransac(pts, est, **okw2)
                print("DEBUG", tag, "neighbors", kw.get("neighbors"), "| with flags: gpu", {k: r[k] for k in ("inliers", "iterations", "best_hyp", "rounds", "samples_drawn")},
                      "oracle", {k: ref[k] for k in ("inliers", "iterations", "best_hyp")}, "| plain fit: gpu", {k: r2[k] for k in ("inliers", "iterations", "best_hyp")},
                      "oracle", {k: ref2[k] for k in ("inliers", "iterations", "best_hyp")})
                for K2 in (1, 16, 512):
                    r3 = ctx.fit(**dict(kw, round_size=K2))[0]
                    print("DEBUG   same flags, round size", K2, {k: r3[k] for k in ("inliers", "iterations", "best_hyp", "rounds")})
        # Quality::getInliers and the refit loop on the result
        if r["inliers"] > 0 and cfg != 1:
            ids = ctx.get_inliers(r["model"], thr)
            oc = O.score(est, pts, r["model"], thr, want_inliers=True)
            if len(ids) != oc[0] or not np.array_equal(ids, oc[3][:oc[0]]):
                bad.append((tag, ["get_inliers"], {}))
            rf = ctx.refit(r["model"], r["inliers"], thr)
            orf = O.refit(est, pts, r["model"], r["inliers"], thr)
            if rf["inliers"] != orf["inliers"] or not np.array_equal(bits(rf["model"]), bits(np.asarray(orf["model"], np.float32))):
                bad.append((tag, ["refit"], {"inliers": (rf["inliers"], orf["inliers"])}))
    except Exception as e:   # noqa: BLE001
        bad.append((tag, ["exception"], {"error": repr(e)}))

# batches of ragged problems in flight: the plain path (device-side select) and the SPRT batch path (host replay per problem)
for case in range(max(4, CASES // 15)):
    cfg = int(g.choice([1, 2, 3, 4]))
    est = EST[cfg]
    B = int(g.integers(2, 9))
    sizes = [int(g.integers(100, 3000)) for _ in range(B)]
    sets = [gen.make(cfg, seed_offset=5000 + 31 * case + i, n=sizes[i], inlier_ratio=float(g.choice([0.05, 0.3, 0.6])))[0] for i in range(B)]
    sizes = [len(x) for x in sets]
    thr, conf = gen.CONFIGS[cfg]["threshold"], gen.CONFIGS[cfg]["confidence"]
    seed, K, max_it = int(g.integers(1, 1000)), int(g.choice([16, 100, 256])), int(g.choice([300, 1000]))
    sprt = bool(g.random() < 0.6)
    ctx.set_points(est, np.concatenate(sets), sizes)
    kw = dict(threshold=thr, confidence=conf, max_iterations=max_it, seed=seed, round_size=K)
    okw = dict(threshold=thr, confidence=conf, max_iterations=max_it, seed=seed, rng=O.RNG_PHILOX)
    if sprt:
        for p_ in range(B):
            ctx.set_sprt_pool(p_, O.sprt_pool(seed, sizes[p_]))
        kw["sprt"] = True; okw.update(sprt=True, batch=min(K, max_it))
    tag = f"batch {case}: cfg {cfg} sizes {sizes} sprt {sprt} K {K} max_it {max_it} seed {seed}"
    try:
        res = ctx.fit(**kw)
        for p_ in range(B):
            ref = O.ransac(sets[p_], est, **okw)
            keys = ["inliers", "iterations", "best_hyp", "best_model_idx"] + (["samples_drawn", "evals"] if sprt else [])
            diff = [k for k in keys if res[p_][k] != ref[k]]
            if not np.array_equal(bits(res[p_]["model"]), bits(ref["model"])):
                diff.append("model")
            if diff:
                bad.append((tag + f" problem {p_}", diff, {k: (res[p_][k], ref[k]) for k in diff if k != "model"}))
    except Exception as e:   # noqa: BLE001
        bad.append((tag, ["exception"], {"error": repr(e)}))

# the ordered inlier list across the multi-launch switch (32768 points) and at sizes that are not multiples of 1024
for n in (32767, 32768, 40001, 131073):
    pts, H, mask = gen.homography(n=n, inlier_ratio=0.3, seed=n)
    ctx.set_points(O.EST_HOMOGRAPHY, pts)
    ids = ctx.get_inliers(np.asarray(H, np.float32), 2.0)
    oc = O.score(O.EST_HOMOGRAPHY, pts, np.asarray(H, np.float32), 2.0, want_inliers=True)
    if len(ids) != oc[0] or not np.array_equal(ids, oc[3][:oc[0]]):
        bad.append((f"get_inliers n={n}", ["ids"], {"count": (len(ids), oc[0])}))
    rf = ctx.refit(np.asarray(H, np.float32), int(oc[0]), 2.0)
    orf = O.refit(O.EST_HOMOGRAPHY, pts, np.asarray(H, np.float32), int(oc[0]), 2.0)
    if rf["inliers"] != orf["inliers"] or not np.array_equal(bits(rf["model"]), bits(np.asarray(orf["model"], np.float32))):
        bad.append((f"refit n={n}", ["refit"], {"inliers": (rf["inliers"], orf["inliers"])}))
ctx.close()
print(f"stress_parity: {CASES} fits + {max(4, CASES // 15)} ragged batches + 4 large inlier lists in {time.time() - t0:.1f} s, {len(bad)} mismatches")
for b in bad:
    print("MISMATCH", b)
sys.exit(1 if bad else 0)
