"""Why BASELINE config 4 (essential, 20 % inliers, SPRT) ends with a chance-level model - CPU only, through the oracle (test infrastructure).
Output kept in profiles/r2_c4_diagnosis.txt."""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from ransac_b200 import generator as gen
pts, E, mask = gen.make(4)
print(len(pts), mask.sum(), gen.CONFIGS[4])
thr = 2.5e-3
cgt = O.score(O.EST_ESSENTIAL, pts, E, thr)
print("GT model inliers", cgt[0])
inl = np.where(mask)[0]
# all-inlier samples in the philox stream
S = O.Sampler(O.SAMPLER_UNIFORM, O.RNG_PHILOX, len(pts), 5, 1)
tab = S.table(10000)
allin = np.where(mask[tab].all(axis=1))[0]
print("all-inlier samples:", allin)
for j in allin:
    mods = O.solve_minimal(O.EST_ESSENTIAL, pts, tab[j])
    cands, valid = O.essential5_candidates(pts, tab[j])
    print(j, "nmodels", len(mods), "cands", len(cands), valid)
    for m in mods:
        print("   returned model inliers", O.score(O.EST_ESSENTIAL, pts, m, thr)[0])
    for c in cands:
        print("   cand inliers", O.score(O.EST_ESSENTIAL, pts, c.astype(np.float32).ravel(), thr)[0])
# random all-inlier samples
g = np.random.default_rng(0)
res=[]
for t in range(40):
    s = g.choice(inl, 5, replace=False).astype(np.int32)
    mods = O.solve_minimal(O.EST_ESSENTIAL, pts, s)
    cands, valid = O.essential5_candidates(pts, s)
    best = max([O.score(O.EST_ESSENTIAL, pts, c.astype(np.float32).ravel(), thr)[0] for c in cands], default=0)
    ret = O.score(O.EST_ESSENTIAL, pts, mods[0], thr)[0] if len(mods) else -1
    res.append((ret,best,len(cands)))
print(res)
for sprt in (False, True):
    t0=time.time()
    r = O.ransac(pts, O.EST_ESSENTIAL, rng=O.RNG_PHILOX, threshold=thr, confidence=0.95, max_iterations=10000, seed=1, sprt=sprt, batch=512 if sprt else 0)
    print("sprt", sprt, r["inliers"], r["iterations"], r["best_hyp"], time.time()-t0)
# the reference's own driver (oracle/_ref: Ransac::run compiled from /root/reference) on the same data, for comparison
try:
    from oracle import ref as R
    for sprt in (False, True):
        for seed in (1, 2, 3):
            t0 = time.time()
            r = R.ransac_run(O.EST_ESSENTIAL, pts, thr, conf=0.95, max_it=10000, sprt=sprt, seed=seed)
            print("reference Ransac::run sprt", sprt, "seed", seed, "inliers after refit", r["inliers"], "iterations", r["iterations"], "GT inliers among them",
                  int(mask[r["ids"]].sum()), f"{time.time() - t0:.1f} s")
except Exception as e:   # noqa: BLE001
    print("oracle/_ref not available:", e)
