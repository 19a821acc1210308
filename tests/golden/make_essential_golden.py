"""Golden vectors for the five-point solver: for 60 five-point samples (noise-free and noisy calibrated scenes) the complete set
of real essential matrices as computed by OpenCV's own five-point solver (cv2.findEssentialMat on exactly five points returns
every real solution stacked as 3k x 3). OpenCV (cv2 4.13) stands in for the reference's dependency here; the reference's
solver (usac/estimator/essential/five_points.cpp) solves the same polynomial system, so its candidate set is the same.
Run in this container:  python tests/golden/make_essential_golden.py  -> tests/golden/essential5_cv.npz"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ransac_b200 import generator as gen  # noqa: E402

samples, sols, counts = [], [], []
for noise, seed in ((0.0, 11), (0.0, 12), (0.5, 13)):
    pts, E_gt, mask = gen.essential(n=2000, noise=noise, seed=seed)
    inl = np.where(mask)[0]
    g = np.random.default_rng(seed)
    for _ in range(20):
        s = g.choice(inl, 5, replace=False)
        p = pts[s].astype(np.float64)
        E, _ = cv2.findEssentialMat(p[:, :2].copy(), p[:, 2:].copy(), np.eye(3), method=cv2.RANSAC, prob=0.999, threshold=1e-3)
        k = 0 if E is None else E.shape[0] // 3
        out = np.zeros((10, 3, 3))
        for i in range(k):
            out[i] = E[3 * i:3 * i + 3]
        samples.append(pts[s])
        sols.append(out)
        counts.append(k)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "essential5_cv.npz"), points=np.stack(samples).astype(np.float32),
                    solutions=np.stack(sols), counts=np.array(counts, np.int32), cv_version=cv2.__version__)
print("wrote", len(samples), "samples; solutions per sample:", counts)
