#!/usr/bin/env python3
"""Golden vectors for SURVEY Appendix B quirk 1: what the reference's minimal homography solver DLt::DLT4p (usac/estimator/dlt/dlt.cpp:7-53)
returns when it runs on top of the real OpenCV. The solver stacks the 8 x 9 DLT system from raw pixel coordinates in float32 (:17-41),
calls cv::SVD::compute (:43) and takes vt.row(vt.rows - 1) / h33 (:48-50); for an 8 x 9 matrix vt is 8 x 9 (thin), so that row is the
8th right singular vector and not the null vector. This script does exactly that with opencv-python-headless (cv2.SVDecomp on the
float32 matrix) and stores the samples and models in tests/golden/dlt4p_cv.npz:

  pts   (256, 4, 4) float32   four correspondences x1 y1 x2 y2 per sample (pixels; the first 128 samples are noisy inliers of a
                              homography, the rest four unrelated correspondences)
  H     (256, 9)    float32   vt[7] / vt[7][8]
  w     (256, 8)    float32   the singular values (sigma_8 / sigma_1 ~ 1e-7: why the float32 route matters)

Checked by tests/test_oracle.py (oracle switch orc_solve_homography_dlt4p_thin) and tests/test_ref_build.py (the compiled reference on
the stand-in SVD of oracle/ref_shim/cvshim.hpp). Run in the build container: python tests/golden/make_dlt4p_golden.py"""
import os

import cv2
import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
g = np.random.Generator(np.random.Philox(20261019))
pts_all, H_all, w_all = [], [], []
while len(pts_all) < 256:
    p1 = g.uniform(0, 1000, (4, 2))
    if len(pts_all) < 128:
        H = np.array([[g.uniform(.7, 1.3), g.uniform(-.3, .3), g.uniform(-100, 100)],
                      [g.uniform(-.3, .3), g.uniform(.7, 1.3), g.uniform(-100, 100)],
                      [g.uniform(-2e-4, 2e-4), g.uniform(-2e-4, 2e-4), 1.0]])
        q = np.c_[p1, np.ones(4)] @ H.T
        p2 = q[:, :2] / q[:, 2:3] + g.normal(0, 0.5, (4, 2))
    else:
        p2 = g.uniform(0, 1000, (4, 2))
    pts = np.c_[p1, p2].astype(np.float32)
    A = np.zeros((8, 9), np.float32)
    for i, (x1, y1, x2, y2) in enumerate(pts):
        A[2 * i] = [-x1, -y1, -1, 0, 0, 0, x2 * x1, x2 * y1, x2]          # float32 products, like the Mat_<float> of dlt.cpp:12
        A[2 * i + 1] = [0, 0, 0, -x1, -y1, -1, y2 * x1, y2 * y1, y2]
    w, u, vt = cv2.SVDecomp(A)
    assert vt.shape == (8, 9) and vt.dtype == np.float32
    if abs(vt[7][8]) < 1e-6:
        continue
    pts_all.append(pts)
    H_all.append(vt[7] / vt[7][8])
    w_all.append(w.ravel())
np.savez_compressed(os.path.join(OUT, "dlt4p_cv.npz"), pts=np.stack(pts_all), H=np.stack(H_all).astype(np.float32), w=np.stack(w_all).astype(np.float32),
                    cv_version=cv2.__version__)
print("wrote dlt4p_cv.npz:", len(pts_all), "samples, OpenCV", cv2.__version__)
