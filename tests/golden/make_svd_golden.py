#!/usr/bin/env python3
"""Golden vectors for the stand-in cv::SVD of oracle/ref_shim/cvshim.hpp: the real OpenCV (opencv-python-headless, cv2.SVDecomp) on
the kinds of matrices the reference's estimators feed cv::SVD, so that the compiled reference (oracle/_ref) can be trusted where its
results go through a singular value decomposition.

  kind 0   7 x 9 float32, FULL_UV   seven_points.cpp:88        rows 7, 8 of vt: the basis of the 2-D null space (any basis: compared as a projector)
  kind 1   5 x 9 float64, FULL_UV   five_points.cpp:65         rows 5..8 of vt: the 4-D null space
  kind 2   8 x 9 float32, thin      dlt.cpp:43 (DLT4p)         vt is 8 x 9; its LAST ROW (8th singular vector) is what DLT4p takes
  kind 3  2N x 9 float32, thin      dlt.cpp:92 (DLT, N >= 5)   last row of vt = null vector of the (normalised) system
  kind 4   N x 9 float32, thin      eight_points.cpp:38        last row of vt
  kind 5   4 x 4 float64            five_points.cpp:327        last row of vt (triangulation)
  kind 6   3 x 3 float64            five_points.cpp:341        singular values of an essential matrix

Also eig_cv.npz: cv2.eigen on 64 symmetric 2 x 2 float32 scatter matrices (line2d_estimator.hpp:86-93).

Stored per case: the matrix (float64 copy of the values in its own depth), kind, depth, flags, cv2's w and vt (zero padded to 9 x 9).
Run in the build container: python tests/golden/make_svd_golden.py"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ransac_b200 import generator as gen  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
g = np.random.Generator(np.random.Philox(20261020))
ptsF, _, maskF = gen.make(3, n=4000)
ptsH, _, maskH = gen.make(2, n=4000)
ptsE, _, maskE = gen.make(4, n=4000)
inlF, inlH, inlE = np.where(maskF)[0], np.where(maskH)[0], np.where(maskE)[0]


def epi_rows(p, order):
    rows = []
    for x1, y1, x2, y2 in p:
        rows.append([x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1] if order == 7 else [x1 * x2, x2 * y1, x2, x1 * y2, y1 * y2, y2, x1, y1, 1])
    return np.array(rows)


def dlt_rows(p):
    rows = []
    for x1, y1, x2, y2 in p:
        rows.append([-x1, -y1, -1, 0, 0, 0, x2 * x1, x2 * y1, x2])
        rows.append([0, 0, 0, -x1, -y1, -1, y2 * x1, y2 * y1, y2])
    return np.array(rows)


def normalise(p):
    q = p.astype(np.float64).copy()
    for c in (0, 2):
        m = q[:, c:c + 2].mean(0)
        d = np.sqrt(((q[:, c:c + 2] - m) ** 2).sum(1)).mean()
        q[:, c:c + 2] = (q[:, c:c + 2] - m) * (np.sqrt(2) / d)
    return q.astype(np.float32)


cases = []
for t in range(40):
    s = g.choice(inlF, 7, replace=False) if t % 2 else g.choice(len(ptsF), 7, replace=False)
    cases.append((0, np.float32, cv2.SVD_FULL_UV, epi_rows(ptsF[s].astype(np.float32), 7).astype(np.float32)))
for t in range(40):
    s = g.choice(inlE, 5, replace=False) if t % 2 else g.choice(len(ptsE), 5, replace=False)
    cases.append((1, np.float64, cv2.SVD_FULL_UV, epi_rows(ptsE[s].astype(np.float64), 5)))
for t in range(40):
    s = g.choice(inlH, 4, replace=False) if t % 2 else g.choice(len(ptsH), 4, replace=False)
    cases.append((2, np.float32, 0, dlt_rows(ptsH[s].astype(np.float32)).astype(np.float32)))
for t in range(30):
    s = g.choice(inlH, int(g.choice([5, 8, 20, 60])), replace=False)
    cases.append((3, np.float32, 0, dlt_rows(normalise(ptsH[s])).astype(np.float32)))
for t in range(30):
    s = g.choice(inlF, int(g.choice([9, 12, 30, 80])), replace=False)
    cases.append((4, np.float32, 0, epi_rows(normalise(ptsF[s]), 7).astype(np.float32)))
for t in range(20):
    cases.append((5, np.float64, 0, g.normal(size=(4, 4))))
for t in range(20):
    R, _ = np.linalg.qr(g.normal(size=(3, 3)))
    tx = g.normal(size=3)
    E = np.array([[0, -tx[2], tx[1]], [tx[2], 0, -tx[0]], [-tx[1], tx[0], 0]]) @ R
    cases.append((6, np.float64, 0, E))

kinds, depths, flags, shapes, mats, ws, vts = [], [], [], [], [], [], []
for kind, dt, fl, A in cases:
    A = np.ascontiguousarray(A, dtype=dt)
    w, u, vt = cv2.SVDecomp(A, flags=fl)
    M = np.zeros((160, 9)); M[:A.shape[0], :A.shape[1]] = A
    W = np.zeros(9); W[:w.size] = w.ravel()
    V = np.zeros((9, 9)); V[:vt.shape[0], :vt.shape[1]] = vt
    kinds.append(kind); depths.append(64 if dt == np.float64 else 32); flags.append(fl); shapes.append(list(A.shape) + list(vt.shape))
    mats.append(M); ws.append(W); vts.append(V)
np.savez_compressed(os.path.join(OUT, "svd_cv.npz"), kind=np.array(kinds), depth=np.array(depths), flags=np.array(flags), shape=np.array(shapes),
                    A=np.stack(mats), w=np.stack(ws), vt=np.stack(vts), cv_version=cv2.__version__)
print("wrote svd_cv.npz:", len(cases), "decompositions, OpenCV", cv2.__version__)

# cv::eigen on the symmetric float32 matrices of the path: the 2 x 2 scatter matrix of the non-minimal line fit (line2d_estimator.hpp:86-93)
eig_A, eig_vals, eig_vecs = [], [], []
for t in range(64):
    n_pts = int(g.choice([2, 5, 50, 500]))
    ang = g.uniform(0, np.pi)
    tt = g.uniform(-500, 500, n_pts)
    xy = np.c_[500 + tt * np.cos(ang), 500 + tt * np.sin(ang)] + g.normal(0, 3.0, (n_pts, 2))
    xy = xy.astype(np.float32)
    m = xy.mean(0)
    cov = ((xy - m).T @ (xy - m)).astype(np.float32)
    cov[1, 0] = cov[0, 1]
    ok, vals, vecs = cv2.eigen(cov)
    eig_A.append(cov); eig_vals.append(vals.ravel()); eig_vecs.append(vecs)
np.savez_compressed(os.path.join(OUT, "eig_cv.npz"), A=np.stack(eig_A), vals=np.stack(eig_vals), vecs=np.stack(eig_vecs), cv_version=cv2.__version__)
print("wrote eig_cv.npz:", len(eig_A), "2 x 2 float32 eigen-decompositions")
