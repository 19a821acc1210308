#!/usr/bin/env python3
"""Golden vectors for the OpenCV primitives the reference's path calls (third-party code, absent from /root/reference).

OpenCV is un-pinned in the reference (find_package(OpenCV REQUIRED), CMakeLists.txt:5); the stand-in available here is
the Python wheel opencv-python-headless 4.13.0. This script calls the same primitives the reference calls and stores
input/output pairs in tests/golden/cv_primitives.npz:

  inv_in/inv_out        cv2.invert on 3x3 float32            <- homography_estimator.hpp:35  (model.inv())
  cubic_in/cubic_out/n  cv2.solveCubic                        <- seven_points.cpp:131
  h4_pts/h4_H           null vector of the Hartley-normalised 4-point DLT matrix via cv2.SVDecomp(FULL_UV) in float64
                        (dlt.cpp:55-101 + normalized_dlt.cpp:7-23 semantics with the true null vector)
  f7_pts/f7_F/f7_n      cv2.findFundamentalMat(FM_7POINT)     <- same Hartley-Zisserman 7-point algorithm the
                        reference copies in seven_points.cpp:49-156 (OpenCV normalises the points first)

Run in the build container: python tests/golden/make_cv_golden.py
"""
import os

import cv2
import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
g = np.random.Generator(np.random.Philox(20261018))

# --- 3x3 inverse -------------------------------------------------------------------------------------------------
inv_in = []
for i in range(512):
    if i % 2 == 0:
        M = g.normal(size=(3, 3))
    else:   # homography-like
        M = np.array([[g.uniform(.5, 1.5), g.uniform(-.3, .3), g.uniform(-200, 200)],
                      [g.uniform(-.3, .3), g.uniform(.5, 1.5), g.uniform(-200, 200)],
                      [g.uniform(-1e-3, 1e-3), g.uniform(-1e-3, 1e-3), 1.0]])
    inv_in.append(M.astype(np.float32))
inv_in.append(np.array([[1, 2, 3], [2, 4, 6], [1, 0, 1]], np.float32))   # singular -> zeros
inv_in = np.stack(inv_in)
inv_out = np.stack([cv2.invert(M)[1] if cv2.invert(M)[0] != 0 else np.zeros((3, 3), np.float32) for M in inv_in]).astype(np.float32)

# --- cubic ---------------------------------------------------------------------------------------------------------
cubic_in, cubic_out, cubic_n = [], [], []
for i in range(512):
    if i % 3 == 0:      # three real roots
        r = np.sort(g.normal(size=3) * 10 ** g.uniform(-1, 2))
        a = g.normal()
        c = a * np.poly(r)
    elif i % 3 == 1:    # one real root + complex pair
        r0 = g.normal() * 5
        re, im = g.normal() * 3, abs(g.normal()) * 3 + 0.1
        c = g.normal() * np.real(np.poly([r0, re + 1j * im, re - 1j * im]))
    else:
        c = g.normal(size=4) * 10 ** g.uniform(-3, 3, 4)
    c = np.asarray(c, np.float64)
    n, roots = cv2.solveCubic(c.reshape(1, 4))
    cubic_in.append(c)
    cubic_out.append(roots.ravel())
    cubic_n.append(n)

# --- 4-point normalised DLT ----------------------------------------------------------------------------------------
def norm_T(p):
    m = p.mean(0)
    d = np.sqrt(((p - m) ** 2).sum(1)).mean()
    s = np.sqrt(2) / d
    return np.array([[s, 0, -m[0] * s], [0, s, -m[1] * s], [0, 0, 1]])

h4_pts, h4_H = [], []
while len(h4_pts) < 256:
    H = np.array([[g.uniform(.7, 1.3), g.uniform(-.3, .3), g.uniform(-100, 100)],
                  [g.uniform(-.3, .3), g.uniform(.7, 1.3), g.uniform(-100, 100)],
                  [g.uniform(-2e-4, 2e-4), g.uniform(-2e-4, 2e-4), 1.0]])
    p1 = g.uniform(0, 1000, (4, 2))
    q = np.c_[p1, np.ones(4)] @ H.T
    p2 = q[:, :2] / q[:, 2:3] + g.normal(0, 0.5, (4, 2))
    pts = np.c_[p1, p2].astype(np.float32).astype(np.float64)
    T1, T2 = norm_T(pts[:, :2]), norm_T(pts[:, 2:])
    a = np.c_[pts[:, :2], np.ones(4)] @ T1.T
    b = np.c_[pts[:, 2:], np.ones(4)] @ T2.T
    A = []
    for (x1, y1, _), (x2, y2, _) in zip(a, b):
        A.append([-x1, -y1, -1, 0, 0, 0, x2 * x1, x2 * y1, x2])
        A.append([0, 0, 0, -x1, -y1, -1, y2 * x1, y2 * y1, y2])
    A = np.array(A)
    w, u, vt = cv2.SVDecomp(A, flags=cv2.SVD_FULL_UV)
    if w[-1, 0] / w[0, 0] < 1e-3:     # keep well-conditioned samples (SURVEY.md hard part 4)
        continue
    Hn = vt[8].reshape(3, 3)
    Hd = np.linalg.inv(T2) @ Hn @ T1
    h4_pts.append(pts.astype(np.float32))
    h4_H.append(Hd / Hd[2, 2])

# --- 7-point ---------------------------------------------------------------------------------------------------------
f7_pts, f7_F, f7_n = [], [], []
while len(f7_pts) < 128:
    f = 800.0
    K = np.array([[f, 0, 500], [0, f, 500], [0, 0, 1.0]])
    ax = g.normal(size=3); ax /= np.linalg.norm(ax)
    ang = np.deg2rad(g.uniform(5, 20))
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    t = g.normal(size=3); t /= np.linalg.norm(t)
    z = g.uniform(4, 12, 7)
    P = np.c_[g.uniform(-.5, .5, (7, 2)) * z[:, None], z]
    a = P @ K.T; a = a[:, :2] / a[:, 2:3]
    Q = P @ R.T + t
    b = Q @ K.T; b = b[:, :2] / b[:, 2:3]
    pts = np.c_[a, b].astype(np.float32)
    F, _ = cv2.findFundamentalMat(pts[:, :2].astype(np.float64), pts[:, 2:].astype(np.float64), cv2.FM_7POINT)
    if F is None:
        continue
    k = F.shape[0] // 3
    Fs = np.zeros((3, 3, 3))
    Fs[:k] = F.reshape(k, 3, 3)
    f7_pts.append(pts); f7_F.append(Fs); f7_n.append(k)

np.savez_compressed(f"{OUT}/cv_primitives.npz", inv_in=inv_in, inv_out=inv_out,
                    cubic_in=np.stack(cubic_in), cubic_out=np.stack(cubic_out), cubic_n=np.array(cubic_n, np.int32),
                    h4_pts=np.stack(h4_pts), h4_H=np.stack(h4_H), f7_pts=np.stack(f7_pts), f7_F=np.stack(f7_F),
                    f7_n=np.array(f7_n, np.int32), cv_version=np.array(cv2.__version__))
print("wrote cv_primitives.npz", cv2.__version__)
