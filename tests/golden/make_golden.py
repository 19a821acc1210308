#!/usr/bin/env python3
"""Generate the known-answer scoring fixtures from the reference's shipped artefacts.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/scoring_kat.npz   46 scoring known-answer vectors
    tests/golden/line2d_gt.npz     the 8 synthetic line files with their ground-truth (a,b,c)

What pins what
--------------
* homogr (12): dataset/homography/sift_update/<n>_pts.txt + dataset/homography/<n>_model.txt, threshold 2.
  Expected = "GT Inl" column of results/homography/uniform_gc_Grid_c_sz_50.csv. The reference keeps whichever of
  H / H^-1 has more inliers (dataset/GetImage.h:250-264), so the test takes max(count(H), count(inv H)).
* EVD (15): dataset/EVD/EVD_tentatives/<n>.png_m.txt (first 4 CSV columns) + dataset/EVD/h/<n>.txt, threshold 2.
  Expected = "GT Inl" column of results/EVD/uniform_gc_Nanoflann_c_sz_50.csv (same max rule).
* Sampson (19): dataset/fundamental/<n>.txt (columns 1,2,4,5) + F = rows 1-3 of results/fundamental/<n>.csv,
  threshold 3 on the SQUARED Sampson value. Expected = row 4 of the same csv.

Each vector is checked here with an independent NumPy float32 restatement of the reference's GetError
(usac/estimator/homography_estimator.hpp:85-110, fundamental_estimator.hpp:101-117) before it is written, so a
fixture that does not reproduce the published count is never committed silently (the script prints the mismatch).
"""
import csv
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def h_err(pts, H, Hi):
    """float32, one rounding per operator, no FMA (NumPy does not contract)."""
    x1, y1, x2, y2 = (pts[:, i].astype(f32) for i in range(4))
    h = H.astype(f32).ravel()
    hi = Hi.astype(f32).ravel()
    ex = h[0] * x1 + h[1] * y1 + h[2]
    ey = h[3] * x1 + h[4] * y1 + h[5]
    ez = h[6] * x1 + h[7] * y1 + h[8]
    with np.errstate(all="ignore"):
        ex = ex / ez
        ey = ey / ez
        fx = hi[0] * x2 + hi[1] * y2 + hi[2]
        fy = hi[3] * x2 + hi[4] * y2 + hi[5]
        fz = hi[6] * x2 + hi[7] * y2 + hi[8]
        fx = fx / fz
        fy = fy / fz
        e = np.sqrt((x2 - ex) * (x2 - ex) + (y2 - ey) * (y2 - ey)) + np.sqrt((x1 - fx) * (x1 - fx) + (y1 - fy) * (y1 - fy))
    return (e / f32(2)).astype(f32)


def cv_inv3(M):
    """cv::Mat::inv() for 3x3 CV_32F: cofactors in double, scaled by 1/det, rounded to float."""
    m = M.astype(f32).astype(np.float64)
    d = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
         + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    if d == 0:
        return np.zeros((3, 3), f32)
    d = 1.0 / d
    t = np.array([
        (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d, (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d, (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d,
        (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d, (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d, (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d,
        (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d, (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d, (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d])
    return t.astype(f32).reshape(3, 3)


def h_count(pts, H, thr):
    Hi = cv_inv3(H)
    return int(np.sum(h_err(pts, H, Hi) < f32(thr)))


def sampson_err(pts, F):
    x1, y1, x2, y2 = (pts[:, i].astype(f32) for i in range(4))
    f = F.astype(f32).ravel()
    a = f[0] * x1 + f[1] * y1 + f[2]
    b = f[3] * x1 + f[4] * y1 + f[5]
    c = f[0] * x2 + f[3] * y2 + f[6]
    d = f[1] * x2 + f[4] * y2 + f[7]
    n = x2 * a + y2 * b + f[6] * x1 + f[7] * y1 + f[8]
    with np.errstate(all="ignore"):
        return ((n * n) / (a * a + b * b + c * c + d * d)).astype(f32)


def read_pts_txt(path):
    """detector/Reader.cpp:182-213: first line N, then 'x1 y1 x2 y2' rows parsed as float."""
    with open(path) as fh:
        lines = fh.read().split("\n")
    n = int(lines[0].strip())
    rows = []
    for ln in lines[1:]:
        if len(rows) == n:
            break
        tok = ln.split()
        if len(tok) >= 4:
            rows.append([f32(t) for t in tok[:4]])
    return np.array(rows, f32)


def read_mat3(path):
    """detector/Reader.cpp:129-145: nine floats."""
    tok = open(path).read().split()
    return np.array([f32(t) for t in tok[:9]], f32).reshape(3, 3)


def read_evd(path):
    """detector/Reader.cpp:219-268: CSV with a header line, first four columns."""
    rows = []
    with open(path) as fh:
        next(fh)
        for ln in fh:
            tok = ln.strip().split(",")
            if len(tok) >= 4:
                rows.append([f32(t) for t in tok[:4]])
    return np.array(rows, f32)


def gt_inl_column(path):
    out = {}
    with open(path) as fh:
        rd = csv.reader(fh)
        seen = False
        for row in rd:
            if not row:
                continue
            if row[0] == "Filename":
                seen = True
                continue
            if seen and len(row) > 1:
                out[row[0]] = int(float(row[1]))
    return out


def main():
    names, kinds, thrs, expected, pts_all, models = [], [], [], [], [], []
    bad = 0

    gt = gt_inl_column(f"{REF}/results/homography/uniform_gc_Grid_c_sz_50.csv")
    for n, exp in gt.items():
        pts = read_pts_txt(f"{REF}/dataset/homography/sift_update/{n}_pts.txt")
        H = read_mat3(f"{REF}/dataset/homography/{n}_model.txt")
        got = max(h_count(pts, H, 2), h_count(pts, cv_inv3(H), 2))
        ok = got == exp
        bad += not ok
        print(f"homogr {n:14s} N={len(pts):5d} expected {exp:5d} numpy {got:5d} {'ok' if ok else 'MISMATCH'}")
        names.append("homogr/" + n); kinds.append(0); thrs.append(2.0); expected.append(exp); pts_all.append(pts); models.append(H)

    gt = gt_inl_column(f"{REF}/results/EVD/uniform_gc_Nanoflann_c_sz_50.csv")
    for n, exp in gt.items():
        pts = read_evd(f"{REF}/dataset/EVD/EVD_tentatives/{n}.png_m.txt")
        H = read_mat3(f"{REF}/dataset/EVD/h/{n}.txt")
        got = max(h_count(pts, H, 2), h_count(pts, cv_inv3(H), 2))
        ok = got == exp
        bad += not ok
        print(f"EVD    {n:14s} N={len(pts):5d} expected {exp:5d} numpy {got:5d} {'ok' if ok else 'MISMATCH'}")
        names.append("EVD/" + n); kinds.append(0); thrs.append(2.0); expected.append(exp); pts_all.append(pts); models.append(H)

    fdir = f"{REF}/results/fundamental"
    for fn in sorted(os.listdir(fdir)):
        if fn == "ALL.csv" or not fn.endswith(".csv"):
            continue
        n = fn[:-4]
        lines = open(f"{fdir}/{fn}").read().split("\n")
        F = np.array([[f32(t) for t in lines[r].split(",")[:3]] for r in range(3)], f32)
        exp = int(lines[3].strip())
        raw = np.loadtxt(f"{REF}/dataset/fundamental/{n}.txt", dtype=np.float64)
        pts = raw[:, [0, 1, 3, 4]].astype(f32)
        got = int(np.sum(sampson_err(pts, F) < f32(3)))
        ok = got == exp
        bad += not ok
        print(f"F      {n:18s} N={len(pts):5d} expected {exp:5d} numpy {got:5d} {'ok' if ok else 'MISMATCH'}")
        names.append("fundamental/" + n); kinds.append(1); thrs.append(3.0); expected.append(exp); pts_all.append(pts); models.append(F)

    offs = np.cumsum([0] + [len(p) for p in pts_all]).astype(np.int64)
    np.savez_compressed(
        f"{OUT}/scoring_kat.npz",
        names=np.array(names), kind=np.array(kinds, np.int32), threshold=np.array(thrs, f32),
        expected=np.array(expected, np.int32), offsets=offs, points=np.concatenate(pts_all).astype(f32),
        models=np.stack(models).astype(f32))
    print(f"wrote scoring_kat.npz: {len(names)} vectors, {offs[-1]} points, {bad} mismatches")

    # line2d ground-truth files (dataset/GetImage.h: width height noise a b c N then N rows) - no published count.
    ldir = f"{REF}/dataset/line2d"
    lnames, lmodels, lpts = [], [], []
    for fn in sorted(os.listdir(ldir)):
        if not fn.endswith(".txt") or fn == "dataset.txt":
            continue
        tok = open(f"{ldir}/{fn}").read().split()
        a, b, c = (f32(t) for t in tok[3:6])
        n = int(tok[6])
        p = np.array([f32(t) for t in tok[7:7 + 2 * n]], f32).reshape(n, 2)
        lnames.append(fn[:-4]); lmodels.append([a, b, c]); lpts.append(p)
    loffs = np.cumsum([0] + [len(p) for p in lpts]).astype(np.int64)
    np.savez_compressed(f"{OUT}/line2d_gt.npz", names=np.array(lnames), models=np.array(lmodels, f32), offsets=loffs,
                        points=np.concatenate(lpts).astype(f32))
    print(f"wrote line2d_gt.npz: {len(lnames)} files, {loffs[-1]} points")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
