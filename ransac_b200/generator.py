"""Deterministic synthetic inputs for the five BASELINE.json configurations.

The reference only generates 2-D line data (generator/generator.cpp:98-148); the correspondence generators below
are new. Every generator returns (points float32 [N, 2|4], gt_model float32, gt_inlier_mask).
Seeds are explicit; numpy's Philox bit generator keeps the streams reproducible across machines.
"""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.Philox(seed))


def line2d(n=1000, inlier_ratio=0.5, width=1000, height=1000, noise=3.0, seed=0xC0FFEE + 1):
    """Config 1. Follows Generate2DLinePoints (generator/generator.cpp:98-148): random orientation through the image
    centre, outliers uniform in the image, inliers at centre + t*tangent*diag (t in [-.5,.5], rejected outside the
    image) with `normal*noise*u - noise/2` added per coordinate. Outliers first, inliers after (generator.cpp:116-146)."""
    g = _rng(seed)
    n_in = int(round(n * inlier_ratio))
    n_out = n - n_in
    alpha = np.float32(np.pi) * np.float32(g.random())
    nx, ny = np.float32(np.sin(alpha)), np.float32(np.cos(alpha))
    tx, ty = -ny, nx
    cx, cy = np.float32(width // 2), np.float32(height // 2)
    c = -(nx * cx + ny * cy)
    pts = np.empty((n, 2), np.float32)
    pts[:n_out, 0] = width * g.random(n_out)
    pts[:n_out, 1] = height * g.random(n_out)
    diag = np.float32(np.sqrt(width * width + height * height))
    i = n_out
    while i < n:
        t = np.float32(g.random()) - np.float32(0.5)
        x = cx + t * tx * diag
        if x < 0 or x > width:
            continue
        y = cy + t * ty * diag
        if y < 0 or y > height:
            continue
        x = x + nx * np.float32(noise) * np.float32(g.random()) - np.float32(noise / 2)
        y = y + ny * np.float32(noise) * np.float32(g.random()) - np.float32(noise / 2)
        pts[i] = (x, y)
        i += 1
    mask = np.zeros(n, bool)
    mask[n_out:] = True
    return pts, np.array([nx, ny, c], np.float32), mask


def _random_homography(g, size):
    ang = np.deg2rad(g.uniform(-15, 15))
    s = g.uniform(0.8, 1.2)
    tx, ty = g.uniform(-100, 100, 2)
    c = size / 2.0
    R = np.array([[s * np.cos(ang), -s * np.sin(ang), 0], [s * np.sin(ang), s * np.cos(ang), 0], [0, 0, 1.0]])
    T0 = np.array([[1, 0, -c], [0, 1, -c], [0, 0, 1.0]])
    T1 = np.array([[1, 0, c + tx], [0, 1, c + ty], [0, 0, 1.0]])
    P = np.eye(3)
    P[2, 0], P[2, 1] = g.uniform(-1e-4, 1e-4, 2)
    H = T1 @ P @ R @ T0
    return H / H[2, 2]


def homography(n=4000, inlier_ratio=0.3, size=1000.0, noise=0.5, seed=0xC0FFEE + 2, clustered=False, shuffle=True):
    """Configs 2 and 5. Mild projective GT (rotation <=15 deg, scale .8-1.2, translation <=100 px, h31,h32 ~1e-4);
    inliers p2 = pi(H p1) + N(0, noise); outliers uniform and independent in both images. With clustered=True the
    inliers concentrate in a few spatial blobs so that grid/NAPSAC neighbourhoods are inlier-rich (config 5)."""
    g = _rng(seed)
    H = _random_homography(g, size)
    n_in = int(round(n * inlier_ratio))
    if clustered:
        n_blobs = 12
        centres = g.uniform(0.2 * size, 0.8 * size, (n_blobs, 2))
        which = g.integers(0, n_blobs, n_in)
        p1_in = centres[which] + g.normal(0, 0.03 * size, (n_in, 2))
        p1_in = np.clip(p1_in, 0, size)
    else:
        p1_in = g.uniform(0, size, (n_in, 2))
    ph = np.c_[p1_in, np.ones(n_in)] @ H.T
    p2_in = ph[:, :2] / ph[:, 2:3] + g.normal(0, noise, (n_in, 2))
    p_out = g.uniform(0, size, (n - n_in, 4))
    pts = np.r_[np.c_[p1_in, p2_in], p_out].astype(np.float32)
    mask = np.zeros(n, bool)
    mask[:n_in] = True
    if shuffle:
        perm = g.permutation(n)
        pts, mask = pts[perm], mask[perm]
    return np.ascontiguousarray(pts), H.astype(np.float32), mask


def _two_view_scene(g, n_in, f=800.0, size=1000.0, noise=0.5):
    """3-D points at depth 4-12 seen by two pinhole cameras (baseline 1, rotation <= 20 deg)."""
    K = np.array([[f, 0, size / 2], [0, f, size / 2], [0, 0, 1.0]])
    ax = g.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = np.deg2rad(g.uniform(5, 20))
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    t = g.normal(size=3)
    t[2] *= 0.3
    t /= np.linalg.norm(t)
    X = np.empty((0, 3))
    x1 = np.empty((0, 2))
    x2 = np.empty((0, 2))
    while len(X) < n_in:
        z = g.uniform(4, 12, 2 * n_in)
        xy = g.uniform(-0.6, 0.6, (2 * n_in, 2)) * z[:, None]
        P = np.c_[xy, z]
        a = P @ K.T
        a = a[:, :2] / a[:, 2:3]
        Q = P @ R.T + t
        b = Q @ K.T
        ok = (Q[:, 2] > 0.5)
        b = b[:, :2] / b[:, 2:3]
        ok &= (a >= 0).all(1) & (a <= size).all(1) & (b >= 0).all(1) & (b <= size).all(1)
        X = np.r_[X, P[ok]]
        x1 = np.r_[x1, a[ok]]
        x2 = np.r_[x2, b[ok]]
    x1, x2 = x1[:n_in], x2[:n_in]
    x1 = x1 + g.normal(0, noise, x1.shape)
    x2 = x2 + g.normal(0, noise, x2.shape)
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = tx @ R
    return K, E, x1, x2


def fundamental(n=10000, inlier_ratio=0.25, size=1000.0, noise=0.5, seed=0xC0FFEE + 3, sort_by_quality=True):
    """Config 3. Rows are sorted by a synthetic match-quality score correlated with inlierness (PROSAC expects
    descending quality, prosac_sampler.hpp:75; the reference's real data is ordered by ascending Lowe ratio)."""
    g = _rng(seed)
    n_in = int(round(n * inlier_ratio))
    K, E, x1, x2 = _two_view_scene(g, n_in, size=size, noise=noise)
    Ki = np.linalg.inv(K)
    F = Ki.T @ E @ Ki
    F = F / F[2, 2]
    p_out = g.uniform(0, size, (n - n_in, 4))
    pts = np.r_[np.c_[x1, x2], p_out].astype(np.float32)
    mask = np.zeros(n, bool)
    mask[:n_in] = True
    ratio = np.where(mask, g.beta(2, 5, n) * 0.7, g.beta(5, 2, n) * 0.7 + 0.3 * g.random(n))   # Lowe-ratio-like
    order = np.argsort(ratio, kind="stable") if sort_by_quality else g.permutation(n)
    return np.ascontiguousarray(pts[order]), F.astype(np.float32), mask[order]


def essential(n=20000, inlier_ratio=0.2, size=1000.0, f=800.0, noise=0.5, seed=0xC0FFEE + 4):
    """Config 4. Same scene, coordinates calibrated with K^-1 (|x| <~ 1); use threshold ~ 2/f."""
    g = _rng(seed)
    n_in = int(round(n * inlier_ratio))
    K, E, x1, x2 = _two_view_scene(g, n_in, f=f, size=size, noise=noise)
    p_out = g.uniform(0, size, (n - n_in, 4))
    pix = np.r_[np.c_[x1, x2], p_out]
    Ki = np.linalg.inv(K)
    a = np.c_[pix[:, :2], np.ones(n)] @ Ki.T
    b = np.c_[pix[:, 2:], np.ones(n)] @ Ki.T
    pts = np.c_[a[:, :2], b[:, :2]].astype(np.float32)
    mask = np.zeros(n, bool)
    mask[:n_in] = True
    perm = g.permutation(n)
    E = E / np.linalg.norm(E)
    return np.ascontiguousarray(pts[perm]), E.astype(np.float32), mask[perm]


CONFIGS = {
    1: dict(name="C1 line2d N=1000 50% uniform", estimator="line2d", threshold=8.0, confidence=0.99),
    2: dict(name="C2 homography N=4000 30% uniform", estimator="homography", threshold=2.0, confidence=0.95),
    3: dict(name="C3 fundamental N=10000 25% PROSAC+SPRT", estimator="fundamental", threshold=2.0, confidence=0.95),
    4: dict(name="C4 essential N=20000 20% uniform+SPRT", estimator="essential", threshold=2.5e-3, confidence=0.95),
    5: dict(name="C5 homography N=1M 10% NAPSAC", estimator="homography", threshold=2.0, confidence=0.95),
}


def make(config_id, seed_offset=0, **kw):
    if config_id == 1:
        return line2d(seed=0xC0FFEE + 1 + seed_offset, **kw)
    if config_id == 2:
        return homography(seed=0xC0FFEE + 2 + seed_offset, **kw)
    if config_id == 3:
        return fundamental(seed=0xC0FFEE + 3 + seed_offset, **kw)
    if config_id == 4:
        return essential(seed=0xC0FFEE + 4 + seed_offset, **kw)
    if config_id == 5:
        kw.setdefault("n", 1000000)
        kw.setdefault("inlier_ratio", 0.1)
        kw.setdefault("clustered", True)
        return homography(seed=0xC0FFEE + 5 + seed_offset, **kw)
    raise ValueError(config_id)
