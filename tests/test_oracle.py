"""CPU tests that pin the oracle: reference known-answer vectors, OpenCV primitives, RNG streams, identities."""
import ctypes
import os

import numpy as np
import pytest

from oracle import oracle as O
from ransac_b200 import generator as gen


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "scoring_kat.npz"))


@pytest.fixture(scope="module")
def cvp(golden_dir):
    return np.load(os.path.join(golden_dir, "cv_primitives.npz"))


def test_scoring_known_answers(kat):
    """46 vectors: 12 homogr + 15 EVD homography counts (thr 2), 19 squared-Sampson counts (thr 3)."""
    offs = kat["offsets"]
    n_h = n_f = 0
    for i, name in enumerate(kat["names"]):
        pts = kat["points"][offs[i]:offs[i + 1]]
        model, thr, exp = kat["models"][i], float(kat["threshold"][i]), int(kat["expected"][i])
        if kat["kind"][i] == 0:
            c1 = O.score(O.EST_HOMOGRAPHY, pts, model, thr)[0]
            inv, _ = O.inv3x3(model)
            c2 = O.score(O.EST_HOMOGRAPHY, pts, inv, thr)[0]
            assert max(c1, c2) == exp, name      # dataset/GetImage.h:250-264 keeps the better of H / H^-1
            n_h += 1
        else:
            assert O.score(O.EST_FUNDAMENTAL, pts, model, thr)[0] == exp, name
            n_f += 1
    assert (n_h, n_f) == (27, 19)


def test_score_matches_errors_and_inlier_list(kat):
    offs = kat["offsets"]
    pts = kat["points"][offs[1]:offs[2]]
    model = kat["models"][1]
    e = O.errors(O.EST_HOMOGRAPHY, pts, model)
    cnt, s, flagged, ids = O.score(O.EST_HOMOGRAPHY, pts, model, 2.0, want_inliers=True)
    assert cnt == int((e < np.float32(2.0)).sum())
    assert np.array_equal(ids, np.where(e < np.float32(2.0))[0])
    seq = np.float32(0)
    for v in e[ids]:
        seq = np.float32(seq + v)
    assert s == seq                                  # sequential float32 sum, quality.hpp:81-96
    assert flagged == int((np.abs(e.astype(np.float64) - 2.0) <= 2e-6).sum())


def test_glibc_random_stream():
    assert O.glibc_random(1, 5) == [1804289383, 846930886, 1681692777, 1714636915, 1957747793]
    assert O.glibc_random(12345, 3) == [383100999, 858300821, 357768173]
    libc = ctypes.CDLL("libc.so.6")
    libc.random.restype = ctypes.c_long
    for seed in (1, 7, 20261018):
        libc.srandom(seed)
        assert O.glibc_random(seed, 2000) == [libc.random() for _ in range(2000)]


def test_philox_known_answers():
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_inv3x3_bit_exact_vs_cv2(cvp):
    for m, ref in zip(cvp["inv_in"], cvp["inv_out"]):
        got, _ = O.inv3x3(m)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_cubic_vs_cv2(cvp):
    """Same root count and order as cv::solveCubic; same values wherever cv's trigonometric formula is itself accurate
    (it loses small roots when the roots differ by many orders of magnitude - the bisection here does not)."""
    def rel_residual(c, r):
        pw = np.array([r ** 3, r ** 2, r, 1.0])
        return abs(np.dot(c, pw)) / max(np.dot(np.abs(c), np.abs(pw)), 1e-300)

    compared = 0
    for c, ref, n in zip(cvp["cubic_in"], cvp["cubic_out"], cvp["cubic_n"]):
        k, r = O.solve_cubic(c)
        assert k == n
        for got, exp in zip(r[:k], ref[:n]):
            assert rel_residual(c, got) < 1e-9, (c, got)
            if rel_residual(c, exp) < 1e-12:
                compared += 1
                assert abs(got - exp) <= 1e-6 * max(1.0, abs(exp)), (c, got, exp)
    assert compared > 600


def test_cubic_degenerate_branches():
    assert O.solve_cubic([0, 0, 2, -4]) == (1, [2.0, 0.0, 0.0])          # linear
    k, r = O.solve_cubic([0, 1, -3, 2])                                   # quadratic
    assert k == 2 and sorted(r[:2]) == [1.0, 2.0]
    assert O.solve_cubic([0, 1, 0, 1])[0] == 0
    assert O.solve_cubic([0, 0, 0, 1])[0] == 0 and O.solve_cubic([0, 0, 0, 0])[0] == -1


def _unit(h):
    h = np.asarray(h, np.float64).ravel()
    h = h / np.linalg.norm(h)
    return h if h[np.argmax(np.abs(h))] > 0 else -h


def test_h4_solver_vs_cv2_null_vector(cvp):
    idx = np.arange(4, dtype=np.int32)
    worst = 0
    for pts, H in zip(cvp["h4_pts"], cvp["h4_H"]):
        got = O.solve_minimal(O.EST_HOMOGRAPHY, pts, idx)
        assert got.shape == (1, 9)
        worst = max(worst, np.abs(_unit(got[0]) - _unit(H)).max())
        # the model annihilates its own sample: transfer error ~ 0
        assert O.errors(O.EST_HOMOGRAPHY, pts, got[0]).max() < 1e-2
    assert worst < 1e-4


def test_f7_solver_vs_cv2(cvp):
    idx = np.arange(7, dtype=np.int32)
    matched = total = 0
    for pts, Fs, n in zip(cvp["f7_pts"], cvp["f7_F"], cvp["f7_n"]):
        got = O.solve_minimal(O.EST_FUNDAMENTAL, pts, idx)   # only roots passing the oriented-epipolar test
        for g in got:
            F = g.reshape(3, 3).astype(np.float64)
            assert abs(np.linalg.det(F / np.linalg.norm(F))) < 1e-6
            x1 = np.c_[pts[:, :2], np.ones(7)]
            x2 = np.c_[pts[:, 2:], np.ones(7)]
            assert np.abs(np.einsum("ij,jk,ik->i", x2, F / np.linalg.norm(F), x1)).max() < 1e-3
            d = min(np.abs(_unit(g) - _unit(Fs[k])).max() for k in range(n))
            total += 1
            matched += d < 1e-4
    assert total > 100 and matched >= 0.95 * total     # cv2 normalises the points first; conditioning differs


def test_null_space_annihilates():
    g = np.random.default_rng(3)
    for rows in (5, 7, 8):
        A = g.normal(size=(rows, 9))
        B = O.null_space(A)
        assert B.shape == (9 - rows, 9)
        assert np.abs(A @ B.T).max() < 1e-12
    assert O.null_space(np.zeros((8, 9))) is None


def test_line_solver():
    pts = np.array([[0, 0], [10, 10], [3, 7]], np.float32)
    m = O.solve_minimal(O.EST_LINE2D, pts, np.array([0, 1], np.int32))[0]
    assert np.allclose(m, [-2 ** -0.5, 2 ** -0.5, 0], atol=1e-6)
    assert np.allclose(O.errors(O.EST_LINE2D, pts, m), [0, 0, 4 / 2 ** 0.5], atol=1e-5)


def test_uniform_sampler_replays_reference_stream():
    """uniform_sampler.hpp:42-54: persistent shrinking pool over glibc random()."""
    n, m = 50, 4
    s = O.Sampler(O.SAMPLER_UNIFORM, O.RNG_GLIBC, n, m, 1)
    stream = iter(O.glibc_random(1, 1000))
    pool, mx = list(range(n)), n
    for _ in range(40):        # crosses the max==0 wrap
        exp = []
        for _ in range(m):
            if mx == 0:
                mx = n
            i = next(stream) % mx
            v = pool[i]
            mx -= 1
            pool[i], pool[mx] = pool[mx], v
            exp.append(v)
        assert list(s.generate()) == exp


def test_philox_sampler_unique_and_uniform():
    n, m = 37, 7
    seen = np.zeros(n)
    for h in range(4000):
        s = O.philox_unique(99, h, 0, n, m)
        assert len(set(s.tolist())) == m and s.min() >= 0 and s.max() < n
        seen[s] += 1
    assert seen.min() > 0.8 * seen.mean() and seen.max() < 1.2 * seen.mean()
    assert np.array_equal(O.philox_unique(99, 5, 0, n, m), O.philox_unique(99, 5, 0, n, m))
    assert not np.array_equal(O.philox_unique(99, 5, 0, n, m), O.philox_unique(100, 5, 0, n, m))


def test_prosac_sampler_growth():
    n, m = 500, 4
    s = O.Sampler(O.SAMPLER_PROSAC, O.RNG_PHILOX, n, m, 5)
    g = s.growth(n)
    assert (g[:m] == 1).all() and (np.diff(g.astype(np.int64)) >= 0).all()
    first = s.generate(0)
    assert sorted(first.tolist()) == [0, 1, 2, 3]               # t=1: the m best-ranked points
    prev_top = 3
    for h in range(1, 300):
        smp = s.generate(h)
        assert len(set(smp.tolist())) == m
        assert smp[-1] >= prev_top and smp[:-1].max() < smp[-1]  # newest point + m-1 from the prefix
        prev_top = smp[-1]
    s.set_termination_length(10)
    for h in range(300, 330):
        assert s.generate(h).max() <= 10                         # closed range [0, termination_length]


def test_grid_cells_and_napsac():
    pts, _, _ = gen.homography(n=3000, seed=5)
    cell, members, start = O.grid_cells(pts, 200)
    key = (pts / 200).astype(np.int32)                           # truncation toward zero, nearest_neighbors.cpp:172
    for c in range(len(start) - 1):
        mem = members[start[c]:start[c + 1]]
        assert (np.diff(mem) > 0).all()
        assert (key[mem] == key[mem[0]]).all(1).all()
    assert sorted(members.tolist()) == list(range(3000))
    s = O.Sampler(O.SAMPLER_NAPSAC, O.RNG_PHILOX, 3000, 4, 11, points=pts, cell_size=200)
    for h in range(200):
        smp = s.generate(h)
        assert (cell[smp] == cell[smp[0]]).all() and smp[0] not in smp[1:]


def test_napsac_knn_cursor():
    n, k, m = 20, 5, 4
    table = np.array([[(p + j + 1) % n for j in range(k)] for p in range(n)], np.int32)
    s = O.Sampler(O.SAMPLER_NAPSAC, O.RNG_GLIBC, n, m, 1, knn_table=table)
    uses = {}
    for _ in range(60):
        smp = s.generate()
        p = int(smp[0])
        c = uses.get(p, 0)
        exp = [table[p][k - 1 - ((c + i) % k)] for i in range(m - 1)]   # farthest first, cyclic (napsac_sampler.hpp:82-91)
        assert smp[1:].tolist() == exp
        uses[p] = c + m - 1


def test_knn_build_vs_kdtree_and_ties():
    """orc_knn_build (nearest_neighbors.cpp:69-128): the k+1 nearest minus the first. Pinned against an independent exact
    KD-tree (scipy cKDTree, float64 - nanoflann itself is not available) on tie-free data; tie rule = ascending index."""
    from scipy.spatial import cKDTree
    g = np.random.default_rng(4)
    for dim, n, k in ((4, 3000, 5), (2, 1200, 9), (4, 64, 8)):
        pts = g.uniform(0, 1000, (n, dim)).astype(np.float32)
        table = O.knn_build(pts, k)
        _, ref = cKDTree(pts.astype(np.float64)).query(pts.astype(np.float64), k=k + 1)
        assert np.array_equal(table, ref[:, 1:].astype(np.int32))
    lattice = np.stack(np.meshgrid(np.arange(6.0), np.arange(6.0), [0.0], [0.0]), -1).reshape(-1, 4).astype(np.float32)
    t = O.knn_build(lattice, 4)
    assert t[0].tolist() == [1, 6, 7, 2]                     # d^2 = 1, 1, 2, 4 (tie 2 vs 12 -> lower index)
    dup = np.concatenate([lattice, lattice[:3]])
    t = O.knn_build(dup, 2)
    assert t[36].tolist() == [36, 1]                         # its duplicate (point 0) takes rank 0; the query itself is kept
    with pytest.raises(ValueError):
        O.knn_build(lattice[:4], 4)


def test_standard_termination():
    """standard_termination_criteria.hpp:52-62 (float32 power by repeated multiply, truncation)."""
    assert O.standard_termination(1200, 4000, 4, 0.95, 10000) == int(np.log(np.float32(0.05)) / np.log(1 - np.float32(0.3) ** 4))
    assert O.standard_termination(100, 4000, 4, 0.95, 10000) == 10000        # w^m < 0.0005
    assert O.standard_termination(2500, 10000, 7, 0.95, 10000) == 10000      # SURVEY finding 9
    assert O.standard_termination(500, 1000, 2, 0.99, 10000) == 16
    assert O.standard_termination(4000, 4000, 4, 0.95, 10000) == 0


@pytest.mark.parametrize("est,cfg", [(O.EST_HOMOGRAPHY, 2), (O.EST_LINE2D, 1)])
def test_batched_equals_sequential_without_sprt(est, cfg):
    pts, _, mask = gen.make(cfg, n=1500) if cfg == 2 else gen.make(cfg)
    thr, conf = gen.CONFIGS[cfg]["threshold"], gen.CONFIGS[cfg]["confidence"]
    for rng in (O.RNG_PHILOX, O.RNG_GLIBC):
        seq = O.ransac(pts, est, rng=rng, threshold=thr, confidence=conf, seed=3)
        for K in (1, 7, 64, 1000):
            b = O.ransac(pts, est, rng=rng, threshold=thr, confidence=conf, seed=3, batch=K)
            for key in ("inliers", "score", "iterations", "best_hyp"):
                assert b[key] == seq[key], (rng, K, key)
            assert np.array_equal(b["model"], seq["model"])
        assert seq["inliers"] >= 0.7 * mask.sum()


def test_sprt_batched_one_equals_sequential():
    pts, _, mask = gen.homography(n=2000, seed=9)
    seq = O.ransac(pts, O.EST_HOMOGRAPHY, rng=O.RNG_GLIBC, sprt=True, seed=1)
    b1 = O.ransac(pts, O.EST_HOMOGRAPHY, rng=O.RNG_GLIBC, sprt=True, seed=1, batch=1)
    # batched(1) differs from the reference only in where a hypothesis starts in the shuffled pool
    # (cursor + 32*q instead of "where the previous one stopped"), so compare outcomes, not traces
    for r in (seq, b1):
        assert r["inliers"] >= 0.6 * mask.sum()
        assert r["evals"] < 0.5 * r["iterations"] * 2000          # SPRT rejects early
    bk = O.ransac(pts, O.EST_HOMOGRAPHY, rng=O.RNG_PHILOX, sprt=True, seed=1, batch=256)
    assert bk["inliers"] >= 0.6 * mask.sum()


def test_table_sampler_equals_philox_run():
    pts, _, _ = gen.homography(n=1000, seed=4)
    s = O.Sampler(O.SAMPLER_UNIFORM, O.RNG_PHILOX, 1000, 4, 21)
    table = s.table(2000)
    a = O.ransac(pts, O.EST_HOMOGRAPHY, rng=O.RNG_PHILOX, seed=21)
    b = O.ransac(pts, O.EST_HOMOGRAPHY, rng=O.RNG_TABLE, sample_table=table)
    assert a["inliers"] == b["inliers"] and a["iterations"] == b["iterations"] and np.array_equal(a["model"], b["model"])


def test_prosac_run_finds_model_early():
    pts, F, mask = gen.fundamental(n=3000, seed=8)
    r = O.ransac(pts, O.EST_FUNDAMENTAL, sampler=O.SAMPLER_PROSAC, threshold=2.0, seed=2, max_iterations=10000)
    assert r["inliers"] >= 0.5 * mask.sum()
    assert r["iterations"] < 10000                                  # PROSAC's prefix criterion stops early


# ---- five-point essential solver -------------------------------------------------------------------------------------
def _unit(E):
    E = np.asarray(E, np.float64).reshape(3, 3)
    E = E / np.linalg.norm(E)
    return E * np.sign(E.ravel()[np.argmax(np.abs(E))])


def test_essential5_candidates_match_opencv_solution_sets(golden_dir):
    """The oracle's candidate set equals the set of real solutions OpenCV's five-point solver returns (scale/sign free) on
    at least 57 of the 60 samples; on the rest (nearly double roots) every oracle candidate still has to be a solution of
    the five-point system - nothing spurious is ever returned."""
    z = np.load(os.path.join(golden_dir, "essential5_cv.npz"))
    dist = lambda a, b: min(np.abs(a - b).max(), np.abs(a + b).max())   # noqa: E731
    exact, total = 0, 0
    for pts, sols, k in zip(z["points"], z["solutions"], z["counts"]):
        Es, valid = O.essential5_candidates(pts, np.arange(5, dtype=np.int32))
        cv_set = [_unit(sols[i]) for i in range(k)]
        mine = [_unit(E) for E in Es]
        p = pts.astype(np.float64)
        for En in mine:
            res = [np.array([x2, y2, 1.0]) @ En @ np.array([x1, y1, 1.0]) for x1, y1, x2, y2 in p]
            assert np.abs(res).max() < 1e-9 and np.abs(2 * En @ En.T @ En - np.trace(En @ En.T) * En).max() < 1e-7
        ok = all(min(dist(c, m) for m in mine) < 1e-6 for c in cv_set) and all(min(dist(c, m) for c in cv_set) < 1e-6 for m in mine)
        exact += ok
        total += k
    assert exact >= 57 and total > 200


def test_essential5_identities_and_selection():
    from ransac_b200 import generator as gen
    pts, E_gt, mask = gen.essential(n=3000, noise=0.0, seed=21)
    inl = np.where(mask)[0]
    g = np.random.default_rng(3)
    found = 0
    for _ in range(100):
        s = g.choice(inl, 5, replace=False).astype(np.int32)
        Es, valid = O.essential5_candidates(pts, s)
        m = O.solve_minimal(O.EST_ESSENTIAL, pts, s)
        if valid.any():                                               # EstimateModel returns the FIRST candidate that passes the vote
            assert len(m) == 1
            first = Es[np.argmax(valid)]
            assert np.array_equal(m[0], first.astype(np.float32).ravel())
        else:
            assert len(m) == 0
        for E in Es:
            En = E / np.linalg.norm(E)
            p = pts[s].astype(np.float64)
            res = [np.array([x2, y2, 1.0]) @ En @ np.array([x1, y1, 1.0]) for x1, y1, x2, y2 in p]
            assert np.abs(res).max() < 1e-9                           # annihilates its own sample
            assert abs(np.linalg.det(En)) < 1e-9 and np.abs(2 * En @ En.T @ En - np.trace(En @ En.T) * En).max() < 1e-9
        d = [np.abs(_unit(E) - _unit(E_gt)).max() for E in Es]
        if d and min(d) < 1e-5:
            found += 1
            assert valid[int(np.argmin(d))]                           # the true E passes the cheirality vote
    assert found >= 95


# ---- non-minimal estimation and the final refit (ransac.cpp:157-207) ---------------------------------------------------------
def test_nonminimal_recovers_ground_truth_on_noise_free_inliers():
    from ransac_b200 import generator as gen

    def dist(a, b):
        a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
        a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
        return min(np.abs(a - b).max(), np.abs(a + b).max())
    for est, make, tol in ((O.EST_HOMOGRAPHY, gen.homography, 1e-7), (O.EST_FUNDAMENTAL, gen.fundamental, 1e-6), (O.EST_ESSENTIAL, gen.essential, 1e-5)):
        pts, gt, mask = make(n=3000, noise=0.0, seed=3)
        m = O.nonminimal(est, pts, np.where(mask)[0])
        assert m is not None and dist(m, gt) < tol
    pts, line, mask = gen.line2d(n=2000, noise=0.0, seed=3)
    m = O.nonminimal(O.EST_LINE2D, pts, np.where(mask)[0])
    assert dist(m[:2], line[:2]) < 1e-5 and abs(abs(m[2]) - abs(line[2])) < 1e-2
    assert O.nonminimal(O.EST_HOMOGRAPHY, pts.repeat(2, axis=1), np.arange(3)) is None     # fewer than four points


def test_refit_loop_improves_or_keeps_the_minimal_model():
    from ransac_b200 import generator as gen
    for seed in (4, 5, 6):
        pts, H, mask = gen.homography(n=4000, seed=seed)
        r = O.ransac(pts, O.EST_HOMOGRAPHY, seed=1)
        rf = O.refit(O.EST_HOMOGRAPHY, pts, r["model"], r["inliers"], 2.0)
        assert rf["inliers"] >= r["inliers"] and 0 <= rf["accepted"] <= 4
        assert len(rf["ids"]) == O.score(O.EST_HOMOGRAPHY, pts, rf["model"], 2.0)[0]
        assert rf["inliers"] > 0.97 * mask.sum()                      # noise 0.5 px, threshold 2 px: essentially every true inlier


def test_local_optimisation_lifts_minimal_models():
    """LO-RANSAC in the oracle (inner_local_optimization.hpp:74-133): more inliers, not more iterations; counters as RansacOutput reports."""
    from ransac_b200 import generator as gen
    pts, H, mask = gen.homography(n=4000, seed=7)
    base = O.ransac(pts, O.EST_HOMOGRAPHY, seed=3)
    for lo in (1, 2):
        r = O.ransac(pts, O.EST_HOMOGRAPHY, seed=3, lo=lo)
        assert r["inliers"] >= base["inliers"] and r["inliers"] >= 0.97 * mask.sum()
        assert r["iterations"] <= base["iterations"] and r["lo_inner"] >= 20 and r["lo_iterative"] >= r["lo_inner"]
    pts, F, mask = gen.fundamental(n=4000, seed=7)
    base = O.ransac(pts, O.EST_FUNDAMENTAL, seed=3, max_iterations=2000)
    r = O.ransac(pts, O.EST_FUNDAMENTAL, seed=3, max_iterations=2000, lo=1)
    assert r["inliers"] > base["inliers"]


def test_reference_dlt4p_switch_matches_opencv(golden_dir):
    """SURVEY Appendix B quirk 1 as a documented switch: orc_solve_homography_dlt4p_thin restates the reference's OWN minimal
    homography solver (dlt.cpp:7-53: raw pixel coordinates, float32 system, last row of the THIN SVD = the 8th singular vector).
    Pinned against that code path on the real OpenCV: tests/golden/dlt4p_cv.npz holds cv2.SVDecomp's row for 256 samples
    (make_dlt4p_golden.py). sigma_8 / sigma_1 ~ 1e-7, so agreement is to float32 conditioning: median 1e-5, worst case < 2 %.
    The vector is NOT the null vector: the model misses its own four points by pixels, where the default solver (normalised DLT,
    true null vector: what BASELINE.json names and the GPU computes) passes through them."""
    d = np.load(os.path.join(golden_dir, "dlt4p_cv.npz"))
    s = np.arange(4, dtype=np.int32)
    rel, own_thin, own_null = [], [], []
    for pts, H in zip(d["pts"], d["H"]):
        o = O.solve_homography_dlt4p_thin(pts, s)
        assert len(o) == 1
        rel.append(np.abs(o[0] - H).max() / np.abs(H).max())
        own_thin.append(O.errors(O.EST_HOMOGRAPHY, pts, o[0]).max())
        m = O.solve_minimal(O.EST_HOMOGRAPHY, pts, s)
        if len(m):
            own_null.append(O.errors(O.EST_HOMOGRAPHY, pts, m[0]).max())
    rel = np.array(rel)
    assert np.median(rel) < 1e-4 and np.quantile(rel, 0.9) < 1e-3 and rel.max() < 2e-2
    assert np.median(own_thin[:128]) > 0.5 and np.median(own_null[:128]) < 1e-2


def test_reference_dlt4p_switch_in_the_whole_fit():
    """orc_config::ref_thin_svd runs Ransac::run with the reference's solver: the termination criterion never fires (no sample model
    reaches the inlier count of a correct one), while the default stops early with the full inlier set."""
    pts, H, mask = gen.make(2)
    a = O.ransac(pts, O.EST_HOMOGRAPHY, threshold=2.0, confidence=0.95, max_iterations=2000, seed=5, ref_thin_svd=True)
    b = O.ransac(pts, O.EST_HOMOGRAPHY, threshold=2.0, confidence=0.95, max_iterations=2000, seed=5)
    assert a["iterations"] == 2000 and a["inliers"] < 0.5 * mask.sum()
    assert b["iterations"] < 1000 and b["inliers"] > 0.75 * mask.sum()


def test_sprt_logwalk_model_agrees_with_the_sequential_chain():
    """The log-domain decision rule of the SPRT tail kernel (sprt.cuh, USAC_SPRT_LOGWALK), restated step for step on the CPU, against
    the sequential chain of double multiplications of sprt.hpp:205-234: every walk the rule decides gives the chain's decision,
    tested points and tested inliers (tools/sprt_logwalk_sim.py exits non-zero on any difference)."""
    import subprocess
    import sys
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sprt_logwalk_sim.py")
    r = subprocess.run([sys.executable, tool, "11", "500"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatches 0" in r.stdout
