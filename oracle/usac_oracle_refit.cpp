/*
 * usac_oracle_refit.cpp - non-minimal estimation and the final refit loop of the CPU oracle. TEST INFRASTRUCTURE ONLY.
 *
 * Restates Estimator::EstimateModelNonMinimalSample for the four estimators
 *   homography   homography_estimator.hpp:67-75 -> DLt::NormalizedDLT dlt/normalized_dlt.cpp:7-23, DLT dlt/dlt.cpp:55-101,
 *                GetNormalizingTransformation dlt/normalizing_transformation.cpp:7-112
 *   fundamental  fundamental_estimator.hpp:65-75 -> EightPointsAlgorithm fundamental/eight_points.cpp:4-100 (no rank-2 step, :47-70)
 *   essential    essential_estimator.hpp:64-74 (the same eight-point solver, no essential-constraint projection)
 *   line2d       line2d_estimator.hpp:59-106 (PCA; `sum_xy` is uninitialised there, :70 - zero here)
 * and the loop that follows the main loop of Ransac::run (usac/ransac/ransac.cpp:157-207): up to four rounds of
 * "estimate from all inliers, re-score, keep if it did not lose more than 20 % of the inliers and improved".
 *
 * Deterministic choices shared with the CUDA path (the reference uses sequential float sums and cv::SVD on A):
 *  - every sum over points is a "lane sum": lane t of 256 adds elements t, t+256, ... in order, then the 256 partial sums are
 *    combined by a fixed binary tree (stride 128, 64, ..., 1) - the order a 256-thread block reduces in;
 *  - the null vector of A is the eigenvector of the smallest eigenvalue of A'A (A in float like the reference, A'A in double),
 *    found by cyclic Jacobi sweeps (at most 12; rotations below 1e-17 relative are skipped, a sweep without a rotation ends the loop);  - H = T2^-1 Hn T1 and F = T2' Fn T1 are evaluated in double, then rounded to float.
 */
#include "oracle_internal.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

const int LANES = 256;

template <class T, class F>
T lane_sum(int n, F value) {
    T part[LANES];
    for (int t = 0; t < LANES; t++) {
        T acc = 0;
        for (int i = t; i < n; i += LANES) acc = acc + value(i);
        part[t] = acc;
    }
    for (int s = LANES / 2; s > 0; s >>= 1)
        for (int t = 0; t < s; t++) part[t] = part[t] + part[t + s];
    return part[0];
}

/* eigenvector of the smallest eigenvalue of the symmetric n x n matrix S (row-major, destroyed; n odd): Jacobi rotations in the
 * round-robin order - round r = 0..n-1 rotates the (n-1)/2 DISJOINT pairs {(r+k) mod n, (r-k) mod n}, k = 1..(n-1)/2 (index r sits
 * out; every pair occurs once per sweep). The angles of a round come from the matrix as it is at the start of the round, then
 * the column updates of all its pairs, then the row updates, then V - the order the device code uses, where the pairs of a
 * round run side by side. At most 12 sweeps; rotations with |apq| <= 1e-12 max(|app|, |aqq|) are skipped and the first sweep
 * without a rotation ends the loop. */
void smallest_eigenvector(double* S, int n, double* vec) {
    double V[81];
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j) ? 1.0 : 0.0;
    const int half = (n - 1) / 2;
    for (int sweep = 0; sweep < 12; sweep++) {
        int rotations = 0;                                        /* a sweep without a rotation: converged */
        for (int r = 0; r < n; r++) {
            int P[8], Q[8];
            double C[8], Sn[8];
            bool on[8];
            for (int k = 1; k <= half; k++) {
                int p = (r + k) % n, q = (r - k + n) % n;
                if (p > q) { const int t = p; p = q; q = t; }
                const int e = k - 1;
                P[e] = p; Q[e] = q; on[e] = false;
                const double apq = S[p * n + q];
                if (apq == 0.0) continue;
                {   /* |apq| <= 1e-12 max(|app|, |aqq|): what the EIGENVECTOR needs (it is rounded to float); a bound relative to
                     * sqrt(|app aqq|) would keep polishing the pairs of the near-null direction for the sake of its eigenvalue */
                    const double big = std::fmax(std::fabs(S[p * n + p]), std::fabs(S[q * n + q]));
                    if (apq * apq <= 1e-24 * (big * big)) continue;
                }
                /* tan of the rotation angle in the hypot form, t = sgn(d) b / (|d| + sqrt(d^2 + b^2)) with d = aqq - app, b = 2 apq,
                 * and c = sqrt(w) (1 / w), w = t^2 + 1: three dependent divisions / square roots instead of five */
                const double d = S[q * n + q] - S[p * n + p], b = 2.0 * apq;
                const double rr = std::sqrt(d * d + b * b);
                const double t0 = b / (std::fabs(d) + rr);
                const double t = d < 0.0 ? -t0 : t0;
                const double w = t * t + 1.0;
                const double c = std::sqrt(w) * (1.0 / w), s = t * c;
                if (!std::isfinite(c) || !std::isfinite(s)) continue;
                C[e] = c; Sn[e] = s; on[e] = true;
                rotations++;
            }
            for (int e = 0; e < half; e++) {                      /* columns p, q of S */
                if (!on[e]) continue;
                const int p = P[e], q = Q[e];
                const double c = C[e], s = Sn[e];
                for (int k = 0; k < n; k++) {
                    const double skp = S[k * n + p], skq = S[k * n + q];
                    S[k * n + p] = c * skp - s * skq;
                    S[k * n + q] = s * skp + c * skq;
                }
            }
            for (int e = 0; e < half; e++) {                      /* rows p, q of S; columns p, q of V */
                if (!on[e]) continue;
                const int p = P[e], q = Q[e];
                const double c = C[e], s = Sn[e];
                for (int k = 0; k < n; k++) {
                    const double spk = S[p * n + k], sqk = S[q * n + k];
                    S[p * n + k] = c * spk - s * sqk;
                    S[q * n + k] = s * spk + c * sqk;
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
        }
        if (!rotations) break;
    }
    int best = 0;
    for (int i = 1; i < n; i++) if (S[i * n + i] < S[best * n + best]) best = i;
    for (int k = 0; k < n; k++) vec[k] = V[k * n + best];
}

/* normalizing_transformation.cpp:7-112: T = [s 0 tx; 0 s ty; 0 0 1] per image, floats */
struct Norm { float s1, t1x, t1y, s2, t2x, t2y; };

bool normalizing(const float* pts, const int* ids, int n, Norm& T) {
    const float fn = (float)n;
    float m[4];
    for (int c = 0; c < 4; c++) m[c] = lane_sum<float>(n, [&](int i) { return pts[4 * (size_t)ids[i] + c]; }) / fn;
    const float d1 = lane_sum<float>(n, [&](int i) { const float* p = pts + 4 * (size_t)ids[i]; float a = p[0] - m[0], b = p[1] - m[1]; return sqrtf(a * a + b * b); });
    const float d2 = lane_sum<float>(n, [&](int i) { const float* p = pts + 4 * (size_t)ids[i]; float a = p[2] - m[2], b = p[3] - m[3]; return sqrtf(a * a + b * b); });
    T.s1 = (float)(M_SQRT2 / (double)(d1 / fn));
    T.s2 = (float)(M_SQRT2 / (double)(d2 / fn));
    T.t1x = -m[0] * T.s1; T.t1y = -m[1] * T.s1; T.t2x = -m[2] * T.s2; T.t2y = -m[3] * T.s2;
    return std::isfinite(T.s1) && std::isfinite(T.s2);
}

bool nonminimal_two_view(int est, const float* pts, const int* ids, int n, float* out) {
    Norm T;
    if (n < 4 || !normalizing(pts, ids, n, T)) return false;
    auto rows = [&](int i, float r[2][9]) {
        const float* p = pts + 4 * (size_t)ids[i];
        const float x1 = T.s1 * p[0] + T.t1x, y1 = T.s1 * p[1] + T.t1y, x2 = T.s2 * p[2] + T.t2x, y2 = T.s2 * p[3] + T.t2y;
        if (est == ORC_EST_HOMOGRAPHY) {                          /* dlt.cpp:66-88 */
            const float a[9] = {-x1, -y1, -1, 0, 0, 0, x2 * x1, x2 * y1, x2}, b[9] = {0, 0, 0, -x1, -y1, -1, y2 * x1, y2 * y1, y2};
            memcpy(r[0], a, sizeof(a)); memcpy(r[1], b, sizeof(b));
        } else {                                                  /* eight_points.cpp:21-35 */
            const float a[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1};
            memcpy(r[0], a, sizeof(a));
        }
    };
    const int nrows = est == ORC_EST_HOMOGRAPHY ? 2 : 1;
    double S[81];
    for (int i = 0; i < 9; i++)
        for (int j = i; j < 9; j++) {
            const double v = lane_sum<double>(n, [&](int k) {
                float r[2][9];
                rows(k, r);
                double acc = 0;
                for (int q = 0; q < nrows; q++) { double t = (double)r[q][i] * (double)r[q][j]; acc = acc + t; }
                return acc;
            });
            S[i * 9 + j] = v; S[j * 9 + i] = v;
        }
    double h[9];
    smallest_eigenvector(S, 9, h);
    double M[9], R[9];
    const double s1 = T.s1, t1x = T.t1x, t1y = T.t1y, s2 = T.s2, t2x = T.t2x, t2y = T.t2y;
    for (int i = 0; i < 3; i++) {                                 /* M = Xn * T1 */
        M[3 * i] = h[3 * i] * s1; M[3 * i + 1] = h[3 * i + 1] * s1;
        M[3 * i + 2] = (h[3 * i] * t1x + h[3 * i + 1] * t1y) + h[3 * i + 2];
    }
    if (est == ORC_EST_HOMOGRAPHY) {                              /* H = T2^-1 M, / h33 (normalized_dlt.cpp:18-20) */
        const double is2 = 1.0 / s2, ux = -(t2x * is2), uy = -(t2y * is2);
        for (int j = 0; j < 3; j++) { R[j] = is2 * M[j] + ux * M[6 + j]; R[3 + j] = is2 * M[3 + j] + uy * M[6 + j]; R[6 + j] = M[6 + j]; }
        const double inv = 1.0 / R[8];
        for (int i = 0; i < 9; i++) { const double v = R[i] * inv; if (!std::isfinite(v)) return false; out[i] = (float)v; }
        out[8] = 1.f;
    } else {                                                      /* F = T2' M, / f33 if |f33| > FLT_EPSILON (eight_points.cpp:72-97) */
        for (int j = 0; j < 3; j++) { R[j] = s2 * M[j]; R[3 + j] = s2 * M[3 + j]; R[6 + j] = (t2x * M[j] + t2y * M[3 + j]) + M[6 + j]; }
        const bool scale = std::fabs((float)R[8]) > FLT_EPSILON;
        const double inv = scale ? 1.0 / R[8] : 1.0;
        for (int i = 0; i < 9; i++) { const double v = R[i] * inv; if (!std::isfinite(v)) return false; out[i] = (float)v; }
    }
    return true;
}

bool nonminimal_line(const float* pts, const int* ids, int n, float* out) {     /* line2d_estimator.hpp:59-106 */
    if (n < 2) return false;
    const float fn = (float)n;
    const float sx = lane_sum<float>(n, [&](int i) { return pts[2 * (size_t)ids[i]]; });
    const float sy = lane_sum<float>(n, [&](int i) { return pts[2 * (size_t)ids[i] + 1]; });
    const float sxy = lane_sum<float>(n, [&](int i) { const float* p = pts + 2 * (size_t)ids[i]; return p[0] * p[1]; });
    const float sx2 = lane_sum<float>(n, [&](int i) { const float* p = pts + 2 * (size_t)ids[i]; return p[0] * p[0]; });
    const float sy2 = lane_sum<float>(n, [&](int i) { const float* p = pts + 2 * (size_t)ids[i]; return p[1] * p[1]; });
    const float mx = sx / fn, my = sy / fn;
    const float c00 = sx2 - 2 * sx * mx + fn * mx * mx;
    const float c01 = sxy - sx * my - sy * mx + fn * mx * my;
    const float c11 = sy2 - 2 * sy * my + fn * my * my;
    /* eigenvector of the smaller eigenvalue of [[c00 c01][c01 c11]] (the line normal), in double */
    const double p = c00, q = c01, r = c11;
    const double half = 0.5 * (p - r), rad = std::sqrt(half * half + q * q), lam = 0.5 * (p + r) - rad;
    double vx = q, vy = lam - p;
    const double wx = lam - r, wy = q;
    if (wx * wx + wy * wy > vx * vx + vy * vy) { vx = wx; vy = wy; }
    const double nn = std::sqrt(vx * vx + vy * vy);
    if (!(nn > 0.0)) { vx = 1.0; vy = 0.0; } else { vx = vx / nn; vy = vy / nn; }
    const float a = (float)vx, b = (float)vy;
    out[0] = a; out[1] = b; out[2] = -a * mx - b * my;
    return std::isfinite(out[0]) && std::isfinite(out[1]) && std::isfinite(out[2]);
}

}  // namespace

extern "C" int orc_nonminimal(int est, const float* pts, const int* ids, int n, float* model_out) {
    return est == ORC_EST_LINE2D ? nonminimal_line(pts, ids, n, model_out) : nonminimal_two_view(est, pts, ids, n, model_out);
}

/* ransac.cpp:157-207. model_io: best minimal model in, final model out; returns the final inlier count, the final inlier ids in
 * ids_out (>= n_points entries) and the number of refits that were accepted in *accepted_out. */
extern "C" int orc_refit(int est, const float* pts, int n_points, float thr, float* model_io, int best_inliers, int* ids_out, int* accepted_out) {
    const int w = est == ORC_EST_LINE2D ? 3 : 9;
    std::vector<int> cur((size_t)n_points);
    int cnt = 0;
    float sum = 0;
    orc_score(est, pts, n_points, model_io, thr, &cnt, &sum, cur.data(), 0, nullptr);          /* quality->getInliers, :163 */
    int prev = 0, accepted = 0, best = best_inliers < cnt ? best_inliers : cnt;   /* never read past the list */
    for (int norm = 0; norm < 4; norm++) {
        float m[9];
        if (!orc_nonminimal(est, pts, cur.data(), best, m)) break;                               /* :173 uses best_score->inlier_number ids */
        int c2 = 0;
        std::vector<int> ids2((size_t)n_points);
        orc_score(est, pts, n_points, m, thr, &c2, &sum, ids2.data(), 0, nullptr);               /* :180 */
        if ((float)c2 / best < 0.8f) break;                                                      /* :187 */
        if ((unsigned)c2 <= (unsigned)prev) break;                                               /* :195 */
        prev = c2;
        best = c2; cur.swap(ids2);
        memcpy(model_io, m, sizeof(float) * w);
        accepted++;
    }
    int fin = 0;
    orc_score(est, pts, n_points, model_io, thr, &fin, &sum, ids_out, 0, nullptr);               /* :210 */
    if (accepted_out) *accepted_out = accepted;
    return best;
}

/* ======================================================================================================
 * LO-RANSAC: inner + iterative local optimisation (local_optimization/inner_local_optimization.hpp:74-133,
 * iterative_local_optimization.hpp:61-135), kinds InItLORsc (unlimited) and InItFLORsc (limited samples).
 * Deterministic choices shared with the CUDA path: the random subsets come from Philox keyed by (seed, call counter) instead of
 * mt19937 seeded from random_device (uniform_random_generator.hpp:11-47); error sums are lane sums over 1024 lanes.
 * Quirks kept: lo_model->threshold is multiplied on every inner iteration, also after a `continue` that skipped the iterative
 * stage (inner_local_optimization.hpp:101-105); the threshold of a failed iterative stage is reset, a successful one ends at
 * theta again only up to float rounding.
 * ====================================================================================================== */
extern "C" void orc_philox_unique(uint64_t seed, uint64_t hyp_id, uint32_t stream, int n, int m, int* out);

namespace {
const int SLANES = 1024;

struct LoState {
    int est, n, m, kind, sample_limit, inner_iters, iter_iters, mult;
    const float* pts;
    float theta, lo_thr, step;
    uint64_t seed, calls;
    std::vector<int> lo_inliers, max_inliers;
    unsigned inner_done, iterative_done;
};

/* Quality::getNumberInliers(score, model, thr, get_inliers=true, ids): strict errors, ascending ids, lane-summed errors */
void lo_score(const LoState& L, const float* model, float thr, int& cnt, float& sum, std::vector<int>& ids) {
    ErrFn f;
    f.set(L.est, model);
    ids.resize((size_t)L.n);
    cnt = 0;
    float part[SLANES];
    for (int t = 0; t < SLANES; t++) part[t] = 0.f;
    for (int i = 0; i < L.n; i++) {
        const float e = f(L.pts, (unsigned)i);
        if (e < thr) { ids[cnt++] = i; part[i % SLANES] = part[i % SLANES] + e; }
    }
    for (int s = SLANES / 2; s > 0; s >>= 1)
        for (int t = 0; t < s; t++) part[t] = part[t] + part[t + s];
    sum = part[0];
}

/* draws 14 distinct positions in [0, count-1] with generalised philox draws (two blocks of 8) */
void lo_random_subset(LoState& L, int count, const std::vector<int>& from, int* sample) {
    int pos[16];
    /* sample_limit <= 16: first 8 from the whole range, the rest from the remaining indices, mapped past the first 8 */
    const int k = L.sample_limit;
    int a[8], b[8];
    orc_philox_unique(L.seed, L.calls, 7, count, k < 8 ? k : 8, a);
    for (int i = 0; i < (k < 8 ? k : 8); i++) pos[i] = a[i];
    if (k > 8) {
        orc_philox_unique(L.seed, L.calls, 8, count - 8, k - 8, b);
        int sorted[8];
        for (int i = 0; i < 8; i++) sorted[i] = a[i];
        for (int i = 1; i < 8; i++) { int v = sorted[i], j = i - 1; while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; j--; } sorted[j + 1] = v; }
        for (int i = 0; i < k - 8; i++) {
            int v = b[i];
            for (int q = 0; q < 8; q++) if (v >= sorted[q]) v++;
            pos[8 + i] = v;
        }
    }
    L.calls++;
    for (int i = 0; i < k; i++) sample[i] = from[pos[i]];
}

bool bigger2(int ia, float sa, int ib, float sb) { return ia > ib || (ia == ib && sa > sb); }

bool lo_iterative(LoState& L, int& lo_inl, float& lo_sum, float* lo_model, int best_inl, float best_sum) {
    int sample[16];
    for (int it = 0; it < L.iter_iters; it++) {
        L.lo_thr -= L.step;
        if (lo_inl <= L.m) break;
        if (L.kind == 2) {                                            /* GetScoreLimited, iterative_local_optimization.hpp:61-100 */
            if (lo_inl > L.sample_limit) {
                lo_random_subset(L, lo_inl, L.lo_inliers, sample);
                if (!orc_nonminimal(L.est, L.pts, sample, L.sample_limit, lo_model)) continue;
            } else if (!orc_nonminimal(L.est, L.pts, L.lo_inliers.data(), lo_inl, lo_model)) break;
            lo_score(L, lo_model, L.lo_thr, lo_inl, lo_sum, L.lo_inliers);
        } else {                                                      /* GetScoreUnlimited, :102-135 */
            if (!orc_nonminimal(L.est, L.pts, L.lo_inliers.data(), lo_inl, lo_model)) break;
            lo_score(L, lo_model, L.lo_thr, lo_inl, lo_sum, L.lo_inliers);
            if (bigger2(best_inl, best_sum, lo_inl, lo_sum)) break;
        }
        L.iterative_done++;
    }
    bool fail = false;
    if (fabsf(L.lo_thr - L.theta) > 0.00001) { fail = true; L.lo_thr = L.theta; }
    return fail;
}
}  // namespace

struct orc_lo { LoState L; };   /* opaque handle of usac_oracle.h */

extern "C" orc_lo* orc_lo_new(int est, const float* pts, int n, float theta, int kind, uint64_t seed) {
    orc_lo* o = new orc_lo;
    LoState& L = o->L;
    L.est = est; L.pts = pts; L.n = n; L.kind = kind; L.seed = seed; L.calls = 0;
    L.m = est == ORC_EST_LINE2D ? 2 : est == ORC_EST_HOMOGRAPHY ? 4 : est == ORC_EST_FUNDAMENTAL ? 7 : 5;
    L.sample_limit = 14; L.inner_iters = 20; L.iter_iters = 4; L.mult = 10;       /* model.hpp:26-29 */
    L.theta = theta; L.lo_thr = theta;
    L.step = (theta * (unsigned)L.mult - theta) / (unsigned)L.iter_iters;          /* iterative_local_optimization.hpp:43 */
    L.inner_done = L.iterative_done = 0;
    return o;
}
extern "C" void orc_lo_free(orc_lo* o) { delete o; }
extern "C" void orc_lo_counters(orc_lo* o, unsigned* inner, unsigned* iterative, unsigned long long* calls) {
    *inner = o->L.inner_done; *iterative = o->L.iterative_done; *calls = o->L.calls;
}

/* InnerLocalOptimization::GetModelScore, inner_local_optimization.hpp:74-133; model/score updated in place */
extern "C" void orc_lo_get_model_score(orc_lo* o, float* best_model, int* best_inl_io, float* best_sum_io) {
    LoState& L = o->L;
    int best_inl = *best_inl_io;
    float best_sum = *best_sum_io;
    if (best_inl < 12) return;
    const int w = 9;
    int cnt;
    float sum;
    lo_score(L, best_model, L.theta, cnt, sum, L.max_inliers);                       /* quality->getInliers(best_model), :79 */
    float lo_model[9];
    int sample[16];
    int avail = best_inl < cnt ? best_inl : cnt;                                      /* ids present in max_inliers */
    for (int it = 0; it < L.inner_iters; it++) {
        if (avail > L.sample_limit) {
            lo_random_subset(L, avail, L.max_inliers, sample);
            if (!orc_nonminimal(L.est, L.pts, sample, L.sample_limit, lo_model)) continue;
        } else if (!orc_nonminimal(L.est, L.pts, L.max_inliers.data(), avail, lo_model)) {
            break;
        }
        L.lo_thr = (unsigned)L.mult * L.lo_thr;                                      /* :101 */
        int lo_inl;
        float lo_sum;
        lo_score(L, lo_model, L.lo_thr, lo_inl, lo_sum, L.lo_inliers);
        if (lo_inl <= L.m) continue;                                                 /* :105 */
        const bool fail = lo_iterative(L, lo_inl, lo_sum, lo_model, best_inl, best_sum);
        if (!fail && bigger2(lo_inl, lo_sum, best_inl, best_sum)) {
            memcpy(best_model, lo_model, sizeof(float) * w);
            best_inl = lo_inl; best_sum = lo_sum; avail = lo_inl;
            L.max_inliers.resize((size_t)L.n);
            for (int i = 0; i < lo_inl; i++) L.max_inliers[i] = L.lo_inliers[i];
        }
        L.inner_done++;
    }
    *best_inl_io = best_inl; *best_sum_io = best_sum;
}
