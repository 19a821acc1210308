/*
 * usac_oracle.cpp - CPU restatement of the reference's hypothesize-and-verify path.
 *
 * TEST INFRASTRUCTURE ONLY (see usac_oracle.h). Build with -ffp-contract=off so that every float/double operator
 * rounds once, like the reference's x86-64 SSE build (no -march/-O flags, CMakeLists.txt:102-107).
 *
 * Citations are file:line under /root/reference.
 */
#include "oracle_internal.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

/* ======================================================================================================
 * RNG streams
 * ====================================================================================================== */

/* glibc random()/srandom(), TYPE_3 (x^31 + x^3 + 1). The reference's UniformSampler, ArrayRandomGenerator and SPRT
 * pool shuffle all consume this stream (uniform_sampler.hpp:47, array_random_generator.hpp:35, sprt.hpp:101). */
struct orc_glibc_rand {
    int32_t r[31];
    int f, b;
};

static void glibc_seed(orc_glibc_rand* g, unsigned seed) {
    if (seed == 0) seed = 1;
    g->r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        long hi = g->r[i - 1] / 127773, lo = g->r[i - 1] % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        g->r[i] = (int32_t)w;
    }
    g->f = 3;
    g->b = 0;
    for (int i = 0; i < 310; i++) orc_glibc_rand_next(g);
}

extern "C" orc_glibc_rand* orc_glibc_rand_new(unsigned seed) {
    orc_glibc_rand* g = new orc_glibc_rand;
    glibc_seed(g, seed);
    return g;
}
extern "C" void orc_glibc_rand_free(orc_glibc_rand* g) { delete g; }
extern "C" int32_t orc_glibc_rand_next(orc_glibc_rand* g) {
    uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
    g->r[g->f] = (int32_t)v;
    int32_t out = (int32_t)(v >> 1);
    g->f = g->f + 1 == 31 ? 0 : g->f + 1;
    g->b = g->b + 1 == 31 ? 0 : g->b + 1;
    return out;
}

/* Philox4x32-10 (Salmon et al., SC'11) - the counter-based stream of the "Philox" sampler mode. */
extern "C" void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* m distinct indices in [0,n): the i-th draw picks the j-th smallest still-unused index, j = mulhi(r_i, n-i).
 * Random words: Philox(counter = {lo(hyp), hi(hyp), block, stream}, key = {lo(seed), hi(seed)}), 4 words a block. */
extern "C" void orc_philox_unique(uint64_t seed, uint64_t hyp_id, uint32_t stream, int n, int m, int* out) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t words[8];
    for (int blk = 0; blk * 4 < m; blk++) {
        uint32_t ctr[4] = {(uint32_t)hyp_id, (uint32_t)(hyp_id >> 32), (uint32_t)blk, stream};
        orc_philox4x32_10(ctr, key, words + 4 * blk);
    }
    int sorted[8];
    for (int i = 0; i < m; i++) {
        int j = (int)(((uint64_t)words[i] * (uint64_t)(uint32_t)(n - i)) >> 32);
        int pos = 0;
        while (pos < i && j >= sorted[pos]) { j++; pos++; }
        for (int q = i; q > pos; q--) sorted[q] = sorted[q - 1];
        sorted[pos] = j;
        out[i] = j;
    }
}

/* ======================================================================================================
 * Scoring - Estimator::GetError + Quality::getNumberInliers
 * ====================================================================================================== */

/* cv::Mat::inv() for 3x3 CV_32F as called by HomographyEstimator::setModelParameters (homography_estimator.hpp:35).
 * OpenCV core (lapack.cpp, invert(), n==3 branch): determinant and cofactors in double, result rounded to float;
 * det==0 leaves an all-zero matrix. Checked bit-for-bit against cv2.invert in tests/golden/cv_primitives.npz. */
extern "C" int orc_inv3x3(const float* m, float* out) {
    double a00 = m[0], a01 = m[1], a02 = m[2], a10 = m[3], a11 = m[4], a12 = m[5], a20 = m[6], a21 = m[7], a22 = m[8];
    double d = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
    if (d == 0.0) {
        for (int i = 0; i < 9; i++) out[i] = 0.f;
        return 0;
    }
    d = 1.0 / d;
    out[0] = (float)((a11 * a22 - a12 * a21) * d);
    out[1] = (float)((a02 * a21 - a01 * a22) * d);
    out[2] = (float)((a01 * a12 - a02 * a11) * d);
    out[3] = (float)((a12 * a20 - a10 * a22) * d);
    out[4] = (float)((a00 * a22 - a02 * a20) * d);
    out[5] = (float)((a02 * a10 - a00 * a12) * d);
    out[6] = (float)((a10 * a21 - a11 * a20) * d);
    out[7] = (float)((a01 * a20 - a00 * a21) * d);
    out[8] = (float)((a00 * a11 - a01 * a10) * d);
    return 1;
}


extern "C" void orc_errors(int estimator, const float* points, int n, const float* model, float* err_out) {
    ErrFn f;
    f.set(estimator, model);
    for (int i = 0; i < n; i++) err_out[i] = f(points, (unsigned)i);
}

/* Quality::getNumberInliers, quality.hpp:60-101: strict `err < threshold`, float sum in point order. */
extern "C" void orc_score(int estimator, const float* points, int n, const float* model, float thr, int* count_out,
                          float* sum_out, int* inliers_out, double band_rel, int* flagged_out) {
    ErrFn f;
    f.set(estimator, model);
    int cnt = 0, flagged = 0;
    float sum = 0;
    for (int i = 0; i < n; i++) {
        float e = f(points, (unsigned)i);
        if (e < thr) {
            if (inliers_out) inliers_out[cnt] = i;
            cnt++;
            sum += e;
        }
        if (flagged_out && std::fabs((double)e - (double)thr) <= band_rel * (double)thr) flagged++;
    }
    if (count_out) *count_out = cnt;
    if (sum_out) *sum_out = sum;
    if (flagged_out) *flagged_out = flagged;
}

/* ======================================================================================================
 * Minimal solvers
 * ====================================================================================================== */

/* Null space of a rows x 9 system by Gauss-Jordan elimination with partial (row) pivoting in double. This replaces
 * cv::SVD (dlt.cpp:43, seven_points.cpp:88): the reference takes the last rows of Vt; any basis of the same null
 * space gives the same models. Operation order is part of the contract with the CUDA solvers:
 *   for k: pivot = first row r>=k maximising |A[r][k]|; swap; scale row k by 1/pivot; for every other row r:
 *   A[r][j] -= A[r][k]*A[k][j] (product rounded, then difference rounded), j = k+1..8.
 * Free columns are rows..8; basis vector q has v[rows+q]=1, v[i<rows] = -A[i][rows+q]. Returns 0 on a zero/NaN pivot. */
extern "C" int orc_null_space(double* A, int rows, double* basis_out) {
    const int C = 9;
    for (int k = 0; k < rows; k++) {
        int piv = k;
        double best = std::fabs(A[k * C + k]);
        for (int r = k + 1; r < rows; r++) {
            double v = std::fabs(A[r * C + k]);
            if (v > best) { best = v; piv = r; }
        }
        if (!(best > 0.0) || !std::isfinite(best)) return 0;
        if (piv != k)
            for (int j = 0; j < C; j++) std::swap(A[k * C + j], A[piv * C + j]);
        double inv = 1.0 / A[k * C + k];
        for (int j = k + 1; j < C; j++) A[k * C + j] = A[k * C + j] * inv;
        for (int r = 0; r < rows; r++) {
            if (r == k) continue;
            double f = A[r * C + k];
            for (int j = k + 1; j < C; j++) {
                double t = f * A[k * C + j];
                A[r * C + j] = A[r * C + j] - t;
            }
        }
    }
    int nfree = C - rows;
    for (int q = 0; q < nfree; q++) {
        double* v = basis_out + q * C;
        for (int i = 0; i < rows; i++) v[i] = -A[i * C + rows + q];
        for (int i = rows; i < C; i++) v[i] = (i == rows + q) ? 1.0 : 0.0;
    }
    return 1;
}

/* Line through two points, line2d_estimator.hpp:36-54 (float, sqrtf). */
static int solve_line(const float* pts, const int* s, float* out) {
    const int i1 = s[0], i2 = s[1];
    float a = pts[2 * i1 + 1] - pts[2 * i2 + 1];
    float b = pts[2 * i2] - pts[2 * i1];
    float mag = sqrtf(a * a + b * b);
    a /= mag;
    b /= mag;
    float c = (pts[2 * i1] * pts[2 * i2 + 1] - pts[2 * i2] * pts[2 * i1 + 1]) / mag;
    out[0] = a; out[1] = b; out[2] = c;
    return 1;
}

/* Homography from 4 correspondences: Hartley normalisation exactly as GetNormalizingTransformation
 * (normalizing_transformation.cpp:7-112: float means, float sqrt distances, scale = M_SQRT2/(avg) evaluated in
 * double and stored as float, float normalised points), DLT rows as dlt.cpp:55-101, then the TRUE null vector
 * (the reference's thin-SVD row is not the null vector, SURVEY.md finding 5) and H = T2^-1 Hn T1, /h33
 * (normalized_dlt.cpp:18-20) in double, rounded to float at the end. */
static int solve_homography4(const float* pts, const int* s, float* out) {
    const unsigned sample_number = 4;
    float m1x = 0, m1y = 0, m2x = 0, m2y = 0;
    for (unsigned i = 0; i < sample_number; i++) {
        const float* p = pts + 4 * s[i];
        m1x += p[0]; m1y += p[1]; m2x += p[2]; m2y += p[3];
    }
    m1x /= sample_number; m1y /= sample_number; m2x /= sample_number; m2y /= sample_number;
    float d1 = 0, d2 = 0;
    for (unsigned i = 0; i < sample_number; i++) {
        const float* p = pts + 4 * s[i];
        float ax = p[0] - m1x, ay = p[1] - m1y, bx = p[2] - m2x, by = p[3] - m2y;
        d1 += sqrtf(ax * ax + ay * ay);
        d2 += sqrtf(bx * bx + by * by);
    }
    const float s1 = (float)(M_SQRT2 / (double)(d1 / sample_number));
    const float s2 = (float)(M_SQRT2 / (double)(d2 / sample_number));
    const float t1x = -m1x * s1, t1y = -m1y * s1, t2x = -m2x * s2, t2y = -m2y * s2;

    double A[8 * 9];
    for (unsigned i = 0; i < sample_number; i++) {
        const float* p = pts + 4 * s[i];
        /* normalised coordinates and DLT rows in double (the reference rounds both to float32, which only adds
         * noise: H = T2^-1 Hn T1 does not depend on the exact T) */
        const double x1 = (double)s1 * (double)p[0] + (double)t1x, y1 = (double)s1 * (double)p[1] + (double)t1y;
        const double x2 = (double)s2 * (double)p[2] + (double)t2x, y2 = (double)s2 * (double)p[3] + (double)t2y;
        double* r0 = A + (2 * i) * 9;
        double* r1 = r0 + 9;
        r0[0] = -x1; r0[1] = -y1; r0[2] = -1; r0[3] = 0; r0[4] = 0; r0[5] = 0;
        r0[6] = x2 * x1; r0[7] = x2 * y1; r0[8] = x2;
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = -x1; r1[4] = -y1; r1[5] = -1;
        r1[6] = y2 * x1; r1[7] = y2 * y1; r1[8] = y2;
    }
    double h[9];
    if (!orc_null_space(A, 8, h)) return 0;

    /* M = Hn * T1, T1 = [s1 0 t1x; 0 s1 t1y; 0 0 1] */
    double M[9];
    for (int i = 0; i < 3; i++) {
        M[3 * i] = h[3 * i] * (double)s1;
        M[3 * i + 1] = h[3 * i + 1] * (double)s1;
        M[3 * i + 2] = h[3 * i] * (double)t1x + h[3 * i + 1] * (double)t1y + h[3 * i + 2];
    }
    /* H = T2^-1 * M, T2^-1 = [1/s2 0 -t2x/s2; 0 1/s2 -t2y/s2; 0 0 1] */
    const double is2 = 1.0 / (double)s2;
    const double ux = -((double)t2x * is2), uy = -((double)t2y * is2);
    double H[9];
    for (int j = 0; j < 3; j++) {
        H[j] = is2 * M[j] + ux * M[6 + j];
        H[3 + j] = is2 * M[3 + j] + uy * M[6 + j];
        H[6 + j] = M[6 + j];
    }
    const double inv = 1.0 / H[8];
    for (int i = 0; i < 9; i++) {
        double v = H[i] * inv;
        if (!std::isfinite(v)) return 0;
        out[i] = (float)v;
    }
    out[8] = 1.f;
    return 1;
}

static inline double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

static inline double cubic_eval(double a, double b, double c, double x) { return ((x + a) * x + b) * x + c; }

/* root of x^3+ax^2+bx+c in [l,r] (signs of p differ at the ends) by bisection down to adjacent doubles */
static double cubic_bisect(double a, double b, double c, double l, double r) {
    double pl = cubic_eval(a, b, c, l);
    if (pl == 0.0) return l;
    double pr = cubic_eval(a, b, c, r);
    if (pr == 0.0) return r;
    const bool neg_left = pl < 0.0;
    for (int it = 0; it < 2200; it++) {
        double m = 0.5 * (l + r);
        if (m == l || m == r) break;
        double pm = cubic_eval(a, b, c, m);
        if (pm == 0.0) return m;
        if ((pm < 0.0) == neg_left) l = m; else r = m;
    }
    return 0.5 * (l + r);
}

/* Real roots of c0 x^3 + c1 x^2 + c2 x + c3. Semantics follow cv::solveCubic (OpenCV core mathfuncs; call site
 * seven_points.cpp:131): returns the number of distinct real roots; with three roots the order is
 * [smallest, largest, middle] (that is what the trigonometric branch of cv::solveCubic produces), c0==0 degrades
 * to the quadratic/linear case. The arithmetic is bracketing + bisection on the monotone pieces so that only
 * + - * / sqrt are used (bit-reproducible on CPU and GPU); agreement with cv2.solveCubic is checked to 1e-9
 * relative in tests (tests/golden/cv_primitives.npz). */
extern "C" int orc_solve_cubic(const double co[4], double roots[3]) {
    roots[0] = roots[1] = roots[2] = 0.0;
    const double a0 = co[0];
    if (a0 == 0.0) {
        const double a1 = co[1], a2 = co[2], a3 = co[3];
        if (a1 == 0.0) {
            if (a2 == 0.0) return a3 == 0.0 ? -1 : 0;
            roots[0] = -a3 / a2;
            return 1;
        }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d < 0.0) return 0;
        d = std::sqrt(d);
        double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (std::fabs(q1) > std::fabs(q2)) { roots[0] = q1 / a1; roots[1] = a3 / q1; }
        else { roots[0] = q2 / a1; roots[1] = a3 / q2; }
        return d > 0.0 ? 2 : 1;
    }
    const double inv = 1.0 / a0;
    const double a = co[1] * inv, b = co[2] * inv, c = co[3] * inv;
    if (!std::isfinite(a) || !std::isfinite(b) || !std::isfinite(c)) return 0;
    double B = std::fabs(a);
    if (std::fabs(b) > B) B = std::fabs(b);
    if (std::fabs(c) > B) B = std::fabs(c);
    B = B + 1.0;                                   /* Cauchy bound: all roots lie in [-B, B] */
    const double disc = a * a - 3.0 * b;           /* p'(x) = 3x^2 + 2ax + b */
    if (!(disc > 0.0)) {                           /* monotone: one real root */
        roots[0] = cubic_bisect(a, b, c, -B, B);
        return 1;
    }
    const double sq = std::sqrt(disc);
    const double xlo = (-a - sq) / 3.0, xhi = (-a + sq) / 3.0;   /* local max, local min */
    const double plo = cubic_eval(a, b, c, xlo), phi = cubic_eval(a, b, c, xhi);
    if (plo < 0.0) { roots[0] = cubic_bisect(a, b, c, xhi, B); return 1; }
    if (phi > 0.0) { roots[0] = cubic_bisect(a, b, c, -B, xlo); return 1; }
    if (plo == 0.0 && phi == 0.0) { roots[0] = xlo; return 1; }                 /* triple root */
    if (plo == 0.0) { roots[0] = cubic_bisect(a, b, c, xhi, B); roots[1] = xlo; return 2; }   /* [single, double] */
    if (phi == 0.0) { roots[0] = cubic_bisect(a, b, c, -B, xlo); roots[1] = xhi; return 2; }
    const double r_small = cubic_bisect(a, b, c, -B, xlo);
    const double r_mid = cubic_bisect(a, b, c, xlo, xhi);
    const double r_large = cubic_bisect(a, b, c, xhi, B);
    roots[0] = r_small; roots[1] = r_large; roots[2] = r_mid;
    return 3;
}

/* Oriented epipolar constraint, fundamental_estimator.hpp:189-231 (float). */
extern "C" int orc_fundamental_is_valid(const float* pts, const float* F, const int* sample) {
    /* epipole(): ec = row0 x row2, falling back to row1 x row2 when all |ec_i| <= 1.9984e-15 */
    float ec[3];
    ec[0] = F[1] * F[8] - F[2] * F[7];
    ec[1] = F[2] * F[6] - F[0] * F[8];
    ec[2] = F[0] * F[7] - F[1] * F[6];
    bool big = false;
    for (int i = 0; i < 3; i++)
        if (ec[i] > 1.9984e-15 || ec[i] < -1.9984e-15) { big = true; break; }
    if (!big) {
        ec[0] = F[4] * F[8] - F[5] * F[7];
        ec[1] = F[5] * F[6] - F[3] * F[8];
        ec[2] = F[3] * F[7] - F[4] * F[6];
    }
    float sig1 = 0;
    for (int i = 0; i < 7; i++) {
        const float* p = pts + 4 * sample[i];
        const float y1 = p[1], x2 = p[2], y2 = p[3];
        float s1 = F[0] * x2 + F[3] * y2 + F[6];
        float s2 = ec[1] - ec[2] * y1;
        float sig = s1 * s2;
        if (i == 0) sig1 = sig;
        else if (sig1 * sig < 0) return 0;
    }
    return 1;
}

/* Seven-point algorithm, seven_points.cpp:49-156 + FundamentalEstimator::EstimateModel (fundamental_estimator.hpp:48-63).
 * A (:68-84), null space, cubic and the F = lambda*f1 + mu*f2 combination run in
 * double (the reference does them in float32 on pixel coordinates; SURVEY.md finding 6 / hard part 4). */
static int solve_fundamental7(const float* pts, const int* s, float* out) {
    double A[7 * 9];
    for (int i = 0; i < 7; i++) {
        const float* p = pts + 4 * s[i];
        const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];   /* products of two floats are exact in double */
        double* r = A + 9 * i;
        r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2;
        r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2;
        r[6] = x1; r[7] = y1; r[8] = 1;
    }
    double basis[2 * 9];
    if (!orc_null_space(A, 7, basis)) return 0;
    double f1[9], f2[9];
    for (int i = 0; i < 9; i++) { f2[i] = basis[9 + i]; f1[i] = basis[i] - f2[i]; }   /* :99-101 f1 -= f2 */

    /* det(l*f1 + f2) = c0 l^3 + c1 l^2 + c2 l + c3 by multilinearity in the rows */
    double c[4];
    c[0] = det3(f1);
    c[3] = det3(f2);
    double m[9];
    c[1] = 0; c[2] = 0;
    for (int row = 0; row < 3; row++) {
        for (int i = 0; i < 9; i++) m[i] = (i / 3 == row) ? f2[i] : f1[i];
        c[1] += det3(m);
        for (int i = 0; i < 9; i++) m[i] = (i / 3 == row) ? f1[i] : f2[i];
        c[2] += det3(m);
    }
    double r[3];
    int nroots = orc_solve_cubic(c, r);
    if (nroots < 1) return 0;
    int valid = 0;
    for (int k = 0; k < nroots; k++) {
        double lambda = r[k], mu = 1.0;
        double sc = f1[8] * r[k] + f2[8];
        float F[9];
        if (std::fabs(sc) > DBL_EPSILON) {           /* :144 */
            mu = 1.0 / sc;
            lambda *= mu;
            F[8] = 1.f;
        } else {
            F[8] = 0.f;
        }
        bool finite = true;
        for (int i = 0; i < 8; i++) {
            double v = f1[i] * lambda + f2[i] * mu;
            if (!std::isfinite(v)) finite = false;
            F[i] = (float)v;
        }
        if (!finite) continue;
        if (orc_fundamental_is_valid(pts, F, s)) {
            for (int i = 0; i < 9; i++) out[9 * valid + i] = F[i];
            valid++;
        }
    }
    return valid;
}

/* SURVEY.md Appendix B, quirk 1 (documented switch, orc_config::ref_thin_svd / orc_solve_homography_dlt4p_thin): what the reference's
 * own minimal homography solver computes. DLt::DLT4p (dlt.cpp:7-53) stacks the 8 x 9 system from RAW pixel coordinates in float32
 * (:17-41), calls cv::SVD::compute on it and takes vt.row(vt.rows - 1) (:43-48). For an 8 x 9 matrix OpenCV's SVD is THIN: vt is
 * 8 x 9, its last row is the right singular vector of the SMALLEST OF THE 8 singular values - not the null vector (the 9th, which a
 * thin SVD never returns). H = that row / h33 (:50). The product path and the oracle's default solve the normalised DLT for the true
 * null vector (what BASELINE.json names); this function exists so that the difference is a switch one can turn, not a silent fix.
 * Numerics: one-sided (Hestenes) Jacobi in double on the 8 rows of A - rotating pairs of rows until they are mutually orthogonal
 * leaves sigma_i v_i' in row i; the row of smallest norm, normalised, is the wanted vector (sign cancels in the division by h33). */
static int solve_homography4_thin_svd(const float* pts, const int* s, float* out) {
    double R[8][9];
    for (int i = 0; i < 4; i++) {
        const float* p = pts + 4 * (size_t)s[i];
        const float x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
        const float r0[9] = {-x1, -y1, -1.f, 0.f, 0.f, 0.f, x2 * x1, x2 * y1, x2};      /* float products, like the Mat_<float> A */
        const float r1[9] = {0.f, 0.f, 0.f, -x1, -y1, -1.f, y2 * x1, y2 * y1, y2};
        for (int k = 0; k < 9; k++) { R[2 * i][k] = r0[k]; R[2 * i + 1][k] = r1[k]; }
    }
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotations = 0;
        for (int a = 0; a < 7; a++) {
            for (int b = a + 1; b < 8; b++) {
                double aa = 0, bb = 0, ab = 0;
                for (int k = 0; k < 9; k++) { aa += R[a][k] * R[a][k]; bb += R[b][k] * R[b][k]; ab += R[a][k] * R[b][k]; }
                if (ab == 0.0 || std::fabs(ab) <= 1e-15 * std::sqrt(aa * bb)) continue;
                const double zeta = (bb - aa) / (2.0 * ab);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
                for (int k = 0; k < 9; k++) {
                    const double ra = R[a][k], rb = R[b][k];
                    R[a][k] = c * ra - sn * rb;
                    R[b][k] = sn * ra + c * rb;
                }
                rotations++;
            }
        }
        if (!rotations) break;
    }
    int last = 0;
    double least = -1;
    for (int i = 0; i < 8; i++) {
        double nn = 0;
        for (int k = 0; k < 9; k++) nn += R[i][k] * R[i][k];
        if (least < 0 || nn < least) { least = nn; last = i; }
    }
    /* vt.row(7) is a unit vector in float32; H = H / H.at<float>(2,2) is a float division (dlt.cpp:48-50) */
    const double norm = std::sqrt(least);
    if (!(norm > 0)) return 0;
    float v[9];
    for (int k = 0; k < 9; k++) v[k] = (float)(R[last][k] / norm);
    for (int k = 0; k < 9; k++) {
        out[k] = v[k] / v[8];
        if (!std::isfinite(out[k])) return 0;
    }
    return 1;
}

extern "C" int orc_solve_homography_dlt4p_thin(const float* points, const int* sample, float* model_out) {
    return solve_homography4_thin_svd(points, sample, model_out);
}

extern "C" int orc_solve_minimal(int estimator, const float* points, const int* sample, float* models_out) {
    switch (estimator) {
        case ORC_EST_LINE2D: return solve_line(points, sample, models_out);
        case ORC_EST_HOMOGRAPHY: return solve_homography4(points, sample, models_out);
        case ORC_EST_FUNDAMENTAL: return solve_fundamental7(points, sample, models_out);
        case ORC_EST_ESSENTIAL: return orc_solve_essential5(points, sample, models_out);
    }
    return 0;
}

/* ======================================================================================================
 * k nearest neighbours, nearest_neighbors.cpp:69-128
 * ====================================================================================================== */

/* For every point the k+1 closest points by squared L2 distance over all columns (the nanoflann result set, ascending),
 * minus the first one (the query itself). Brute force with the total order (distance bits, point index), so that the
 * result is defined for equidistant points too. */
extern "C" int orc_knn_build(const float* points, int n, int dim, int k, int* table_out) {
    if (n < k + 1 || k < 1) return -1;
    std::vector<uint64_t> best((size_t)k + 1);
    for (int q = 0; q < n; q++) {
        int have = 0;
        const float* pq = points + (size_t)q * dim;
        for (int p = 0; p < n; p++) {
            const float* pp = points + (size_t)p * dim;
            float d = 0.f;
            for (int c = 0; c < dim; c++) {
                const float diff = pq[c] - pp[c];
                const float sq = diff * diff;
                d = c == 0 ? sq : d + sq;
            }
            uint32_t bits;
            std::memcpy(&bits, &d, 4);
            bits &= 0x7fffffffu;                       /* d >= 0 or NaN: the bit pattern orders like the value, NaN last */
            if (d != d) bits = 0x7fc00000u;            /* one NaN pattern, whatever payload the hardware produced */
            uint64_t key = ((uint64_t)bits << 32) | (uint32_t)p;
            if (have == k + 1 && key >= best[k]) continue;
            int j = have < k + 1 ? have++ : k;
            while (j > 0 && best[j - 1] > key) { best[j] = best[j - 1]; j--; }
            best[j] = key;
        }
        for (int j = 0; j < k; j++) table_out[(size_t)q * k + j] = (int)(uint32_t)best[j + 1];
    }
    return 0;
}

/* ======================================================================================================
 * Neighbourhood grid, nearest_neighbors.cpp:160-201
 * ====================================================================================================== */

/* Points are neighbours iff they share the 4-D cell (int(x1/c), int(y1/c), int(x2/c), int(y2/c)) (truncation toward
 * zero, float division by the int cell size, :172-174). The reference's per-point list is "the other members of my
 * cell in ascending index order" (:190-200); here it is stored as CSR: members sorted by (cell, index). Cells are
 * numbered in std::map key order (c1x, c1y, c2x, c2y lexicographic, nearest_neighbors.hpp CellCoord::operator<). */
extern "C" void orc_grid_cells(const float* points, int n, int cell_size, int* cell_of_point, int* members,
                               int* cell_start, int* ncells_out) {
    typedef std::tuple<int, int, int, int> Key;
    std::map<Key, std::vector<int>> cells;
    for (int i = 0; i < n; i++) {
        const float* p = points + 4 * i;
        Key k((int)(p[0] / cell_size), (int)(p[1] / cell_size), (int)(p[2] / cell_size), (int)(p[3] / cell_size));
        cells[k].push_back(i);
    }
    int c = 0, pos = 0;
    for (auto& kv : cells) {
        cell_start[c] = pos;
        for (int idx : kv.second) {
            members[pos++] = idx;
            cell_of_point[idx] = c;
        }
        c++;
    }
    cell_start[c] = pos;
    *ncells_out = c;
}

/* ======================================================================================================
 * Samplers
 * ====================================================================================================== */

struct orc_sampler {
    int kind, rng, n, m;
    uint64_t seed;
    orc_glibc_rand* g = nullptr;
    /* uniform_sampler.hpp:17-18 / array_random_generator.hpp:13-15: persistent shrinking pool */
    std::vector<unsigned> pool;
    int max = 0;
    /* PROSAC, prosac_sampler.hpp */
    std::vector<unsigned> growth;
    unsigned subset_size = 0, largest_sample_size = 0, hyp_count = 1, growth_max_samples = 200000;
    unsigned termination_length = 0;
    /* NAPSAC, napsac_sampler.hpp */
    int neighbors_type = ORC_NEIGH_NONE, knn = 0;
    std::vector<int> knn_table, next_neighbors;
    std::vector<int> cell_of_point, members, cell_start, rank_in_cell;
    uint64_t draws = 0;   /* Philox: counter for seed-point redraws */
};

/* One step of the persistent Fisher-Yates pool (uniform_sampler.hpp:42-54 / array_random_generator.hpp:31-44). */
static unsigned pool_draw(orc_sampler* s) {
    if (s->max == 0) s->max = s->n;
    unsigned idx = (unsigned)orc_glibc_rand_next(s->g) % (unsigned)s->max;
    unsigned v = s->pool[idx];
    s->max--;
    s->pool[idx] = s->pool[s->max];
    s->pool[s->max] = v;
    return v;
}

extern "C" orc_sampler* orc_sampler_new(int kind, int rng, int n, int m, uint64_t seed) {
    orc_sampler* s = new orc_sampler;
    s->kind = kind; s->rng = rng; s->n = n; s->m = m; s->seed = seed;
    /* reset_random_generator=false leaves glibc at its default seed 1; `seed` selects it here */
    s->g = orc_glibc_rand_new((unsigned)seed);
    s->pool.resize(n);
    for (int i = 0; i < n; i++) s->pool[i] = i;
    s->max = n;
    if (kind == ORC_SAMPLER_PROSAC) {
        /* initProsacSampler, prosac_sampler.hpp:62-114 */
        s->growth.resize(n);
        double T_n = s->growth_max_samples;
        for (int i = 0; i < m; i++) T_n *= (double)(m - i) / (n - i);
        unsigned T_n_prime = 1;
        for (int i = 0; i < n; i++) {
            if (i + 1 <= m) { s->growth[i] = T_n_prime; continue; }
            double Tn_plus1 = (double)(i + 1) * T_n / (i + 1 - m);
            s->growth[i] = T_n_prime + (unsigned)std::ceil(Tn_plus1 - T_n);
            T_n = Tn_plus1;
            T_n_prime = s->growth[i];
        }
        s->largest_sample_size = m;
        s->subset_size = m;
        s->hyp_count = 1;
        s->termination_length = n;    /* prosac_termination_criteria.hpp:59 */
    }
    if (kind == ORC_SAMPLER_NAPSAC) s->next_neighbors.assign(n, 0);
    return s;
}
extern "C" void orc_sampler_free(orc_sampler* s) {
    if (!s) return;
    orc_glibc_rand_free(s->g);
    delete s;
}
extern "C" void orc_sampler_set_knn(orc_sampler* s, const int* neighbors, int k) {
    s->neighbors_type = ORC_NEIGH_KNN;
    s->knn = k;
    s->knn_table.assign(neighbors, neighbors + (size_t)s->n * k);
}
extern "C" void orc_sampler_set_grid(orc_sampler* s, const float* points, int cell_size) {
    s->neighbors_type = ORC_NEIGH_GRID;
    s->cell_of_point.resize(s->n); s->members.resize(s->n); s->cell_start.resize(s->n + 1); s->rank_in_cell.resize(s->n);
    int nc = 0;
    orc_grid_cells(points, s->n, cell_size, s->cell_of_point.data(), s->members.data(), s->cell_start.data(), &nc);
    for (int c = 0; c < nc; c++)
        for (int q = s->cell_start[c]; q < s->cell_start[c + 1]; q++) s->rank_in_cell[s->members[q]] = q - s->cell_start[c];
}
extern "C" void orc_sampler_set_termination_length(orc_sampler* s, unsigned len) { s->termination_length = len; }
extern "C" unsigned orc_sampler_largest_sample_size(orc_sampler* s) { return s->largest_sample_size; }
extern "C" const unsigned* orc_sampler_growth_function(orc_sampler* s) { return s->growth.data(); }

/* distinct draws in the closed range [0, maxv] (uniform_random_generator.hpp:36-47), range clipped to the data
 * (the reference can ask for index n). The reference's mt19937 is seeded from
 * std::random_device and cannot be replayed; here the stream is Philox keyed by (seed, hypothesis, stream). */
static void unique_closed(orc_sampler* s, uint64_t hyp, uint32_t stream, unsigned maxv, int cnt, int* out) {
    int range = (int)std::min<unsigned>(maxv + 1, (unsigned)s->n);
    orc_philox_unique(s->seed, hyp, stream, range, cnt, out);
}

static int napsac_nbr_count(orc_sampler* s, int p) {
    if (s->neighbors_type == ORC_NEIGH_KNN) return s->knn;
    int c = s->cell_of_point[p];
    return s->cell_start[c + 1] - s->cell_start[c] - 1;
}
/* j-th neighbour of p: kNN row entry, or j-th other member of p's cell in ascending index order */
static int napsac_nbr(orc_sampler* s, int p, int j) {
    if (s->neighbors_type == ORC_NEIGH_KNN) return s->knn_table[(size_t)s->knn * p + j];
    int c = s->cell_of_point[p];
    int r = s->rank_in_cell[p];
    return s->members[s->cell_start[c] + j + (j >= r ? 1 : 0)];
}

extern "C" void orc_sampler_generate(orc_sampler* s, uint64_t hyp_id, int* out) {
    const int m = s->m;
    if (s->kind == ORC_SAMPLER_UNIFORM) {
        if (s->rng == ORC_RNG_GLIBC) {
            for (int i = 0; i < m; i++) out[i] = (int)pool_draw(s);      /* uniform_sampler.hpp:42-54 */
        } else {
            orc_philox_unique(s->seed, hyp_id, 0, s->n, m, out);
        }
        return;
    }
    if (s->kind == ORC_SAMPLER_PROSAC) {                                  /* prosac_sampler.hpp:117-172 */
        if (s->hyp_count > s->growth_max_samples) { unique_closed(s, hyp_id, 1, s->n, m, out); return; }
        if (s->subset_size > s->termination_length) { unique_closed(s, hyp_id, 2, s->termination_length, m, out); return; }
        if (s->hyp_count > s->growth[s->subset_size - 1]) {
            ++s->subset_size;
            if (s->subset_size > (unsigned)s->n) s->subset_size = s->n;
            if (s->largest_sample_size < s->subset_size) s->largest_sample_size = s->subset_size;
        }
        unique_closed(s, hyp_id, 3, s->subset_size - 2, m - 1, out);
        out[m - 1] = (int)s->subset_size - 1;
        s->hyp_count++;
        return;
    }
    if (s->kind == ORC_SAMPLER_NAPSAC) {
        int p = -1;
        if (s->neighbors_type == ORC_NEIGH_KNN) {                         /* napsac_sampler.hpp:76-98 */
            if (s->rng == ORC_RNG_GLIBC) p = (int)pool_draw(s);
            else { int t; orc_philox_unique(s->seed, hyp_id, 4, s->n, 1, &t); p = t; }
            out[0] = p;
            for (int i = 1; i < m; i++) {
                out[i] = s->knn_table[(size_t)s->knn * p + s->next_neighbors[p] + s->knn - 1];
                s->next_neighbors[p]--;
                if (s->next_neighbors[p] == -s->knn) s->next_neighbors[p] = 0;
            }
            return;
        }
        /* grid, napsac_sampler.hpp:100-128: redraw the seed until it has >= m neighbours */
        int tries = 0;
        for (; tries < s->n; tries++) {
            if (s->rng == ORC_RNG_GLIBC) p = (int)pool_draw(s);
            else { int t; orc_philox_unique(s->seed, hyp_id, 16 + (uint32_t)tries, s->n, 1, &t); p = t; }
            if (napsac_nbr_count(s, p) < m) continue;
            break;
        }
        if (tries == s->n) {   /* reference falls back to uniform sampling for the rest of the run */
            s->kind = ORC_SAMPLER_UNIFORM;
            orc_sampler_generate(s, hyp_id, out);
            return;
        }
        out[0] = p;
        const int cnt = napsac_nbr_count(s, p);
        for (int i = 1; i < m; i++) {
            out[i] = napsac_nbr(s, p, s->next_neighbors[p]);
            s->next_neighbors[p]++;
            if (s->next_neighbors[p] >= cnt) s->next_neighbors[p] = 0;
        }
        return;
    }
}

/* ======================================================================================================
 * Termination, standard_termination_criteria.hpp:52-62
 * ====================================================================================================== */
extern "C" unsigned orc_standard_termination(unsigned inliers, unsigned n, int m, float confidence, unsigned max_iterations) {
    const float log_1_p = (float)logf(1 - confidence);       /* :27, float overload (see ErrFn note) */
    float inl_ratio = (float)inliers / n;
    float inl_prob = inl_ratio * inl_ratio;
    int k = m;
    while (k > 2) { inl_prob *= inl_ratio; k--; }
    if (inl_prob < 0.0005f) return max_iterations;
    return (unsigned)(log_1_p / logf(1 - inl_prob));
}

/* the sampler's glibc stream, shared with the SPRT pool shuffle (Ransac ctor order, ransac.hpp:50-92) */
extern "C" int32_t orc_sampler_glibc_next(orc_sampler* s) { return orc_glibc_rand_next(s->g); }
