/*
 * usac_oracle_essential.cpp - five-point essential-matrix solver of the CPU oracle. TEST INFRASTRUCTURE ONLY (usac_oracle.h).
 *
 * Restates EssentialSolver::FivePoints / Solve5PointEssential (usac/estimator/essential/five_points.cpp:13-274): 5x9 design
 * matrix (:48-63), 4-D null space E = x*B0 + y*B1 + z*B2 + B3 (:65-105), the ten cubic constraints 2EE'E - tr(EE')E = 0 and
 * det E = 0 arranged as a 10x10 matrix M(z) over the monomials [x^3 y^3 x^2y xy^2 x^2 y^2 xy x y 1] (mblock.hpp, generated
 * code there; polynomial algebra here), det M(z) recovered from its values at z = -5..5 (:117-138), its real roots (:140-157),
 * per root the null vector of M(z) -> x, y -> E (:181-204), and the cheirality vote over the five points for the four
 * (R, t) decompositions (:206-252); one E (or none) is returned, cast to float (:28).
 *
 * Parity: "partial". The reference's own solver is compiled in oracle/_ref (five_points.cpp + rpoly.cpp + mblock.hpp on the
 * stand-in cv::SVD, which agrees with the real OpenCV's null SPACE to 4e-14): tests/test_ref_build.py shows that the ONE matrix it
 * returns is a member of this file's candidate set and cheirality-valid here (53 of 55 samples), and that rpoly's real roots are
 * this file's roots (1e-6). What cannot be pinned is WHICH candidate comes first: that depends on the basis cv::SVD picks inside
 * the null space and on Jenkins-Traub's visiting order ("parity unpinned" for the root order). Where the
 * result is basis/ordering dependent this file fixes a deterministic choice, shared with the CUDA solver:
 *   - null spaces by Gauss-Jordan with partial pivoting (orc_null_space) instead of SVD rows (any basis spans the same E's);
 *   - polynomial coefficients by Newton divided differences instead of inverting the Vandermonde matrix;
 *   - real roots by derivative bracketing + bisection, visited in order of ascending |z| (Jenkins-Traub finds zeros in
 *     roughly increasing modulus; the reference keeps the first root whose decomposition passes the vote);
 *   - x, y from Gauss-Jordan on [M(z) | last column] instead of the last right-singular vector;
 *   - (R, t) in closed form (R = cof(E) -/+ [t]x E for unit-scale E) and depths from the 2x2 least-squares triangulation
 *     instead of SVD decomposition / DLT triangulation - the same four cameras and, for consistent points, the same signs.
 * All arithmetic is double, one rounding per operator (-ffp-contract=off), only + - * / sqrt and comparisons.
 */
#include "oracle_internal.h"

#include <cmath>
#include <cstring>

namespace {

/* monomials x^i y^j z^k, i+j+k <= 3, grouped so that the xy-part indexes the columns of M(z) */
struct Mono { int i, j, k; };
static const Mono MONO[20] = {
    {3, 0, 0}, {0, 3, 0}, {2, 1, 0}, {1, 2, 0},                                  /* x^3 y^3 x^2y xy^2            (z^0)      */
    {2, 0, 0}, {2, 0, 1}, {0, 2, 0}, {0, 2, 1}, {1, 1, 0}, {1, 1, 1},            /* x^2, y^2, xy                 (z^0, z^1) */
    {1, 0, 0}, {1, 0, 1}, {1, 0, 2}, {0, 1, 0}, {0, 1, 1}, {0, 1, 2},            /* x, y                         (z^0..z^2) */
    {0, 0, 0}, {0, 0, 1}, {0, 0, 2}, {0, 0, 3}};                                 /* 1                            (z^0..z^3) */
/* column c of M(z) owns MONO[COL_FIRST[c]] .. (COL_DEG[c] + 1 entries, ascending power of z) */
static const int COL_FIRST[10] = {0, 1, 2, 3, 4, 6, 8, 10, 13, 16};
static const int COL_DEG[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};

static int mono_index(int i, int j, int k) {
    for (int m = 0; m < 20; m++) if (MONO[m].i == i && MONO[m].j == j && MONO[m].k == k) return m;
    return -1;
}

/* linear forms: [x y z 1]; quadratics: 10 coefficients indexed through QMONO; cubics: 20 coefficients (MONO) */
static const Mono QMONO[10] = {{2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
static const Mono LMONO[4] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};

struct Tables {
    int ll[4][4];      /* linear x linear -> quadratic index */
    int ql[10][4];     /* quadratic x linear -> cubic index  */
    Tables() {
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++)
                for (int q = 0; q < 10; q++)
                    if (QMONO[q].i == LMONO[a].i + LMONO[b].i && QMONO[q].j == LMONO[a].j + LMONO[b].j && QMONO[q].k == LMONO[a].k + LMONO[b].k) ll[a][b] = q;
        for (int q = 0; q < 10; q++)
            for (int b = 0; b < 4; b++) ql[q][b] = mono_index(QMONO[q].i + LMONO[b].i, QMONO[q].j + LMONO[b].j, QMONO[q].k + LMONO[b].k);
    }
};
static const Tables T;

/* q += s * a*b (linear forms) */
static void acc_ll(double* q, const double* a, const double* b, double s) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) { double t = a[i] * b[j]; t = t * s; q[T.ll[i][j]] = q[T.ll[i][j]] + t; }
}
/* c += s * q*l */
static void acc_ql(double* c, const double* q, const double* l, double s) {
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 4; j++) { double t = q[i] * l[j]; t = t * s; c[T.ql[i][j]] = c[T.ql[i][j]] + t; }
}

/* determinant of an n x n matrix (row-major, destroyed): LU with partial pivoting */
static double det_lu(double* a, int n) {
    double det = 1.0;
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = std::fabs(a[k * n + k]);
        for (int r = k + 1; r < n; r++) { double v = std::fabs(a[r * n + k]); if (v > best) { best = v; piv = r; } }
        if (!(best > 0.0)) return 0.0;
        if (piv != k) { for (int j = 0; j < n; j++) { double t = a[k * n + j]; a[k * n + j] = a[piv * n + j]; a[piv * n + j] = t; } det = -det; }
        det = det * a[k * n + k];
        double inv = 1.0 / a[k * n + k];
        for (int r = k + 1; r < n; r++) {
            double f = a[r * n + k] * inv;
            for (int j = k + 1; j < n; j++) { double t = f * a[k * n + j]; a[r * n + j] = a[r * n + j] - t; }
        }
    }
    return det;
}

static double horner(const double* c, int deg, double x) {   /* c[0] + c[1] x + ... */
    double v = c[deg];
    for (int i = deg - 1; i >= 0; i--) { v = v * x; v = v + c[i]; }
    return v;
}

static double horner_rev(const double* c, int deg, double w) {   /* c[0] w^deg + c[1] w^(deg-1) + ... + c[deg] */
    double v = c[0];
    for (int i = 1; i <= deg; i++) { v = v * w; v = v + c[i]; }
    return v;
}

/* root of p (degree deg, derivative dc of degree deg-1) in [l, r] with p(l), p(r) of opposite sign (or zero): Newton steps kept inside
 * the bracket, a bisection step whenever Newton leaves it or stops halving it (the classical safeguarded iteration), until the step is
 * below two ulps of the root; <= 200 steps. ~7 evaluations of p and p' instead of ~55 evaluations of p for plain bisection. */
static double bisect(const double* c, const double* dc, int deg, double l, double r, double pl) {
    if (pl == 0.0) return l;
    double xl = pl < 0.0 ? l : r, xh = pl < 0.0 ? r : l;          /* p(xl) < 0 < p(xh) */
    double x = 0.5 * (l + r), dxold = std::fabs(r - l), dx = dxold;
    double f = horner(c, deg, x), df = horner(dc, deg - 1, x);
    for (int it = 0; it < 200; it++) {
        if (f == 0.0) return x;
        const double a = (x - xh) * df - f, b = (x - xl) * df - f;
        const bool newton = (a * b <= 0.0) && (std::fabs(2.0 * f) <= std::fabs(dxold * df));   /* false for NaN: bisect */
        dxold = dx;
        if (!newton) {
            dx = 0.5 * (xh - xl);
            x = xl + dx;
            if (xl == x) return x;
        } else {
            dx = f / df;
            const double t = x;
            x = x - dx;
            if (t == x) return x;
        }
        if (std::fabs(dx) <= 4.4e-16 * std::fabs(x)) return x;
        f = horner(c, deg, x);
        df = horner(dc, deg - 1, x);
        if (f < 0.0) xl = x; else xh = x;
    }
    return x;
}

/* real roots of c[0..deg] (ascending powers, c[deg] != 0) in ascending order: the roots of p' bracket the roots of p */
static int real_roots(const double* c, int deg, double* roots) {
    if (deg < 1) return 0;
    double bound = 0.0;
    for (int i = 0; i < deg; i++) { double v = std::fabs(c[i] / c[deg]); if (v > bound) bound = v; }
    bound = bound + 1.0;
    if (!(bound < 1e300)) return 0;
    double der[11][11];                              /* der[d] = coefficients of the (deg-d)-th derivative, degree d */
    for (int i = 0; i <= deg; i++) der[deg][i] = c[i];
    for (int d = deg; d > 1; d--)
        for (int i = 0; i < d; i++) der[d - 1][i] = der[d][i + 1] * (double)(i + 1);
    double prev[11], cur[11];
    int nprev = 0;
    prev[0] = -(der[1][0] / der[1][1]);
    nprev = 1;
    if (!(std::fabs(prev[0]) <= bound)) nprev = 0;
    for (int d = 2; d <= deg; d++) {
        const double* p = der[d];
        int ncur = 0;
        double left = -bound, pl = horner(p, d, left);
        for (int s = 0; s <= nprev; s++) {
            double right = s < nprev ? prev[s] : bound;
            double pr = horner(p, d, right);
            if (pl == 0.0) { cur[ncur++] = left; }
            else if (pr != 0.0 && ((pl < 0.0) != (pr < 0.0))) { cur[ncur++] = bisect(p, der[d - 1], d, left, right, pl); }
            left = right; pl = pr;
        }
        if (pl == 0.0) cur[ncur++] = left;
        for (int i = 0; i < ncur; i++) prev[i] = cur[i];
        nprev = ncur;
    }
    for (int i = 0; i < nprev; i++) roots[i] = prev[i];
    return nprev;
}

/* candidate E's of one sample, in the order the reference's loop would visit them (ascending |z| here); valid[] marks
 * the ones whose decomposition puts all five points in front of both cameras */
static int essential_candidates(const float* pts, const int* s, double (*Es)[9], int* valid, double* coef_out = nullptr) {
    double x1[5], y1[5], x2[5], y2[5], A[5 * 9];
    for (int i = 0; i < 5; i++) {
        const float* p = pts + 4 * (size_t)s[i];
        x1[i] = p[0]; y1[i] = p[1]; x2[i] = p[2]; y2[i] = p[3];
        double* r = A + 9 * i;                                           /* five_points.cpp:48-63 */
        r[0] = x1[i] * x2[i]; r[1] = x2[i] * y1[i]; r[2] = x2[i]; r[3] = x1[i] * y2[i]; r[4] = y1[i] * y2[i]; r[5] = y2[i];
        r[6] = x1[i]; r[7] = y1[i]; r[8] = 1.0;
    }
    double Bs[4 * 9];
    if (!orc_null_space(A, 5, Bs)) return 0;
    /* orthonormal basis of the same null space (modified Gram-Schmidt, last vector first so that B3 keeps the direction of
     * the Gauss-Jordan vector with E22 = 1): the SVD basis of the reference is orthonormal too, which keeps (x, y, z) of
     * typical solutions O(1) - with the raw Gauss-Jordan basis z = E21/E22 is often large and the interpolated
     * polynomial (nodes -5..5) loses those roots. */
    for (int a = 3; a >= 0; a--) {
        double* v = Bs + 9 * a;
        for (int b = 3; b > a; b--) {
            const double* u = Bs + 9 * b;
            double dot = 0.0;
            for (int e = 0; e < 9; e++) { double t = v[e] * u[e]; dot = dot + t; }
            for (int e = 0; e < 9; e++) { double t = dot * u[e]; v[e] = v[e] - t; }
        }
        double nn = 0.0;
        for (int e = 0; e < 9; e++) { double t = v[e] * v[e]; nn = nn + t; }
        if (!(nn > 0.0)) return 0;
        const double inv = 1.0 / std::sqrt(nn);
        for (int e = 0; e < 9; e++) v[e] = v[e] * inv;
    }
    /* E_ij as linear forms in (x, y, z, 1) */
    double L[9][4];
    for (int e = 0; e < 9; e++) for (int q = 0; q < 4; q++) L[e][q] = Bs[q * 9 + e];
    /* Q = E E' (quadratics) */
    double Q[3][3][10];
    std::memset(Q, 0, sizeof(Q));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) acc_ll(Q[i][j], L[3 * i + k], L[3 * j + k], 1.0);
    double tr[10];
    for (int m = 0; m < 10; m++) tr[m] = (Q[0][0][m] + Q[1][1][m]) + Q[2][2][m];
    /* rows 0..8: 2 (E E' E)_ij - tr * E_ij ; row 9: det E */
    double C[10][20];
    std::memset(C, 0, sizeof(C));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double* c = C[3 * i + j];
            for (int k = 0; k < 3; k++) acc_ql(c, Q[i][k], L[3 * k + j], 2.0);
            acc_ql(c, tr, L[3 * i + j], -1.0);
        }
    {
        double m0[10], m1[10], m2[10];
        std::memset(m0, 0, sizeof(m0)); std::memset(m1, 0, sizeof(m1)); std::memset(m2, 0, sizeof(m2));
        acc_ll(m0, L[4], L[8], 1.0); acc_ll(m0, L[5], L[7], -1.0);       /* E11 E22 - E12 E21 */
        acc_ll(m1, L[3], L[8], 1.0); acc_ll(m1, L[5], L[6], -1.0);       /* E10 E22 - E12 E20 */
        acc_ll(m2, L[3], L[7], 1.0); acc_ll(m2, L[4], L[6], -1.0);       /* E10 E21 - E11 E20 */
        acc_ql(C[9], m0, L[0], 1.0); acc_ql(C[9], m1, L[1], -1.0); acc_ql(C[9], m2, L[2], 1.0);
    }
    /* det M(z) is a polynomial of degree 10 in z. The reference samples it at z = -5..5 (:117-138); with an orthonormal basis
     * the wanted roots are O(1) or smaller and those nodes bury the low-order coefficients in rounding noise (measured: no
     * real root found for 10-40 % of noise-free all-inlier samples). Two well-conditioned interpolations instead, both on 11
     * equispaced nodes in [-1, 1]: p(z) for the roots with |z| <= 1.05, and q(w) = w^10 p(1/w) = det(M(1/w) diag(w^deg_c))
     * for the roots with |w| < 1/1.05, i.e. |z| > 1.05. */
    double roots[20];
    int nroots = 0;
    for (int pass = 0; pass < 2; pass++) {
        double zs[11], dd[11], Mz[100];
        for (int t = 0; t < 11; t++) {
            const double z = (double)(t - 5) / 5.0;
            for (int r = 0; r < 10; r++)
                for (int c = 0; c < 10; c++)
                    Mz[r * 10 + c] = pass == 0 ? horner(&C[r][COL_FIRST[c]], COL_DEG[c], z) : horner_rev(&C[r][COL_FIRST[c]], COL_DEG[c], z);
            zs[t] = z;
            dd[t] = det_lu(Mz, 10);
        }
        for (int lev = 1; lev < 11; lev++)                                   /* Newton divided differences */
            for (int t = 10; t >= lev; t--) dd[t] = (dd[t] - dd[t - 1]) / (zs[t] - zs[t - lev]);
        double coef[11];
        for (int i = 0; i < 11; i++) coef[i] = 0.0;
        coef[0] = dd[10];
        for (int t = 9; t >= 0; t--) {                                       /* coef = coef * (z - zs[t]) + dd[t] */
            for (int i = 10; i >= 1; i--) { double a = coef[i] * zs[t]; coef[i] = coef[i - 1] - a; }
            { double a = coef[0] * zs[t]; coef[0] = dd[t] - a; }
        }
        if (coef_out && pass == 0) for (int i = 0; i < 11; i++) coef_out[i] = coef[i];
        int deg = 10;
        double cmax = 0.0;
        bool finite = true;
        for (int i = 0; i <= 10; i++) { if (!std::isfinite(coef[i])) finite = false; if (std::fabs(coef[i]) > cmax) cmax = std::fabs(coef[i]); }
        if (!finite) return 0;
        while (deg > 0 && !(std::fabs(coef[deg]) > 1e-13 * cmax)) deg--;
        if (deg < 1) continue;
        double rr[10];
        const int nr = real_roots(coef, deg, rr);
        for (int i = 0; i < nr; i++) {
            if (pass == 0) { if (std::fabs(rr[i]) <= 1.05) roots[nroots++] = rr[i]; }
            else if (std::fabs(rr[i]) < 1.0 / 1.05 && rr[i] != 0.0) roots[nroots++] = 1.0 / rr[i];
        }
    }
    if (nroots > 10) nroots = 10;
    /* visit in order of ascending |z| (stable insertion sort) */
    for (int i = 1; i < nroots; i++) {
        double v = roots[i];
        int j = i - 1;
        while (j >= 0 && std::fabs(roots[j]) > std::fabs(v)) { roots[j + 1] = roots[j]; j--; }
        roots[j + 1] = v;
    }
    double Mz[100];
    int ncand = 0;
    for (int ri = 0; ri < nroots; ri++) {
        const double z = roots[ri];
        /* [M(z)] v = 0 with v[9] = 1: Gauss-Jordan on the 10 x 10 system, pivots in columns 0..8 (:181-191) */
        for (int r = 0; r < 10; r++)
            for (int c = 0; c < 10; c++) Mz[r * 10 + c] = horner(&C[r][COL_FIRST[c]], COL_DEG[c], z);
        bool ok = true;
        for (int k = 0; k < 9 && ok; k++) {
            int piv = k;
            double best = std::fabs(Mz[k * 10 + k]);
            for (int r = k + 1; r < 10; r++) { double v = std::fabs(Mz[r * 10 + k]); if (v > best) { best = v; piv = r; } }
            if (!(best > 0.0) || !std::isfinite(best)) { ok = false; break; }
            if (piv != k) for (int j = 0; j < 10; j++) { double t = Mz[k * 10 + j]; Mz[k * 10 + j] = Mz[piv * 10 + j]; Mz[piv * 10 + j] = t; }
            double inv = 1.0 / Mz[k * 10 + k];
            for (int j = k + 1; j < 10; j++) Mz[k * 10 + j] = Mz[k * 10 + j] * inv;
            for (int r = 0; r < 10; r++) {
                if (r == k) continue;
                double f = Mz[r * 10 + k];
                for (int j = k + 1; j < 10; j++) { double t = f * Mz[k * 10 + j]; Mz[r * 10 + j] = Mz[r * 10 + j] - t; }
            }
        }
        if (!ok) continue;
        double u[3] = {-Mz[7 * 10 + 9], -Mz[8 * 10 + 9], z};
        /* eight Gauss-Newton steps on the ten cubic constraints: removes the error the interpolated determinant leaves in z
         * (and hence in x, y) - the reference has no such step; it only makes the returned E more accurate */
        for (int it = 0; it < 8; it++) {
            double pw[3][4];
            for (int a = 0; a < 3; a++) { pw[a][0] = 1.0; pw[a][1] = u[a]; pw[a][2] = u[a] * u[a]; pw[a][3] = pw[a][2] * u[a]; }
            double JtJ[6] = {0, 0, 0, 0, 0, 0}, Jtr[3] = {0, 0, 0};
            for (int r = 0; r < 10; r++) {
                double val = 0.0, g[3] = {0, 0, 0};
                for (int m = 0; m < 20; m++) {
                    const int e[3] = {MONO[m].i, MONO[m].j, MONO[m].k};
                    const double c = C[r][m];
                    { double t = (pw[0][e[0]] * pw[1][e[1]]) * pw[2][e[2]]; t = c * t; val = val + t; }
                    for (int a = 0; a < 3; a++) {
                        if (e[a] == 0) continue;
                        double t = (double)e[a];
                        for (int b = 0; b < 3; b++) t = t * pw[b][b == a ? e[b] - 1 : e[b]];
                        t = c * t; g[a] = g[a] + t;
                    }
                }
                { double t;
                  t = g[0] * g[0]; JtJ[0] = JtJ[0] + t; t = g[0] * g[1]; JtJ[1] = JtJ[1] + t; t = g[0] * g[2]; JtJ[2] = JtJ[2] + t;
                  t = g[1] * g[1]; JtJ[3] = JtJ[3] + t; t = g[1] * g[2]; JtJ[4] = JtJ[4] + t; t = g[2] * g[2]; JtJ[5] = JtJ[5] + t;
                  t = g[0] * val; Jtr[0] = Jtr[0] + t; t = g[1] * val; Jtr[1] = Jtr[1] + t; t = g[2] * val; Jtr[2] = Jtr[2] + t; }
            }
            /* solve the symmetric 3x3 system by cofactors */
            const double a = JtJ[0], b = JtJ[1], c = JtJ[2], d = JtJ[3], e = JtJ[4], f = JtJ[5];
            const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d, c11 = a * f - c * c, c12 = b * c - a * e, c22 = a * d - b * b;
            const double det = (a * c00 + b * c01) + c * c02;
            if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) break;
            const double dx = ((c00 * Jtr[0] + c01 * Jtr[1]) + c02 * Jtr[2]) / det;
            const double dy = ((c01 * Jtr[0] + c11 * Jtr[1]) + c12 * Jtr[2]) / det;
            const double dz = ((c02 * Jtr[0] + c12 * Jtr[1]) + c22 * Jtr[2]) / det;
            if (!std::isfinite(dx) || !std::isfinite(dy) || !std::isfinite(dz)) break;
            u[0] = u[0] - dx; u[1] = u[1] - dy; u[2] = u[2] - dz;
        }
        const double x = u[0], y = u[1], zz = u[2];
        double E[9];
        bool finite = true;
        for (int e = 0; e < 9; e++) {                                    /* :194-204 */
            double v = L[e][0] * x; double w = L[e][1] * y; v = v + w; w = L[e][2] * zz; v = v + w; v = v + L[e][3];
            E[e] = v;
            if (!std::isfinite(v)) finite = false;
        }
        if (!finite) continue;
        /* cheirality vote (:206-252) on the unit-scale E: t = left null vector, R = cof(E) -/+ [t]x E */
        double n2 = 0.0;
        for (int e = 0; e < 9; e++) { double t = E[e] * E[e]; n2 = n2 + t; }
        double sc = std::sqrt(0.5 * n2);
        if (!(sc > 0.0)) continue;
        double En[9];
        for (int e = 0; e < 9; e++) En[e] = E[e] / sc;
        /* left null vector: cross products of the columns, the longest one */
        double col[3][3];
        for (int c = 0; c < 3; c++) for (int r = 0; r < 3; r++) col[c][r] = En[3 * r + c];
        double tv[3] = {0, 0, 0}, tbest = -1.0;
        for (int a = 0; a < 3; a++) {
            const int b = (a + 1) % 3;
            double cx = col[a][1] * col[b][2] - col[a][2] * col[b][1];
            double cy = col[a][2] * col[b][0] - col[a][0] * col[b][2];
            double cz = col[a][0] * col[b][1] - col[a][1] * col[b][0];
            double nn = (cx * cx + cy * cy) + cz * cz;
            if (nn > tbest) { tbest = nn; tv[0] = cx; tv[1] = cy; tv[2] = cz; }
        }
        if (!(tbest > 0.0)) continue;
        { double inv = 1.0 / std::sqrt(tbest); tv[0] = tv[0] * inv; tv[1] = tv[1] * inv; tv[2] = tv[2] * inv; }
        double cof[9], tE[9];
        cof[0] = En[4] * En[8] - En[5] * En[7]; cof[1] = En[5] * En[6] - En[3] * En[8]; cof[2] = En[3] * En[7] - En[4] * En[6];
        cof[3] = En[2] * En[7] - En[1] * En[8]; cof[4] = En[0] * En[8] - En[2] * En[6]; cof[5] = En[1] * En[6] - En[0] * En[7];
        cof[6] = En[1] * En[5] - En[2] * En[4]; cof[7] = En[2] * En[3] - En[0] * En[5]; cof[8] = En[0] * En[4] - En[1] * En[3];
        for (int c = 0; c < 3; c++) {                                    /* [t]x E, column by column */
            tE[0 + c] = tv[1] * En[6 + c] - tv[2] * En[3 + c];
            tE[3 + c] = tv[2] * En[0 + c] - tv[0] * En[6 + c];
            tE[6 + c] = tv[0] * En[3 + c] - tv[1] * En[0 + c];
        }
        bool pass = false;
        for (int cam = 0; cam < 4 && !pass; cam++) {
            double R[9], t[3];
            const double rs = (cam < 2) ? -1.0 : 1.0, ts = (cam & 1) ? -1.0 : 1.0;
            for (int e = 0; e < 9; e++) R[e] = cof[e] + rs * tE[e];
            for (int e = 0; e < 3; e++) t[e] = ts * tv[e];
            int infront = 0;
            for (int k = 0; k < 5; k++) {
                /* lambda2 * p2 = lambda1 * R p1 + t in least squares */
                double a0 = (R[0] * x1[k] + R[1] * y1[k]) + R[2], a1 = (R[3] * x1[k] + R[4] * y1[k]) + R[5], a2 = (R[6] * x1[k] + R[7] * y1[k]) + R[8];
                double b0 = x2[k], b1 = y2[k], b2 = 1.0;
                double aa = (a0 * a0 + a1 * a1) + a2 * a2, bb = (b0 * b0 + b1 * b1) + b2 * b2, ab = (a0 * b0 + a1 * b1) + a2 * b2;
                double at = (a0 * t[0] + a1 * t[1]) + a2 * t[2], bt = (b0 * t[0] + b1 * t[1]) + b2 * t[2];
                double det = aa * bb - ab * ab;
                double l1 = (ab * bt - bb * at) / det, l2 = (aa * bt - ab * at) / det;
                if (l1 > 0.0 && l2 > 0.0) infront++; else break;
            }
            if (infront == 5) pass = true;
        }
        for (int e = 0; e < 9; e++) Es[ncand][e] = E[e];
        valid[ncand] = pass ? 1 : 0;
        ncand++;
    }
    return ncand;
}

}  // namespace

/* EssentialEstimator::EstimateModel (essential_estimator.hpp:52-62): 0 or 1 model */
int orc_solve_essential5(const float* pts, const int* s, float* out) {
    double Es[10][9];
    int valid[10];
    const int n = essential_candidates(pts, s, Es, valid);
    for (int i = 0; i < n; i++)
        if (valid[i]) {
            for (int e = 0; e < 9; e++) out[e] = (float)Es[i][e];
            return 1;
        }
    return 0;
}

/* all candidates (for the unit tests): Es_out[10][9] doubles, valid_out[10]; returns the count */
extern "C" int orc_essential5_poly(const float* pts, const int* sample, double* coef_out) {
    double Es[10][9];
    int valid[10];
    for (int i = 0; i < 11; i++) coef_out[i] = 0.0;
    return essential_candidates(pts, sample, Es, valid, coef_out);
}

extern "C" int orc_essential5_candidates(const float* pts, const int* sample, double* Es_out, int* valid_out) {
    double Es[10][9];
    int valid[10];
    const int n = essential_candidates(pts, sample, Es, valid);
    for (int i = 0; i < n; i++) { for (int e = 0; e < 9; e++) Es_out[9 * i + e] = Es[i][e]; valid_out[i] = valid[i]; }
    return n;
}
