// lo.cuh - LO-RANSAC as ONE device-side loop: InnerLocalOptimization::GetModelScore with its IterativeLocalOptimization
// (usac/local_optimization/inner_local_optimization.hpp:74-133, iterative_local_optimization.hpp:61-135) for a so-far-the-best
// model, one launch of one CTA of 1024 threads per call. In round 1 the control flow was host code between device calls
// (~100 stream synchronisations per LO call, 27 ms for a homography fit with LO); here the host uploads the model and reads
// the result back - one synchronisation.
//
// The arithmetic is that of refit.cuh / score.cuh, order included, so the results stay bit-identical to the CPU restatement the parity tests use:
//  * non-minimal estimation (cta_nonminimal): lanes = the first 256 threads; every sum over the points is thread t adding
//    elements t, t + 256, ... in order, then the fixed binary tree stride 128 ... 1 over the 256 partials. The 45 entries of A'A
//    are accumulated in ONE pass over the points (each entry still sees its addends in the same order) and go through the tree
//    together: 5 CTA barriers per fit instead of 45 x 9.
//  * scoring with the ordered inlier list (cta_score): lanes = all 1024 threads, errors in the reference's arithmetic
//    (strict_error), error sum = lane sums + fixed tree stride 512 ... 1.
#pragma once
#include "pipeline.cuh"
#include "refit.cuh"
#include "score.cuh"

#define LO_THREADS 1024

struct LoIO {                    // in/out record of one LO call (global memory)
    float model[9];
    int inliers;
    float score;
    float lo_thr;                // IterativeLocalOptimization's running threshold (a member in the reference: it survives calls)
    unsigned long long calls;    // keys the random inlier subsets
    unsigned inner_done, iterative_done;
};

struct LoArgs {
    const float* aos;            // points of the problem
    const ProblemDesc* prob;
    int problem, n, m, kind, sample_limit, inner_iters, iter_iters, mult;
    float theta, step;
    unsigned long long seed;
    LoIO* io;
    int* A;                      // inliers of the best model (n entries)
    int* B;                      // inliers of the candidate (n entries)
};

struct LoShared {
    double tree[45 * REFIT_THREADS];     // A'A partials of the 256 lanes (also used for the float trees)
    double S[81], V[81], h[9];
    float part[LO_THREADS];              // error-sum partials of the 1024 scoring lanes
    float rec[USAC_REC_STRIDE];
    float model[9], best_model[9], stats[8];
    int warp_tot[32];
    int sample[16];
    int base, ok, cnt;
    float sum;
};

// fixed binary tree over 256 float / double partials of `count` quantities laid out tree[q * 256 + t]; result q in tree[q * 256].
// Pairing and order are those of refit_tree_f / refit_tree_d (sm[t] += sm[t + s], s = 128 ... 1).
template <class T>
__device__ __forceinline__ void lo_tree256(T* tree, int count) {
    const int t = threadIdx.x;
    __syncthreads();
    for (int s = REFIT_THREADS / 2; s >= 32; s >>= 1) {               // partners live in different warps
        for (int i = t; i < count * s; i += LO_THREADS) {
            const int q = i / s, l = i % s;
            if (sizeof(T) == 8) tree[q * REFIT_THREADS + l] = (T)__dadd_rn((double)tree[q * REFIT_THREADS + l], (double)tree[q * REFIT_THREADS + l + s]);
            else tree[q * REFIT_THREADS + l] = (T)__fadd_rn((float)tree[q * REFIT_THREADS + l], (float)tree[q * REFIT_THREADS + l + s]);
        }
        __syncthreads();
    }
    // s = 16 ... 1: one warp per quantity, partners in the same warp
    const int warp = t >> 5, lane = t & 31;
    for (int q = warp; q < count; q += LO_THREADS / 32) {
        for (int s = 16; s > 0; s >>= 1) {
            if (lane < s) {
                if (sizeof(T) == 8) tree[q * REFIT_THREADS + lane] = (T)__dadd_rn((double)tree[q * REFIT_THREADS + lane], (double)tree[q * REFIT_THREADS + lane + s]);
                else tree[q * REFIT_THREADS + lane] = (T)__fadd_rn((float)tree[q * REFIT_THREADS + lane], (float)tree[q * REFIT_THREADS + lane + s]);
            }
            __syncwarp();
        }
    }
    __syncthreads();
}

// Estimator::EstimateModelNonMinimalSample on `n` point ids -> sh.model, sh.ok. Same arithmetic as nonminimal_kernel (refit.cuh).
template <int EST>
__device__ void cta_nonminimal(const float* __restrict__ pts, const int* ids, int n, LoShared& sh) {
    const int t = threadIdx.x;
    const bool lane_on = t < REFIT_THREADS;
    const float fn = (float)n;
    float* treef = reinterpret_cast<float*>(sh.tree);
    __syncthreads();                                                   // everyone has read the previous call's sh.ok / sh.model
    if (t == 0) sh.ok = 0;
    if (n < 4) { __syncthreads(); return; }                            // uniform
    // ---- normalising transformations ----
    if (lane_on) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = t; i < n; i += REFIT_THREADS) {
            const float4 p = reinterpret_cast<const float4*>(pts)[ids[i]];
            a[0] = __fadd_rn(a[0], p.x); a[1] = __fadd_rn(a[1], p.y); a[2] = __fadd_rn(a[2], p.z); a[3] = __fadd_rn(a[3], p.w);
        }
        for (int k = 0; k < 4; k++) treef[k * REFIT_THREADS + t] = a[k];
    }
    lo_tree256<float>(treef, 4);
    float m[4];
    for (int k = 0; k < 4; k++) m[k] = __fdiv_rn(treef[k * REFIT_THREADS], fn);
    __syncthreads();
    if (lane_on) {
        float d1 = 0.f, d2 = 0.f;
        for (int i = t; i < n; i += REFIT_THREADS) {
            const float4 p = reinterpret_cast<const float4*>(pts)[ids[i]];
            const sf a = sf(p.x) - sf(m[0]), b = sf(p.y) - sf(m[1]), c = sf(p.z) - sf(m[2]), d = sf(p.w) - sf(m[3]);
            d1 = __fadd_rn(d1, ssqrt(a * a + b * b).v);
            d2 = __fadd_rn(d2, ssqrt(c * c + d * d).v);
        }
        treef[t] = d1; treef[REFIT_THREADS + t] = d2;
    }
    lo_tree256<float>(treef, 2);
    const float d1 = treef[0], d2 = treef[REFIT_THREADS];
    __syncthreads();
    const double SQRT2 = 1.41421356237309504880;
    const float s1 = (float)(sd(SQRT2) / sd((double)__fdiv_rn(d1, fn))).v, s2 = (float)(sd(SQRT2) / sd((double)__fdiv_rn(d2, fn))).v;
    const float t1x = (-sf(m[0]) * sf(s1)).v, t1y = (-sf(m[1]) * sf(s1)).v, t2x = (-sf(m[2]) * sf(s2)).v, t2y = (-sf(m[3]) * sf(s2)).v;
    if (!isfinite(s1) || !isfinite(s2)) return;                        // uniform (every thread holds the same values)
    // ---- A'A: one pass over the points, 45 accumulators per lane (each entry adds its terms in the same order as a pass of its own) ----
    constexpr int NROWS = (EST == USAC_EST_HOMOGRAPHY) ? 2 : 1;
    if (lane_on) {
        double acc[45];
#pragma unroll
        for (int e = 0; e < 45; e++) acc[e] = 0.0;
        for (int k = t; k < n; k += REFIT_THREADS) {
            const float4 p = reinterpret_cast<const float4*>(pts)[ids[k]];
            const float x1 = (sf(s1) * sf(p.x) + sf(t1x)).v, y1 = (sf(s1) * sf(p.y) + sf(t1y)).v;
            const float x2 = (sf(s2) * sf(p.z) + sf(t2x)).v, y2 = (sf(s2) * sf(p.w) + sf(t2y)).v;
            float r[2][9];
            if (EST == USAC_EST_HOMOGRAPHY) {
                r[0][0] = -x1; r[0][1] = -y1; r[0][2] = -1.f; r[0][3] = 0.f; r[0][4] = 0.f; r[0][5] = 0.f;
                r[0][6] = __fmul_rn(x2, x1); r[0][7] = __fmul_rn(x2, y1); r[0][8] = x2;
                r[1][0] = 0.f; r[1][1] = 0.f; r[1][2] = 0.f; r[1][3] = -x1; r[1][4] = -y1; r[1][5] = -1.f;
                r[1][6] = __fmul_rn(y2, x1); r[1][7] = __fmul_rn(y2, y1); r[1][8] = y2;
            } else {
                r[0][0] = __fmul_rn(x2, x1); r[0][1] = __fmul_rn(x2, y1); r[0][2] = x2; r[0][3] = __fmul_rn(y2, x1); r[0][4] = __fmul_rn(y2, y1);
                r[0][5] = y2; r[0][6] = x1; r[0][7] = y1; r[0][8] = 1.f;
                for (int q = 0; q < 9; q++) r[1][q] = 0.f;
            }
            int e = 0;
#pragma unroll
            for (int i = 0; i < 9; i++)
#pragma unroll
                for (int j = i; j < 9; j++) {
                    double a2 = 0.0;
#pragma unroll
                    for (int q = 0; q < NROWS; q++) a2 = __dadd_rn(a2, __dmul_rn((double)r[q][i], (double)r[q][j]));
                    acc[e] = __dadd_rn(acc[e], a2);
                    e++;
                }
        }
#pragma unroll
        for (int e = 0; e < 45; e++) sh.tree[e * REFIT_THREADS + t] = acc[e];
    }
    lo_tree256<double>(sh.tree, 45);
    if (t < 45) {
        int i = 0, e = t;
        while (e >= 9 - i) { e -= 9 - i; i++; }
        const int j = i + e;
        const double v = sh.tree[t * REFIT_THREADS];
        sh.S[i * 9 + j] = v; sh.S[j * 9 + i] = v;
    }
    __syncthreads();
    if (t < 32) refit_smallest_eigenvector_warp(sh.S, 9, sh.h, sh.V);
    if (t == 0) {
        const double* h = sh.h;
        sd M[9], R[9];
        const sd S1((double)s1), T1x((double)t1x), T1y((double)t1y), S2((double)s2), T2x((double)t2x), T2y((double)t2y);
        for (int i = 0; i < 3; i++) {
            M[3 * i] = sd(h[3 * i]) * S1; M[3 * i + 1] = sd(h[3 * i + 1]) * S1;
            M[3 * i + 2] = (sd(h[3 * i]) * T1x + sd(h[3 * i + 1]) * T1y) + sd(h[3 * i + 2]);
        }
        bool ok = true;
        if (EST == USAC_EST_HOMOGRAPHY) {
            const sd is2 = sd(1.0) / S2, ux = -(T2x * is2), uy = -(T2y * is2);
            for (int j = 0; j < 3; j++) { R[j] = is2 * M[j] + ux * M[6 + j]; R[3 + j] = is2 * M[3 + j] + uy * M[6 + j]; R[6 + j] = M[6 + j]; }
            const sd inv = sd(1.0) / R[8];
            for (int i = 0; i < 9; i++) { const double v = (R[i] * inv).v; if (!dfinite(v)) ok = false; sh.model[i] = (float)v; }
            sh.model[8] = 1.f;
        } else {
            for (int j = 0; j < 3; j++) { R[j] = S2 * M[j]; R[3 + j] = S2 * M[3 + j]; R[6 + j] = (T2x * M[j] + T2y * M[3 + j]) + M[6 + j]; }
            const bool scale = fabsf((float)R[8].v) > 1.1920929e-07f;     // FLT_EPSILON
            const sd inv = scale ? sd(1.0) / R[8] : sd(1.0);
            for (int i = 0; i < 9; i++) { const double v = (R[i] * inv).v; if (!dfinite(v)) ok = false; sh.model[i] = (float)v; }
        }
        sh.ok = ok ? 1 : 0;
    }
    __syncthreads();
}

// Quality::getNumberInliers(score, model, thr, get_inliers = true, ids) (quality.hpp:60-101): sh.cnt, sh.sum, ids in ascending order.
// Same arithmetic as inliers_sum_kernel (score.cuh).
template <int EST>
__device__ void cta_score(const float* __restrict__ aos, int n, const float* model, float thr, const ProblemDesc& pd, int* ids, LoShared& sh) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) {
        float rec[USAC_REC_STRIDE];
        make_record(EST, model, thr, pd, rec);
        for (int i = 0; i < USAC_REC_STRIDE; i++) sh.rec[i] = rec[i];
        sh.base = 0;
    }
    __syncthreads();
    float acc = 0.f;
    for (int start = 0; start < n; start += LO_THREADS) {
        const int i = start + t;
        bool in = false;
        if (i < n) {
            const float4 p = reinterpret_cast<const float4*>(aos)[i];
            const float e = strict_error<EST>(sh.rec, p.x, p.y, p.z, p.w);
            in = e < thr;
            if (in) acc = __fadd_rn(acc, e);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) sh.warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; w++) { const int v = sh.warp_tot[w]; if (w < warp) before += v; total += v; }
        if (in) ids[sh.base + before + __popc(bal & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (t == 0) sh.base += total;
        __syncthreads();
    }
    sh.part[t] = acc;
    __syncthreads();
    for (int s = LO_THREADS / 2; s >= 32; s >>= 1) {
        if (t < s) sh.part[t] = __fadd_rn(sh.part[t], sh.part[t + s]);
        __syncthreads();
    }
    if (warp == 0) {
        for (int s = 16; s > 0; s >>= 1) {
            if (lane < s) sh.part[lane] = __fadd_rn(sh.part[lane], sh.part[lane + s]);
            __syncwarp();
        }
    }
    __syncthreads();
    if (t == 0) { sh.cnt = sh.base; sh.sum = sh.part[0]; }
    __threadfence_block();
    __syncthreads();
}

// 14 distinct positions in [0, count): two Philox draws (8 + 6), the second mapped past the first (LoRunner::fit_random_subset)
__device__ void lo_subset(unsigned long long seed, unsigned long long calls, int k, int count, const int* from, int* sample) {
    int pos[16], a[8], b[8];
    philox_unique(seed, calls, 7, count, k < 8 ? k : 8, a);
    for (int i = 0; i < (k < 8 ? k : 8); i++) pos[i] = a[i];
    if (k > 8) {
        philox_unique(seed, calls, 8, count - 8, k - 8, b);
        int sorted[8];
        for (int i = 0; i < 8; i++) sorted[i] = a[i];
        for (int i = 1; i < 8; i++) { const int v = sorted[i]; int j = i - 1; while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; j--; } sorted[j + 1] = v; }
        for (int i = 0; i < k - 8; i++) {
            int v = b[i];
            for (int q = 0; q < 8; q++) if (v >= sorted[q]) v++;
            pos[8 + i] = v;
        }
    }
    for (int i = 0; i < k; i++) sample[i] = from[pos[i]];
}

__device__ __forceinline__ bool lo_bigger(int ia, float sa, int ib, float sb) { return ia > ib || (ia == ib && sa > sb); }

template <int EST>
__global__ void __launch_bounds__(LO_THREADS) lo_kernel(const LoArgs a) {
    extern __shared__ __align__(16) unsigned char lo_smem[];
    LoShared& sh = *reinterpret_cast<LoShared*>(lo_smem);
    const int t = threadIdx.x;
    const ProblemDesc pd = a.prob[a.problem];
    // every thread follows the same control flow: the decisions depend on values that all threads read from shared / global memory
    // behind a CTA barrier
    int best_inl = a.io->inliers;
    float best_sum = a.io->score, lo_thr = a.io->lo_thr;
    unsigned long long calls = a.io->calls;
    unsigned inner_done = 0, iterative_done = 0;
    if (t < 9) sh.best_model[t] = a.io->model[t];
    __syncthreads();
    if (best_inl >= 12) {                                              // inner_local_optimization.hpp:76
        cta_score<EST>(a.aos, a.n, sh.best_model, a.theta, pd, a.A, sh);   // quality->getInliers(best_model)
        int avail = min(best_inl, sh.cnt);                             // ids present in A (never index past the list)
        for (int it = 0; it < a.inner_iters; it++) {
            if (avail > a.sample_limit) {
                if (t == 0) lo_subset(a.seed, calls, a.sample_limit, avail, a.A, sh.sample);
                calls++;
                __syncthreads();
                cta_nonminimal<EST>(a.aos, sh.sample, a.sample_limit, sh);
                if (!sh.ok) continue;
            } else {
                cta_nonminimal<EST>(a.aos, a.A, avail, sh);
                if (!sh.ok) break;
            }
            lo_thr = (unsigned)a.mult * lo_thr;                         // inner_local_optimization.hpp:101
            cta_score<EST>(a.aos, a.n, sh.model, lo_thr, pd, a.B, sh);
            int lo_inl = sh.cnt;
            float lo_sum = sh.sum;
            if (lo_inl <= a.m) continue;
            // ---- IterativeLocalOptimization::GetModelScore ----
            for (int k = 0; k < a.iter_iters; k++) {
                lo_thr -= a.step;
                if (lo_inl <= a.m) break;
                if (a.kind == 2) {                                     // GetScoreLimited
                    if (lo_inl > a.sample_limit) {
                        if (t == 0) lo_subset(a.seed, calls, a.sample_limit, lo_inl, a.B, sh.sample);
                        calls++;
                        __syncthreads();
                        cta_nonminimal<EST>(a.aos, sh.sample, a.sample_limit, sh);
                        if (!sh.ok) continue;
                    } else {
                        cta_nonminimal<EST>(a.aos, a.B, lo_inl, sh);
                        if (!sh.ok) break;
                    }
                    cta_score<EST>(a.aos, a.n, sh.model, lo_thr, pd, a.B, sh);
                    lo_inl = sh.cnt; lo_sum = sh.sum;
                } else {                                               // GetScoreUnlimited
                    cta_nonminimal<EST>(a.aos, a.B, lo_inl, sh);
                    if (!sh.ok) break;
                    cta_score<EST>(a.aos, a.n, sh.model, lo_thr, pd, a.B, sh);
                    lo_inl = sh.cnt; lo_sum = sh.sum;
                    if (lo_bigger(best_inl, best_sum, lo_inl, lo_sum)) break;
                }
                iterative_done++;
            }
            bool fail = false;
            if (fabsf(lo_thr - a.theta) > 0.00001) { fail = true; lo_thr = a.theta; }
            if (!fail && lo_bigger(lo_inl, lo_sum, best_inl, best_sum)) {
                __syncthreads();
                if (t < 9) sh.best_model[t] = sh.model[t];
                for (int i = t; i < lo_inl; i += LO_THREADS) a.A[i] = a.B[i];
                best_inl = lo_inl; best_sum = lo_sum; avail = lo_inl;
                __threadfence_block();
                __syncthreads();
            }
            inner_done++;
        }
    }
    __syncthreads();
    if (t == 0) {
        for (int i = 0; i < 9; i++) a.io->model[i] = sh.best_model[i];
        a.io->inliers = best_inl; a.io->score = best_sum; a.io->lo_thr = lo_thr; a.io->calls = calls;
        a.io->inner_done = inner_done; a.io->iterative_done = iterative_done;
    }
}
