// sprt.cuh - Wald SPRT verification (usac/sprt.hpp) on the device.
//
// Two kernels per round. sprt_walk_kernel: one thread per model of the round walks the first SPRT_HEAD points of its stretch of the
// shuffled point pool (sprt.hpp:93-107; the points are stored once in pool order, so the walk is a contiguous stream) from its start
// offset (cursor + 32*q) mod N, updating the likelihood ratio lambda in double exactly as sprt.hpp:205-234 does (lambda *= delta/eps
// for an inlier, (1-delta)/(1-eps) otherwise; reject when lambda > A), under the test (eps, delta, A) frozen for the round. Almost all
// models of a round are rejected there. The few that are not (and the rejected models of the first hypotheses, which count every
// point) are handed to sprt_tail_kernel: one WARP per model, 64 points per step (lane = point, exact decisions, two ballots), then
// the same multiplications in the same order, replicated in every lane - a model that passes the test walks all N points in
// N/64 warp steps instead of N thread steps (C3: 2.5 ms -> 0.1 ms for the round). Errors come from the packed fast
// evaluator with the same guard band + strict re-evaluation as the scoring kernel, so every inlier decision - hence
// tested_pts / tested_inl - is bit-exact. Models rejected during the first 20 hypotheses finish counting their inliers
// (sprt.hpp:243-257).
#pragma once
#include "pipeline.cuh"
#include "score.cuh"

struct SprtModelResult { int good, tested_inl, tested_pts, full_inl; };
// a walk handed from the thread-per-model head to the warp-per-model tail
struct SprtCarry { int q, tp, tin, rejected; double lambda; unsigned start; int count_all; };
#ifndef USAC_SPRT_LOGWALK
#define USAC_SPRT_LOGWALK 1          // 1: the warp tail decides 64-point steps in the log domain (fixed point), exact chain only when too close to call
#endif
#ifndef USAC_SPRT_HEAD
#define USAC_SPRT_HEAD 64            // points a model walks in the thread-per-model kernel (even)
#endif

__host__ inline void sprt_init_state(FitState& s, int est) {
    // sprt.hpp:114-153 initial (epsilon, delta); A is designed on the host in usac_gpu_fit
    if (est == USAC_EST_HOMOGRAPHY) { s.sprt_delta = 0.01; s.sprt_eps = 0.1; }
    else if (est == USAC_EST_LINE2D) { s.sprt_delta = 0.0001; s.sprt_eps = 0.001; }
    else { s.sprt_delta = 0.05; s.sprt_eps = 0.2; }
    s.sprt_A = 0; s.sprt_cursor = 0; s.sprt_last_update = 0; s.sprt_ntests = 0;
}

// pool-ordered copy of the points: dst[i] = aos[pool[i]]
__global__ void pool_gather_kernel(const float* __restrict__ aos, const int* __restrict__ pool, int n, int dim, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int src = pool[i];
    if (dim == 4) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(aos)[src];
    else reinterpret_cast<float2*>(dst)[i] = reinterpret_cast<const float2*>(aos)[src];
}

template <int EST>
struct PoolPoint {
    float4 v;
    __device__ __forceinline__ void load(const float* __restrict__ P, int i) {
        if (EST == USAC_EST_LINE2D) { const float2 p = reinterpret_cast<const float2*>(P)[i]; v = make_float4(p.x, p.y, 0.f, 0.f); }
        else v = reinterpret_cast<const float4*>(P)[i];
    }
};

// inlier decisions of two pool points (positions i0, i1) for one model: fast value, strict when undecided
template <int EST>
__device__ __forceinline__ void decide_pair(const FastModel<EST>& fm, const float* __restrict__ rec, const float* __restrict__ P, int n,
                                            const PoolPoint<EST>& a, const PoolPoint<EST>& b, int i0, int i1, bool& in0, bool& in1) {
    float4 A, B;
    if (EST == USAC_EST_LINE2D) { A = make_float4(a.v.x, b.v.x, a.v.y, b.v.y); B = A; }
    else { A = make_float4(a.v.x, b.v.x, a.v.y, b.v.y); B = make_float4(a.v.z, b.v.z, a.v.w, b.v.w); }
    float2 t, s, w;
    fm.eval(A, B, t, s, w);
    in0 = t.x < 0.f; in1 = t.y < 0.f;
    if (!(fabsf(t.x) > s.x)) in0 = __float_as_uint(strict_em<EST>(rec, P, i0, n)) >> 31;
    if (!(fabsf(t.y) > s.y)) in1 = __float_as_uint(strict_em<EST>(rec, P, i1, n)) >> 31;
}

// One model's walk: the likelihood-ratio test of sprt.hpp:205-234 from pool position `start`, then - when `count_all` and the model
// was rejected - the rest of the pool (sprt.hpp:243-257).
// Returns true when the walk is complete (`res` final); false when it has to be continued by a warp (`carry` filled, carry.q not set).
template <int EST>
__device__ __forceinline__ bool sprt_walk_one(const float* __restrict__ rec, const float* __restrict__ P, int n, unsigned start,
                                              double eps, double delta, double A, bool count_all, int head, SprtModelResult& res, SprtCarry& carry) {
    FastModel<EST> fm;
    fm.load(rec);
    res = SprtModelResult{0, 0, 0, 0};
    const double r_in = __ddiv_rn(delta, eps), r_out = __ddiv_rn(__dsub_rn(1.0, delta), __dsub_rn(1.0, eps));
    auto wrap = [n](int v) { while (v >= n) v -= n; return v; };      // n may be as small as the minimal sample
    int idx = (int)(start % (unsigned)n);
    double lambda = 1.0;
    int tp = 0, tin = 0;
    bool good = true;
    PoolPoint<EST> pa, pb, na, nb;
    pa.load(P, idx); pb.load(P, wrap(idx + 1));
#pragma unroll 1
    while (tp < n && tp < head) {
        const int i0 = idx, i1 = wrap(idx + 1);
        na.load(P, wrap(idx + 2)); nb.load(P, wrap(idx + 3));          // next pair in flight while this one is evaluated
        bool in0, in1;
        decide_pair<EST>(fm, rec, P, n, pa, pb, i0, i1, in0, in1);
        double ln = __dmul_rn(lambda, in0 ? r_in : r_out);
        tin += in0; tp++;
        if (ln > A) { good = false; break; }
        lambda = ln;
        if (tp >= n) break;
        ln = __dmul_rn(lambda, in1 ? r_in : r_out);
        tin += in1; tp++;
        if (ln > A) { good = false; break; }
        lambda = ln;
        pa = na; pb = nb; idx = wrap(idx + 2);
    }
    res.good = good; res.tested_inl = tin; res.tested_pts = tp; res.full_inl = tin;
    if ((good && tp < n) || (!good && count_all && head < n)) {       // unfinished test, or a long count: a warp takes over
        carry.tp = tp; carry.tin = tin; carry.rejected = good ? 0 : 1; carry.lambda = lambda; carry.start = start; carry.count_all = count_all ? 1 : 0;
        return false;
    }
    if (!good && count_all) {
        // sprt.hpp:243-257: keep counting from the point after the rejecting one
        int pos = (int)(((unsigned long long)start + (unsigned long long)tp) % (unsigned long long)n);
        int rest = n - tp, c = 0;
#pragma unroll 1
        while (rest > 0) {
            const int i0 = pos, i1 = wrap(pos + 1);
            pa.load(P, i0); pb.load(P, i1);
            bool in0, in1;
            decide_pair<EST>(fm, rec, P, n, pa, pb, i0, i1, in0, in1);
            c += in0; rest--;
            if (rest > 0) { c += in1; rest--; }
            pos = wrap(pos + 2);
        }
        res.full_inl = tin + c;
    }
    return true;
}

// The rest of a walk by one warp: 64 pool points per step (lane l decides points base + l and base + 32 + l), then the 64
// multiplications of sprt.hpp:205-234 in pool order, identically in every lane.
template <int EST>
__device__ __forceinline__ SprtModelResult sprt_walk_warp(const float* __restrict__ rec, const float* __restrict__ P, int n, const SprtCarry& cy,
                                                          double eps, double delta, double A) {
    const int lane = threadIdx.x & 31;
    FastModel<EST> fm;
    fm.load(rec);
    const double r_in = __ddiv_rn(delta, eps), r_out = __ddiv_rn(__dsub_rn(1.0, delta), __dsub_rn(1.0, eps));
    const unsigned un = (unsigned)n;
    int tp = cy.tp, tin = cy.tin;
    double lambda = cy.lambda;
    bool good = !cy.rejected;
    auto decide64 = [&](int done) -> unsigned long long {
        const unsigned base = (unsigned)(((unsigned long long)cy.start + (unsigned long long)done) % un);
        const int i0 = (int)((base + (unsigned)lane) % un), i1 = (int)((base + 32u + (unsigned)lane) % un);
        PoolPoint<EST> a, b;
        a.load(P, i0); b.load(P, i1);
        bool in0, in1;
        decide_pair<EST>(fm, rec, P, n, a, b, i0, i1, in0, in1);
        return (unsigned long long)__ballot_sync(0xffffffffu, in0) | ((unsigned long long)__ballot_sync(0xffffffffu, in1) << 32);
    };
    // four 64-point steps at once: the 8 point loads of a lane are in flight together (one warp walks alone: without this every step
    // waits for its own two loads, ~160 dependent global-memory latencies for N = 10 000)
    auto decide256 = [&](int done, unsigned long long* mm) {
        PoolPoint<EST> pa[4], pb[4];
        int i0[4], i1[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const unsigned base = (unsigned)(((unsigned long long)cy.start + (unsigned long long)done + 64ull * b) % un);
            i0[b] = (int)((base + (unsigned)lane) % un); i1[b] = (int)((base + 32u + (unsigned)lane) % un);
            pa[b].load(P, i0[b]); pb[b].load(P, i1[b]);
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            bool in0, in1;
            decide_pair<EST>(fm, rec, P, n, pa[b], pb[b], i0[b], i1[b], in0, in1);
            mm[b] = (unsigned long long)__ballot_sync(0xffffffffu, in0) | ((unsigned long long)__ballot_sync(0xffffffffu, in1) << 32);
        }
    };
#if USAC_SPRT_LOGWALK
    // ---- the decision of a 64-point step WITHOUT the 64 dependent double multiplications (C3: a model that passes the test walks all
    // N points; its N multiplications were 2/3 of the fit). While lambda is a normal double, the sequential product differs from
    // the exact product r_in^i r_out^o lambda_0 by at most k 2^-53 relative after k steps, so log(lambda_k) = S_k +- err with
    // S_k = log(lambda_0) + i_k log(r_in) + o_k log(r_out) kept in 64-bit FIXED POINT (2^-36; integer sums do not accumulate rounding:
    // err <= 0.51 k units incl. the 1-ulp errors of log()). `margin` = 4 N + 1024 units. Per step:
    //   * up = S + (positive parts of the step's terms) bounds every prefix of the step: up < log A - margin => no rejection in the
    //     step, S moves by the step's total (warp-uniform integer work only);
    //   * otherwise the lanes evaluate S_k for the 64 positions: the first k with S_k > log A + margin is THE rejecting point if no
    //     earlier position lies within the margin of log A;
    //   * once a prefix may fall below e^-600 (lambda about to leave the normal range: what a model that passes does after a few
    //     thousand points) S becomes an UPPER BOUND of log(lambda): D' = max(D f (1 + 2^-53), e^-600) dominates the sequential value
    //     (rounding and multiplication by a positive factor are monotone), and per step D_end <= max(S + total, -600 + positive
    //     parts). The bound can only prove "no rejection".
    // Anything else - a position inside the margin, the bound reaching log A, parameters outside the fixed-point range, a NaN - replays
    // the walk with the exact chain below, from the carry. Checked against the sequential chain for every decision / tested_pts /
    // tested_inl: tools/sprt_logwalk_sim.py (CPU model of exactly this logic) and the parity suite.
    if (good) {
        const double Lin = log(r_in), Lout = log(r_out), LA = log(A), L0 = log(lambda);
        const bool in_range = r_in > 0.0 && r_out > 0.0 && A > 0.0 && lambda > 1e-200 && lambda < 1e200 && fabs(Lin) < 64.0 && fabs(Lout) < 64.0 &&
                              fabs(LA) < 64.0 && n <= (1 << 22);
        if (in_range) {
            const double sc = 68719476736.0;                                  // 2^36
            const long long Lin_f = __double2ll_rn(Lin * sc), Lout_f = __double2ll_rn(Lout * sc), LA_f = __double2ll_rn(LA * sc);
            const long long floor_f = __double2ll_rn(-600.0 * sc), margin = 4ll * n + 1024ll;
            const long long LinP = max(Lin_f, 0ll), LinN = min(Lin_f, 0ll), LoutP = max(Lout_f, 0ll), LoutN = min(Lout_f, 0ll);
            long long S = __double2ll_rn(L0 * sc);
            bool bound_mode = false, replay = false, rejected = false;
            int ftp = tp, ftin = tin;
            unsigned long long mm[4];
            int blk = 4;
#pragma unroll 1
            while (ftp < n) {
                if (blk == 4) { decide256(ftp, mm); blk = 0; }                // blocks blk..3 of mm are the steps at ftp, ftp + 64, ...
                const int cnt = min(64, n - ftp);
                unsigned long long m = blk == 0 ? mm[0] : blk == 1 ? mm[1] : blk == 2 ? mm[2] : mm[3];
                blk++;
                if (cnt < 64) m &= (1ull << cnt) - 1ull;
                const long long i = __popcll(m), o = cnt - i;
                const long long up = S + i * LinP + o * LoutP, low = S + i * LinN + o * LoutN;
                if (up < LA_f - margin) {                                     // no prefix of the step can reach log A
                    if (low < floor_f) bound_mode = true;
                    S = bound_mode ? max(S + i * Lin_f + o * Lout_f, floor_f + i * LinP + o * LoutP) : S + i * Lin_f + o * Lout_f;
                    ftp += cnt; ftin += (int)i;
                    continue;
                }
                if (bound_mode || low < floor_f) { replay = true; break; }    // an upper bound cannot place a rejection
                unsigned long long hi = 0ull, band = 0ull;
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int k = lane + 32 * h;
                    const long long ik = __popcll(m & ((2ull << k) - 1ull)), ok = (k + 1) - ik;
                    const long long Sk = S + ik * Lin_f + ok * Lout_f;
                    const bool valid = k < cnt;
                    const unsigned bh = __ballot_sync(0xffffffffu, valid && Sk > LA_f + margin);
                    const unsigned bb = __ballot_sync(0xffffffffu, valid && Sk >= LA_f - margin && Sk <= LA_f + margin);
                    hi |= (unsigned long long)bh << (32 * h); band |= (unsigned long long)bb << (32 * h);
                }
                if (!hi && !band) { S += i * Lin_f + o * Lout_f; ftp += cnt; ftin += (int)i; continue; }
                const int first_hi = hi ? __ffsll((long long)hi) - 1 : 64, first_band = band ? __ffsll((long long)band) - 1 : 64;
                if (first_band < first_hi) { replay = true; break; }          // too close to call in fixed point
                ftin += __popcll(m & ((2ull << first_hi) - 1ull)); ftp += first_hi + 1; rejected = true;
                break;
            }
            if (!replay) { tp = ftp; tin = ftin; good = !rejected; }
            if (!replay && good) tp = n;                                      // (the loop below is skipped: tp == n)
        }
    }
    if (good) {
#else
    if (good) {
#endif
#pragma unroll 1
        while (tp < n) {
            const int cnt = min(64, n - tp);
            const unsigned long long m = decide64(tp);
            int rej_k = -1;
#pragma unroll
            for (int k = 0; k < 64; k++) {
                const double ln = __dmul_rn(lambda, ((m >> k) & 1ull) ? r_in : r_out);
                const bool live = k < cnt && rej_k < 0;
                if (live && ln > A) rej_k = k;
                else if (live) lambda = ln;
            }
            if (rej_k >= 0) {
                tin += __popcll(m & ((2ull << rej_k) - 1ull)); tp += rej_k + 1; good = false;
                break;
            }
            tin += __popcll(cnt == 64 ? m : (m & ((1ull << cnt) - 1ull))); tp += cnt;
        }
    }
    SprtModelResult res;
    res.good = good; res.tested_inl = tin; res.tested_pts = tp; res.full_inl = tin;
    if (!good && cy.count_all) {                                      // sprt.hpp:243-257
        int done = tp, c = 0;
#pragma unroll 1
        while (done < n) {
            unsigned long long mm[4];
            decide256(done, mm);
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int cnt = min(64, n - done);
                if (cnt > 0) { c += __popcll(cnt == 64 ? mm[b] : (mm[b] & ((1ull << cnt) - 1ull))); done += cnt; }
            }
        }
        res.full_inl = tin + c;
    }
    return res;
}

template <int EST>
__global__ void __launch_bounds__(64) sprt_walk_kernel(const RoundArgs a, const float* __restrict__ pool_pts, SprtCarry* __restrict__ carry,
                                                       unsigned* __restrict__ carry_count) {
    const int slot = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.K * a.S) return;
    const int j = q / a.S, i = q % a.S;
    const int pid = a.active[slot];
    const ProblemDesc pd = a.prob[pid];
    const FitState& st = a.state[pid];
    SprtModelResult* out = a.sprt_res + (size_t)slot * a.mstride + q;
    if (i >= a.nmodels[(size_t)slot * a.K + j]) { *out = SprtModelResult{0, 0, 0, 0}; return; }
    const float* rec = a.recs + ((size_t)slot * a.mstride + a.offsets[(size_t)slot * a.K + j] + i) * USAC_REC_STRIDE;
    const float* P = pool_pts + (size_t)pd.aos_off * (EST == USAC_EST_LINE2D ? 2 : 4);
    const unsigned start = (unsigned)(((unsigned long long)st.sprt_cursor + 32ull * (unsigned long long)q) % (unsigned long long)pd.n);
    const bool count_all = (unsigned long long)st.samples_drawn + (unsigned long long)j < (unsigned long long)a.before_sprt;
    SprtModelResult res;
    SprtCarry cy;
    if (sprt_walk_one<EST>(rec, P, pd.n, start, st.sprt_eps, st.sprt_delta, st.sprt_A, count_all, USAC_SPRT_HEAD, res, cy)) { *out = res; return; }
    cy.q = slot * a.mstride + q;
    carry[atomicAdd(carry_count, 1u)] = cy;
}

// the handed-over walks of a round: one warp each (grid-stride over the list)
template <int EST>
__global__ void __launch_bounds__(128) sprt_tail_kernel(const RoundArgs a, const float* __restrict__ pool_pts, const SprtCarry* __restrict__ carry,
                                                        const unsigned* __restrict__ carry_count) {
    const int warps = gridDim.x * (blockDim.x >> 5), lane = threadIdx.x & 31;
    const unsigned count = *carry_count;
    for (unsigned e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < count; e += warps) {
        const SprtCarry cy = carry[e];
        const int slot = cy.q / a.mstride, q = cy.q % a.mstride;
        const int j = q / a.S, i = q % a.S;
        const int pid = a.active[slot];
        const ProblemDesc pd = a.prob[pid];
        const FitState& st = a.state[pid];
        const float* rec = a.recs + ((size_t)slot * a.mstride + a.offsets[(size_t)slot * a.K + j] + i) * USAC_REC_STRIDE;
        const float* P = pool_pts + (size_t)pd.aos_off * (EST == USAC_EST_LINE2D ? 2 : 4);
        const SprtModelResult res = sprt_walk_warp<EST>(rec, P, pd.n, cy, st.sprt_eps, st.sprt_delta, st.sprt_A);
        if (lane == 0) a.sprt_res[cy.q] = res;
    }
}

// SPRT::verifyModelAndGetModelScore for caller-supplied models (usac_gpu_sprt_verify): the same two stages
template <int EST>
__global__ void __launch_bounds__(64) sprt_verify_kernel(const float* __restrict__ recs, int M, const float* __restrict__ P, int n, const unsigned* __restrict__ start,
                                                         const int* __restrict__ count_all, double eps, double delta, double A, SprtModelResult* __restrict__ out,
                                                         SprtCarry* __restrict__ carry, unsigned* __restrict__ carry_count) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= M) return;
    SprtModelResult res;
    SprtCarry cy;
    if (sprt_walk_one<EST>(recs + (size_t)q * USAC_REC_STRIDE, P, n, start[q], eps, delta, A, count_all ? count_all[q] != 0 : false, USAC_SPRT_HEAD, res, cy)) {
        out[q] = res;
        return;
    }
    cy.q = q;
    carry[atomicAdd(carry_count, 1u)] = cy;
}
template <int EST>
__global__ void __launch_bounds__(128) sprt_verify_tail_kernel(const float* __restrict__ recs, const float* __restrict__ P, int n, double eps, double delta, double A,
                                                               SprtModelResult* __restrict__ out, const SprtCarry* __restrict__ carry,
                                                               const unsigned* __restrict__ carry_count) {
    const int warps = gridDim.x * (blockDim.x >> 5), lane = threadIdx.x & 31;
    const unsigned count = *carry_count;
    for (unsigned e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < count; e += warps) {
        const SprtCarry cy = carry[e];
        const SprtModelResult res = sprt_walk_warp<EST>(recs + (size_t)cy.q * USAC_REC_STRIDE, P, n, cy, eps, delta, A);
        if (lane == 0) out[cy.q] = res;
    }
}

// per-model (count, sum) of a fully scored round, in (sample, root) slot order q = j*S + i (count -1 = no such model):
// the input of the host replay used with PROSAC termination
__global__ void model_scores_kernel(const RoundArgs a, int2* __restrict__ out) {
    const int slot = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.K * a.S) return;
    const int j = q / a.S, i = q % a.S;
    int c = -1;
    float s = 0.f;
    if (i < a.nmodels[(size_t)slot * a.K + j]) {
        const int off = a.offsets[(size_t)slot * a.K + j] + i;
        c = 0;
        for (int ch = 0; ch < a.nchunks; ch++) {
            const size_t o = ((size_t)slot * a.nchunks + ch) * a.mstride + off;
            c += a.part_cnt[o];
            s += a.part_sum[o];
        }
    }
    out[(size_t)slot * a.mstride + q] = make_int2(c, __float_as_int(s));
}
