// score.cuh - the fused scoring kernel: M models x N points -> (inlier count, error sum) per model.
//
// Replaces Quality::getNumberInliers (quality.hpp:60-101) looping the virtual Estimator::GetError.
//
// Mapping: one thread owns one model (register resident, every parameter duplicated into both halves of a 64-bit
// register pair); a CTA of USAC_SCORE_THREADS models walks a chunk of the point set that is staged through shared memory
// in tiles of USAC_TILE_PAIRS point pairs by 1-D bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier),
// USAC_STAGES deep. Points are stored pair-interleaved ([x1a x1b y1a y1b x2a x2b y2a y2b] per pair) so that one broadcast
// LDS.128 feeds two points straight into the packed FP32x2 pipe (FFMA2/FMUL2/FADD2, sm_100): the residual costs ~28
// packed instructions per pair for a homography, which makes the kernel FMA-pipe bound rather than issue bound.
//
// Exactness: the packed path uses FMA contraction and approximate MUFU rcp/sqrt, so its error value differs from the
// reference's one-rounding-per-operator float32 value in the low bits. Each model record carries a rigorous bound on
// that difference (prepare_kernel, "guard band"); a point whose fast value is closer to the threshold than the bound
// is re-evaluated with strict_error<EST>() - the reference's exact arithmetic - so the inlier COUNT is bit-exact. The
// error SUM is accumulated from the fast values (tolerance 1e-4 relative, BASELINE.json).
#pragma once
#include "strict_math.cuh"

__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float2 dup(float x) { return make_float2(x, x); }

// ---- mbarrier / bulk-copy helpers (PTX ISA 8.x, sm_90+) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- per-estimator fast evaluators --------------------------------------------------------------------------------
// eval(): given one pair of points returns, per lane, `t` (< 0 <=> inlier by the fast path), `bnd` (guard band for
// |t|) and `e` (the value accumulated into the error sum for an inlier, in the units of finish()).

template <int EST> struct FastModel;

template <> struct FastModel<USAC_EST_HOMOGRAPHY> {
    // rows 0,1 negated so that dx = x2 - nx/nz is a single FFMA2
    float2 a11, a12, a13, a21, a22, a23, h31, h32, h33;
    float2 b11, b12, b13, b21, b22, b23, g31, g32, g33;
    float2 negT, c0, c1;
    __device__ __forceinline__ void load(const float* r) {
        a11 = dup(-r[0]); a12 = dup(-r[1]); a13 = dup(-r[2]); a21 = dup(-r[3]); a22 = dup(-r[4]); a23 = dup(-r[5]);
        h31 = dup(r[6]); h32 = dup(r[7]); h33 = dup(r[8]);
        b11 = dup(-r[9]); b12 = dup(-r[10]); b13 = dup(-r[11]); b21 = dup(-r[12]); b22 = dup(-r[13]); b23 = dup(-r[14]);
        g31 = dup(r[15]); g32 = dup(r[16]); g33 = dup(r[17]);
        negT = dup(-2.f * r[REC_THR]); c0 = dup(r[REC_BAND]); c1 = dup(r[REC_BAND + 1]);
    }
    __device__ __forceinline__ void eval(const float4 A, const float4 B, float2& t, float2& bnd, float2& e) const {
        const float2 X1 = make_float2(A.x, A.y), Y1 = make_float2(A.z, A.w), X2 = make_float2(B.x, B.y), Y2 = make_float2(B.z, B.w);
        const float2 nz = __ffma2_rn(h31, X1, __ffma2_rn(h32, Y1, h33));
        const float2 nx = __ffma2_rn(a11, X1, __ffma2_rn(a12, Y1, a13));   // -(h11 x1 + h12 y1 + h13)
        const float2 ny = __ffma2_rn(a21, X1, __ffma2_rn(a22, Y1, a23));
        const float2 mz = __ffma2_rn(g31, X2, __ffma2_rn(g32, Y2, g33));
        const float2 mx = __ffma2_rn(b11, X2, __ffma2_rn(b12, Y2, b13));
        const float2 my = __ffma2_rn(b21, X2, __ffma2_rn(b22, Y2, b23));
        const float2 q = __fmul2_rn(nz, mz);
        const float2 r = make_float2(fast_rcp(q.x), fast_rcp(q.y));        // one reciprocal serves both projections
        const float2 r1 = __fmul2_rn(r, mz), r2 = __fmul2_rn(r, nz);       // 1/nz, 1/mz
        const float2 dx = __ffma2_rn(nx, r1, X2), dy = __ffma2_rn(ny, r1, Y2);
        const float2 ex = __ffma2_rn(mx, r2, X1), ey = __ffma2_rn(my, r2, Y1);
        const float2 sa = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
        const float2 sb = __ffma2_rn(ey, ey, __fmul2_rn(ex, ex));
        const float2 d1 = make_float2(fast_sqrt(sa.x), fast_sqrt(sa.y)), d2 = make_float2(fast_sqrt(sb.x), fast_sqrt(sb.y));
        e = __fadd2_rn(d1, d2);                                             // 2 * error
        t = __fadd2_rn(e, negT);
        bnd = __ffma2_rn(c1, __fmul2_rn(r, r), c0);
    }
    static __device__ __forceinline__ float finish(float sum) { return 0.5f * sum; }
    static __device__ __forceinline__ float strict_to_e(float err) { return 2.f * err; }
};

template <> struct FastModel<USAC_EST_FUNDAMENTAL> {
    float2 f11, f12, f13, f21, f22, f23, f31, f32, f33, negthr, b1, b0;
    __device__ __forceinline__ void load(const float* r) {
        f11 = dup(r[0]); f12 = dup(r[1]); f13 = dup(r[2]); f21 = dup(r[3]); f22 = dup(r[4]); f23 = dup(r[5]);
        f31 = dup(r[6]); f32 = dup(r[7]); f33 = dup(r[8]);
        negthr = dup(-r[REC_THR]); b1 = dup(r[REC_BAND]); b0 = dup(r[REC_BAND + 1]);
    }
    __device__ __forceinline__ void eval(const float4 A, const float4 B, float2& t, float2& bnd, float2& e) const {
        const float2 X1 = make_float2(A.x, A.y), Y1 = make_float2(A.z, A.w), X2 = make_float2(B.x, B.y), Y2 = make_float2(B.z, B.w);
        const float2 a = __ffma2_rn(f11, X1, __ffma2_rn(f12, Y1, f13));
        const float2 b = __ffma2_rn(f21, X1, __ffma2_rn(f22, Y1, f23));
        const float2 c = __ffma2_rn(f11, X2, __ffma2_rn(f21, Y2, f31));
        const float2 d = __ffma2_rn(f12, X2, __ffma2_rn(f22, Y2, f32));
        const float2 n = __ffma2_rn(X2, a, __ffma2_rn(Y2, b, __ffma2_rn(f31, X1, __ffma2_rn(f32, Y1, f33))));
        const float2 n2 = __fmul2_rn(n, n);
        const float2 den = __ffma2_rn(d, d, __ffma2_rn(c, c, __ffma2_rn(b, b, __fmul2_rn(a, a))));
        t = __ffma2_rn(negthr, den, n2);          // n^2 - thr*den  (< 0 <=> n^2/den < thr, no division)
        bnd = __ffma2_rn(b1, den, b0);
        e = make_float2(n2.x * fast_rcp(den.x), n2.y * fast_rcp(den.y));   // only the error sum needs the quotient
    }
    static __device__ __forceinline__ float finish(float sum) { return sum; }
    static __device__ __forceinline__ float strict_to_e(float err) { return err; }
};

template <> struct FastModel<USAC_EST_ESSENTIAL> {
    float2 e11, e12, e13, e21, e22, e23, e31, e32, e33, negT, ka, kb, k0;
    __device__ __forceinline__ void load(const float* r) {
        e11 = dup(r[0]); e12 = dup(r[1]); e13 = dup(r[2]); e21 = dup(r[3]); e22 = dup(r[4]); e23 = dup(r[5]);
        e31 = dup(r[6]); e32 = dup(r[7]); e33 = dup(r[8]);
        negT = dup(-2.f * r[REC_THR]); ka = dup(r[REC_BAND]); kb = dup(r[REC_BAND + 1]); k0 = dup(r[REC_BAND + 2]);
    }
    __device__ __forceinline__ void eval(const float4 A, const float4 B, float2& t, float2& bnd, float2& e) const {
        const float2 X1 = make_float2(A.x, A.y), Y1 = make_float2(A.z, A.w), X2 = make_float2(B.x, B.y), Y2 = make_float2(B.z, B.w);
        const float2 l1 = __ffma2_rn(e11, X2, __ffma2_rn(e21, Y2, e31));
        const float2 l2 = __ffma2_rn(e12, X2, __ffma2_rn(e22, Y2, e32));
        const float2 l3 = __ffma2_rn(e13, X2, __ffma2_rn(e23, Y2, e33));
        const float2 t1 = __ffma2_rn(e11, X1, __ffma2_rn(e12, Y1, e13));
        const float2 t2 = __ffma2_rn(e21, X1, __ffma2_rn(e22, Y1, e23));
        const float2 t3 = __ffma2_rn(e31, X1, __ffma2_rn(e32, Y1, e33));
        const float2 a1 = __ffma2_rn(l1, X1, __ffma2_rn(l2, Y1, l3));
        const float2 b1 = __ffma2_rn(t1, X2, __ffma2_rn(t2, Y2, t3));
        const float2 a2 = __ffma2_rn(l2, l2, __fmul2_rn(l1, l1));
        const float2 b2 = __ffma2_rn(t2, t2, __fmul2_rn(t1, t1));
        const float2 ra = make_float2(fast_rsqrt(a2.x), fast_rsqrt(a2.y)), rb = make_float2(fast_rsqrt(b2.x), fast_rsqrt(b2.y));
        const float2 aa = make_float2(fabsf(a1.x), fabsf(a1.y)), bb = make_float2(fabsf(b1.x), fabsf(b1.y));
        e = __ffma2_rn(aa, ra, __fmul2_rn(bb, rb));     // 2 * error
        t = __fadd2_rn(e, negT);
        bnd = __ffma2_rn(ka, ra, __ffma2_rn(kb, rb, k0));
    }
    static __device__ __forceinline__ float finish(float sum) { return 0.5f * sum; }
    static __device__ __forceinline__ float strict_to_e(float err) { return 2.f * err; }
};

template <> struct FastModel<USAC_EST_LINE2D> {
    float2 a, b, c, negthr, band;
    __device__ __forceinline__ void load(const float* r) {
        a = dup(r[0]); b = dup(r[1]); c = dup(r[2]); negthr = dup(-r[REC_THR]); band = dup(r[REC_BAND]);
    }
    // line pairs are [xa xb ya yb]: one float4 per pair (B unused)
    __device__ __forceinline__ void eval(const float4 A, const float4, float2& t, float2& bnd, float2& e) const {
        const float2 X = make_float2(A.x, A.y), Y = make_float2(A.z, A.w);
        const float2 v = __ffma2_rn(a, X, __ffma2_rn(b, Y, c));
        e = make_float2(fabsf(v.x), fabsf(v.y));
        t = __fadd2_rn(e, negthr);
        bnd = band;
    }
    static __device__ __forceinline__ float finish(float sum) { return sum; }
    static __device__ __forceinline__ float strict_to_e(float err) { return err; }
};

struct ScoreArgs {
    const float* pairs;          // pair-interleaved points of all problems
    const float* aos;            // original AoS points (strict re-evaluation reads these)
    const ProblemDesc* prob;
    const int* active;           // blockIdx.z -> problem id (NULL: identity)
    const float* recs;           // [slot][mstride][USAC_REC_STRIDE]
    const int* mvalid;           // [slot] models to score (NULL: M for all)
    int M, mstride;              // models per problem (upper bound) and record/partial stride
    int chunk_pairs, nchunks;    // point pairs per CTA along y
    int* part_cnt;               // [slot][nchunks][mstride]
    float* part_sum;
};

// Slow path of one lane: the reference's exact arithmetic for point `idx` of the problem.
template <int EST>
__device__ __noinline__ void strict_fix(const float* __restrict__ rec, const float* __restrict__ aos, int idx, int n, float thr,
                                        bool& in, float& e) {
    if (idx >= n) { in = false; e = 0.f; return; }
    float err;
    if (EST == USAC_EST_LINE2D) {
        const float2 p = reinterpret_cast<const float2*>(aos)[idx];
        err = strict_error<EST>(rec, p.x, p.y, 0.f, 0.f);
    } else {
        const float4 p = reinterpret_cast<const float4*>(aos)[idx];
        err = strict_error<EST>(rec, p.x, p.y, p.z, p.w);
    }
    in = err < thr;
    e = FastModel<EST>::strict_to_e(err);
}

template <int EST>
__global__ void __launch_bounds__(USAC_SCORE_THREADS) score_kernel(const ScoreArgs a) {
    constexpr int PAIR_FLOATS = (EST == USAC_EST_LINE2D) ? 4 : 8;
    constexpr int TILE_BYTES = USAC_TILE_PAIRS * PAIR_FLOATS * 4;
    __shared__ __align__(128) float tile[USAC_STAGES][USAC_TILE_PAIRS * PAIR_FLOATS];
    __shared__ __align__(8) uint64_t full[USAC_STAGES];

    const int slot = blockIdx.z;
    const int M = a.mvalid ? a.mvalid[slot] : a.M;
    if ((int)(blockIdx.x * USAC_SCORE_THREADS) >= M) return;        // uniform per CTA
    const ProblemDesc pd = a.prob[a.active ? a.active[slot] : slot];
    const int m = blockIdx.x * USAC_SCORE_THREADS + threadIdx.x;
    const bool live = m < M;
    const float* rec = a.recs + ((size_t)slot * a.mstride + (live ? m : 0)) * USAC_REC_STRIDE;

    const int pair_begin = blockIdx.y * a.chunk_pairs;
    const int pair_end = min(pair_begin + a.chunk_pairs, pd.n_pairs);
    const int npairs = pair_end - pair_begin;
    if (npairs <= 0) {                                              // ragged batch: this problem is shorter than the chunk grid
        if (live) {
            const size_t o = ((size_t)slot * a.nchunks + blockIdx.y) * a.mstride + m;
            a.part_cnt[o] = 0;
            a.part_sum[o] = 0.f;
        }
        return;
    }
    const int ntiles = (npairs + USAC_TILE_PAIRS - 1) / USAC_TILE_PAIRS;
    const float* src = a.pairs + ((size_t)pd.pair_off + pair_begin) * PAIR_FLOATS;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < USAC_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int s = 0; s < USAC_STAGES && s < ntiles; s++) {
            const int np = min(USAC_TILE_PAIRS, npairs - s * USAC_TILE_PAIRS);
            const uint32_t bytes = np * PAIR_FLOATS * 4;
            mbar_expect_tx(&full[s], bytes);
            bulk_copy_g2s(tile[s], src + (size_t)s * USAC_TILE_PAIRS * PAIR_FLOATS, bytes, &full[s]);
        }
    }

    FastModel<EST> fm;
    fm.load(rec);
    const float thr = rec[REC_THR];
    const float* aos = a.aos + (size_t)pd.aos_off * (EST == USAC_EST_LINE2D ? 2 : 4);
    int cnt = 0;
    float2 sum = make_float2(0.f, 0.f);

    for (int tI = 0; tI < ntiles; tI++) {
        const int s = tI % USAC_STAGES;
        const uint32_t parity = (tI / USAC_STAGES) & 1;
        mbar_wait(&full[s], parity);
        const int np = min(USAC_TILE_PAIRS, npairs - tI * USAC_TILE_PAIRS);
        const float4* tp = reinterpret_cast<const float4*>(tile[s]);
#pragma unroll 2
        for (int j = 0; j < np; j++) {
            float4 A, B;
            if (EST == USAC_EST_LINE2D) { A = tp[j]; B = A; }
            else { A = tp[2 * j]; B = tp[2 * j + 1]; }
            float2 t, bnd, e;
            fm.eval(A, B, t, bnd, e);
            bool inx = t.x < 0.f, iny = t.y < 0.f;
            // "not clearly decided" (also catches NaN): re-evaluate with the reference's arithmetic
            if (!(fabsf(t.x) > bnd.x) || !(fabsf(t.y) > bnd.y)) {
                const int idx = 2 * (pair_begin + tI * USAC_TILE_PAIRS + j);
                if (!(fabsf(t.x) > bnd.x)) strict_fix<EST>(rec, aos, idx, pd.n, thr, inx, e.x);
                if (!(fabsf(t.y) > bnd.y)) strict_fix<EST>(rec, aos, idx + 1, pd.n, thr, iny, e.y);
            }
            cnt += (int)inx + (int)iny;
            sum = __fadd2_rn(sum, make_float2(inx ? e.x : 0.f, iny ? e.y : 0.f));
        }
        __syncthreads();                                   // every warp is done with stage s
        if (threadIdx.x == 0 && tI + USAC_STAGES < ntiles) {
            const int nt = tI + USAC_STAGES;
            const int np2 = min(USAC_TILE_PAIRS, npairs - nt * USAC_TILE_PAIRS);
            const uint32_t bytes = np2 * PAIR_FLOATS * 4;
            mbar_expect_tx(&full[s], bytes);
            bulk_copy_g2s(tile[s], src + (size_t)nt * USAC_TILE_PAIRS * PAIR_FLOATS, bytes, &full[s]);
        }
    }
    if (live) {
        const size_t o = ((size_t)slot * a.nchunks + blockIdx.y) * a.mstride + m;
        a.part_cnt[o] = cnt;
        a.part_sum[o] = FastModel<EST>::finish(sum.x + sum.y);
    }
}

// Reference-arithmetic scoring of every point (no fast path): used by usac_gpu_errors and as the in-library
// cross-check of the fast kernel in tests. One thread per point.
template <int EST>
__global__ void errors_kernel(const float* __restrict__ aos, int n, const float* __restrict__ rec, float* __restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (EST == USAC_EST_LINE2D) {
        const float2 p = reinterpret_cast<const float2*>(aos)[i];
        err[i] = strict_error<EST>(rec, p.x, p.y, 0.f, 0.f);
    } else {
        const float4 p = reinterpret_cast<const float4*>(aos)[i];
        err[i] = strict_error<EST>(rec, p.x, p.y, p.z, p.w);
    }
}

// Quality::getInliers (quality.hpp:108-121): ids of the inliers in ascending order. One CTA, ordered block scan.
template <int EST>
__global__ void __launch_bounds__(1024) inliers_kernel(const float* __restrict__ aos, int n, const float* __restrict__ rec, float thr,
                                                       int* __restrict__ ids, int* __restrict__ count) {
    __shared__ int warp_tot[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < n; start += 1024) {
        const int i = start + threadIdx.x;
        bool in = false;
        if (i < n) {
            float e;
            if (EST == USAC_EST_LINE2D) { const float2 p = reinterpret_cast<const float2*>(aos)[i]; e = strict_error<EST>(rec, p.x, p.y, 0.f, 0.f); }
            else { const float4 p = reinterpret_cast<const float4*>(aos)[i]; e = strict_error<EST>(rec, p.x, p.y, p.z, p.w); }
            in = e < thr;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; w++) { const int v = warp_tot[w]; if (w < warp) before += v; total += v; }
        if (in) ids[base + before + __popc(bal & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (threadIdx.x == 0) base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}
