// score_queue.cuh - EXPERIMENTAL variant of the scoring kernel for the two-phase evaluators (homography, essential):
// a per-warp SURVIVOR QUEUE instead of the warp-wide slow path. Selected at run time with USAC_GPU_SCORE_QUEUE=1
// (=2: the variant with an out-of-line drain) in launch_score (usac_gpu.cu); the default path is score_kernel<EST> in
// score.cuh and is not touched by this file.
//
// Why: in score_kernel one evaluation out of the 64 of a pair of points x 32 models that the forward test cannot reject
// sends the whole warp through the backward half for that pair (56 % of a launch on hard problems, profiles/README.md),
// although < 0.3 % of the evaluations need it. Here the forward phase is the same, but a surviving (model lane, point) is
// only RECORDED: 32-bit entries in a per-warp shared-memory queue (ballot + prefix popcount). The queue is drained in dense
// batches of 32 survivors - any lane evaluates any of the warp's 32 models from a shared-memory copy of their records, with
// the very same arithmetic (the packed evaluator on a duplicated point), undecided ones through strict_em - and the results
// travel back to the owner lanes in entry order (shuffles), so counts stay exact and sums deterministic.
//
// Status (end of round 1): parity-green - tests/test_gpu_parity.py passes with USAC_GPU_SCORE_QUEUE=1 (58 tests: exact
// counts, fits, stress) - and performance-NEUTRAL so far: C2 bench roofline.frac 0.775 vs 0.781 (value 9.9e11 vs 1.0e12),
// outlier-only rounds 0.86 - 0.92 vs 0.91 - 0.94 (the shared copy of the records and the larger shared-memory footprint
// cost what the dense batches save). Not the default; it needs an ncu pass before it is worth more work. First suspect: code
// size - `drain` is inlined at five call sites, the kernel is 2288 SASS instructions (36 KB, score_kernel: 1100), so the
// instruction cache may be what limits it; make the drain a single out-of-line call site.
#pragma once
#include "score.cuh"

#define USAC_Q_CAP 128           // queue entries per warp (a pair of points adds at most 64)
#define USAC_Q_REC 23            // leading floats of a model record kept in shared memory (odd stride: conflict-free rows)

// The dense batches as ONE out-of-line function (variant 2, USAC_GPU_SCORE_QUEUE=2): the inlined form below is instantiated
// at five call sites. Written after the last GPU run of round 1: compiles, not yet run.
template <int EST>
__device__ __noinline__ void queue_drain_fn(const unsigned* qbuf, int qn, const float4* tp, const float* rec_s, const float* rec_group,
                                            const float* aos, int idx0, int n, unsigned* cnt_io, float* sum_io) {
    const int lane = threadIdx.x & 31;
    unsigned cnt = *cnt_io;
    float sum = *sum_io;
    __syncwarp();
    for (int base = 0; base < qn; base += 32) {
        const int e = base + lane;
        float em = 0.f;
        int owner = -1;
        if (e < qn) {
            const unsigned ent = qbuf[e];
            owner = (int)(ent >> 16);
            const int pj = (int)((ent >> 1) & 0x7fffu), h = (int)(ent & 1u);
            const float4 A = tp[2 * pj], B = tp[2 * pj + 1];
            const float x1 = h ? A.y : A.x, y1 = h ? A.w : A.z, x2 = h ? B.y : B.x, y2 = h ? B.w : B.z;
            FastModel<EST> fq;
            fq.load(rec_s + owner * USAC_Q_REC);
            float2 t, sb, w;
            fq.eval(make_float4(x1, x1, y1, y1), make_float4(x2, x2, y2, y2), t, sb, w);
            em = fminf(t.x, 0.f);
            if (!(fabsf(t.x) > sb.x)) em = strict_em<EST>(rec_group + (size_t)owner * USAC_REC_STRIDE, aos, idx0 + 2 * pj + h, n);
        }
        __syncwarp();
#pragma unroll 4
        for (int l = 0; l < 32; l++) {
            const int o = __shfl_sync(0xffffffffu, owner, l);
            const float v = __shfl_sync(0xffffffffu, em, l);
            if (o == lane) { cnt += __float_as_uint(v) >> 31; sum += v; }
        }
    }
    __syncwarp();
    *cnt_io = cnt;
    *sum_io = sum;
}

template <int EST, bool OUTLINE>
__global__ void __launch_bounds__(USAC_SCORE_THREADS, USAC_SCORE_MIN_CTAS) score_queue_kernel(const ScoreArgs a) {
    static_assert(FastModel<EST>::TWO_PHASE, "the survivor queue needs a forward-only outlier test");
    static_assert(REC_THR < USAC_Q_REC && REC_BAND + 2 < USAC_Q_REC, "FastModel::load must find its fields in the shared copy");
    constexpr int PAIR_FLOATS = 8;
    constexpr int NWARPS = USAC_SCORE_THREADS / 32;
    __shared__ __align__(128) float tile_all[NWARPS][USAC_WARP_STAGES][USAC_TILE_PAIRS * PAIR_FLOATS];
    __shared__ __align__(8) uint64_t full_all[NWARPS][USAC_WARP_STAGES];
    __shared__ float rec_all[NWARPS][32 * USAC_Q_REC];
    __shared__ unsigned q_all[NWARPS][USAC_Q_CAP];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float (*tile)[USAC_TILE_PAIRS * PAIR_FLOATS] = tile_all[warp];
    uint64_t* full = full_all[warp];
    float* rec_s = rec_all[warp];
    unsigned* qbuf = q_all[warp];
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < USAC_WARP_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const int mgroups = a.mblocks * NWARPS;
    const unsigned total = (unsigned)a.slots * (unsigned)a.nchunks * (unsigned)mgroups;
    uint32_t g = 0;

    for (;;) {
        unsigned item = 0;
        if (lane == 0) item = atomicAdd(a.work, 1u) - a.work_base;
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        const int mgroup = (int)(item % (unsigned)mgroups);
        const unsigned rest = item / (unsigned)mgroups;
        const int chunk = (int)(rest % (unsigned)a.nchunks), slot = (int)(rest / (unsigned)a.nchunks);
        const int M = a.mvalid ? a.mvalid[slot] : a.M;
        if (mgroup * 32 >= M) continue;
        const ProblemDesc pd = a.prob[a.active ? a.active[slot] : slot];
        const int m = mgroup * 32 + lane;
        const bool live = m < M;
        const float* rec = a.recs + ((size_t)slot * a.mstride + (live ? m : 0)) * USAC_REC_STRIDE;
        const float* rec_group = a.recs + ((size_t)slot * a.mstride + (size_t)mgroup * 32) * USAC_REC_STRIDE;
        const size_t out = ((size_t)slot * a.nchunks + chunk) * a.mstride + m;

        const int pair_begin = chunk * a.chunk_pairs;
        const int npairs = min(pair_begin + a.chunk_pairs, pd.n_pairs) - pair_begin;
        if (npairs <= 0) {
            if (live) { a.part_cnt[out] = 0; a.part_sum[out] = 0.f; }
            continue;
        }
        const int ntiles = (npairs + USAC_TILE_PAIRS - 1) / USAC_TILE_PAIRS;
        const float* src = a.pairs + ((size_t)pd.pair_off + pair_begin) * PAIR_FLOATS;

        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int k = 0; k < USAC_WARP_STAGES && k < ntiles; k++) {
                const int s = (g + k) % USAC_WARP_STAGES;
                const uint32_t bytes = min(USAC_TILE_PAIRS, npairs - k * USAC_TILE_PAIRS) * PAIR_FLOATS * 4;
                mbar_expect_tx(&full[s], bytes);
                bulk_copy_g2s(tile[s], src + (size_t)k * USAC_TILE_PAIRS * PAIR_FLOATS, bytes, &full[s]);
            }
        }

        FastModel<EST> fm;
        fm.load(rec);
        // shared copy of the warp's records for the dense batches (row = lane, stride USAC_Q_REC)
#pragma unroll
        for (int i = 0; i < USAC_Q_REC; i++) rec_s[lane * USAC_Q_REC + i] = rec[i];
        __syncwarp();
        const float* aos = a.aos + (size_t)pd.aos_off * 4;
        unsigned cnt = 0;
        float sum = 0.f;
        int qn = 0;                                                  // queue fill (uniform)

        for (int k = 0; k < ntiles; k++) {
            const int s = (g + k) % USAC_WARP_STAGES;
            mbar_wait(&full[s], ((g + k) / USAC_WARP_STAGES) & 1);
            const int np = min(USAC_TILE_PAIRS, npairs - k * USAC_TILE_PAIRS);
            const float4* tp = reinterpret_cast<const float4*>(tile[s]);
            const int idx0 = 2 * (pair_begin + k * USAC_TILE_PAIRS);

            // one dense batch per 32 queued survivors; entries = lane << 16 | pair in tile << 1 | half
            auto drain = [&]() {
                if constexpr (OUTLINE) {
                    queue_drain_fn<EST>(qbuf, qn, tp, rec_s, rec_group, aos, idx0, pd.n, &cnt, &sum);
                    qn = 0;
                    return;
                }
                __syncwarp();                                         // queue stores visible to every lane
                for (int base = 0; base < qn; base += 32) {
                    const int e = base + lane;
                    float em = 0.f;
                    int owner = -1;
                    if (e < qn) {
                        const unsigned ent = qbuf[e];
                        owner = (int)(ent >> 16);
                        const int pj = (int)((ent >> 1) & 0x7fffu), h = (int)(ent & 1u);
                        const float4 A = tp[2 * pj], B = tp[2 * pj + 1];
                        const float x1 = h ? A.y : A.x, y1 = h ? A.w : A.z, x2 = h ? B.y : B.x, y2 = h ? B.w : B.z;
                        FastModel<EST> fq;
                        fq.load(rec_s + owner * USAC_Q_REC);
                        float2 t, sb, w;
                        fq.eval(make_float4(x1, x1, y1, y1), make_float4(x2, x2, y2, y2), t, sb, w);   // same arithmetic as the packed path
                        em = fminf(t.x, 0.f);
                        if (!(fabsf(t.x) > sb.x)) em = strict_em<EST>(rec_group + (size_t)owner * USAC_REC_STRIDE, aos, idx0 + 2 * pj + h, pd.n);
                    }
                    __syncwarp();
#pragma unroll 4
                    for (int l = 0; l < 32; l++) {                    // results back to the owners, in entry order
                        const int o = __shfl_sync(0xffffffffu, owner, l);
                        const float v = __shfl_sync(0xffffffffu, em, l);
                        if (o == lane) { cnt += __float_as_uint(v) >> 31; sum += v; }
                    }
                }
                qn = 0;
                __syncwarp();                                         // every lane has read its entries before the queue refills
            };

            int j = 0;
#pragma unroll 1
            for (; j + USAC_PPI <= np; j += USAC_PPI) {
                typename FastModel<EST>::P1 st[USAC_PPI];
                bool lane_any = false;
#pragma unroll
                for (int q = 0; q < USAC_PPI; q++) {
                    const float4 A = tp[2 * (j + q)], B = tp[2 * (j + q) + 1];
                    fm.phase1(A, B, st[q]);
                    bool ox, oy;
                    fm.sure(st[q], ox, oy);
                    lane_any = lane_any || !ox || !oy;
                }
                if (!__any_sync(0xffffffffu, live && lane_any)) continue;
#pragma unroll
                for (int q = 0; q < USAC_PPI; q++) {
                    bool ox, oy;
                    fm.sure(st[q], ox, oy);
                    const bool sx = live && !ox, sy = live && !oy;
                    if (__any_sync(0xffffffffu, sx || sy)) {          // warp-uniform
                        if (qn + 64 > USAC_Q_CAP) drain();
                        const unsigned bx = __ballot_sync(0xffffffffu, sx), by = __ballot_sync(0xffffffffu, sy);
                        const unsigned lt = (1u << lane) - 1u;
                        const unsigned ent = ((unsigned)lane << 16) | ((unsigned)(j + q) << 1);
                        if (sx) qbuf[qn + __popc(bx & lt)] = ent;
                        qn += __popc(bx);
                        if (sy) qbuf[qn + __popc(by & lt)] = ent | 1u;
                        qn += __popc(by);
                    }
                }
            }
            // remainder of the tile (fewer than USAC_PPI pairs): every lane evaluates, as in score_kernel
#pragma unroll 1
            for (; j < np; j++) {
                const float4 A = tp[2 * j], B = tp[2 * j + 1];
                float2 t, sb, w;
                fm.eval(A, B, t, sb, w);
                float2 e1 = make_float2(fminf(t.x, 0.f), fminf(t.y, 0.f));
                if (!(fabsf(t.x) > sb.x)) e1.x = strict_em<EST>(rec, aos, idx0 + 2 * j, pd.n);
                if (!(fabsf(t.y) > sb.y)) e1.y = strict_em<EST>(rec, aos, idx0 + 2 * j + 1, pd.n);
                if (live) {
                    cnt += (__float_as_uint(e1.x) >> 31) + (__float_as_uint(e1.y) >> 31);
                    sum += e1.x + e1.y;
                }
            }
            if (qn) drain();                                         // the survivors' points are still in this stage
            __syncwarp();
            if (lane == 0 && k + USAC_WARP_STAGES < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const int nk = k + USAC_WARP_STAGES;
                const uint32_t bytes = min(USAC_TILE_PAIRS, npairs - nk * USAC_TILE_PAIRS) * PAIR_FLOATS * 4;
                mbar_expect_tx(&full[s], bytes);
                bulk_copy_g2s(tile[s], src + (size_t)nk * USAC_TILE_PAIRS * PAIR_FLOATS, bytes, &full[s]);
            }
        }
        g += (uint32_t)ntiles;
        if (live) {
            a.part_cnt[out] = (int)cnt;
            a.part_sum[out] = FastModel<EST>::finish(sum, (int)cnt, rec[REC_THR]);
        }
        __syncwarp();
    }
}
