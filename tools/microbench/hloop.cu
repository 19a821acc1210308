// Ablation microbenchmark of the homography scoring loop body (score.cuh FastModel<HOMOGRAPHY>::eval + accumulation):
// a register-resident model per thread, a 128-pair tile in shared memory re-read ITERS times (no TMA, no strict path).
// Variants (VAR): 0 full body; 1 no guard band; 2 no MUFU (rcp/sqrt replaced by FMUL); 3 no band + no MUFU; 4 projections only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../ransac_b200/csrc -o hloop hloop.cu
#include <cstdio>
#include "score.cuh"

template <int VAR>
__global__ void __launch_bounds__(128, 5) body(const float* __restrict__ recs, const float* __restrict__ pairs, float* out, int iters) {
    __shared__ __align__(16) float tile[128 * 8];
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) tile[i] = pairs[i];
    __syncthreads();
    const float* r = recs + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 32;
    FastModel<USAC_EST_HOMOGRAPHY> fm;
    fm.load(r);
    const float4* tp = reinterpret_cast<const float4*>(tile);
    unsigned cnt = 0;
    float2 sum[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    for (int it = 0; it < iters; it++) {
#pragma unroll 1
        for (int j = 0; j < 128; j += 4) {
            bool unsure = false;
            float2 em[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 A = tp[2 * (j + q)], B = tp[2 * (j + q) + 1];
                const float2 X1 = make_float2(A.x, A.y), Y1 = make_float2(A.z, A.w), X2 = make_float2(B.x, B.y), Y2 = make_float2(B.z, B.w);
                const float2 nz = __ffma2_rn(dup(fm.h31), X1, __ffma2_rn(dup(fm.h32), Y1, dup(fm.h33)));
                const float2 nx = __ffma2_rn(dup(fm.a11), X1, __ffma2_rn(dup(fm.a12), Y1, dup(fm.a13)));
                const float2 ny = __ffma2_rn(dup(fm.a21), X1, __ffma2_rn(dup(fm.a22), Y1, dup(fm.a23)));
                const float2 mz = __ffma2_rn(dup(fm.g31), X2, __ffma2_rn(dup(fm.g32), Y2, dup(fm.g33)));
                const float2 mx = __ffma2_rn(dup(fm.b11), X2, __ffma2_rn(dup(fm.b12), Y2, dup(fm.b13)));
                const float2 my = __ffma2_rn(dup(fm.b21), X2, __ffma2_rn(dup(fm.b22), Y2, dup(fm.b23)));
                float2 t, s = make_float2(0.f, 0.f);
                if (VAR == 4) {
                    t = __fadd2_rn(__fadd2_rn(__fadd2_rn(nz, nx), __fadd2_rn(ny, mz)), __fadd2_rn(mx, my));
                } else {
                    const float2 q2 = __fmul2_rn(nz, mz);
                    const float2 rr = (VAR == 2 || VAR == 3) ? __fmul2_rn(q2, q2) : make_float2(fast_rcp(q2.x), fast_rcp(q2.y));
                    const float2 r1 = __fmul2_rn(rr, mz), r2 = __fmul2_rn(rr, nz);
                    const float2 dx = __ffma2_rn(nx, r1, X2), dy = __ffma2_rn(ny, r1, Y2);
                    const float2 ex = __ffma2_rn(mx, r2, X1), ey = __ffma2_rn(my, r2, Y1);
                    const float2 sa = __ffma2_rn(dy, dy, __fmul2_rn(dx, dx));
                    const float2 sb = __ffma2_rn(ey, ey, __fmul2_rn(ex, ex));
                    const float2 d1 = (VAR == 2 || VAR == 3) ? __fmul2_rn(sa, sa) : make_float2(fast_sqrt(sa.x), fast_sqrt(sa.y));
                    const float2 d2 = (VAR == 2 || VAR == 3) ? __fmul2_rn(sb, sb) : make_float2(fast_sqrt(sb.x), fast_sqrt(sb.y));
                    t = __fadd2_rn(__fadd2_rn(d1, dup(fm.negT)), d2);
                    if (VAR == 0 || VAR == 2) {
                        const float2 p1 = __fmul2_rn(dup(fm.k1), r1), p2 = __fmul2_rn(dup(fm.k2), r2);
                        s = make_float2(fabsf(p1.x) + fabsf(p2.x), fabsf(p1.y) + fabsf(p2.y));
                    }
                    if (VAR == 5) {   // band as max(|2 k1 r1|, |2 k2 r2|): FMNMX (ALU pipe) instead of FADD (FMA pipe)
                        const float2 p1 = __fmul2_rn(dup(fm.k1), r1), p2 = __fmul2_rn(dup(fm.k2), r2);
                        s = make_float2(fmaxf(fabsf(p1.x), fabsf(p2.x)), fmaxf(fabsf(p1.y), fabsf(p2.y)));
                    }
                    if (VAR == 6) {   // four compares, no add
                        const float2 p1 = __fmul2_rn(dup(fm.k1), r1), p2 = __fmul2_rn(dup(fm.k2), r2);
                        unsure = unsure || !(fabsf(t.x) > fabsf(p1.x)) || !(fabsf(t.x) > fabsf(p2.x)) || !(fabsf(t.y) > fabsf(p1.y)) || !(fabsf(t.y) > fabsf(p2.y));
                    }
                    if (VAR == 7) {   // one product: band = |k r|, r = 1/(nz mz)
                        const float2 p1 = __fmul2_rn(dup(fm.k1), rr);
                        unsure = unsure || !(fabsf(t.x) > fabsf(p1.x)) || !(fabsf(t.y) > fabsf(p1.y));
                    }
                }
                em[q] = make_float2(fminf(t.x, 0.f), fminf(t.y, 0.f));
                if (VAR == 0 || VAR == 2 || VAR == 5) unsure = unsure || !(fabsf(t.x) > s.x) || !(fabsf(t.y) > s.y);
            }
            if (__any_sync(0xffffffffu, unsure)) { cnt += 1000000; }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                cnt += (__float_as_uint(em[q].x) >> 31) + (__float_as_uint(em[q].y) >> 31);
                sum[q] = __fadd2_rn(sum[q], em[q]);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)cnt + sum[0].x + sum[0].y + sum[1].x + sum[1].y + sum[2].x + sum[2].y + sum[3].x + sum[3].y;
}

template <int VAR>
void run(const char* name, const float* recs, const float* pairs, float* out, int sms) {
    const int grid = sms * 5, iters = 200;
    body<VAR><<<grid, 128>>>(recs, pairs, out, iters);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    body<VAR><<<grid, 128>>>(recs, pairs, out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double evals = (double)grid * 128 * iters * 256.0;
    const double cyc_per_pair_warp = (ms * 1e-3 * 1.965e9) / ((double)iters * 128.0 * 5 /*ctas*/ * 4 /*warps*/ / 4 /*smsp*/);
    printf("%-34s %.3f ms  %.3f T evals/s  frac(42 flop) %.3f  cycles per warp-pair per SMSP %.1f\n", name, ms, evals / (ms * 1e-3) / 1e12,
           42 * evals / (ms * 1e-3) / 1e12 / (2 * 128 * sms * 1.965e-3), cyc_per_pair_warp);
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount, nmod = sms * 5 * 128;
    float *recs, *pairs, *out;
    cudaMalloc(&recs, sizeof(float) * 32 * nmod); cudaMalloc(&pairs, sizeof(float) * 8 * 128); cudaMalloc(&out, sizeof(float) * nmod);
    float* h = new float[32 * (size_t)nmod];
    for (int m = 0; m < nmod; m++) {
        float* r = h + 32 * (size_t)m;
        for (int i = 0; i < 32; i++) r[i] = 0.f;
        const float H[9] = {1.01f + 1e-5f * m, 0.02f, 3.f, -0.01f, 0.98f, -2.f, 1e-5f, -2e-5f, 1.f};
        for (int i = 0; i < 9; i++) { r[i] = H[i]; r[9 + i] = H[i]; }
        r[11] = -3.f; r[14] = 2.f; r[18] = 1e-4f; r[19] = 1e-4f; r[22] = 2.f;
    }
    cudaMemcpy(recs, h, sizeof(float) * 32 * nmod, cudaMemcpyHostToDevice);
    float hp[8 * 128];
    for (int i = 0; i < 8 * 128; i++) hp[i] = 100.f + (float)((i * 7919) % 800);
    cudaMemcpy(pairs, hp, sizeof(hp), cudaMemcpyHostToDevice);
    run<0>("full body", recs, pairs, out, sms);
    run<1>("no guard band", recs, pairs, out, sms);
    run<2>("no MUFU (FMUL instead)", recs, pairs, out, sms);
    run<3>("no band, no MUFU", recs, pairs, out, sms);
    run<4>("projections + accumulate only", recs, pairs, out, sms);
    run<5>("band via FMNMX", recs, pairs, out, sms);
    run<6>("band via 4 compares", recs, pairs, out, sms);
    run<7>("band = |k r| (one product)", recs, pairs, out, sms);
    return 0;
}
