// Pipe-throughput microbenchmark for sm_100a: FFMA, FFMA2 (packed f32x2), MUFU.RCP, MUFU.SQRT, DFMA and mixes.
// Prints thread-ops per cycle per SM for each, measured with clock64 inside one resident wave.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

constexpr int ITERS = 4096;

template<int MODE>
__global__ void __launch_bounds__(256) bench(float* out, long long* cycles, float seed) {
    float a[8], b = seed, c = seed * 0.5f;
    float2 p[8], q[8], r[8]; double d[4];
    #pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x; p[i] = make_float2(a[i], a[i] + 1.f); q[i] = make_float2(0.999f + 1e-6f * a[i], 0.998f); r[i] = make_float2(1e-3f * a[i], 2e-3f); }
    #pragma unroll
    for (int i = 0; i < 4; i++) d[i] = seed + i;
    float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {            // scalar FFMA x8
            #pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
        } else if (MODE == 1) {     // packed FFMA2 x8  (16 fma lane-ops)
            #pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], b2, c2);
        } else if (MODE == 2) {     // MUFU.RCP x8
            #pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if (MODE == 3) {     // MUFU.SQRT x8
            #pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if (MODE == 4) {     // DFMA x4
            #pragma unroll
            for (int i = 0; i < 4; i++) d[i] = fma(d[i], 1.0000001, 1e-9);
        } else if (MODE == 5) {     // mix: 8 FFMA2 + 2 MUFU
            #pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], b2, c2);
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[0]));
            asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[1]));
        } else if (MODE == 6) {     // mix: 8 FFMA + 8 IADD-ish ALU (LOP3)
            #pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
            #pragma unroll
            for (int i = 0; i < 8; i++) { int v = __float_as_int(p[i].x); v = (v ^ (v >> 3)) + it; p[i].x = __int_as_float(v); }
        } else if (MODE == 7) {     // mix: 8 FFMA + 1 MUFU (issue-slot share)
            #pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(p[0].x));
        } else if (MODE == 9) {     // FFMA2 with three distinct vector operands
            #pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], q[i], r[i]);
        } else if (MODE == 10) {    // FFMA2 with two vector operands + one scalar
            #pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], q[i], c2);
        } else if (MODE == 11) {    // FMUL2 two vector operands
            #pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __fmul2_rn(p[i], q[i]);
        } else if (MODE == 8) {     // 8 FFMA2 + 8 scalar FFMA
            #pragma unroll
            for (int i = 0; i < 8; i++) p[i] = __ffma2_rn(p[i], b2, c2);
            #pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(a[i], b, c);
        }
    }
    long long t1 = clock64();
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 8; i++) s += a[i] + p[i].x + p[i].y + q[i].x + r[i].y;
    #pragma unroll
    for (int i = 0; i < 4; i++) s += (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template<int MODE>
int run(const char* name, double ops_per_iter_per_thread, int ctas_per_sm, int sms, float* out, long long* cyc) {
    int grid = sms * ctas_per_sm;
    bench<MODE><<<grid, 256>>>(out, cyc, 1.0f);
    CHK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<grid, 256>>>(out, cyc, 1.0f);
    cudaEventRecord(e1);
    CHK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096]; CHK(cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid; i++) mean += h[i]; mean /= grid;
    double ops_sm = ops_per_iter_per_thread * ITERS * 256.0 * ctas_per_sm;
    printf("%-28s ctas/sm=%d  cycles=%.0f  thread-ops/cycle/SM=%.1f  ms=%.3f  (Gops/s total %.1f)\n",
           name, ctas_per_sm, mean, ops_sm / mean, ms, ops_sm * sms / (ms * 1e6));
    return 0;
}

int main() {
    cudaDeviceProp pr; CHK(cudaGetDeviceProperties(&pr, 0));
    int sms = pr.multiProcessorCount;
    printf("device %s sms=%d clock=%d kHz\n", pr.name, sms, pr.clockRate);
    float* out; long long* cyc;
    CHK(cudaMalloc(&out, sizeof(float) * 4096 * 256)); CHK(cudaMalloc(&cyc, sizeof(long long) * 4096));
    for (int c : {1, 2, 4}) {
        run<0>("FFMA x8", 8, c, sms, out, cyc);
        run<1>("FFMA2 x8 (16 lane-fma)", 16, c, sms, out, cyc);
        run<2>("MUFU.RCP x8", 8, c, sms, out, cyc);
        run<3>("MUFU.SQRT x8", 8, c, sms, out, cyc);
        run<4>("DFMA x4", 4, c, sms, out, cyc);
        run<5>("8 FFMA2 + 2 MUFU (18)", 18, c, sms, out, cyc);
        run<6>("8 FFMA + ~24 ALU (8 fma)", 8, c, sms, out, cyc);
        run<7>("8 FFMA + 1 MUFU (9)", 9, c, sms, out, cyc);
        run<8>("8 FFMA2 + 8 FFMA (24)", 24, c, sms, out, cyc);
        run<9>("FFMA2 3 vector operands", 16, c, sms, out, cyc);
        run<10>("FFMA2 2 vector + scalar", 16, c, sms, out, cyc);
        run<11>("FMUL2 2 vector", 16, c, sms, out, cyc);
    }
    return 0;
}
