// knn.cuh - exact k nearest neighbours on the device.
//
// Replaces NearestNeighbors::getNearestNeighbors_nanoflann (usac/utils/nearest_neighbors.cpp:69-128): for every point the
// k+1 closest points by squared L2 distance over all columns, ascending, minus the first one (the query itself). The
// reference walks a KD-tree (nanoflann, not vendored); here the points are binned into a uniform G x G grid over the first
// two columns and every query searches Chebyshev rings of cells around its own cell until the (k+1)-th best distance is
// provably smaller than anything outside the rings visited (a point r+1 rings away is at least r cell widths away in one of
// the first two coordinates, hence in the full metric). Exact, and deterministic: candidates are ordered by the 64-bit key
// (distance bits, point index), i.e. equidistant neighbours come in ascending index order (the order a brute-force search by that key gives).
//
// Distance arithmetic: float32, squared differences accumulated in column order, one rounding per operator (no FMA).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace knn {

struct GridDesc {
    float x0, y0;        // lower corner of the bounding box of the finite points (columns 0, 1)
    float inv_cell;      // cells per unit length
    float cell;          // cell width
};

__device__ __forceinline__ int ord_enc(float f) {            // monotone float -> int map (for atomicMin/atomicMax)
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord_dec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// bbox[0..3] = enc(min x), enc(min y), enc(max x), enc(max y) over finite coordinates; initialise to INT_MAX, INT_MAX, INT_MIN, INT_MIN
__global__ void bbox_kernel(const float* __restrict__ pts, int n, int dim, int* __restrict__ bbox) {
    int lo_x = 0x7fffffff, lo_y = 0x7fffffff, hi_x = (int)0x80000000, hi_y = (int)0x80000000;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = pts[(size_t)i * dim], y = pts[(size_t)i * dim + 1];
        if (isfinite(x)) { lo_x = min(lo_x, ord_enc(x)); hi_x = max(hi_x, ord_enc(x)); }
        if (isfinite(y)) { lo_y = min(lo_y, ord_enc(y)); hi_y = max(hi_y, ord_enc(y)); }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lo_x = min(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); lo_y = min(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o));
        hi_x = max(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o)); hi_y = max(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bbox[0], lo_x); atomicMin(&bbox[1], lo_y); atomicMax(&bbox[2], hi_x); atomicMax(&bbox[3], hi_y);
    }
}

__global__ void bbox_init_kernel(int* bbox) {
    bbox[0] = bbox[1] = 0x7fffffff;
    bbox[2] = bbox[3] = (int)0x80000000;
}

__global__ void grid_desc_kernel(const int* __restrict__ bbox, int G, GridDesc* __restrict__ out) {
    GridDesc g;
    const bool any = bbox[0] <= bbox[2] && bbox[1] <= bbox[3];
    const float x0 = any ? ord_dec(bbox[0]) : 0.f, y0 = any ? ord_dec(bbox[1]) : 0.f;
    const float x1 = any ? ord_dec(bbox[2]) : 0.f, y1 = any ? ord_dec(bbox[3]) : 0.f;
    float ext = fmaxf(x1 - x0, y1 - y0);
    if (!(ext > 0.f) || !isfinite(ext)) ext = 1.f;           // all points in one spot (or an overflowing range): one row of cells suffices
    g.x0 = x0; g.y0 = y0;
    g.cell = ext / (float)G;
    g.inv_cell = (float)G / ext;
    if (!(g.cell > 0.f) || !isfinite(g.inv_cell)) { g.cell = 0.f; g.inv_cell = 0.f; }   // degenerate: everything in cell 0, lower bounds 0
    *out = g;
}

__device__ __forceinline__ int cell_coord(float v, float v0, float inv, int G) {
    const float t = (v - v0) * inv;
    int c = (t >= 0.f) ? (t < (float)G ? (int)t : G - 1) : 0;   // NaN -> 0
    return c;
}

__global__ void cell_keys_kernel(const float* __restrict__ pts, int n, int dim, const GridDesc* __restrict__ gd, int G,
                                 unsigned* __restrict__ keys, int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GridDesc g = *gd;
    const int cx = cell_coord(pts[(size_t)i * dim], g.x0, g.inv_cell, G), cy = cell_coord(pts[(size_t)i * dim + 1], g.y0, g.inv_cell, G);
    keys[i] = (unsigned)(cy * G + cx);
    idx[i] = i;
}

// cell_start[c] = first sorted position whose key >= c, for c in [0, ncells]
__global__ void cell_start_kernel(const unsigned* __restrict__ skeys, int n, int ncells, int* __restrict__ cell_start) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncells) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (skeys[mid] < (unsigned)c) lo = mid + 1; else hi = mid;
    }
    cell_start[c] = lo;
}

__global__ void gather_kernel(const float* __restrict__ pts, int n, int dim, const int* __restrict__ sidx, float4* __restrict__ spts) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float* p = pts + (size_t)sidx[s] * dim;
    spts[s] = dim == 4 ? make_float4(p[0], p[1], p[2], p[3]) : make_float4(p[0], p[1], 0.f, 0.f);
}

template <int DIM>
__device__ __forceinline__ float dist2(const float4 a, const float4 b) {
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
    float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (DIM == 4) {
        const float dz = __fsub_rn(a.z, b.z), dw = __fsub_rn(a.w, b.w);
        d = __fadd_rn(d, __fmul_rn(dz, dz));
        d = __fadd_rn(d, __fmul_rn(dw, dw));
    }
    return d;
}

// candidates at sorted positions [a, b): insertion into the sorted list by a compare-exchange chain (static indices only,
// so that the list stays in registers). The k+1 live entries occupy the TOP slots of the list - the slots below them hold
// the key 0, which no candidate displaces - so the (k+1)-th best is always best[CAP-1].
template <int DIM, int CAP>
__device__ __forceinline__ void scan_range(const float4* __restrict__ spts, const int* __restrict__ sidx, const float4 Q,
                                           unsigned long long (&best)[CAP], unsigned long long& worst, const int a, const int b) {
    for (int t = a; t < b; t++) {
        const float d = dist2<DIM>(Q, spts[t]);
        const unsigned db = d != d ? 0x7fc00000u : __float_as_uint(d) & 0x7fffffffu;      // NaN: one pattern, sorts last
        unsigned long long key = ((unsigned long long)db << 32) | (unsigned)sidx[t];
        if (key < worst) {
#pragma unroll
            for (int j = 0; j < CAP; j++) {
                const unsigned long long cur = best[j];
                const bool sw = key < cur;
                best[j] = sw ? key : cur;
                key = sw ? cur : key;
            }
            worst = best[CAP - 1];
        }
    }
}

// One thread per query, queries in cell order (neighbouring threads search the same cells). CAP >= k + 1 is the capacity
// of the register-resident sorted candidate list.
template <int DIM, int CAP>
__global__ void __launch_bounds__(128) query_kernel(const float4* __restrict__ spts, const int* __restrict__ sidx, const unsigned* __restrict__ skeys,
                                                    const int* __restrict__ cell_start, const GridDesc* __restrict__ gd, int G, int n, int k,
                                                    int* __restrict__ table) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const GridDesc g = *gd;
    const float4 Q = spts[s];
    const int q = sidx[s];
    const int qc = (int)skeys[s], qx = qc % G, qy = qc / G;
    unsigned long long best[CAP];
    const int first = CAP - 1 - k;                           // live slots: [first, CAP-1]
#pragma unroll
    for (int j = 0; j < CAP; j++) best[j] = j < first ? 0ull : ~0ull;
    unsigned long long worst = ~0ull;                       // best[CAP-1]: the (k+1)-th candidate so far

    const int rmax = max(max(qx, G - 1 - qx), max(qy, G - 1 - qy));
    for (int r = 0; r <= rmax; r++) {
        const int xa = max(qx - r, 0), xb = min(qx + r, G - 1);
        for (int cy = qy - r; cy <= qy + r; cy++) {
            if (cy < 0 || cy >= G) continue;
            if (cy == qy - r || cy == qy + r) {               // a whole row of the ring: contiguous cells
                scan_range<DIM, CAP>(spts, sidx, Q, best, worst, cell_start[cy * G + xa], cell_start[cy * G + xb + 1]);
            } else {
                if (qx - r >= 0) scan_range<DIM, CAP>(spts, sidx, Q, best, worst, cell_start[cy * G + qx - r], cell_start[cy * G + qx - r + 1]);
                if (qx + r < G) scan_range<DIM, CAP>(spts, sidx, Q, best, worst, cell_start[cy * G + qx + r], cell_start[cy * G + qx + r + 1]);
            }
        }
        // everything not visited yet lies >= r cell widths away in x or y; the slack covers the rounding of the cell
        // coordinates (|error| <= a few G * 2^-24 cells, G <= 1024) and of the float32 distances
        const float lb = fmaxf((float)r - 0.01f, 0.f) * g.cell;
        const float lb2 = lb * lb * 0.999f;
        if (worst != ~0ull && __uint_as_float((unsigned)(worst >> 32)) < lb2) break;
    }
    // neighbour j (1..k; 0 is the query itself) sits in slot first + j
#pragma unroll
    for (int j = 1; j < CAP; j++) {
        const int out = j - first - 1;                       // rank among the k neighbours
        if (out >= 0) table[(size_t)q * k + out] = (int)(unsigned)best[j];
    }
}

}  // namespace knn
