// sprt.cuh - Wald SPRT verification (usac/sprt.hpp) on the device.
#pragma once
#include "pipeline.cuh"

struct SprtModelResult { int good, tested_inl, tested_pts, full_inl; };

__host__ inline void sprt_init_state(FitState& s, int est) {
    // sprt.hpp:114-153 initial (epsilon, delta); A is designed on the host in usac_gpu_fit
    if (est == USAC_EST_HOMOGRAPHY) { s.sprt_delta = 0.01; s.sprt_eps = 0.1; }
    else if (est == USAC_EST_LINE2D) { s.sprt_delta = 0.0001; s.sprt_eps = 0.001; }
    else { s.sprt_delta = 0.05; s.sprt_eps = 0.2; }
    s.sprt_A = 0; s.sprt_cursor = 0; s.sprt_last_update = 0; s.sprt_ntests = 0;
}

static void launch_sprt(int est, const RoundArgs& a, int slots, cudaStream_t stream) {
    (void)est; (void)a; (void)slots; (void)stream;
}
