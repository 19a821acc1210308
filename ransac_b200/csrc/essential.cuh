// essential.cuh - five-point essential-matrix solver (EssentialEstimator::EstimateModel, essential_estimator.hpp:52-62).
#pragma once
#include "strict_math.cuh"

__device__ int solve_essential5(const float* __restrict__ pts, const int* s, float* out) {
    (void)pts; (void)s; (void)out;
    return 0;   // built in a later milestone (essential row)
}
