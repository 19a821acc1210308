// ransac_output.hpp - result record, same accessors as usac/ransac/ransac_output.hpp:10-108 (deep copies, :43-44).
#pragma once
#include <vector>

#include "model.hpp"

class RansacOutput {
    Model* model;
    std::vector<int> inliers;
    long time_mcs;
    unsigned int number_inliers, number_iterations, lo_inner_iters, lo_iterative_iters, gc_iters;
public:
    RansacOutput(const Model* const model_, const int* const inliers_, long time_mcs_, unsigned int number_inliers_,
                 unsigned int number_iterations_, unsigned int lo_inner_iters_, unsigned int lo_iterative_iters_, unsigned int gc_iters_)
        : model(new Model(model_)), inliers(inliers_, inliers_ + number_inliers_), time_mcs(time_mcs_), number_inliers(number_inliers_),
          number_iterations(number_iterations_), lo_inner_iters(lo_inner_iters_), lo_iterative_iters(lo_iterative_iters_), gc_iters(gc_iters_) {
        model->setDescriptor(model_->returnDescriptor());
    }
    ~RansacOutput() { delete model; }
    RansacOutput(const RansacOutput&) = delete;
    std::vector<int> getInliers() { return inliers; }
    long getTimeMicroSeconds() { return time_mcs; }
    unsigned int getNumberOfInliers() { return number_inliers; }
    unsigned int getNumberOfMainIterations() { return number_iterations; }
    unsigned int getLOIters() { return lo_inner_iters + lo_iterative_iters + gc_iters; }
    unsigned int getLOInnerIters() { return lo_inner_iters; }
    unsigned int getLOIterativeIters() { return lo_iterative_iters; }
    unsigned int getGCIters() { return gc_iters; }
    Model* getModel() { return model; }
};
