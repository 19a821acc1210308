// prosac_termination_criteria.hpp - PROSAC's non-randomness / maximality stopping rule behind the reference's class
// (usac/termination_criteria/prosac_termination_criteria.hpp:10-203): same constructor and getUpBoundIterations(hypCount, model).
// The inlier mask of the model over the quality-sorted points comes from the device (usac_gpu_get_inliers); the table updates are
// host arithmetic shared with the fused path (csrc/host_replay.hpp, ProsacTermHost).
#pragma once
#include "../csrc/host_replay.hpp"
#include "gpu_plugins.hpp"

class ProsacTerminationCriteria : public TerminationCriteria {
    GpuDevice* dev;
    ProsacTermHost host;
    StandardTerminationCriteria standart_termination_criteria;
    unsigned int* largest_sample_size = nullptr;
    unsigned int points_size;
    float threshold;
    std::vector<int> ids;
    std::vector<unsigned char> mask;

public:
    ProsacTerminationCriteria(unsigned int* growth_function_, const Model* const model, unsigned int points_size_, Estimator* estimator_)
        : standart_termination_criteria(model, points_size_), points_size(points_size_), threshold(model->threshold), ids(points_size_), mask(points_size_) {
        GpuEstimator* ge = dynamic_cast<GpuEstimator*>(estimator_);
        if (!ge) throw std::runtime_error("ProsacTerminationCriteria: needs a GpuEstimator");
        dev = ge->device();
        std::vector<unsigned> growth(growth_function_, growth_function_ + points_size_);
        host.init(growth, points_size_, model->sample_size, model->desired_prob, model->max_iterations);
        isinit = true;
    }
    void setLargestSampleSize(unsigned int* p) { largest_sample_size = p; }
    unsigned int* getStoppingLength() { return &host.termination_length; }
    unsigned int getUpBoundIterations(unsigned int inlier_size) override { return standart_termination_criteria.getUpBoundIterations(inlier_size); }
    unsigned int getUpBoundIterations(unsigned int inlier_size, unsigned int n) override { return standart_termination_criteria.getUpBoundIterations(inlier_size, n); }
    // prosac_termination_criteria.hpp:148-201
    unsigned int getUpBoundIterations(unsigned int hypCount, const cv::Mat& model) {
        int n_in = 0;
        dev->check(usac_gpu_get_inliers(dev->ctx, 0, model.ptr(), threshold, ids.data(), &n_in), "usac_gpu_get_inliers");
        std::fill(mask.begin(), mask.end(), (unsigned char)0);
        for (int i = 0; i < n_in; i++) mask[(size_t)ids[i]] = 1;
        return host.update(hypCount, mask, largest_sample_size ? *largest_sample_size : points_size);
    }
};
