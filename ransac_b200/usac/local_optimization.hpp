// local_optimization.hpp - LO-RANSAC behind the reference's classes: InnerLocalOptimization (with its IterativeLocalOptimization,
// usac/local_optimization/inner_local_optimization.hpp:16-140, iterative_local_optimization.hpp:15-140) for LocOpt::InItLORsc and
// InItFLORsc. GetModelScore (local_optimization.hpp:19) forwards to usac_gpu_lo_model_score: non-minimal fits, re-scoring at the
// shrinking thresholds and the inlier lists all stay on the device; the same code serves the fused path (usac_fit_cfg::lo).
#pragma once
#include "gpu_plugins.hpp"

class InnerLocalOptimization : public LocalOptimization {
    GpuDevice* dev;
    usac_fit_cfg cfg{};
    uint64_t calls = 0;                  // keys the random inlier subsets (Philox; the reference seeds mt19937 from random_device)
    float lo_threshold = 0;              // IterativeLocalOptimization's running threshold survives calls (iterative_local_optimization.hpp:20-58)
public:
    unsigned int lo_inner_iters = 0, lo_iterative_iters = 0;

    InnerLocalOptimization(Model* model, Estimator* estimator_, Quality* /*quality*/, unsigned int /*points_size*/) {
        GpuEstimator* ge = dynamic_cast<GpuEstimator*>(estimator_);
        if (!ge) throw std::runtime_error("InnerLocalOptimization: needs a GpuEstimator");
        dev = ge->device();
        cfg.threshold = model->threshold;
        cfg.lo = model->lo == InItFLORsc ? 2 : 1;
        cfg.sampler.seed = model->seed;
        cfg.lo_sample_size = model->lo_sample_size;                       // model.hpp:26-29
        cfg.lo_inner_iterations = model->lo_inner_iterations;
        cfg.lo_iterative_iterations = model->lo_iterative_iterations;
        cfg.lo_threshold_multiplier = model->lo_threshold_multiplier;
    }
    void GetModelScore(Model* best_model, Score* best_score) override {
        const cv::Mat d = best_model->returnDescriptor();
        float params[9] = {0};
        for (int k = 0; k < d.rows * d.cols; k++) params[k] = d.ptr()[k];
        dev->check(usac_gpu_lo_model_score(dev->ctx, 0, &cfg, &calls, &lo_threshold, params, &best_score->inlier_number, &best_score->score, &lo_inner_iters, &lo_iterative_iters),
                   "usac_gpu_lo_model_score");
        cv::Mat out(d.rows, d.cols);
        for (int k = 0; k < d.rows * d.cols; k++) out.ptr()[k] = params[k];
        best_model->setDescriptor(out);
    }
};
