// ransac.hpp - the driver, same construction and accessors as usac/ransac/ransac.hpp:17-118; run() has two forms:
//   run()            the GPU hypothesis batch: the whole loop of ransac.cpp:58-139 in usac_gpu_fit (rounds of K samples, one
//                    host sync per round), the refit loop of ransac.cpp:157-207 in usac_gpu_refit, then the final inlier list;
//   run_sequential() the reference's one-hypothesis-at-a-time loop (ransac.cpp:58-139) over the virtual plugin interfaces,
//                    each call forwarding to the C ABI - the same results, used to show the drop-in at plugin granularity.
#pragma once
#include <chrono>

#include "gpu_plugins.hpp"
#include "ransac_output.hpp"

class Ransac {
protected:
    Model* model;
    GpuDevice* device = nullptr;
    Quality* quality = nullptr;
    Sampler* sampler = nullptr;
    TerminationCriteria* termination_criteria = nullptr;
    RansacOutput* ransac_output = nullptr;
    Estimator* estimator = nullptr;
    unsigned int points_size;
    usac_fit_result last_fit{};
public:
    Ransac(Model* model_, cv::InputArray points_, int gpu = 0) : model(model_) {
        assert(model != nullptr);
        const cv::Mat& pts = points_.getMat();
        points_size = (unsigned int)pts.rows;
        device = new GpuDevice(gpu, model->estimator, pts);                           // initEstimator, init.cpp:3-21
        estimator = new GpuEstimator(device);
        if (model->sampler == Napsac) {                                               // ransac.hpp:61-78
            if (model->neighborsType == Grid) device->check(usac_gpu_set_neighbors_grid(device->ctx, 0, model->cell_size), "usac_gpu_set_neighbors_grid");
            else device->check(usac_gpu_build_neighbors_knn(device->ctx, 0, (int)model->k_nearest_neighbors),      // getNearestNeighbors_nanoflann
                               "usac_gpu_build_neighbors_knn");
        }
        sampler = new GpuSampler(device, model);                                       // initSampler, init.cpp:23-50
        quality = new GpuQuality(device);
        quality->init(points_size, model->threshold, estimator);
        termination_criteria = new StandardTerminationCriteria(model, points_size);   // initTerminationCriteria, init.cpp:52-55
    }
    ~Ransac() { delete sampler; delete quality; delete estimator; delete termination_criteria; delete ransac_output; delete device; }
    Ransac(const Ransac&) = delete;

    void setSampler(Sampler* s) { sampler = s; }
    void setModel(Model* m) { model = m; }
    void setTerminationCriteria(TerminationCriteria* t) { termination_criteria = t; }
    void setQuality(Quality* q) { quality = q; }
    RansacOutput* getRansacOutput() { return ransac_output; }
    const usac_fit_result& lastFit() const { return last_fit; }

    void run() {
        auto t0 = std::chrono::steady_clock::now();
        usac_fit_cfg cfg{};
        cfg.sampler = static_cast<GpuSampler*>(sampler)->config();
        cfg.threshold = model->threshold; cfg.confidence = model->desired_prob; cfg.max_iterations = model->max_iterations;
        cfg.sprt = model->sprt; cfg.round_size = model->gpu_round_size; cfg.rank = 0; cfg.nranks = 1;
        if (model->lo == GC || model->lo == IRLS) throw std::runtime_error("Ransac: graph-cut / IRLS local optimisation is not part of the GPU layer");
        cfg.lo = (int)model->lo;                                                         // NullLO 0, InItLORsc 1, InItFLORsc 2 (model.hpp:13)
        device->check(usac_gpu_fit(device->ctx, &cfg, &last_fit), "usac_gpu_fit");
        if (last_fit.inliers <= 0) throw std::runtime_error("Ransac: best score is 0");           // ransac.cpp:143-147
        usac_refit_result rf{};                                                                   // ransac.cpp:157-207 on the device
        device->check(usac_gpu_refit(device->ctx, 0, last_fit.model, last_fit.inliers, model->threshold, &rf), "usac_gpu_refit");
        finish(rf.model, rf.inliers, last_fit.iterations, t0, last_fit.lo_inner_iters, last_fit.lo_iterative_iters);
    }

    void run_sequential() {
        auto t0 = std::chrono::steady_clock::now();
        Score best, cur;
        std::vector<Model*> models;
        const int nmod = model->estimator == Fundamental ? 3 : 1;
        for (int i = 0; i < nmod; i++) models.push_back(new Model(model));
        Model best_model(model);
        std::vector<int> sample((size_t)estimator->SampleNumber());
        unsigned int iters = 0, max_iters = model->max_iterations;
        while (iters < max_iters) {
            sampler->generateSample(sample.data());
            const unsigned int n = estimator->EstimateModel(sample.data(), models);
            for (unsigned int i = 0; i < n; i++) {
                quality->getNumberInliers(&cur, models[i]->returnDescriptor());
                if (cur.bigger(&best)) {
                    best.copyFrom(&cur);
                    best_model.setDescriptor(models[i]->returnDescriptor());
                    max_iters = termination_criteria->getUpBoundIterations((unsigned int)best.inlier_number);
                }
            }
            iters++;
        }
        for (Model* m : models) delete m;
        if (best.inlier_number == 0) throw std::runtime_error("Ransac: best score is 0");   // the reference exit(111)s, ransac.cpp:143-147
        last_fit = usac_fit_result{};
        last_fit.inliers = best.inlier_number; last_fit.score = best.score; last_fit.iterations = iters;
        const cv::Mat d = best_model.returnDescriptor();
        for (int k = 0; k < d.rows * d.cols; k++) last_fit.model[k] = d.ptr()[k];
        // the refit loop of ransac.cpp:157-207 over the virtual plugin calls
        std::vector<int> max_inliers(points_size);
        quality->getInliers(best_model.returnDescriptor(), max_inliers.data());
        Model non_minimal(model);
        unsigned int previous = 0;
        for (unsigned int norm = 0; norm < 4; norm++) {
            if (!estimator->EstimateModelNonMinimalSample(max_inliers.data(), (unsigned int)best.inlier_number, non_minimal)) break;
            quality->getNumberInliers(&cur, non_minimal.returnDescriptor(), model->threshold, true, max_inliers.data());
            if ((float)cur.inlier_number / best.inlier_number < 0.8) break;
            if ((unsigned int)cur.inlier_number <= previous) break;
            previous = (unsigned int)cur.inlier_number;
            best.copyFrom(&cur);
            best_model.setDescriptor(non_minimal.returnDescriptor());
        }
        const cv::Mat fd = best_model.returnDescriptor();
        float params[9] = {0};
        for (int k = 0; k < fd.rows * fd.cols; k++) params[k] = fd.ptr()[k];
        finish(params, best.inlier_number, iters, t0);
    }

private:
    void finish(const float* params, int inlier_number, unsigned int iters, std::chrono::steady_clock::time_point t0,
                unsigned int lo_inner = 0, unsigned int lo_iterative = 0) {
        if (inlier_number <= 0) throw std::runtime_error("Ransac: best score is 0");
        const bool line = model->estimator == Line2d;
        cv::Mat d = line ? cv::Mat(1, 3) : cv::Mat(3, 3);
        for (int k = 0; k < (line ? 3 : 9); k++) d.ptr()[k] = params[k];
        Model best(model);
        best.setDescriptor(d);
        std::vector<int> inliers(points_size);
        quality->getInliers(d, inliers.data());                                       // ransac.cpp:210
        const long us = (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        delete ransac_output;
        ransac_output = new RansacOutput(&best, inliers.data(), us, (unsigned int)inlier_number, iters, lo_inner, lo_iterative, 0);
    }
};
