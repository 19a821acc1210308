// sprt.hpp - Wald's sequential probability ratio test behind the reference's class (usac/sprt.hpp:40-493): same constructor,
// verifyModelAndGetModelScore (sprt.hpp:191-317) and getUpperBoundIterations (sprt.hpp:371-393) signatures and semantics.
// The point walk runs on the device (usac_gpu_sprt_verify: likelihood ratio in double, exact inlier decisions); the adaptive test
// design - (epsilon, delta, A) history, threshold A, the iteration bound with its log/exp/pow arithmetic - is host code shared
// with the fused path (csrc/host_replay.hpp). The shuffled point pool (sprt.hpp:93-107) is drawn from glibc's random() generator
// (TYPE_3, through the re-entrant random_r so that Model::seed, not the process-wide state, decides it) and uploaded once.
#pragma once
#include <cstdlib>

#include "../csrc/host_replay.hpp"
#include "gpu_plugins.hpp"

class SPRT {
    GpuDevice* dev;
    SprtHost host;                       // test history, A, iteration bound
    unsigned int points_size, random_pool_idx = 0;
    float threshold;
    int max_hypothesis_test_before_sprt;
    std::vector<int> points_random_pool;

public:
    // glibc random() stream -> Fisher-Yates pool, sprt.hpp:93-107
    static std::vector<int> shuffledPool(unsigned long long seed, unsigned int n) {
        struct random_data buf;
        char state[128];
        std::memset(&buf, 0, sizeof(buf));
        initstate_r((unsigned int)seed, state, sizeof(state), &buf);
        std::vector<int> pool(n);
        for (unsigned int i = 0; i < n; i++) pool[i] = (int)i;
        int max = (int)n;
        for (unsigned int i = 0; i < n; i++) {
            int32_t r;
            random_r(&buf, &r);
            const unsigned int idx = (unsigned int)r % (unsigned int)max;
            const int t = pool[idx];
            max--;
            pool[idx] = pool[max];
            pool[max] = t;
        }
        return pool;
    }

    SPRT(Model* model, Estimator* estimator_, unsigned int points_size_) : points_size(points_size_), threshold(model->threshold),
          max_hypothesis_test_before_sprt((int)model->max_hypothesis_test_before_sprt) {
        GpuEstimator* ge = dynamic_cast<GpuEstimator*>(estimator_);
        if (!ge) throw std::runtime_error("SPRT: needs a GpuEstimator (the point walk runs on the device)");
        dev = ge->device();
        points_random_pool = shuffledPool(model->seed, points_size);
        dev->check(usac_gpu_set_sprt_pool(dev->ctx, 0, points_random_pool.data()), "usac_gpu_set_sprt_pool");
        host.init(dev->estimator, points_size, model->sample_size, model->max_iterations);
    }

    bool verifyModelAndGetModelScore(Model* model, int current_hypothese, unsigned int maximum_score, Score* score) {
        const SprtTestH t = host.current();
        const cv::Mat d = model->returnDescriptor();
        float params[9] = {0};
        for (int k = 0; k < d.rows * d.cols; k++) params[k] = d.ptr()[k];
        if (random_pool_idx >= points_size) random_pool_idx = 0;
        const int count_all = current_hypothese < max_hypothesis_test_before_sprt;
        usac_sprt_result r{};
        dev->check(usac_gpu_sprt_verify(dev->ctx, 0, params, 1, threshold, t.epsilon, t.delta, t.A, &random_pool_idx, &count_all, &r), "usac_gpu_sprt_verify");
        const bool good = r.good != 0;
        // the cursor continues behind the points this call consumed (sprt.hpp:212-223, 246-255)
        const unsigned int consumed = (!good && count_all) ? points_size : (unsigned int)r.tested_points;
        random_pool_idx = (random_pool_idx + consumed) % points_size;
        if (good || count_all) { score->inlier_number = r.inliers; score->score = (float)r.inliers; }     // sprt.hpp:236-257
        if (good) {
            if ((unsigned int)r.tested_inliers > maximum_score) host.push((float)r.tested_inliers / points_size, t.delta, current_hypothese);   // :266-282
        } else {
            const float delta_estimated = (float)r.tested_inliers / r.tested_points;                        // :291
            if (delta_estimated > 0 && std::fabs(t.delta - delta_estimated) / t.delta > 0.05) host.push(t.epsilon, delta_estimated, current_hypothese);
        }
        return good;
    }

    unsigned int getUpperBoundIterations(int inliers_size) { return host.upper_bound(inliers_size); }
    const std::vector<int>& pool() const { return points_random_pool; }
};
