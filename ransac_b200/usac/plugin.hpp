// plugin.hpp - the abstract plugin surface of the reference, re-authored with the same names and signatures:
//   Estimator            usac/estimator/estimator.hpp:14-41
//   Sampler              usac/sampler/sampler.hpp:12-35
//   Score, Quality       usac/quality/quality.hpp:16-121 (the entry points are virtual HERE so that a GPU Quality can subclass;
//                        in the reference Quality is a concrete class looping the virtual Estimator::GetError)
//   TerminationCriteria  usac/termination_criteria/termination_criteria.hpp:10-18
//   LocalOptimization    usac/local_optimization/local_optimization.hpp:12-20
#pragma once
#include <cassert>
#include <iostream>
#include <vector>

#include "model.hpp"

class Estimator {
public:
    virtual ~Estimator() = default;
    virtual unsigned int EstimateModel(const int* const sample, std::vector<Model*>& models) = 0;
    virtual bool EstimateModelNonMinimalSample(const int* const sample, unsigned int sample_size, Model& model) = 0;
    virtual bool LeastSquaresFitting(const int* const sample, unsigned int sample_size, Model& model) {
        return EstimateModelNonMinimalSample(sample, sample_size, model);
    }
    virtual bool EstimateModelNonMinimalSample(const int* const, unsigned int, const float* const, Model&) {
        std::cout << "NOT IMPLEMENTED EstimateModelNonMinimalSample in estimator\n";
        return false;
    }
    virtual float GetError(unsigned int pidx) = 0;
    virtual void GetError(float*, float, int*, unsigned int*) { std::cout << "NOT IMPLEMENTED GetError (float * weights) in estimator\n"; }
    virtual int SampleNumber() = 0;
    virtual void setModelParameters(const cv::Mat& model) = 0;
    virtual bool isModelValid(const cv::Mat&, const int* const) { return true; }
};

class Sampler {
protected:
    unsigned int k_iterations = 0, points_size = 0, sample_size = 0;
public:
    virtual ~Sampler() = default;
    virtual void generateSample(int* sample) = 0;
    unsigned int getNumberOfIterations() { return k_iterations; }
    virtual bool isInit() { return false; }
};

class Score {
public:
    int inlier_number = 0;
    float score = 0;
    // more inliers wins; ties go to the LARGER error sum (quality.hpp:22-31)
    inline bool bigger(const Score* const o) { return inlier_number > o->inlier_number || (inlier_number == o->inlier_number && score > o->score); }
    inline bool bigger(const Score& o) { return bigger(&o); }
    void copyFrom(const Score* const o) { score = o->score; inlier_number = o->inlier_number; }
};

class Quality {
protected:
    unsigned int points_size = 0;
    float threshold = 0;
    Estimator* estimator = nullptr;
    bool isinit = false;
public:
    virtual ~Quality() = default;
    bool isInit() { return isinit; }
    virtual void init(unsigned int points_size_, float threshold_, Estimator* estimator_) {
        points_size = points_size_; threshold = threshold_; estimator = estimator_; isinit = true;
    }
    // count = #{err < threshold}, score = sum of the inliers' errors; quality.hpp:60-101 (generic form over GetError)
    virtual void getNumberInliers(Score* score, const cv::Mat& model, float threshold_ = 0, bool get_inliers = false,
                                  int* inliers = nullptr, bool /*parallel*/ = false) {
        if (threshold_ == 0) threshold_ = threshold;
        estimator->setModelParameters(model);
        unsigned int n = 0;
        float sum = 0;
        for (unsigned int p = 0; p < points_size; p++) {
            const float err = estimator->GetError(p);
            if (err < threshold_) { if (get_inliers) inliers[n] = (int)p; n++; sum += err; }
        }
        score->inlier_number = (int)n;
        score->score = sum;
    }
    virtual void getInliers(const cv::Mat& model, int* inliers) {                   // quality.hpp:108-121
        assert(isinit);
        estimator->setModelParameters(model);
        int n = 0;
        for (unsigned int p = 0; p < points_size; p++) if (estimator->GetError(p) < threshold) inliers[n++] = (int)p;
    }
};

class TerminationCriteria {
protected:
    bool isinit = false;
public:
    virtual ~TerminationCriteria() = default;
    bool isInit() { return isinit; }
    virtual unsigned int getUpBoundIterations(unsigned int inlier_size) = 0;
    virtual unsigned int getUpBoundIterations(unsigned int inlier_size, unsigned int points_size) = 0;
};

class LocalOptimization {
public:
    virtual ~LocalOptimization() = default;
    virtual void GetModelScore(Model* best_model, Score* best_score) = 0;
};
