// host_replay.hpp - host-side pieces of the rounds that use SPRT and/or PROSAC termination.
//
// In those modes the verification / scoring of a round's models runs on the device (sprt_walk_kernel, score_kernel), and
// the short sequential part - replaying the round in hypothesis order with the reference's accounting
// (ransac.cpp:58-139), re-designing the SPRT test (sprt.hpp:259-311, 332-355), the SPRT iteration bound
// (sprt.hpp:371-393, 442-491: log/exp/pow in double) and the PROSAC non-randomness / maximality tables
// (prosac_termination_criteria.hpp:44-201) - runs here between rounds, on the results the round's one host sync brings
// back. libm transcendental functions are used exactly where the reference uses them.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

// standard_termination_criteria.hpp:52-62 (float arithmetic, truncation)
static inline unsigned standard_termination_value(unsigned inliers, unsigned n, int m, float log_1_p, unsigned max_iterations) {
    const float w = (float)inliers / n;
    float p = w * w;
    for (int k = m; k > 2; k--) p *= w;
    if (p < 0.0005f) return max_iterations;
    return (unsigned)(log_1_p / logf(1 - p));
}

struct SprtTestH { double epsilon, delta, A; int k; };

struct SprtHost {
    std::vector<SprtTestH> hist;
    int last_update = 0;
    double t_M = 0, m_S = 0;
    unsigned n = 0, sample_size = 0, max_iterations = 0, cursor = 0;

    double threshold_A(double epsilon, double delta) const {          // sprt.hpp:332-355
        const double C = (1 - delta) * std::log((1 - delta) / (1 - epsilon)) + delta * (std::log(delta / epsilon));
        const double K = (t_M * C) / m_S + 1;
        double prev = K, cur = K;
        for (unsigned i = 0; i < 10; ++i) {
            cur = K + std::log(prev);
            if (std::fabs(cur - prev) < 1.5e-8) break;
            prev = cur;
        }
        return cur;
    }
    void init(int est, unsigned n_, unsigned m, unsigned max_it) {    // sprt.hpp:114-175
        n = n_; sample_size = m; max_iterations = max_it; cursor = 0; last_update = 0;
        SprtTestH t;
        if (est == USAC_EST_HOMOGRAPHY) { t.delta = 0.01; t.epsilon = 0.1; t_M = 200; m_S = 1; }
        else if (est == USAC_EST_FUNDAMENTAL) { t.delta = 0.05; t.epsilon = 0.2; t_M = 200; m_S = 2.48; }
        else if (est == USAC_EST_ESSENTIAL) { t.delta = 0.05; t.epsilon = 0.2; t_M = 300; m_S = 4; }
        else { t.delta = 0.0001; t.epsilon = 0.001; t_M = 100; m_S = 1; }
        t.A = threshold_A(t.epsilon, t.delta);
        t.k = 0;
        hist.assign(1, t);
    }
    const SprtTestH& current() const { return hist.back(); }
    void push(double eps, double delta, int hypothesis) {             // sprt.hpp:266-311 (test history bookkeeping)
        SprtTestH t;
        t.epsilon = eps; t.delta = delta; t.A = threshold_A(eps, delta);
        t.k = hypothesis - last_update;
        last_update = hypothesis;
        hist.push_back(t);
    }
    static double exponent_h(double epsilon, double epsilon_new, double delta) {   // sprt.hpp:442-491
        const double a = std::log(delta / epsilon), b = std::log((1 - delta) / (1 - epsilon));
        const double x0 = std::log(1 / (1 - epsilon_new)) / b;
        const double v0 = epsilon_new * std::exp(x0 * a);
        const double x1 = std::log((1 - 2 * v0) / (1 - epsilon_new)) / b;
        const double v1 = epsilon_new * std::exp(x1 * a) + (1 - epsilon_new) * std::exp(x1 * b);
        const double h = x0 - (x0 - x1) / (1 + v0 - v1) * v0;
        return std::isnan(h) ? 0 : h;
    }
    unsigned upper_bound(int inliers) const {                          // sprt.hpp:371-393
        const double epsilon = (double)inliers / n;
        const double P_g = std::pow(epsilon, (double)sample_size);
        double log_eta = 0;
        for (size_t t = 0; t + 1 < hist.size(); t++) {
            const double h = exponent_h(hist[t].epsilon, epsilon, hist[t].delta);
            log_eta += std::log(1 - P_g * (1 - std::pow(hist[t].A, -h))) * hist[t].k;
        }
        const double num = std::log(0.05) - log_eta;
        if (num >= 0) return 0;
        const double den = std::log(1 - P_g * (1 - 1 / hist.back().A));
        if (std::isnan(den) || std::fabs(den) < 0.00001) return max_iterations;
        return (unsigned)std::min((unsigned)(num / den), max_iterations);
    }
};

struct ProsacTermHost {
    std::vector<unsigned> maximality_samples, non_random_inliers, growth;
    unsigned termination_length = 0, n = 0, m = 0, max_iterations = 0;
    float log_1_p = 0;
    static constexpr unsigned kMinTerminationLength = 20;

    // The initial non-randomness table (prosac_termination_criteria.hpp:58-113) depends on (n, m) only and costs ~0.5 M pow/mul per
    // call: it is computed once per (n, m) and copied for every fit (the table itself is updated while a fit runs).
    static const std::vector<unsigned>& initial_non_random_inliers(unsigned n, unsigned m) {
        static std::mutex mu;
        static std::map<std::pair<unsigned, unsigned>, std::vector<unsigned>> cache;
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find({n, m});
        if (it != cache.end()) return it->second;
        if (cache.size() > 32) cache.clear();
        const float beta = 0.05f, non_randomness = 0.95f;               // float constants: the mixed float/double expressions below follow :58-103
        std::vector<unsigned> table(n, 0);
        std::vector<double> pn(n);
        for (size_t nn = (size_t)m + 1; nn <= n; ++nn) {
            if (nn - 1 > 1000) { table[nn - 1] = table[nn - 2]; continue; }
            std::fill(pn.begin(), pn.begin() + nn, 0.0);              // entries [m, nn) are the only ones read below
            pn[m] = (beta) * std::pow((double)1 - beta, (double)nn - m - 1) * (nn - m);
            double cur = pn[m];
            for (size_t i = (size_t)m + 2; i <= nn; ++i) {
                if (i == nn) { pn[nn - 1] = std::pow((double)beta, (double)nn - m); break; }
                pn[i - 1] = cur * ((beta) / (1 - beta)) * ((double)(nn - i) / (i - m + 1));
                cur = pn[i - 1];
            }
            double acc = 0.0;
            unsigned i_min = 0;
            for (size_t i = nn; i >= (size_t)m + 1; --i) {
                acc += pn[i - 1];
                if (acc < 1 - non_randomness) i_min = (unsigned)i; else break;
            }
            table[nn - 1] = i_min;
        }
        return cache.emplace(std::make_pair(n, m), std::move(table)).first->second;
    }
    void init(const std::vector<unsigned>& growth_, unsigned n_, unsigned m_, float confidence, unsigned max_it) {   // :44-119
        growth = growth_; n = n_; m = m_; max_iterations = max_it; termination_length = n;
        log_1_p = (float)logf(1 - confidence);
        non_random_inliers = initial_non_random_inliers(n, m);
        maximality_samples.assign(n, 10000u);
    }
    // getUpBoundIterations(hypCount, model), :148-201; mask[i] = (GetError(i) < threshold) over the quality-sorted points
    unsigned update(unsigned hyp_count, const std::vector<unsigned char>& mask, unsigned largest_sample_size) {
        unsigned max_samples = maximality_samples[termination_length - 1];
        if (n <= kMinTerminationLength) return max_samples;
        unsigned count = 0;
        for (unsigned i = 0; i < kMinTerminationLength; i++) count += mask[i];
        bool next = false, cur = mask[kMinTerminationLength];
        for (unsigned i = kMinTerminationLength; i < n; ++i) {
            if (i != n - 1) next = mask[i + 1];
            count += cur;
            if (non_random_inliers[i] < count) {
                non_random_inliers[i] = count;
                if (i == n - 1 || (cur && !next)) {
                    unsigned s = standard_termination_value(count, i + 1, (int)m, log_1_p, max_iterations);
                    if (i + 1 < largest_sample_size) s += hyp_count - growth[i];
                    if (s < maximality_samples[i]) {
                        maximality_samples[i] = s;
                        if (s < max_samples || (s == max_samples && i + 1 >= termination_length)) { termination_length = i + 1; max_samples = s; }
                    }
                }
            }
            cur = next;
        }
        return max_samples;
    }
};
