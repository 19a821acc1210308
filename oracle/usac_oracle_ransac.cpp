/*
 * usac_oracle_ransac.cpp - SPRT, PROSAC termination and the main loop of the CPU oracle.
 * TEST INFRASTRUCTURE ONLY (see usac_oracle.h). Citations are file:line under /root/reference.
 */
#include "oracle_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

/* sampler internals needed for the shared glibc stream (defined in usac_oracle.cpp) */
extern "C" int32_t orc_sampler_glibc_next(orc_sampler* s);

namespace {

/* ------------------------------------------------------------------------------------------------------
 * SPRT, usac/sprt.hpp
 * ------------------------------------------------------------------------------------------------------ */
struct SprtTest { double epsilon, delta, A; int k; };

struct Sprt {
    std::vector<SprtTest> hist;
    unsigned cur = 0;
    int last_update = 0;
    double t_M = 0, m_S = 0, threshold = 0;
    unsigned n = 0, sample_size = 0, max_iterations = 0, pool_idx = 0;
    int max_hyp_before = 20;                     /* model.hpp:39 */
    std::vector<unsigned> pool;

    /* sprt.hpp:332-355 */
    double estimateThresholdA(double epsilon, double delta) const {
        double C = (1 - delta) * std::log((1 - delta) / (1 - epsilon)) + delta * (std::log(delta / epsilon));
        double K = (t_M * C) / m_S + 1;
        double An_1 = K, An = K;
        for (unsigned i = 0; i < 10; ++i) {
            An = K + std::log(An_1);
            if (std::fabs(An - An_1) < 1.5e-8) break;
            An_1 = An;
        }
        return An;
    }

    /* ctor, sprt.hpp:89-175. `next_random` yields the glibc random() stream shared with the sampler. */
    template <class Rand>
    void init(int estimator, float thr, unsigned n_, unsigned m, unsigned max_it, Rand next_random) {
        n = n_; sample_size = m; max_iterations = max_it; threshold = thr;
        pool.resize(n);
        for (unsigned i = 0; i < n; i++) pool[i] = i;
        int max = (int)n;
        for (unsigned i = 0; i < n; i++) {                         /* :100-106 */
            unsigned idx = (unsigned)next_random() % (unsigned)max;
            unsigned t = pool[idx];
            max--;
            pool[idx] = pool[max];
            pool[max] = t;
        }
        pool_idx = 0;
        SprtTest t0;
        if (estimator == ORC_EST_HOMOGRAPHY) { t0.delta = 0.01; t0.epsilon = 0.1; t_M = 200; m_S = 1; }
        else if (estimator == ORC_EST_FUNDAMENTAL) { t0.delta = 0.05; t0.epsilon = 0.2; t_M = 200; m_S = 2.48; }
        else if (estimator == ORC_EST_ESSENTIAL) { t0.delta = 0.05; t0.epsilon = 0.2; t_M = 300; m_S = 4; }
        else { t0.delta = 0.0001; t0.epsilon = 0.001; t_M = 100; m_S = 1; }
        t0.A = estimateThresholdA(t0.epsilon, t0.delta);
        t0.k = 0;
        hist.clear();
        hist.push_back(t0);
        cur = 0;
        last_update = 0;
    }

    /* The likelihood-ratio walk of verifyModelAndGetModelScore (sprt.hpp:205-234) from pool offset `start`,
     * with an explicit test. Returns good; tested_point counts the points consumed by the test. */
    bool walk(const ErrFn& f, const float* pts, const SprtTest& t, unsigned start, unsigned& tested_point,
              unsigned& tested_inliers, unsigned& end_idx, unsigned long long& evals) const {
        double lambda = 1, lambda_new;
        unsigned idx = start;
        tested_inliers = 0;
        bool good = true;
        for (tested_point = 0; tested_point < n; tested_point++) {
            if (idx >= n) idx = 0;
            evals++;
            if (f(pts, pool[idx]) < threshold) {
                tested_inliers++;
                lambda_new = lambda * (t.delta / t.epsilon);
            } else {
                lambda_new = lambda * ((1 - t.delta) / (1 - t.epsilon));
            }
            idx++;
            if (lambda_new > t.A) {
                good = false;
                tested_point++;
                break;
            }
            lambda = lambda_new;
        }
        end_idx = idx;
        return good;
    }

    /* finish counting after a rejection during the first max_hyp_before hypotheses (sprt.hpp:243-257) */
    unsigned count_rest(const ErrFn& f, const float* pts, unsigned tested_point, unsigned& idx, unsigned long long& evals) const {
        unsigned c = 0;
        for (unsigned p = tested_point; p < n; p++) {
            if (idx >= n) idx = 0;
            evals++;
            if (f(pts, pool[idx]) < threshold) c++;
            idx++;
        }
        return c;
    }

    void push_test(double eps, double delta, int current_hypothese) {
        SprtTest t;
        t.epsilon = eps; t.delta = delta; t.A = estimateThresholdA(eps, delta);
        t.k = current_hypothese - last_update;
        last_update = current_hypothese;
        cur++;
        hist.push_back(t);
    }

    /* sequential verifyModelAndGetModelScore, sprt.hpp:191-317 */
    bool verify(const ErrFn& f, const float* pts, int current_hypothese, unsigned maximum_score, int& inl, float& score,
                unsigned long long& evals) {
        const SprtTest t = hist[cur];
        unsigned tested_point, tested_inliers, end;
        bool good = walk(f, pts, t, pool_idx, tested_point, tested_inliers, end, evals);
        pool_idx = end;
        if (good) {
            inl = (int)tested_inliers; score = (float)inl;
        } else if (current_hypothese < max_hyp_before) {
            unsigned rest = count_rest(f, pts, tested_point, pool_idx, evals);
            inl = (int)(tested_inliers + rest); score = (float)inl;
        }
        if (good) {
            if (tested_inliers > maximum_score) push_test((float)tested_inliers / n, t.delta, current_hypothese);   /* :266-282 */
        } else {
            float delta_estimated = (float)tested_inliers / tested_point;                                          /* :291 */
            if (delta_estimated > 0 && std::fabs(t.delta - delta_estimated) / t.delta > 0.05)
                push_test(t.epsilon, delta_estimated, current_hypothese);
        }
        return good;
    }

    /* sprt.hpp:442-491 */
    static double computeExponentH(double epsilon, double epsilon_new, double delta) {
        double a = std::log(delta / epsilon);
        double b = std::log((1 - delta) / (1 - epsilon));
        double x0 = std::log(1 / (1 - epsilon_new)) / b;
        double v0 = epsilon_new * std::exp(x0 * a);
        double x1 = std::log((1 - 2 * v0) / (1 - epsilon_new)) / b;
        double v1 = epsilon_new * std::exp(x1 * a) + (1 - epsilon_new) * std::exp(x1 * b);
        double h = x0 - (x0 - x1) / (1 + v0 - v1) * v0;
        if (std::isnan(h)) return 0;
        return h;
    }

    /* sprt.hpp:371-393 */
    unsigned getUpperBoundIterations(int inliers_size) const {
        double epsilon = (double)inliers_size / n;
        double P_g = std::pow(epsilon, (double)sample_size);
        double log_eta_l_1 = 0;
        for (unsigned test = 0; test < cur; test++) {
            double h = computeExponentH(hist[test].epsilon, epsilon, hist[test].delta);
            log_eta_l_1 += std::log(1 - P_g * (1 - std::pow(hist[test].A, -h))) * hist[test].k;
        }
        double numerator = std::log(0.05) - log_eta_l_1;
        if (numerator >= 0) return 0;
        double denumerator = std::log(1 - P_g * (1 - 1 / hist[cur].A));
        if (std::isnan(denumerator) || std::fabs(denumerator) < 0.00001) return max_iterations;
        double kl = numerator / denumerator;
        return (unsigned)std::min((unsigned)kl, max_iterations);
    }
};

/* ------------------------------------------------------------------------------------------------------
 * PROSAC termination, usac/termination_criteria/prosac_termination_criteria.hpp
 * ------------------------------------------------------------------------------------------------------ */
struct ProsacTermination {
    std::vector<unsigned> maximality_samples, non_random_inliers;
    const unsigned* growth = nullptr;
    unsigned termination_length = 0, n = 0, m = 0, max_iterations = 0;
    float threshold = 0, confidence = 0;
    const unsigned min_termination_length = 20;

    void init(const unsigned* growth_, unsigned n_, unsigned m_, float thr, float conf, unsigned max_it) {   /* :44-119 */
        growth = growth_; n = n_; m = m_; threshold = thr; confidence = conf; max_iterations = max_it;
        termination_length = n;
        const float non_randomness = 0.95f, beta = 0.05f;
        non_random_inliers.assign(n, 0);
        std::vector<double> pn_i_vec(n);
        for (size_t nn = m + 1; nn <= n; ++nn) {
            if (nn - 1 > 1000) { non_random_inliers[nn - 1] = non_random_inliers[nn - 2]; continue; }
            std::fill(pn_i_vec.begin(), pn_i_vec.end(), 0.0);
            pn_i_vec[m] = (beta) * std::pow((double)1 - beta, (double)nn - m - 1) * (nn - m);
            double pn_i = pn_i_vec[m];
            for (size_t i = m + 2; i <= nn; ++i) {
                if (i == nn) { pn_i_vec[nn - 1] = std::pow((double)beta, (double)nn - m); break; }
                pn_i_vec[i - 1] = pn_i * ((beta) / (1 - beta)) * ((double)(nn - i) / (i - m + 1));
                pn_i = pn_i_vec[i - 1];
            }
            double acc = 0.0;
            unsigned i_min = 0;
            for (size_t i = nn; i >= (size_t)m + 1; --i) {
                acc += pn_i_vec[i - 1];
                if (acc < 1 - non_randomness) i_min = (unsigned)i; else break;
            }
            non_random_inliers[nn - 1] = i_min;
        }
        maximality_samples.assign(n, 10000u);   /* :65, :115-118 */
    }

    /* getUpBoundIterations(hypCount, model), :148-201. `inl[i]` = GetError(i) < threshold over the sorted points. */
    unsigned update(unsigned hypCount, const std::vector<unsigned char>& inl, unsigned largest_sample_size) {
        unsigned max_samples = maximality_samples[termination_length - 1];
        unsigned inlier_count = 0;
        for (unsigned i = 0; i < min_termination_length; i++) inlier_count += inl[i];
        bool is_inlier_iplus1 = false;
        bool is_inlier_i = inl[min_termination_length];
        for (unsigned i = min_termination_length; i < n; ++i) {
            if (i != n - 1) is_inlier_iplus1 = inl[i + 1];
            inlier_count += is_inlier_i;
            if (non_random_inliers[i] < inlier_count) {
                non_random_inliers[i] = inlier_count;
                if ((i == n - 1) || (is_inlier_i && !is_inlier_iplus1)) {
                    unsigned new_samples = orc_standard_termination(inlier_count, i + 1, (int)m, confidence, max_iterations);
                    if (i + 1 < largest_sample_size) new_samples += hypCount - growth[i];
                    if (new_samples < maximality_samples[i]) {
                        maximality_samples[i] = new_samples;
                        if ((new_samples < max_samples) || ((new_samples == max_samples) && (i + 1 >= termination_length))) {
                            termination_length = i + 1;
                            max_samples = new_samples;
                        }
                    }
                }
            }
            is_inlier_i = is_inlier_iplus1;
        }
        return max_samples;
    }
};

inline bool bigger(int inl_a, float sc_a, int inl_b, float sc_b) {   /* Score::bigger, quality.hpp:22-26 */
    if (inl_a > inl_b) return true;
    if (inl_a == inl_b) return sc_a > sc_b;
    return false;
}

}  // namespace

/* ------------------------------------------------------------------------------------------------------
 * Ransac::run main loop, usac/ransac/ransac.cpp:58-139 (no LO, no final polish - SURVEY.md section 8f "next")
 *
 * batch == 0: the reference's sequential semantics.
 * batch == K: rounds of K samples. All samples of a round are drawn first (PROSAC termination_length frozen), all
 *   models are verified under the state at the start of the round (SPRT test frozen; model q = j*S+i of the round
 *   starts its pool walk at (cursor + 32*q) mod N), then the round is replayed in order with the reference's
 *   best-update / iteration accounting / termination logic, stopping at the first sample whose `iters` reached
 *   `max_iters`. After the round the SPRT test is re-designed once (epsilon from the last accepted improving
 *   model, delta from the pooled rejected models) and the cursor advances by 32*K*S. Without SPRT and PROSAC the
 *   result is identical to batch==0 for every K.
 * ------------------------------------------------------------------------------------------------------ */
/* The SPRT point pool (sprt.hpp:93-107) a fit with this seed uses: Fisher-Yates over the sampler's glibc stream. Exposed so
 * that the tests can hand the identical pool to the CUDA path (usac_gpu_set_sprt_pool). */
extern "C" void orc_sprt_pool(uint64_t seed, int n, int* pool_out) {
    orc_sampler* sampler = orc_sampler_new(ORC_SAMPLER_UNIFORM, ORC_RNG_PHILOX, n, 2, seed);
    Sprt sprt;
    sprt.init(ORC_EST_HOMOGRAPHY, 1.f, (unsigned)n, 2, 1, [&]() { return orc_sampler_glibc_next(sampler); });
    for (int i = 0; i < n; i++) pool_out[i] = (int)sprt.pool[i];
    orc_sampler_free(sampler);
}

extern "C" int orc_ransac(const orc_config* cfg, const float* points, int n, orc_result* out) {
    const int est = cfg->estimator;
    const int m = est == ORC_EST_LINE2D ? 2 : est == ORC_EST_HOMOGRAPHY ? 4 : est == ORC_EST_FUNDAMENTAL ? 7 : 5;
    const int S = est == ORC_EST_FUNDAMENTAL ? 3 : 1;
    const int msize = est == ORC_EST_LINE2D ? 3 : 9;
    memset(out, 0, sizeof(*out));
    out->best_hyp = -1;

    orc_sampler* sampler = orc_sampler_new(cfg->sampler, cfg->rng == ORC_RNG_TABLE ? ORC_RNG_PHILOX : cfg->rng, n, m, cfg->seed);
    if (cfg->sampler == ORC_SAMPLER_NAPSAC) {
        if (cfg->neighbors == ORC_NEIGH_GRID) orc_sampler_set_grid(sampler, points, cfg->cell_size);
        else orc_sampler_set_knn(sampler, cfg->knn_table, cfg->knn);
    }
    const bool is_prosac = cfg->sampler == ORC_SAMPLER_PROSAC;
    ProsacTermination pterm;
    if (is_prosac) pterm.init(orc_sampler_growth_function(sampler), n, m, cfg->threshold, cfg->confidence, cfg->max_iterations);

    Sprt sprt;
    if (cfg->sprt) {
        /* Ransac ctor order (ransac.hpp:50-92): sampler first, SPRT last; both draw from the one glibc stream */
        sprt.init(est, cfg->threshold, n, m, cfg->max_iterations, [&]() { return orc_sampler_glibc_next(sampler); });
    }

    int best_inl = 0;
    float best_score = 0;
    float best_model[9] = {0};
    unsigned iters = 0, max_iters = cfg->max_iterations;
    unsigned long long evals = 0;
    unsigned models_scored = 0;
    uint64_t hyp = 0;   /* samples drawn so far */
    ErrFn f;
    std::vector<unsigned char> inl_mask;

    auto draw = [&](int* sample) {
        if (cfg->rng == ORC_RNG_TABLE) {
            if (hyp < cfg->sample_table_rows) memcpy(sample, cfg->sample_table + (size_t)hyp * m, sizeof(int) * m);
            else orc_sampler_generate(sampler, hyp, sample);
        } else {
            if (is_prosac) orc_sampler_set_termination_length(sampler, pterm.termination_length);
            orc_sampler_generate(sampler, hyp, sample);
        }
        hyp++;
    };
    orc_lo* lo = cfg->lo ? orc_lo_new(est, points, n, cfg->threshold, cfg->lo, cfg->seed) : nullptr;
    auto on_new_best = [&](const float* model_in, int inl, float score, long long h, int mi) {
        float lo_model[9] = {0};
        memcpy(lo_model, model_in, sizeof(float) * msize);
        if (lo) orc_lo_get_model_score(lo, lo_model, &inl, &score);                 /* ransac.cpp:108-110 */
        const float* model = lo_model;
        best_inl = inl; best_score = score;
        memcpy(best_model, model, sizeof(float) * msize);
        out->best_hyp = h; out->best_model_idx = mi;
        if (is_prosac) {                                              /* ransac.cpp:123-125 */
            inl_mask.resize(n);
            f.set(est, model);
            for (int i = 0; i < n; i++) inl_mask[i] = f(points, (unsigned)i) < cfg->threshold;
            max_iters = pterm.update(iters, inl_mask, orc_sampler_largest_sample_size(sampler));
        } else {
            max_iters = orc_standard_termination((unsigned)best_inl, (unsigned)n, m, cfg->confidence, cfg->max_iterations);
        }
        if (cfg->sprt) max_iters = std::min(max_iters, sprt.getUpperBoundIterations(best_inl));   /* :129-133 */
    };

    int sample[8];
    float models[3 * 9];
    auto solve = [&](const int* smp, float* mdl) {                    /* Appendix B quirk 1 as a switch */
        if (cfg->ref_thin_svd && est == ORC_EST_HOMOGRAPHY) return orc_solve_homography_dlt4p_thin(points, smp, mdl);
        return orc_solve_minimal(est, points, smp, mdl);
    };

    if (cfg->batch <= 0) {
        /* ---------------- reference-sequential ---------------- */
        while (iters < max_iters) {
            draw(sample);
            int nm = solve(sample, models);
            for (int i = 0; i < nm; i++) {
                int cur_inl = 0; float cur_score = 0;
                const float* mdl = models + 9 * i;
                if (cfg->sprt) {
                    f.set(est, mdl);
                    /* NOTE: on a rejection after the first 20 hypotheses the reference leaves current_score
                     * untouched and `continue`s (ransac.cpp:77-85), so stale values are never compared. */
                    bool good = sprt.verify(f, points, (int)iters, (unsigned)best_inl, cur_inl, cur_score, evals);
                    models_scored++;
                    if (!good && iters >= 20) { iters++; continue; }
                } else {
                    orc_score(est, points, n, mdl, cfg->threshold, &cur_inl, &cur_score, nullptr, 0, nullptr);
                    evals += n; models_scored++;
                }
                if (bigger(cur_inl, cur_score, best_inl, best_score)) on_new_best(mdl, cur_inl, cur_score, (long long)hyp - 1, i);
            }
            iters++;
        }
    } else {
        /* ---------------- batched(K) ---------------- */
        const int K = cfg->batch;
        std::vector<int> samples((size_t)K * m);
        std::vector<float> rmodels((size_t)K * S * 9);
        std::vector<int> nmodels(K), r_inl((size_t)K * S), r_good((size_t)K * S), r_tested_inl((size_t)K * S), r_tested_pts((size_t)K * S);
        std::vector<float> r_score((size_t)K * S);
        unsigned cursor = 0;
        bool done = false;
        while (!done && iters < max_iters) {
            const uint64_t hyp0 = hyp;
            if (is_prosac) orc_sampler_set_termination_length(sampler, pterm.termination_length);
            for (int j = 0; j < K; j++) {
                if (cfg->rng == ORC_RNG_TABLE && hyp < cfg->sample_table_rows)
                    memcpy(&samples[(size_t)j * m], cfg->sample_table + (size_t)hyp * m, sizeof(int) * m);
                else
                    orc_sampler_generate(sampler, hyp, &samples[(size_t)j * m]);
                hyp++;
                nmodels[j] = solve(&samples[(size_t)j * m], &rmodels[(size_t)j * S * 9]);
            }
            const SprtTest frozen = cfg->sprt ? sprt.hist[sprt.cur] : SprtTest();
            for (int j = 0; j < K; j++) {
                for (int i = 0; i < nmodels[j]; i++) {
                    const size_t q = (size_t)j * S + i;
                    const float* mdl = &rmodels[q * 9];
                    if (cfg->sprt) {
                        f.set(est, mdl);
                        unsigned tp, ti, end;
                        unsigned start = (unsigned)(((unsigned long long)cursor + 32ull * q) % (unsigned)n);
                        bool good = sprt.walk(f, points, frozen, start, tp, ti, end, evals);
                        r_good[q] = good; r_tested_inl[q] = (int)ti; r_tested_pts[q] = (int)tp;
                        r_inl[q] = (int)ti;
                        if (!good && hyp0 + j < 20) r_inl[q] = (int)(ti + sprt.count_rest(f, points, tp, end, evals));
                        r_score[q] = (float)r_inl[q];
                    } else {
                        orc_score(est, points, n, mdl, cfg->threshold, &r_inl[q], &r_score[q], nullptr, 0, nullptr);
                        evals += n;
                        r_good[q] = 1;
                    }
                    models_scored++;
                }
            }
            /* replay */
            const unsigned iters_round_start = iters;
            long long last_improving_inl = -1;
            unsigned long long rej_inl = 0, rej_pts = 0;
            for (int j = 0; j < K; j++) {
                if (!(iters < max_iters)) { done = true; break; }
                for (int i = 0; i < nmodels[j]; i++) {
                    const size_t q = (size_t)j * S + i;
                    if (cfg->sprt) {
                        if (r_good[q]) { if (r_tested_inl[q] > best_inl) last_improving_inl = r_tested_inl[q]; }
                        else { rej_inl += r_tested_inl[q]; rej_pts += r_tested_pts[q]; }
                        if (!r_good[q] && iters >= 20) { iters++; continue; }
                    }
                    if (bigger(r_inl[q], r_score[q], best_inl, best_score))
                        on_new_best(&rmodels[q * 9], r_inl[q], r_score[q], (long long)(hyp0 + j), i);
                }
                iters++;
            }
            if (cfg->sprt) {
                const SprtTest t = sprt.hist[sprt.cur];
                double eps = t.epsilon, delta = t.delta;
                bool redesign = false;
                if (last_improving_inl >= 0) { eps = (float)last_improving_inl / n; redesign = true; }
                if (rej_pts > 0) {
                    float delta_estimated = (float)rej_inl / (unsigned)rej_pts;
                    if (delta_estimated > 0 && std::fabs(t.delta - delta_estimated) / t.delta > 0.05) { delta = delta_estimated; redesign = true; }
                }
                if (redesign) sprt.push_test(eps, delta, (int)iters_round_start);
                cursor = (unsigned)(((unsigned long long)cursor + 32ull * K * S) % (unsigned)n);
            }
        }
    }

    memcpy(out->model, best_model, sizeof(float) * msize);
    out->inliers = best_inl;
    out->score = best_score;
    out->iterations = iters;
    out->samples_drawn = (unsigned)hyp;
    out->evals = evals;
    out->models_scored = models_scored;
    if (lo) { unsigned a, b; unsigned long long c; orc_lo_counters(lo, &a, &b, &c); out->lo_inner = a; out->lo_iterative = b; orc_lo_free(lo); }
    orc_sampler_free(sampler);
    return best_inl > 0 ? 0 : 1;   /* ransac.cpp:143-147: the reference exits(111) when nothing was found */
}

/* ------------------------------------------------------------------------------------------------------
 * Sequences for the cross-check against the compiled reference (oracle/_ref, tests/test_ref_build.py): the same inputs go
 * through the reference's SPRT / ProsacTerminationCriteria classes (ref_driver.cpp) and through the restatements above.
 * ------------------------------------------------------------------------------------------------------ */
extern "C" int orc_sprt_sequence(int est, const float* points, int n, float thr, uint64_t seed, unsigned max_it, const float* models, int M,
                                 const int* hyp, int* good_out, int* inl_out, unsigned* pool_idx_after, unsigned* bound_after,
                                 double* hist_out, int* nhist, int* pool_out) {
    const int m = est == ORC_EST_LINE2D ? 2 : est == ORC_EST_HOMOGRAPHY ? 4 : est == ORC_EST_FUNDAMENTAL ? 7 : 5;
    orc_sampler* sampler = orc_sampler_new(ORC_SAMPLER_UNIFORM, ORC_RNG_GLIBC, n, m, seed);
    Sprt sprt;
    sprt.init(est, thr, (unsigned)n, (unsigned)m, max_it, [&]() { return orc_sampler_glibc_next(sampler); });
    for (int i = 0; i < n; i++) pool_out[i] = (int)sprt.pool[i];
    ErrFn f;
    int best = 0;
    unsigned long long evals = 0;
    for (int q = 0; q < M; q++) {
        f.set(est, models + 9 * q);
        int inl = -1;
        float score = -1;
        const bool good = sprt.verify(f, points, hyp[q], (unsigned)best, inl, score, evals);
        good_out[q] = good;
        inl_out[q] = (good || hyp[q] < sprt.max_hyp_before) ? inl : -1;
        bound_after[q] = 0xffffffffu;
        if (inl_out[q] > best) { best = inl_out[q]; bound_after[q] = sprt.getUpperBoundIterations(best); }
        pool_idx_after[q] = sprt.pool_idx;
    }
    *nhist = (int)sprt.hist.size();
    for (int i = 0; i < *nhist && i < 4096; i++) {
        hist_out[4 * i] = sprt.hist[i].epsilon; hist_out[4 * i + 1] = sprt.hist[i].delta; hist_out[4 * i + 2] = sprt.hist[i].A; hist_out[4 * i + 3] = sprt.hist[i].k;
    }
    orc_sampler_free(sampler);
    return 0;
}

extern "C" int orc_prosac_termination_sequence(int est, const float* points, int n, float thr, float conf, unsigned max_it, const float* models, int M,
                                               const unsigned* hyp_count, const unsigned* largest, unsigned* max_samples_out, unsigned* term_len_out) {
    const int m = est == ORC_EST_LINE2D ? 2 : est == ORC_EST_HOMOGRAPHY ? 4 : est == ORC_EST_FUNDAMENTAL ? 7 : 5;
    orc_sampler* sampler = orc_sampler_new(ORC_SAMPLER_PROSAC, ORC_RNG_PHILOX, n, m, 1);
    ProsacTermination pterm;
    pterm.init(orc_sampler_growth_function(sampler), (unsigned)n, (unsigned)m, thr, conf, max_it);
    ErrFn f;
    std::vector<unsigned char> mask(n);
    for (int q = 0; q < M; q++) {
        f.set(est, models + 9 * q);
        for (int i = 0; i < n; i++) mask[i] = f(points, (unsigned)i) < thr;
        max_samples_out[q] = pterm.update(hyp_count[q], mask, largest[q]);
        term_len_out[q] = pterm.termination_length;
    }
    orc_sampler_free(sampler);
    return 0;
}
