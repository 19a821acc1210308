/*
 * usac_oracle.h - C ABI of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. This library is a CPU restatement of the hypothesize-and-verify path of the reference
 * (MathsionYang/Ransac, `usac/`). It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product (ransac_b200/) never links, imports or
 * falls back to it.
 *
 * Parity status (DESIGN.md section 3 has the table):
 *  - pinned by the reference's own artefacts: the homography and Sampson scoring functions reproduce the 46 known-answer vectors
 *    recovered from the reference's shipped datasets/results (tests/golden/scoring_kat.npz);
 *  - pinned by THE REFERENCE ITSELF, compiled here: oracle/_ref/libusac_ref.so is the reference's own usac/ sources built where they
 *    lie (oracle/Makefile.ref; OpenCV / Eigen / nanoflann answered by the stand-ins of oracle/ref_shim/, whose SVD, eigen, solveCubic
 *    and 3x3 inverse are themselves held against the real OpenCV: tests/golden/svd_cv.npz, eig_cv.npz, cv_primitives.npz).
 *    tests/test_ref_build.py: all four GetError metrics and Quality::getNumberInliers, the uniform sampler's glibc stream, the PROSAC
 *    schedule, standard and PROSAC termination, SPRT (decisions, counts, pool cursor, test history, iteration bound), the
 *    neighbourhood structures and the line solver are BIT-IDENTICAL / index-identical to the compiled reference; Ransac::run as a
 *    whole for line fits (also past max_iterations); the non-minimal, 7-point and 5-point solvers agree to float32 conditioning;
 *  - the reference's quirks are kept, or are switches: orc_config::ref_thin_svd = its own 4-point DLT4p,
 *    pinned against the real OpenCV (tests/golden/dlt4p_cv.npz) and the compiled reference;
 *  - still "parity unpinned": what the reference computes THROUGH an OpenCV numerical routine whose low-order bits cannot be
 *    reproduced here - the basis cv::SVD picks inside a null space (hence the root ORDER of the five-point solver and the float32
 *    low bits of 7-point models). Those rows are capped at "partial" in DESIGN.md.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#ifndef USAC_ORACLE_H
#define USAC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* enums mirror usac/model.hpp:10-13 */
enum { ORC_EST_LINE2D = 1, ORC_EST_HOMOGRAPHY = 2, ORC_EST_FUNDAMENTAL = 3, ORC_EST_ESSENTIAL = 4 };
enum { ORC_SAMPLER_UNIFORM = 1, ORC_SAMPLER_PROGRESSIVE_NAPSAC = 2, ORC_SAMPLER_NAPSAC = 3, ORC_SAMPLER_PROSAC = 4 };
enum { ORC_NEIGH_NONE = 0, ORC_NEIGH_KNN = 1, ORC_NEIGH_GRID = 2 };
/* random source of the samplers */
enum { ORC_RNG_GLIBC = 0 /* reference stream: glibc random(), uniform_sampler.hpp:42-54 */,
       ORC_RNG_PHILOX = 1 /* counter based, keyed by (seed, hypothesis id) */,
       ORC_RNG_TABLE = 2 /* caller supplied K x m index table */ };

/* ---- RNG streams ---- */
typedef struct orc_glibc_rand orc_glibc_rand;
orc_glibc_rand* orc_glibc_rand_new(unsigned seed);             /* srandom(seed) */
void orc_glibc_rand_free(orc_glibc_rand*);
int32_t orc_glibc_rand_next(orc_glibc_rand*);                   /* random() */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---- scoring (usac/quality/quality.hpp:60-101 over Estimator::GetError) ---- */
/* per-point error, float32, one rounding per operator, no FMA */
void orc_errors(int estimator, const float* points, int n, const float* model, float* err_out);
/* count/sum exactly as Quality::getNumberInliers; inliers_out may be NULL; thr==0 is not special here.
 * flagged_out (may be NULL) = number of points with |err-thr| <= band_rel*thr. */
void orc_score(int estimator, const float* points, int n, const float* model, float thr, int* count_out,
               float* sum_out, int* inliers_out, double band_rel, int* flagged_out);
/* cv::Mat::inv() of a 3x3 CV_32F (OpenCV core, closed-form cofactor path): returns 0 when det==0 (all zeros) */
int orc_inv3x3(const float* m, float* out);

/* ---- minimal solvers: returns number of models written to models_out (<= 3 models x 9 floats; line: 3 floats) ---- */
int orc_solve_minimal(int estimator, const float* points, const int* sample, float* models_out);
/* SURVEY.md Appendix B quirk 1 as a switch: the reference's DLt::DLT4p (dlt.cpp:7-53: raw coordinates, last row of a THIN SVD = the 8th
 * singular vector, not the null vector). Returns the number of models (0/1). orc_config::ref_thin_svd runs a whole fit with it. */
int orc_solve_homography_dlt4p_thin(const float* points, const int* sample, float* model_out);
/* pieces exposed for unit tests */
int orc_solve_cubic(const double c[4], double roots[3]);       /* c0 x^3 + c1 x^2 + c2 x + c3, cv::solveCubic order */
int orc_fundamental_is_valid(const float* points, const float* F, const int* sample);
int orc_null_space(double* A, int rows, double* basis_out);    /* A rows x 9 row-major, destroyed; basis (9-rows) x 9 */

/* five-point candidates in visiting order (ascending |z|): Es_out[10*9] doubles, valid_out[10] = passed the cheirality vote */
int orc_essential5_candidates(const float* points, const int* sample, double* Es_out, int* valid_out);

/* ---- samplers ---- */
typedef struct orc_sampler orc_sampler;
orc_sampler* orc_sampler_new(int kind, int rng, int n, int m, uint64_t seed);
void orc_sampler_free(orc_sampler*);
/* neighbourhood structures for NAPSAC */
void orc_sampler_set_knn(orc_sampler*, const int* neighbors, int k);
void orc_sampler_set_grid(orc_sampler*, const float* points, int cell_size);
void orc_sampler_set_termination_length(orc_sampler*, unsigned len);   /* PROSAC coupling */
unsigned orc_sampler_largest_sample_size(orc_sampler*);
const unsigned* orc_sampler_growth_function(orc_sampler*);
/* generate the sample of hypothesis `hyp_id` (sequential samplers ignore hyp_id and advance their state) */
void orc_sampler_generate(orc_sampler*, uint64_t hyp_id, int* sample_out);
void orc_philox_unique(uint64_t seed, uint64_t hyp_id, uint32_t stream, int n, int m, int* out);

/* ---- termination ---- */
unsigned orc_standard_termination(unsigned inliers, unsigned n, int m, float confidence, unsigned max_iterations);

/* ---- k nearest neighbours (usac/utils/nearest_neighbors.cpp:69-128, nanoflann KD-tree, L2_Simple metric over all `dim`
 * columns, k+1 results of which the first - the query itself - is dropped; ascending distance). Brute force here.
 * nanoflann is not in /root/reference (un-vendored, un-pinned): "parity unpinned" for the order of exactly equidistant
 * neighbours, which in the KD-tree depends on the traversal; the oracle breaks such ties by ascending point index.
 * Distance: float32, sum of squared differences accumulated in column order, one rounding per operator.
 * table_out: n x k; returns 0, or -1 when n < k + 1. */
int orc_knn_build(const float* points, int n, int dim, int k, int* table_out);

/* ---- grid neighbours (usac/utils/nearest_neighbors.cpp:160-201) as CSR: cell id per point, members sorted ---- */
void orc_grid_cells(const float* points, int n, int cell_size, int* cell_of_point, int* members, int* cell_start,
                    int* ncells_out);

/* ---- full robust fit (usac/ransac/ransac.cpp:14-238, main loop :58-139; no final polish) ---- */
typedef struct {
    int estimator, sampler, rng;
    float threshold, confidence;
    unsigned max_iterations;
    int sprt;                 /* 0/1 */
    int batch;                /* 0 = reference-sequential; K>0 = batched(K) semantics (state frozen within a round) */
    int neighbors, knn, cell_size;
    uint64_t seed;
    const int* sample_table;  /* rng==ORC_RNG_TABLE: K_total x m */
    unsigned sample_table_rows;
    const int* knn_table;     /* neighbors==KNN: n x knn */
    int lo;                   /* LocOpt: 0 none, 1 InItLORsc, 2 InItFLORsc (model.hpp:13; inner_local_optimization.hpp) */
    int ref_thin_svd;         /* Appendix B quirk 1: 1 = homography samples go through the reference's DLT4p (orc_solve_homography_dlt4p_thin) */
} orc_config;

typedef struct {
    float model[9];
    int inliers;
    float score;
    unsigned iterations;      /* value of `iters` when the loop ended */
    unsigned samples_drawn;
    long long best_hyp;       /* sample index that produced the best model, -1 if none */
    int best_model_idx;       /* which root of that sample */
    unsigned long long evals; /* GetError calls executed */
    unsigned models_scored;
    unsigned lo_inner, lo_iterative;   /* RansacOutput::getLOInnerIters / getLOIterativeIters */
} orc_result;

int orc_ransac(const orc_config* cfg, const float* points, int n, orc_result* out);
/* ---- non-minimal estimation and the final refit (usac_oracle_refit.cpp) ---- */
int orc_nonminimal(int est, const float* pts, const int* ids, int n, float* model_out);
int orc_refit(int est, const float* pts, int n_points, float thr, float* model_io, int best_inliers, int* ids_out, int* accepted_out);
/* LO-RANSAC (usac_oracle_refit.cpp) */
typedef struct orc_lo orc_lo;
orc_lo* orc_lo_new(int est, const float* pts, int n, float theta, int kind, uint64_t seed);
void orc_lo_free(orc_lo*);
void orc_lo_counters(orc_lo*, unsigned* inner, unsigned* iterative, unsigned long long* calls);
void orc_lo_get_model_score(orc_lo*, float* best_model, int* best_inl_io, float* best_sum_io);
/* the SPRT pool permutation orc_ransac uses for this seed (sprt.hpp:93-107) */
void orc_sprt_pool(uint64_t seed, int n, int* pool_out);
/* sequences for the cross-check against the compiled reference (oracle/_ref; see usac_oracle_ransac.cpp) */
int orc_sprt_sequence(int est, const float* points, int n, float thr, uint64_t seed, unsigned max_it, const float* models, int M,
                      const int* hyp, int* good_out, int* inl_out, unsigned* pool_idx_after, unsigned* bound_after,
                      double* hist_out, int* nhist, int* pool_out);
int orc_prosac_termination_sequence(int est, const float* points, int n, float thr, float conf, unsigned max_it, const float* models, int M,
                                    const unsigned* hyp_count, const unsigned* largest, unsigned* max_samples_out, unsigned* term_len_out);

#ifdef __cplusplus
}
#endif
#endif
