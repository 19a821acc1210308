// usac_harness.cpp - the reference's test harness shape (test/test.cpp:5-60: build a Model, construct Ransac(model, points),
// run(), print RansacOutput) over the GPU plugin layer. Points come from a `*_pts.txt`-style file (first line N, then N
// rows `x1 y1 x2 y2`, or `x y` for lines - the format of dataset/homography/sift_update/*_pts.txt).
//   usac_harness <points.txt> <line2d|homography|fundamental> <uniform|prosac|napsac> <threshold> <confidence> [seed] [--sequential] [--both]
// Prints one `key=value` line per result (model as IEEE bit patterns, inlier ids as a hash) for the parity tests.
#include <cstring>
#include <fstream>

#include "ransac.hpp"

static void report(const char* tag, Ransac& r, const Model& m) {
    RansacOutput* o = r.getRansacOutput();
    const cv::Mat d = o->getModel()->returnDescriptor();
    std::printf("%s iterations=%u inliers=%u time_us=%ld model_bits=", tag, o->getNumberOfMainIterations(), o->getNumberOfInliers(), o->getTimeMicroSeconds());
    for (int k = 0; k < d.rows * d.cols; k++) { unsigned u; float f = d.ptr()[k]; std::memcpy(&u, &f, 4); std::printf("%08x%s", u, k + 1 < d.rows * d.cols ? "," : ""); }
    unsigned long long h = 1469598103934665603ull;                                   // FNV-1a over the inlier ids
    for (int id : o->getInliers()) { h ^= (unsigned)id; h *= 1099511628211ull; }
    std::printf(" inlier_hash=%016llx score=%.9g\n", h, (double)r.lastFit().score);
    (void)m;
}

int main(int argc, char** argv) {
    if (argc < 6 || !std::strcmp(argv[1], "--help")) {
        std::fprintf(stderr, "usage: %s <points.txt> <line2d|homography|fundamental> <uniform|prosac|napsac> <threshold> <confidence> [seed] [--sequential|--both]\n", argv[0]);
        return argc < 2 ? 2 : (!std::strcmp(argv[1], "--help") ? 0 : 2);
    }
    const std::string est = argv[2], smp = argv[3];
    const ESTIMATOR e = est == "line2d" ? Line2d : est == "homography" ? Homography : est == "fundamental" ? Fundamental : NullE;
    const SAMPLER s = smp == "uniform" ? Uniform : smp == "prosac" ? Prosac : smp == "napsac" ? Napsac : NullS;
    if (e == NullE || s == NullS) { std::fprintf(stderr, "unknown estimator/sampler\n"); return 2; }
    const int dim = e == Line2d ? 2 : 4;
    std::ifstream in(argv[1]);
    int n = 0;
    if (!(in >> n) || n <= 0) { std::fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    cv::Mat points(n, dim);
    for (int i = 0; i < n * dim; i++) if (!(in >> points.ptr()[i])) { std::fprintf(stderr, "short points file\n"); return 2; }
    const unsigned m = e == Line2d ? 2 : e == Homography ? 4 : 7;
    Model model((float)std::atof(argv[4]), m, (float)std::atof(argv[5]), 5, e, s);
    bool sequential = false, both = false;
    for (int i = 6; i < argc; i++) {
        if (!std::strcmp(argv[i], "--sequential")) sequential = true;
        else if (!std::strcmp(argv[i], "--both")) both = true;
        else model.seed = std::strtoull(argv[i], nullptr, 10);
    }
    model.setCellSize(50);
    if (s == Napsac) model.setNeighborsType(Grid);
    try {
        if (!sequential || both) { Ransac r(&model, points); r.run(); report("fused", r, model); }
        if (sequential || both) { Ransac r(&model, points); r.run_sequential(); report("sequential", r, model); }
    } catch (const std::exception& ex) {
        std::fprintf(stderr, "usac_harness: %s\n", ex.what());
        return 111;                                                                  // the reference's fatal exit code (init.cpp:17-19)
    }
    return 0;
}
