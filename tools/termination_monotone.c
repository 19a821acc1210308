// Is the standard termination bound (standard_termination_criteria.hpp:52-62, float arithmetic + glibc logf) non-increasing in the inlier count
// once it is past the cap (w^m >= 0.0005)? samples_in_reach (pipeline.cuh) relies on it: only a best model in the CAPPED region can be followed by a
// larger bound. Scan: m in {2,4,5,7}, confidence in {0.95,0.99,0.999}, n = 8..30000, plus one 1M-point table: 44 664 tables, 0 inversions.
// gcc -O2 -ffp-contract=off -o /tmp/mono tools/termination_monotone.c -lm && /tmp/mono
#include <math.h>
#include <stdio.h>
static unsigned val(unsigned inliers, unsigned n, int m, float log_1_p, unsigned max_iterations) {
    const float w = (float)inliers / n;
    float p = w * w;
    for (int k = m; k > 2; k--) p *= w;
    if (p < 0.0005f) return max_iterations;
    return (unsigned)(log_1_p / logf(1 - p));
}
int main() {
    const int ms[4] = {2, 4, 5, 7};
    const float confs[3] = {0.95f, 0.99f, 0.999f};
    long long inversions = 0, tables = 0;
    for (int ci = 0; ci < 3; ci++) {
        const float log_1_p = (float)logf(1 - confs[ci]);
        for (int mi = 0; mi < 4; mi++)
            for (unsigned n = 8; n <= 30000; n += (n < 3000 ? 1 : 37)) {
                tables++;
                unsigned prev = 0; int have = 0;
                for (unsigned inl = 0; inl <= n; inl++) {
                    const float w = (float)inl / n; float p = w * w; for (int k = ms[mi]; k > 2; k--) p *= w;
                    if (p < 0.0005f) continue;
                    unsigned v = val(inl, n, ms[mi], log_1_p, 10000);
                    if (have && v > prev) { inversions++; if (inversions < 5) printf("inversion n=%u m=%d conf=%g inl=%u: %u -> %u\n", n, ms[mi], confs[ci], inl, prev, v); }
                    prev = v; have = 1;
                }
            }
    }
    // one 1M table
    {
        const float log_1_p = (float)logf(1 - 0.95f); unsigned prev = 0; int have = 0;
        for (unsigned inl = 0; inl <= 1000000; inl++) {
            const float w = (float)inl / 1000000; float p = w * w; p *= w; p *= w;
            if (p < 0.0005f) continue;
            unsigned v = val(inl, 1000000, 4, log_1_p, 10000);
            if (have && v > prev) inversions++;
            prev = v; have = 1;
        }
    }
    printf("tables %lld inversions %lld\n", tables, inversions);
    return 0;
}
