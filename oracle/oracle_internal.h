/* oracle_internal.h - shared pieces of the CPU oracle (TEST INFRASTRUCTURE ONLY, see usac_oracle.h). */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H
#include "usac_oracle.h"

#include <cmath>

struct ErrFn {
    int est;
    float p[18];
    void set(int estimator, const float* model) {
        est = estimator;
        if (est == ORC_EST_LINE2D) {
            p[0] = model[0]; p[1] = model[1]; p[2] = model[2];          /* line2d_estimator.hpp:162-164 */
        } else {
            for (int i = 0; i < 9; i++) p[i] = model[i];
            if (est == ORC_EST_HOMOGRAPHY) orc_inv3x3(model, p + 9);   /* homography_estimator.hpp:33-45 */
        }
    }
    /* `sqrt`/`fabsf` bind to the float overloads: <math.h> leaks into the reference's TUs through
     * opencv2/flann (hard part 1), so sqrtf is used here. */
    inline float operator()(const float* pts, unsigned i) const {
        if (est == ORC_EST_LINE2D) {                                    /* line2d_estimator.hpp:154-156 */
            return fabsf(p[0] * pts[2 * i] + p[1] * pts[2 * i + 1] + p[2]);
        }
        const float x1 = pts[4 * i], y1 = pts[4 * i + 1], x2 = pts[4 * i + 2], y2 = pts[4 * i + 3];
        if (est == ORC_EST_HOMOGRAPHY) {                                /* homography_estimator.hpp:85-110 */
            float ex = p[0] * x1 + p[1] * y1 + p[2];
            float ey = p[3] * x1 + p[4] * y1 + p[5];
            float ez = p[6] * x1 + p[7] * y1 + p[8];
            ex /= ez; ey /= ez;
            float fx = p[9] * x2 + p[10] * y2 + p[11];
            float fy = p[12] * x2 + p[13] * y2 + p[14];
            float fz = p[15] * x2 + p[16] * y2 + p[17];
            fx /= fz; fy /= fz;
            float e = sqrtf((x2 - ex) * (x2 - ex) + (y2 - ey) * (y2 - ey)) + sqrtf((x1 - fx) * (x1 - fx) + (y1 - fy) * (y1 - fy));
            return e / 2;
        }
        if (est == ORC_EST_FUNDAMENTAL) {                               /* fundamental_estimator.hpp:101-117 */
            float a = p[0] * x1 + p[1] * y1 + p[2];
            float b = p[3] * x1 + p[4] * y1 + p[5];
            float c = p[0] * x2 + p[3] * y2 + p[6];
            float d = p[1] * x2 + p[4] * y2 + p[7];
            float n = x2 * a + y2 * b + p[6] * x1 + p[7] * y1 + p[8];
            return (n * n) / (a * a + b * b + c * c + d * d);
        }
        /* essential_estimator.hpp:76-107 */
        float l1 = p[0] * x2 + p[3] * y2 + p[6];
        float l2 = p[1] * x2 + p[4] * y2 + p[7];
        float l3 = p[2] * x2 + p[5] * y2 + p[8];
        float t1 = p[0] * x1 + p[1] * y1 + p[2];
        float t2 = p[3] * x1 + p[4] * y1 + p[5];
        float t3 = p[6] * x1 + p[7] * y1 + p[8];
        float a1 = l1 * x1 + l2 * y1 + l3;
        float a2 = sqrtf(l1 * l1 + l2 * l2);
        float b1 = t1 * x2 + t2 * y2 + t3;
        float b2 = sqrtf(t1 * t1 + t2 * t2);
        return (fabsf(a1 / a2) + fabsf(b1 / b2)) / 2;
    }
};


int orc_solve_essential5(const float* pts, const int* s, float* out);   /* usac_oracle_essential.cpp */

#endif
