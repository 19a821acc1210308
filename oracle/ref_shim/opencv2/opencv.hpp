// stand-in for <opencv2/opencv.hpp>: see cvshim.hpp (test infrastructure, oracle/_ref build only)
#pragma once
// The real umbrella header pulls the C API headers (core_c.h -> types_c.h: <assert.h> <stdlib.h> <string.h> <float.h> <math.h>),
// and with libstdc++ <math.h> brings the float overloads of sqrt / log / fabs into the global namespace. That decides what the
// reference's unqualified `sqrt(float)` calls bind to (homography_estimator.hpp:104-105, normalizing_transformation.cpp:46-47):
// with <math.h> the float overload, with <cmath> alone the double one (a difference of at most one ulp of the error; the 46
// known-answer inlier counts hold either way). The oracle and the CUDA strict path assume the float overloads (SURVEY.md section
// 7, hard part 1); build with -DUSAC_REF_NO_MATH_H to see the other binding.
#ifndef USAC_REF_NO_MATH_H
#include <math.h>
#endif
#include "../cvshim.hpp"
