// Canonical libm for the path: sinf cosf atan2f asinf logf powf(x,5) -- the only external arithmetic the
// reference's hot path uses (call sites pcg.cpp:92-93, sphere.cpp:7-8, volumes.cpp:24, material.h:109,
// texture.cpp:9, scene_object.cpp:36-37).
//
// The reference's result depends on whichever libm the host provides (glibc picks FMA / non-FMA variants
// per CPU at run time, MSVC's differs again), and a 1-ulp difference in a sampled direction is amplified
// chaotically by a few specular bounces.  To make "same inputs -> same image" well defined, BOTH sides
// use the correctly rounded value: the function is evaluated in double precision (error <= 2 ulp of
// double) and rounded once to float, which yields the IEEE correctly rounded float result except when the
// exact value lies within ~1e-16 relative of a rounding boundary (probability ~1e-8 per call).
// The oracle binary interposes the same six functions (oracle/cr_libm.cpp); glibc's own versions agree
// with these in >99.9% of arguments and always within 1 ulp (tests/test_oracle_pinned.py).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define MRT_LIBM_HD __host__ __device__ __forceinline__
#define MRT_LIBM_FN __host__ __device__ __forceinline__   /* every function below has exactly one call site in the kernels */
#else
#define MRT_LIBM_HD inline
#define MRT_LIBM_FN inline
#endif

namespace mrt {
MRT_LIBM_FN float cr_sinf(float x) { return (float) sin((double) x); }
MRT_LIBM_FN float cr_cosf(float x) { return (float) cos((double) x); }
// sinf and cosf of the same argument (one shared range reduction); identical values to cr_sinf / cr_cosf
MRT_LIBM_FN void cr_sincosf(float x, float *s, float *c) {
    double ds, dc;
    sincos((double) x, &ds, &dc);
    *s = (float) ds;
    *c = (float) dc;
}
MRT_LIBM_FN float cr_logf(float x) { return (float) log((double) x); }
MRT_LIBM_FN float cr_atan2f(float y, float x) { return (float) atan2((double) y, (double) x); }
MRT_LIBM_FN float cr_asinf(float x) { return (float) asin((double) x); }
// powf(x, 5): x^5 with three double multiplications (<= 1.5 ulp of double before the single rounding)
MRT_LIBM_HD float cr_pow5f(float x) {
    double d = (double) x;
    double d2 = d * d;
    double d4 = d2 * d2;
    return (float) (d4 * d);
}
}  // namespace mrt
