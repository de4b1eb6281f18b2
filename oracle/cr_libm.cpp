// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.
//
// Link-time interposition of the six libm functions the reference's path calls (sinf cosf atan2f asinf
// logf powf) by their correctly rounded values: evaluated in double precision by the host libm and
// rounded once to float.  The reference's SOURCES are unchanged; only the symbol the calls bind to is.
// Rationale: see miniraytracer_b200/csrc/mrt_libm.h.  Setting MRT_ORACLE_LIBM=host in the environment
// routes the calls to the host libm's own float functions instead (used by the test that measures how
// often the two differ).
// Compiled with -fno-builtin so that (float)sin((double)x) is not folded back into sinf(x).
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace {
typedef float (*f1_t)(float);
typedef float (*f2_t)(float, float);
typedef void (*sc_t)(float, float *, float *);
struct HostLibm {
    bool use_host;
    f1_t sinf_, cosf_, asinf_, logf_;
    f2_t atan2f_, powf_;
    sc_t sincosf_;
    HostLibm() {
        const char *e = getenv("MRT_ORACLE_LIBM");
        use_host = e && !strcmp(e, "host");
        sinf_ = (f1_t) dlsym(RTLD_NEXT, "sinf");
        cosf_ = (f1_t) dlsym(RTLD_NEXT, "cosf");
        asinf_ = (f1_t) dlsym(RTLD_NEXT, "asinf");
        logf_ = (f1_t) dlsym(RTLD_NEXT, "logf");
        atan2f_ = (f2_t) dlsym(RTLD_NEXT, "atan2f");
        powf_ = (f2_t) dlsym(RTLD_NEXT, "powf");
        sincosf_ = (sc_t) dlsym(RTLD_NEXT, "sincosf");
    }
};
const HostLibm &host() {
    static HostLibm h;
    return h;
}
}  // namespace

extern "C" {
float sinf(float x) { return host().use_host ? host().sinf_(x) : (float) sin((double) x); }
float cosf(float x) { return host().use_host ? host().cosf_(x) : (float) cos((double) x); }
// gcc merges sinf(x) + cosf(x) (pcg.cpp:92-93) into one sincosf call
void sincosf(float x, float *s, float *c) {
    if (host().use_host) { host().sincosf_(x, s, c); return; }
    *s = (float) sin((double) x);
    *c = (float) cos((double) x);
}
float asinf(float x) { return host().use_host ? host().asinf_(x) : (float) asin((double) x); }
float logf(float x) { return host().use_host ? host().logf_(x) : (float) log((double) x); }
float atan2f(float y, float x) { return host().use_host ? host().atan2f_(y, x) : (float) atan2((double) y, (double) x); }
float powf(float x, float y) {
    if (host().use_host) return host().powf_(x, y);
    if (y == 5.0f) {   // the only exponent on the path (fresnel_schlick, material.h:109)
        double d = (double) x, d2 = d * d, d4 = d2 * d2;
        return (float) (d4 * d);
    }
    return (float) pow((double) x, (double) y);
}
}
