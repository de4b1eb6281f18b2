// Per-path tracer core: camera ray generation, scene traversal, primitive hit
// tests, material scatter, mixture-pdf light sampling, textures and the PCG32
// generator -- the body of the reference's hot path (main.cpp:66-188 and
// everything it calls), rebuilt as an iterative, stack-machine tracer over the
// flattened scene of mrt_types.h.
//
// The functions are written once and compiled by nvcc for sm_100a (the product)
// and, for logic tests only, by g++ (tests/host_emul) -- the shipped library
// has no CPU execution path.
//
// Arithmetic contract (what "parity" means): every
// floating-point expression keeps the reference's operation order, divisions
// and square roots are IEEE (no rsqrt / reciprocal shortcuts), and the file is
// compiled with -fmad=false (nvcc) / -ffp-contract=off (g++) so that no
// multiply-add is fused.  The libm calls on the path (sinf cosf atan2f asinf
// logf powf) use the correctly rounded values of mrt_libm.h on both sides.
#pragma once
#include "mrt_types.h"
#include "mrt_libm.h"
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define MRT_HD __host__ __device__ __forceinline__
// Everything is inlined (ABI calls would force Ray/Hit through local memory); the hot loop is kept
// inside the instruction cache by giving every large piece exactly ONE call site (ncu round 1: 55 % of
// warp stalls were stall_no_inst with 154 KB of SASS from duplicated inlined bodies).
#define MRT_FN __host__ __device__ __forceinline__
#else
#define MRT_HD inline
#define MRT_FN inline
#endif

namespace mrt {

// Algorithmic operation counters (op counts of the REFERENCE algorithm, from which the roofline's
// flop per ray is derived -- tools/alg_flops.py).  Compiled in only by the test-side host emulation
// (tests/host_emul/emul_render.cpp defines MRT_COUNT_OPS and owns the thread-local `mrt_ops`); nothing on the device.
#if defined(MRT_COUNT_OPS) && !defined(__CUDACC__)
struct OpCounts {
    unsigned long long paths, ray_ctor, rng, sphere_hit, sphere_moving, rect_hit, tri_hit, translate, rotate, lambert, metal, dielectric,
        isotropic, lightpdf, perlin, image, checker, sky;
};
extern thread_local OpCounts mrt_ops;
#define MRT_OP(x) (mrt::mrt_ops.x++)
#else
#define MRT_OP(x) ((void) 0)
#endif

// Scene-feature specialisation.  Every function below that takes a `feat` argument is force-inlined into a
// kernel instantiated for a compile-time constant mask (MRT_FEAT_* in mrt_types.h); code for object classes,
// materials and textures the scene does not contain is removed by constant folding, which shortens the
// hot loop (fewer branch targets, fewer live registers).  MRT_FEAT_ALL keeps everything.
#define MRT_HAS(feat, bit) (((feat) & (bit)) != 0u)

#define MRT_PI_F 3.14159265358979323846f /* mrt_math.h:11 */

// ------------------------------------------------------------------ utilities
MRT_HD uint32_t f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
MRT_HD float u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
MRT_HD float fsqrt(float x) {  // MRT::sqrt = sqrtss (mrt_math.h:76-78): correctly rounded
#ifdef __CUDA_ARCH__
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
MRT_HD float fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
MRT_HD float frcp(float x) {   // 1 / x, correctly rounded: the same value as fdiv(1, x) from a shorter sequence
#ifdef __CUDA_ARCH__
    return __frcp_rn(x);
#else
    return 1.0f / x;
#endif
}
// x / len with a shortcut for x == +-0: 0 / len is that same signed zero for any positive finite len, and the
// zero-numerator case takes the slow path of __fdiv_rn (ncu round 1: 2 slow-path calls per warp iteration in the
// Cornell box: axis-aligned onb vectors, a light pdf of 0, a scattering pdf of 0).  len_ok = len is positive and finite.
MRT_HD float fdiv_zero_num(float x, float len, bool len_ok) {
    if (len_ok && x == 0.0f) return x;
    return fdiv(x, len);
}
MRT_HD bool is_finite(float x) { return (f2u(x) & 0x7F800000u) != 0x7F800000u; }

MRT_HD MrtF4 ld4(const MrtF4 *p, uint32_t i) {
#ifdef __CUDA_ARCH__
    float4 v = __ldg(reinterpret_cast<const float4 *>(p) + i);
    MrtF4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    return p[i];
#endif
}
MRT_HD uint32_t ldu(const uint32_t *p, uint32_t i) {
#ifdef __CUDA_ARCH__
    return __ldg(p + i);
#else
    return p[i];
#endif
}

struct V3 { float x, y, z; };
MRT_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
MRT_HD V3 v3(const MrtF4 &f) { return v3(f.x, f.y, f.z); }
MRT_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
MRT_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
MRT_HD V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
MRT_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
MRT_HD V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
MRT_HD V3 operator/(V3 a, float s) { return v3(fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)); }
MRT_HD V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
// vec3.h:245-248 : (x + y) + z
MRT_HD float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
MRT_HD float sdot(V3 a) { return dot(a, a); }
// vec3.h:250-266
MRT_HD V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// vec3.h:137-139 : v / sqrt(v.v)
MRT_HD V3 normalize(V3 a) { return a / fsqrt(sdot(a)); }

// ---------------------------------------------------------------------- PCG32
// pcg.cpp:13-35,53-62
struct Rng { uint64_t state, inc; };
MRT_HD uint32_t rng_next(Rng &r) {
    uint64_t old = r.state;
    r.state = old * 6364136223846793005ULL + r.inc;
    uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
}
MRT_HD void rng_seed(Rng &r, uint64_t initstate, uint64_t initseq) {
    r.state = 0u;
    r.inc = (initseq << 1u) | 1u;
    rng_next(r);
    r.state += initstate;
    rng_next(r);
}
MRT_HD float randf(Rng &r) { MRT_OP(rng); return u2f(0x3f800000u | (rng_next(r) & 0x007FFFFFu)) - 1.0f; }
// pcg.cpp:70-77 (draw order x, y, z)
MRT_HD V3 random_in_sphere(Rng &r) {
    V3 p;
    do {
        float rx = randf(r), ry = randf(r), rz = randf(r);
        p = v3(2.0f * rx - 1.0f, 2.0f * ry - 1.0f, 2.0f * rz - 1.0f);
    } while (sdot(p) >= 1.0f);
    return p;
}
// pcg.cpp:112-119
MRT_HD V3 random_in_disk(Rng &r) {
    V3 p;
    do {
        float rx = randf(r), ry = randf(r);
        p = v3(2.0f * rx - 1.0f, 2.0f * ry - 1.0f, 0.0f);
    } while (sdot(p) >= 1.0f);
    return p;
}
// pcg.cpp:87-95 (the factor 2 on x,y is the reference's)
MRT_FN V3 random_cosine_direction(Rng &r) {
    float r1 = randf(r);
    float r2 = randf(r);
    float z = fsqrt(1 - r2);
    float phi = 2 * MRT_PI_F * r1;
    float sr2 = fsqrt(r2);
    float sn, cs;
    cr_sincosf(phi, &sn, &cs);
    float x = cs * 2 * sr2;
    float y = sn * 2 * sr2;
    return v3(x, y, z);
}
// pcg.cpp:125-133
MRT_HD V3 random_towards_sphere(float radius, float dist_sq, Rng &r) {
    float r1 = randf(r);
    float r2 = randf(r);
    float z = 1 + r2 * (fsqrt(1 - fdiv(radius * radius, dist_sq)) - 1);
    float phi = 2 * MRT_PI_F * r1;
    float s = fsqrt(1 - z * z);
    float sn, cs;
    cr_sincosf(phi, &sn, &cs);
    float x = cs * s;
    float y = sn * s;
    return v3(x, y, z);
}

// ------------------------------------------------------------------------ ray
// ray.h:7-52.  inv = 1/dir is what aabb::hit recomputes per test (aabb.h:49).
struct Ray {
    V3 o, d, inv;
    float time;
    int inside;
    uint32_t mask;
};
MRT_HD uint32_t dir_mask(V3 d) {
    uint32_t X = f2u(d.x) >> 31, Y = f2u(d.y) >> 31, Z = f2u(d.z) >> 31;
    return 1u << (Z | (Y << 1) | (X << 2));
}
MRT_FN void ray_set_dir(Ray &r, V3 dir) {  // direction is normalised by the ctor (ray.h:30)
    MRT_OP(ray_ctor);
    r.d = normalize(dir);
    r.inv = v3(frcp(r.d.x), frcp(r.d.y), frcp(r.d.z));
    r.mask = dir_mask(r.d);
}
// ray for a primitive self-test (pdf_value): only origin + normalised direction are used
MRT_HD Ray make_probe_ray(V3 o, V3 dir, float time) {
    MRT_OP(ray_ctor);
    Ray r;
    r.o = o;
    r.time = time;
    r.inside = 0;
    r.d = normalize(dir);
    r.inv = v3(0, 0, 0);
    r.mask = 0;
    return r;
}
MRT_HD V3 ray_eval(const Ray &r, float t) { return r.o + t * r.d; }

// aabb.h:45-76 with the SSE min/max NaN rule (second operand wins) and strict >
MRT_HD bool aabb_hit(const MrtF4 &bmin, const MrtF4 &bmax, const Ray &r, float tmin, float tmax) {
    float t0x = (bmin.x - r.o.x) * r.inv.x, t1x = (bmax.x - r.o.x) * r.inv.x;
    float t0y = (bmin.y - r.o.y) * r.inv.y, t1y = (bmax.y - r.o.y) * r.inv.y;
    float t0z = (bmin.z - r.o.z) * r.inv.z, t1z = (bmax.z - r.o.z) * r.inv.z;
    if (r.inv.x < 0.0f) { float t = t0x; t0x = t1x; t1x = t; }
    if (r.inv.y < 0.0f) { float t = t0y; t0y = t1y; t1y = t; }
    if (r.inv.z < 0.0f) { float t = t0z; t0z = t1z; t1z = t; }
    float a0 = (t0x > t0z) ? t0x : t0z;
    float a1 = (t0y > tmin) ? t0y : tmin;
    float lo = (a0 > a1) ? a0 : a1;
    float b0 = (t1x < t1z) ? t1x : t1z;
    float b1 = (t1y < tmax) ? t1y : tmax;
    float hi = (b0 < b1) ? b0 : b1;
    return hi > lo;
}

// Conservative early-out for box-less transform nodes (translate, scene_object.cpp:9-18): true only if the ray
// certainly misses `lo..hi`, the object's bounds in the parent frame inflated by the flattener by 1e-3 of the scene
// scale (= slack, ~1000x the float32 rounding of the exact path at scene coordinates).  A miss inside such a node
// leaves no trace (ray restored, hit record / tmax / RNG untouched), so skipping the node for these rays does
// not change the result.  Any NaN in the test means "do not skip".
MRT_HD bool cull_miss(const MrtF4 &lo, const MrtF4 &hi, const Ray &r, float tmin, float tmax) {
    const float slack = hi.w;
    float t0 = (lo.x - r.o.x) * r.inv.x, t1 = (hi.x - r.o.x) * r.inv.x;
    float u0 = (lo.y - r.o.y) * r.inv.y, u1 = (hi.y - r.o.y) * r.inv.y;
    float v0 = (lo.z - r.o.z) * r.inv.z, v1 = (hi.z - r.o.z) * r.inv.z;
    // ordered comparisons are false on NaN; "ok" collects that every product is a number
    const bool ok = (t0 == t0) & (t1 == t1) & (u0 == u0) & (u1 == u1) & (v0 == v0) & (v1 == v1);
    float nx = t0 < t1 ? t0 : t1, fx = t0 < t1 ? t1 : t0;
    float ny = u0 < u1 ? u0 : u1, fy = u0 < u1 ? u1 : u0;
    float nz = v0 < v1 ? v0 : v1, fz = v0 < v1 ? v1 : v0;
    float enter = nx > ny ? nx : ny; enter = enter > nz ? enter : nz;
    float leave = fx < fy ? fx : fy; leave = leave < fz ? leave : fz;
    return ok && ((enter > leave) || (leave < tmin - slack) || (enter > tmax + slack));
}

// -------------------------------------------------------------- scene (device)
struct SceneView {
    const MrtF4 *sphere, *rect, *list, *bvh, *node2, *tri, *trin, *xlate, *rot, *vol, *mat, *tex, *perlin_vec;
    const uint32_t *child, *lights, *trileaf;
    const int32_t *perlin_perm;
    const uint8_t *image;
    uint32_t root, n_lights, sky;
    MrtCamera cam;
};

struct Hit {   // hit_record, scene_object.h:10-17
    float t;
    V3 p, n;
    float u, v;
    uint32_t mat;
};

// Traversal stack: 32-bit words, one column per thread.  On the GPU the columns
// of a warp are interleaved in shared memory (word k of lane l at base[k*32+l])
// so that every push/pop is conflict free.
struct Stack {
    uint32_t *base;
    uint32_t stride;
    uint32_t sp;
    MRT_HD void push(uint32_t v) { base[sp * stride] = v; sp++; }
    MRT_HD uint32_t pop() { sp--; return base[sp * stride]; }
    MRT_HD void pushf(float f) { push(f2u(f)); }
    MRT_HD float popf() { return u2f(pop()); }
};

// frame tags (top 3 bits of a stack word); payload = low 28 bits, bit 28 = list "found" flag
#define MRT_F_IF_MISS 1u   /* bvh: visit payload ref only if the closer child missed */
#define MRT_F_LIST 2u      /* object_list continuation: payload = index into child[] */
#define MRT_F_XLATE_END 3u /* translate: restore ray, rec.p += offset */
#define MRT_F_ROT_END 4u   /* rotate_y: restore ray, rotate rec.p / rec.n back */
#define MRT_F_VOL1 5u      /* constant_volume: first boundary hit returned */
#define MRT_F_VOL2 6u      /* constant_volume: second boundary hit returned */
#define MRT_FRAME(tag, payload) (((tag) << 29) | (payload))

struct Counters {   // optional algorithmic op counters
    unsigned long long rays, aabb, sphere, rect, tri, vol, xform;
};

// ----------------------------------------------------------- primitive tests
MRT_HD V3 sphere_center(const MrtF4 &s0, const MrtF4 &s1, const MrtF4 *tab, uint32_t idx, float time) {
    if (f2u(s1.w) >> 31) {  // isMoving, sphere.h:24-31
        MRT_OP(sphere_moving);
        MrtF4 s2 = ld4(tab, 3 * idx + 2);
        float k = fdiv(time - s2.x, s2.y - s2.x);
        return v3(s0) + k * (v3(s1) - v3(s0));
    }
    return v3(s0);
}
MRT_FN void sphere_uv(V3 n, float *u, float *v) {  // sphere.cpp:6-11
    float phi = cr_atan2f(n.z, n.x);
    float theta = cr_asinf(n.y);
    *u = 0.5f - phi * (1.0f / (2.0f * MRT_PI_F));
    *v = 0.5f + theta * (1.0f / MRT_PI_F);
}
// sphere.cpp:13-46.  full=false: only the distance is wanted (volume boundary probe).
MRT_FN bool hit_sphere(const uint32_t feat, const SceneView &sc, uint32_t idx, const Ray &r, float tmin, float tmax, bool full, Hit &rec) {
    MrtF4 s0 = ld4(sc.sphere, 3 * idx), s1 = ld4(sc.sphere, 3 * idx + 1);
    V3 cen = MRT_HAS(feat, MRT_FEAT_MOVING) ? sphere_center(s0, s1, sc.sphere, idx, r.time) : v3(s0);
    float radius = s0.w;
    V3 oc = r.o - cen;
    float b = dot(oc, r.d);
    float c = sdot(oc) - radius * radius;
    float disc = b * b - c;
    if (disc > 0) {
        float sq = fsqrt(disc);
        float t = (-b - sq);
        bool ok = (t < tmax && t > tmin);
        if (!ok && r.inside) {
            t = (-b + sq);
            ok = (t < tmax && t > tmin);
        }
        if (ok) {
            MRT_OP(sphere_hit);
            rec.t = t;
            if (full) {
                uint32_t mat = f2u(s1.w) & 0x7FFFFFFFu;
                rec.p = ray_eval(r, t);
                rec.n = (rec.p - cen) / radius;
                rec.mat = mat;
                if (MRT_HAS(feat, MRT_FEAT_TEX)) { if (f2u(ld4(sc.mat, mat).x) & MRT_MAT_NEEDS_UV) sphere_uv(rec.n, &rec.u, &rec.v); }
            }
            return true;
        }
    }
    return false;
}

// rect.cpp:24-45,69-90,130-152.  axis = 0: xy (k=z), 1: xz (k=y), 2: yz (k=x)
MRT_HD bool hit_rect(const uint32_t feat, const SceneView &sc, uint32_t axis, uint32_t idx, const Ray &r, float tmin, float tmax, bool full, Hit &rec) {
    MrtF4 q0 = ld4(sc.rect, 2 * idx), q1 = ld4(sc.rect, 2 * idx + 1);
    float ok_, dk, oa, da, ob, db;
    if (axis == 0)      { ok_ = r.o.z; dk = r.d.z; oa = r.o.x; da = r.d.x; ob = r.o.y; db = r.d.y; }
    else if (axis == 1) { ok_ = r.o.y; dk = r.d.y; oa = r.o.x; da = r.d.x; ob = r.o.z; db = r.d.z; }
    else                { ok_ = r.o.x; dk = r.d.x; oa = r.o.y; da = r.d.y; ob = r.o.z; db = r.d.z; }
    float ns = q1.y;
    if (dk * ns > 0.0f) return false;   // one-sided
    float t = fdiv(q1.x - ok_, dk);
    if (t < tmin || t > tmax) return false;
    float a = oa + t * da;
    float b = ob + t * db;
    if (a < q0.x || a > q0.y || b < q0.z || b > q0.w) return false;
    MRT_OP(rect_hit);
    rec.t = t;
    if (full) {
        rec.mat = f2u(q1.z);
        if (MRT_HAS(feat, MRT_FEAT_TEX) && (f2u(ld4(sc.mat, rec.mat).x) & MRT_MAT_NEEDS_UV)) {
            rec.u = fdiv(a - q0.x, q0.y - q0.x);
            rec.v = fdiv(b - q0.z, q0.w - q0.z);
        }
        rec.p = ray_eval(r, t);
        rec.n = (axis == 0) ? v3(0, 0, ns) : (axis == 1) ? v3(0, ns, 0) : v3(ns, 0, 0);
    }
    return true;
}

// triangle.cpp:222-266 (Moeller-Trumbore path; NEW_INTERSECT is off, common.h:7)
MRT_FN bool hit_triangle(const SceneView &sc, uint32_t idx, const Ray &r, float tmin, float tmax, bool full, Hit &rec) {
    MrtF4 tm = ld4(sc.tri, 3 * idx), tu = ld4(sc.tri, 3 * idx + 1), tv = ld4(sc.tri, 3 * idx + 2);
    V3 u = v3(tu), v = v3(tv);
    V3 pvec = cross(r.d, v);
    float det = dot(u, pvec);
    float sign = 1.0f;
    if (r.inside) {
        sign = det < 0.0f ? -1.0f : 1.0f;
        det = sign * det;
    }
    if (det < 0.00001f) return false;
    V3 tvec = r.o - v3(tm);
    float uu = dot(tvec, pvec) * sign;
    V3 qvec = cross(tvec, u);
    float vv = dot(r.d, qvec) * sign;
    if ((uu < 0) | (uu > det) | (vv < 0) | ((uu + vv) > det)) return false;
    float invDet = frcp(det);
    float t = dot(v, qvec) * invDet * sign;
    if ((t < tmin) | (t > tmax)) return false;
    MRT_OP(tri_hit);
    rec.t = t;
    if (full) {
        uu *= invDet;
        vv *= invDet;
        V3 mn = v3(ld4(sc.trin, 3 * idx)), un = v3(ld4(sc.trin, 3 * idx + 1)), vn = v3(ld4(sc.trin, 3 * idx + 2));
        rec.p = ray_eval(r, t);
        rec.n = normalize(((mn * (1 - uu - vv)) + (un * uu)) + (vn * vv));
        rec.u = uu;
        rec.v = vv;
        rec.mat = f2u(tm.w);
    }
    return true;
}

// ------------------------------------------------------------------ traversal
// scene.hit(r, tmin, tmax, &rec) of the reference, for any nesting of
// object_list / bvh_node / pod_bvh / translate / rotate_y / constant_volume.
// Invariants that make one hit record and one (tmin, tmax) pair sufficient:
//  * object_list (scene_object.h:79-103) passes its shrinking `closest` down,
//    so any hit reported by a nested container supersedes the list's best;
//  * bvh_node / pod_bvh (scene_object.h:208-244, triangle.h:171-213) return on
//    the first child (front-to-back by node_order & dirMask) that reports a
//    hit and never tighten tmax between children -> IF_MISS frames;
//  * constant_volume (volumes.cpp:5-36) probes its boundary twice with private
//    (tmin, tmax) and only needs the two distances -> `probe` mode, in which
//    hits update tmax only and the main record is untouched.  A volume's
//    boundary must not contain another volume (checked by the flattener).
// State of one scene.hit() call, kept by the caller so that the traversal can be suspended (COOP, see below).
struct Isect {
    float tmin0, tmin, tmax, main_tmax, vol_t1;
    uint32_t cur, sp0;
    bool ret, probe_v;
};
MRT_HD void isect_begin(Isect &s, const SceneView &sc, float tmin0, float tmax0, const Stack &st) {
    s.tmin0 = tmin0; s.tmin = tmin0; s.tmax = tmax0; s.main_tmax = tmax0; s.vol_t1 = 0.0f;
    s.cur = sc.root; s.sp0 = st.sp; s.ret = false; s.probe_v = false;
}
// Runs the traversal.  Returns true when it is finished (s.ret = hit, rec = record).
// COOP = true (warp-cooperative tree traversal, coop_tree.cuh): the per-lane machine does not enter BVH trees.  When it
// stands at the root of a tree whose own box the ray hits it returns FALSE with s.cur = the tree's root child; the
// caller traverses the tree, stores the outcome in s.ret / rec (and s.tmax = rec.t on a hit) and calls again with
// resume = true, which continues with the frame on top of the stack exactly as if the visit had returned.
template <bool COOP>
MRT_HD bool isect_run(const uint32_t feat, const SceneView &sc, Ray &ray, Isect &s, Hit &rec, Rng &rng, Stack &st, bool resume, Counters *cnt) {
    const float tmin0 = s.tmin0;
    float tmin = s.tmin, tmax = s.tmax;
    float main_tmax = s.main_tmax;
    bool probe_v = s.probe_v;
#define probe (MRT_HAS(feat, MRT_FEAT_VOLUMES) && probe_v)
    float vol_t1 = s.vol_t1;
    uint32_t cur = s.cur;
    bool ret = s.ret;
    const uint32_t sp0 = s.sp0;
    for (;;) {
        // ------------------------------------------------------------ visit
        bool descend = !resume;
        resume = false;
        while (descend) {
            descend = false;
            const uint32_t type = MRT_REF_TYPE(cur), idx = MRT_REF_INDEX(cur);
            switch (type) {
            case MRT_T_LIST: {
                MrtF4 l0 = ld4(sc.list, 2 * idx), l1 = ld4(sc.list, 2 * idx + 1);
                ret = false;
                if (f2u(l1.w) >> 31) {
                    if (cnt) cnt->aabb++;
                    if (!aabb_hit(l0, l1, ray, tmin, tmax)) break;
                }
                st.push(MRT_FRAME(MRT_F_LIST, f2u(l0.w)));
                break;   // the LIST frame is popped right away and runs the child loop
            }
            case MRT_T_BVH: if (!MRT_HAS(feat, MRT_FEAT_TREES)) { ret = false; break; } {   // root of a bvh_node / pod_bvh tree: its own box test (scene_object.h:211, triangle.h:175)
                MrtF4 b0 = ld4(sc.bvh, 2 * idx), b1 = ld4(sc.bvh, 2 * idx + 1);
                if (cnt) cnt->aabb++;
                ret = false;
                if (!aabb_hit(b0, b1, ray, tmin, tmax)) break;
                cur = f2u(b0.w);
                if (COOP) {   // suspend: the warp traverses the tree together
                    s.tmin = tmin; s.tmax = tmax; s.main_tmax = main_tmax; s.probe_v = probe_v; s.vol_t1 = vol_t1; s.cur = cur; s.ret = false;
                    return false;
                }
                descend = true;
                break;
            }
            case MRT_T_NODE2: if (COOP || !MRT_HAS(feat, MRT_FEAT_TREES)) { ret = false; break; } {
                // Inner node carrying both children's boxes.  Visit the closer child (node_order & dirMask,
                // scene_object.h:224-231) if its box is hit; the farther one only if the closer reports no hit
                // (IF_MISS frame).  tmin/tmax are constant inside a tree, so the child box tests made here are
                // the ones the children would make on entry.
                ret = false;
                for (;;) {
                    const uint32_t ni = MRT_REF_INDEX(cur);
                    MrtF4 n0 = ld4(sc.node2, 4 * ni), n1 = ld4(sc.node2, 4 * ni + 1);
                    MrtF4 n2 = ld4(sc.node2, 4 * ni + 2), n3 = ld4(sc.node2, 4 * ni + 3);
                    const uint32_t w0 = f2u(n0.w), w1 = f2u(n1.w), flags = f2u(n2.w);
                    const uint32_t order = (w0 >> 28) | ((w1 >> 28) << 4);
                    const uint32_t left = w0 & 0x0FFFFFFFu, right = w1 & 0x0FFFFFFFu;
                    const bool hl = !(flags & 1u) || aabb_hit(n0, n1, ray, tmin, tmax);
                    const bool hr = !(flags & 2u) || aabb_hit(n2, n3, ray, tmin, tmax);
                    const bool lfirst = (order & ray.mask) != 0;
                    const bool h_first = lfirst ? hl : hr, h_second = lfirst ? hr : hl;
                    const uint32_t first = lfirst ? left : right, second = lfirst ? right : left;
                    if (cnt) cnt->aabb += ((lfirst ? flags : flags >> 1) & 1u);   // the closer child's own test
                    if (h_first) {
                        if (h_second) st.push(MRT_FRAME(MRT_F_IF_MISS, second));
                        else if (cnt) cnt->aabb += ((lfirst ? flags >> 1 : flags) & 1u);   // farther child would be entered and miss its box
                        cur = first;
                    } else if (h_second) {
                        if (cnt) cnt->aabb += ((lfirst ? flags >> 1 : flags) & 1u);
                        cur = second;
                    } else {
                        if (cnt) cnt->aabb += ((lfirst ? flags >> 1 : flags) & 1u);
                        break;
                    }
                    if (MRT_REF_TYPE(cur) != MRT_T_NODE2) { descend = true; break; }
                }
                break;
            }
            case MRT_T_TRILEAF: if (COOP || !MRT_HAS(feat, MRT_FEAT_TREES) || !MRT_HAS(feat, MRT_FEAT_TRIS)) { ret = false; break; } {   // pod_bvh leaf: closest hit among its triangles (triangle.h:179-187)
                const uint32_t first = ldu(sc.trileaf, 2 * idx), count = ldu(sc.trileaf, 2 * idx + 1);
                ret = false;
                for (uint32_t i = 0; i < count; i++) {
                    if (cnt) cnt->tri++;
                    if (hit_triangle(sc, first + i, ray, tmin, tmax, !probe, rec)) {
                        ret = true;
                        tmax = rec.t;
                    }
                }
                break;
            }
            case MRT_T_TRANSLATE:     // scene_object.cpp:9-18
            case MRT_T_ROTATE_Y: if (!MRT_HAS(feat, MRT_FEAT_XFORM)) { ret = false; break; } {    // scene_object.cpp:70-98
                ret = false;
                MrtF4 r0, r2;
                if (type == MRT_T_ROTATE_Y) {
                    r0 = ld4(sc.rot, 3 * idx);
                    MrtF4 r1 = ld4(sc.rot, 3 * idx + 1);
                    if (f2u(r1.w)) {   // bbox pre-test
                        if (cnt) cnt->aabb++;
                        if (!aabb_hit(r0, r1, ray, tmin, tmax)) break;
                    }
                    r2 = ld4(sc.rot, 3 * idx + 2);
                } else {
                    r0 = ld4(sc.xlate, 3 * idx);
                    MrtF4 c0 = ld4(sc.xlate, 3 * idx + 1);
                    if (f2u(c0.w) && cull_miss(c0, ld4(sc.xlate, 3 * idx + 2), ray, tmin, tmax)) break;
                    r2 = r0;
                }
                if (cnt) cnt->xform++;
                if (type == MRT_T_ROTATE_Y) MRT_OP(rotate); else MRT_OP(translate);
                st.pushf(ray.o.x); st.pushf(ray.o.y); st.pushf(ray.o.z);
                st.pushf(ray.d.x); st.pushf(ray.d.y); st.pushf(ray.d.z);
                st.pushf(ray.inv.x); st.pushf(ray.inv.y); st.pushf(ray.inv.z);
                st.push((uint32_t) ray.inside);
                st.push(MRT_FRAME(type == MRT_T_ROTATE_Y ? MRT_F_ROT_END : MRT_F_XLATE_END, idx));
                V3 d = ray.d;
                if (type == MRT_T_ROTATE_Y) {
                    const float sin_t = r2.x, cos_t = r2.y;
                    V3 o = ray.o;
                    o.x = cos_t * ray.o.x - sin_t * ray.o.z;
                    o.z = cos_t * ray.o.z + sin_t * ray.o.x;
                    d.x = cos_t * ray.d.x - sin_t * ray.d.z;
                    d.z = cos_t * ray.d.z + sin_t * ray.d.x;
                    ray.o = o;
                } else {
                    ray.o = ray.o - v3(r0);
                }
                ray.inside = 0;
                ray_set_dir(ray, d);   // the ray ctor re-normalises (ray.h:30); isInside resets to 0
                cur = f2u(r0.w);
                descend = true;
                break;
            }
            case MRT_T_VOLUME: if (!MRT_HAS(feat, MRT_FEAT_VOLUMES)) { ret = false; break; } {      // volumes.cpp:5-36, first probe
                MrtF4 vl = ld4(sc.vol, idx);
                if (cnt) cnt->vol++;
                st.push(MRT_FRAME(MRT_F_VOL1, idx));
                main_tmax = tmax;
                probe_v = true;
                tmin = -FLT_MAX;
                tmax = FLT_MAX;
                cur = f2u(vl.x);
                descend = true;
                break;
            }
            default:
                ret = false;
                break;
            }
        }

        // ----------------------------------------------------------- return
        for (;;) {
            if (st.sp == sp0) { s.ret = ret && !probe; return true; }
            uint32_t e = st.pop();
            uint32_t tag = e >> 29;
            if (!COOP && MRT_HAS(feat, MRT_FEAT_TREES) && tag == MRT_F_IF_MISS) {
                if (ret) continue;
                cur = e & 0x0FFFFFFFu;
                break;
            } else if (tag == MRT_F_LIST) {
                // closest-hit loop over the children (scene_object.h:88-95); primitives inline
                uint32_t ci = e & 0x0FFFFFFFu;
                bool found = ((e >> 28) & 1u) | (ret ? 1u : 0u);
                bool composite = false;
                for (;;) {
                    uint32_t c = ldu(sc.child, ci);
                    uint32_t ctype = MRT_REF_TYPE(c);
                    if (ctype == MRT_T_END) break;
                    ci++;
                    if (MRT_HAS(feat, MRT_FEAT_SPHERES) && ctype == MRT_T_SPHERE) {
                        if (cnt) cnt->sphere++;
                        if (hit_sphere(feat, sc, MRT_REF_INDEX(c), ray, tmin, tmax, !probe, rec)) { found = true; tmax = rec.t; }
                    } else if (ctype <= MRT_T_RECT_YZ) {
                        if (cnt) cnt->rect++;
                        if (hit_rect(feat, sc, ctype - MRT_T_RECT_XY, MRT_REF_INDEX(c), ray, tmin, tmax, !probe, rec)) { found = true; tmax = rec.t; }
                    } else if (MRT_HAS(feat, MRT_FEAT_TRI_OBJECT) && ctype == MRT_T_TRI) {   // triangle_scene_object::hit (triangle.cpp:49-91) = triangle::hit
                        if (cnt) cnt->tri++;
                        if (hit_triangle(sc, MRT_REF_INDEX(c), ray, tmin, tmax, !probe, rec)) { found = true; tmax = rec.t; }
                    } else {
                        st.push(MRT_FRAME(MRT_F_LIST, ci) | (found ? (1u << 28) : 0u));
                        cur = c;
                        composite = true;
                        break;
                    }
                }
                if (composite) break;
                ret = found;
                continue;
            } else if (MRT_HAS(feat, MRT_FEAT_XFORM) && (tag == MRT_F_XLATE_END || tag == MRT_F_ROT_END)) {
                ray.inside = (int) st.pop();
                ray.inv.z = st.popf(); ray.inv.y = st.popf(); ray.inv.x = st.popf();
                ray.d.z = st.popf(); ray.d.y = st.popf(); ray.d.x = st.popf();
                ray.o.z = st.popf(); ray.o.y = st.popf(); ray.o.x = st.popf();
                ray.mask = dir_mask(ray.d);
                if (ret && !probe) {
                    uint32_t idx = e & 0x0FFFFFFFu;
                    if (tag == MRT_F_XLATE_END) {
                        rec.p = rec.p + v3(ld4(sc.xlate, 3 * idx));
                    } else {
                        MrtF4 r2 = ld4(sc.rot, 3 * idx + 2);
                        float sin_t = r2.x, cos_t = r2.y;
                        V3 p = rec.p, n = rec.n;
                        p.x = cos_t * rec.p.x + sin_t * rec.p.z;
                        p.z = cos_t * rec.p.z - sin_t * rec.p.x;
                        n.x = cos_t * rec.n.x + sin_t * rec.n.z;
                        n.z = cos_t * rec.n.z - sin_t * rec.n.x;
                        rec.p = p;
                        rec.n = n;
                    }
                }
                continue;
            } else if (MRT_HAS(feat, MRT_FEAT_VOLUMES) && tag == MRT_F_VOL1) {
                uint32_t idx = e & 0x0FFFFFFFu;
                if (!ret) {
                    probe_v = false; tmin = tmin0; tmax = main_tmax;
                    continue;
                }
                vol_t1 = tmax;   // rec1.t
                st.push(MRT_FRAME(MRT_F_VOL2, idx));
                tmin = vol_t1 + 0.0001f;
                tmax = FLT_MAX;
                cur = f2u(ld4(sc.vol, idx).x);
                break;
            } else if (MRT_HAS(feat, MRT_FEAT_VOLUMES)) {   // MRT_F_VOL2
                uint32_t idx = e & 0x0FFFFFFFu;
                float t2 = tmax;
                bool both = ret;
                probe_v = false; tmin = tmin0; tmax = main_tmax;
                ret = false;
                if (!both) continue;
                float t1 = vol_t1;
                if (t1 < tmin) t1 = tmin;
                if (t2 > tmax) t2 = tmax;
                if (t1 >= t2) continue;
                if (t1 < 0) t1 = 0;
                MrtF4 vl = ld4(sc.vol, idx);
                float inside_dist = (t2 - t1);
                float hit_dist = -(frcp(vl.y)) * cr_logf(randf(rng));
                if (hit_dist < inside_dist) {
                    rec.t = t1 + hit_dist;
                    rec.p = ray_eval(ray, rec.t);
                    rec.n = v3(1, 0, 0);
                    rec.mat = f2u(vl.z);
                    tmax = rec.t;
                    ret = true;
                }
                continue;
            }
        }
    }
}
#undef probe

MRT_HD bool intersect(const uint32_t feat, const SceneView &sc, Ray &ray, float tmin0, float tmax0, Hit &rec, Rng &rng, Stack &st,
                      Counters *cnt) {
    Isect s;
    isect_begin(s, sc, tmin0, tmax0, st);
    isect_run<false>(feat, sc, ray, s, rec, rng, st, false, cnt);
    return s.ret;
}

// ------------------------------------------------------------------- textures
// texture.cpp:68-165
MRT_HD float perlin_noise(const SceneView &sc, V3 p) {
    MRT_OP(perlin);
    float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    float u = p.x - fx, v = p.y - fy, w = p.z - fz;
    int i = (int) fx, j = (int) fy, k = (int) fz;
    int x0 = sc.perlin_perm[(i + 0) & 255], x1 = sc.perlin_perm[(i + 1) & 255];
    int y0 = sc.perlin_perm[256 + ((j + 0) & 255)], y1 = sc.perlin_perm[256 + ((j + 1) & 255)];
    int z0 = sc.perlin_perm[512 + ((k + 0) & 255)], z1 = sc.perlin_perm[512 + ((k + 1) & 255)];
    float uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w);
    float acc = 0;
#define MRT_PERLIN_CORNER(XI, YI, ZI, DI, DJ, DK)                                            \
    {                                                                                         \
        V3 c = v3(ld4(sc.perlin_vec, (uint32_t) ((XI) ^ (YI) ^ (ZI))));                       \
        V3 wt = v3(u - (DI), v - (DJ), w - (DK));                                             \
        float ax = (DI) ? uu : (1 - uu), ay = (DJ) ? vv : (1 - vv), az = (DK) ? ww : (1 - ww); \
        acc += ax * ay * az * dot(c, wt);                                                     \
    }
    MRT_PERLIN_CORNER(x0, y0, z0, 0, 0, 0)
    MRT_PERLIN_CORNER(x0, y0, z1, 0, 0, 1)
    MRT_PERLIN_CORNER(x0, y1, z0, 0, 1, 0)
    MRT_PERLIN_CORNER(x0, y1, z1, 0, 1, 1)
    MRT_PERLIN_CORNER(x1, y0, z0, 1, 0, 0)
    MRT_PERLIN_CORNER(x1, y0, z1, 1, 0, 1)
    MRT_PERLIN_CORNER(x1, y1, z0, 1, 1, 0)
    MRT_PERLIN_CORNER(x1, y1, z1, 1, 1, 1)
#undef MRT_PERLIN_CORNER
    return acc;
}
MRT_FN float perlin_turbulence(const SceneView &sc, V3 p) {   // depth 7, texture.cpp:153-165
    float acc = 0;
    float weight = 1.0f;
    for (int i = 0; i < 7; i++) {
        acc += weight * perlin_noise(sc, p);
        weight *= 0.5f;
        p = p * 2.0f;
    }
    return fabsf(acc);
}

MRT_FN V3 tex_sample(const uint32_t feat, const SceneView &sc, uint32_t tex, float u, float v, V3 p) {
    for (;;) {
        MrtF4 t = ld4(sc.tex, tex);
        uint32_t kind = f2u(t.x);
        if (!MRT_HAS(feat, MRT_FEAT_TEX) || kind == MRT_X_COLOR) return v3(t.y, t.z, t.w);   // color_tex only
        if (kind == MRT_X_CHECKER) {   // texture.cpp:7-13
            MRT_OP(checker);
            float s = t.w;
            float sines = 1.0f;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
            for (int a = 0; a < 3; a++) {   // one copy of the double-precision sine body; (sx * sy) * sz
                float v = cr_sinf(s * (a == 0 ? p.x : (a == 1 ? p.y : p.z)));
                sines = (a == 0) ? v : sines * v;
            }
            tex = (sines < 0) ? f2u(t.z) : f2u(t.y);
            continue;
        }
        if (kind == MRT_X_PERLIN) {    // texture.h:56-59
            float turb = perlin_turbulence(sc, p * t.y);
            return v3(1, 1, 1) * turb;
        }
        // image, texture.cpp:207-225
        MRT_OP(image);
        int32_t width = (int32_t) f2u(t.y), height = (int32_t) f2u(t.z);
        int32_t i = (int32_t) (u * width);
        int32_t j = (int32_t) ((1 - v) * height);
        i = i < 0 ? 0 : (i > width - 1 ? width - 1 : i);
        j = j < 0 ? 0 : (j > height - 1 ? height - 1 : j);
        const uint8_t *px = sc.image + f2u(t.w) + ((size_t) i + (size_t) width * j) * 3;
        const float f = (1.0f / 255.0f);
        return v3((float) px[0], (float) px[1], (float) px[2]) * f;
    }
}

// ----------------------------------------------------------------- light pdfs
// object_list::pdf_value / pdf_generate over scene.biased_objects
// (scene_object.h:64-77); sphere (sphere.cpp:63-79) and xz_rect (rect.cpp:92-107)
// have pdfs, every other object the base-class defaults (scene_object.h:24-29).
MRT_FN float light_pdf_value(const uint32_t feat, const SceneView &sc, V3 origin, V3 dir, float time) {
    MRT_OP(lightpdf);
    float sum = 0;
    // every light builds the same probe ray (ray ctor: normalise once more, ray.h:30) -- hoisted out of the loop
    Ray r = make_probe_ray(origin, dir, time);
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (uint32_t i = 0; i < sc.n_lights; i++) {
        uint32_t l = ldu(sc.lights, i);
        uint32_t type = MRT_REF_TYPE(l), idx = MRT_REF_INDEX(l);
        float pv = 0;
        Hit rec;
        if (type == MRT_T_RECT_XZ) {
            if (hit_rect(MRT_FEAT_ALL, sc, 1, idx, r, 0.001f, FLT_MAX, false, rec)) {
                MrtF4 q0 = ld4(sc.rect, 2 * idx), q1 = ld4(sc.rect, 2 * idx + 1);
                float area = (q0.y - q0.x) * (q0.w - q0.z);
                float dist_sq = rec.t * rec.t;
                float cosine = fabsf(dot(dir, v3(0, q1.y, 0)));
                pv = fdiv(dist_sq, (cosine * area));
            }
        } else if (MRT_HAS(feat, MRT_FEAT_LIGHT_SPHERE) && MRT_HAS(feat, MRT_FEAT_SPHERES) && type == MRT_T_SPHERE) {
            if (hit_sphere(feat, sc, idx, r, 0.001f, FLT_MAX, false, rec)) {
                MrtF4 s0 = ld4(sc.sphere, 3 * idx), s1 = ld4(sc.sphere, 3 * idx + 1);
                V3 cen = sphere_center(s0, s1, sc.sphere, idx, time);
                float cos_theta_max = fsqrt(1 - fdiv(s0.w * s0.w, sdot(cen - origin)));
                float solid_angle = 2 * MRT_PI_F * (1 - cos_theta_max);
                pv = frcp(solid_angle);
            }
        }
        sum += pv;
    }
    return fdiv_zero_num(sum, (float) sc.n_lights, true);
}

struct Onb { V3 u, v, w; };   // onb.h:19-27
MRT_HD Onb make_onb(V3 n) {
    Onb o;
    o.w = n;
    V3 a = (fabsf(n.x) > 0.9f) ? v3(0, 1, 0) : v3(1, 0, 0);
    V3 c = cross(o.w, a);
    float len = fsqrt(sdot(c));
    bool ok = (len > 0.0f) && (len < INFINITY);
    o.v = v3(fdiv_zero_num(c.x, len, ok), fdiv_zero_num(c.y, len, ok), fdiv_zero_num(c.z, len, ok));   // normalize(c)
    o.u = cross(o.w, o.v);
    return o;
}
MRT_HD V3 onb_local(const Onb &o, V3 a) { return (a.x * o.u + a.y * o.v) + a.z * o.w; }

MRT_FN V3 light_pdf_generate(const uint32_t feat, const SceneView &sc, V3 origin, float time, Rng &rng) {
    int i = (int) (randf(rng) * (float) sc.n_lights);
    uint32_t l = ldu(sc.lights, (uint32_t) i);
    uint32_t type = MRT_REF_TYPE(l), idx = MRT_REF_INDEX(l);
    if (type == MRT_T_RECT_XZ) {
        MrtF4 q0 = ld4(sc.rect, 2 * idx), q1 = ld4(sc.rect, 2 * idx + 1);
        float rx = randf(rng), rz = randf(rng);
        V3 rnd = v3(q0.x + rx * (q0.y - q0.x), q1.x, q0.z + rz * (q0.w - q0.z));
        return rnd - origin;
    } else if (MRT_HAS(feat, MRT_FEAT_LIGHT_SPHERE) && MRT_HAS(feat, MRT_FEAT_SPHERES) && type == MRT_T_SPHERE) {
        MrtF4 s0 = ld4(sc.sphere, 3 * idx), s1 = ld4(sc.sphere, 3 * idx + 1);
        V3 dir = sphere_center(s0, s1, sc.sphere, idx, time) - origin;
        float dist_sq = sdot(dir);
        Onb uvw = make_onb(normalize(dir));
        return onb_local(uvw, random_towards_sphere(s0.w, dist_sq, rng));
    }
    return v3(1, 0, 0);
}

// ---------------------------------------------------------------------- trace
// One path = the reference's recursive trace() (main.cpp:66-118) unrolled:
//   L = e0 + w0 (e1 + w1 (...))   ->   L += T * e_k ; T *= w_k
// with w = attenuation * scattering_pdf / pdf_v (main.cpp:102) or attenuation
// (specular, main.cpp:83; emitted light is dropped there, as in the reference).
//
// The per-segment work is arranged in phases so that every large piece of code has
// exactly one call site (instruction-cache footprint) and so that all lanes of a warp
// run the same phase at the same time:
//   A  path_begin    : camera ray of a new (pixel, sample)          -- raw direction
//   B  path_advance  : normalise the raw direction (THE ray constructor, ray.h:30)
//   C                : weight of the previous diffuse bounce: pdf_v, scattering_pdf (main.cpp:88-102)
//   D  intersect     : scene.hit
//   E  path_shade    : emission, material scatter -> raw direction of the next segment
struct Path {
    Ray ray;        // between E and B, ray.d holds the raw (un-normalised) direction
    V3 T, L;
    uint32_t depth;
    uint32_t pending;   // 0: nothing deferred; 1: lambertian bounce; 2: isotropic bounce
    V3 p_att, p_n;      // attenuation and surface normal of the deferred bounce
};

// A: camera::get_ray (camera.h:38-45) for sample s of pixel (x, y); regular sub-pixel grid (main.cpp:319-332,156-157)
MRT_HD void path_begin(const SceneView &sc, Path &p, Rng &rng, uint32_t x, uint32_t y, uint32_t s, uint32_t sqrt_n,
                       uint32_t width, uint32_t height, uint64_t seed) {
    MRT_OP(paths);
    uint64_t stream = ((uint64_t) y * width + x) * ((uint64_t) sqrt_n * sqrt_n) + s;
    rng_seed(rng, seed, stream);
    uint32_t i = s / sqrt_n, j = s - i * sqrt_n;
    float sx = fdiv(i + 0.5f, (float) sqrt_n);
    float sy = fdiv(j + 0.5f, (float) sqrt_n);
    float u = fdiv(x + sx, (float) width);
    float v = fdiv(y + sy, (float) height);
    const MrtCamera &c = sc.cam;
    V3 rd = c.lens_radius * random_in_disk(rng);
    V3 cu = v3(c.u[0], c.u[1], c.u[2]), cv = v3(c.v[0], c.v[1], c.v[2]);
    V3 offset = cu * rd.x + cv * rd.y;
    float time = c.time0 + (c.time1 - c.time0) * randf(rng);
    V3 origin = v3(c.origin[0], c.origin[1], c.origin[2]);
    V3 ll = v3(c.llcorner[0], c.llcorner[1], c.llcorner[2]);
    V3 horz = v3(c.horz[0], c.horz[1], c.horz[2]), vert = v3(c.vert[0], c.vert[1], c.vert[2]);
    p.ray.o = origin + offset;
    p.ray.d = (((ll + u * horz) + v * vert) - origin) - offset;
    p.ray.time = time;
    p.ray.inside = 0;
    p.T = v3(1, 1, 1);
    p.L = v3(0, 0, 0);
    p.depth = 0;
    p.pending = 0;
}

// B + C
MRT_HD void path_advance(const uint32_t feat, const SceneView &sc, Path &p) {
    ray_set_dir(p.ray, p.ray.d);
    if (p.pending) {
        const V3 d = p.ray.d;
        float mat_pdf, spdf;
        if (!MRT_HAS(feat, MRT_FEAT_VOLUMES) || p.pending == 1) {   // isotropic only exists as a volume's phase function
            float cosine = dot(d, p.p_n);                    // cosine_pdf::value (pdf.h:24-30); uvw.w == n
            mat_pdf = (cosine > 0) ? fdiv(cosine, MRT_PI_F) : 0.0f;
            spdf = (cosine < 0) ? 0.0f : cosine * (1.0f / MRT_PI_F);   // lambertian::scattering_pdf (material.h:40-46)
        } else {
            mat_pdf = frcp(2 * MRT_PI_F);               // isotropic_pdf::value (pdf.h:41-43)
            spdf = 1.0f / (2.0f * MRT_PI_F);                 // isotropic::scattering_pdf (material.h:64-66)
        }
        float pdf_v = mat_pdf;
        if (sc.n_lights) pdf_v = 0.5f * (light_pdf_value(feat, sc, p.ray.o, d, p.ray.time) + mat_pdf);   // mix_pdf::value
        const V3 num = p.p_att * spdf;
        const bool pdf_ok = (pdf_v > 0.0f) && (pdf_v < INFINITY);
        V3 w = v3(fdiv_zero_num(num.x, pdf_v, pdf_ok), fdiv_zero_num(num.y, pdf_v, pdf_ok), fdiv_zero_num(num.z, pdf_v, pdf_ok));
        p.T = p.T * w;
        p.pending = 0;
    }
}

// E: returns true if the path continues (p.ray then holds the next origin and the raw direction)
MRT_HD bool path_shade(const uint32_t feat, const SceneView &sc, Path &p, bool hit, const Hit &rec, uint32_t max_bounces, Rng &rng) {
    if (!hit) {   // main.cpp:108-117
        if (sc.sky) {
            MRT_OP(sky);
            float t = 0.5f * (p.ray.d.y + 1.0f);
            float a = 1.0f - t;
            V3 bg = v3(a, a, a) + t * v3(0.5f, 0.7f, 1.0f);
            p.L = p.L + p.T * bg;
        }
        return false;
    }
    MrtF4 m = ld4(sc.mat, rec.mat);
    const uint32_t kind = f2u(m.x) & 0xFFu;
    const Ray &r = p.ray;
    // a light emits only towards the side its normal faces and never scatters (material.h:190-199);
    // every other material emits nothing, so hitting the bounce limit ends the path with no contribution
    if (kind == MRT_M_LIGHT ? !(dot(rec.n, r.d) < 0.0f) : !(p.depth < max_bounces)) return false;
    V3 texv = v3(1, 1, 1);
    if (!MRT_HAS(feat, MRT_FEAT_DIELECTRIC) || kind != MRT_M_DIELECTRIC) texv = tex_sample(feat, sc, f2u(m.y), rec.u, rec.v, rec.p);   // albedo / emissive
    if (kind == MRT_M_LIGHT) {
        p.L = p.L + p.T * (m.z * texv);
        return false;
    }
    V3 dir;
    int inside = 0;
    const bool is_metal = MRT_HAS(feat, MRT_FEAT_METAL) && kind == MRT_M_METAL;
    const bool is_diel = MRT_HAS(feat, MRT_FEAT_DIELECTRIC) && kind == MRT_M_DIELECTRIC;
    if (is_metal || is_diel) {
        float dp = 2.0f * dot(r.d, rec.n);          // reflect(), vec3.h:178-181
        dir = r.d - (dp * rec.n);
        if (is_metal) {                             // material.h:84-98
            MRT_OP(metal);
            dir = dir + (1 - m.z) * random_in_sphere(rng);
            p.T = texv * p.T;
        } else {                                    // material.h:106-175
            MRT_OP(dielectric);
            float ref_index = m.z;
            V3 fn;
            float ni_over_nt;
            float cosI = -dot(r.d, rec.n);
            if (cosI < 0) { fn = neg(rec.n); ni_over_nt = ref_index; }
            else          { fn = rec.n;      ni_over_nt = frcp(ref_index); }
            float ncosI = dot(r.d, fn);             // refract(), vec3.h:185-198
            float sinT2 = (ni_over_nt * ni_over_nt) * (1.0f - ncosI * ncosI);
            inside = r.inside;
            if (sinT2 <= 1.0f) {
                float cosT = fsqrt(1.0f - sinT2);
                float k = ni_over_nt * (-ncosI) - cosT;
                float cosine_schlick;
                if (cosI < 0) cosine_schlick = fsqrt(1.0f - ni_over_nt * ni_over_nt * (1.0f - cosI * cosI));
                else          cosine_schlick = cosI;
                float r0 = fdiv(1 - ref_index, 1 + ref_index);   // fresnel_schlick, material.h:106-110
                r0 = r0 * r0;
                float reflect_prob = r0 + (1 - r0) * cr_pow5f((1 - cosine_schlick));
                if (!(randf(rng) < reflect_prob)) {
                    if (cosI < 0) { inside--; if (inside < 0) inside = 0; }
                    else          { inside++; }
                    dir = ni_over_nt * r.d + k * fn;
                }
            }
        }
    } else {
        // lambertian (material.h:40-53) / isotropic (material.h:64-73): direction from the mixture pdf
        // (main.cpp:84-92); its weight needs the NORMALISED direction and is applied in path_advance
        const bool lambert = !MRT_HAS(feat, MRT_FEAT_VOLUMES) || (kind == MRT_M_LAMBERTIAN);
        if (lambert) MRT_OP(lambert); else MRT_OP(isotropic);
        bool use_light = false;
        if (sc.n_lights) use_light = randf(rng) < 0.5f;   // mix_pdf::generate, pdf.h:74-79
        if (use_light) {
            dir = light_pdf_generate(feat, sc, rec.p, r.time, rng);
        } else if (lambert) {
            Onb uvw = make_onb(rec.n);
            dir = onb_local(uvw, random_cosine_direction(rng));
        } else {
            dir = random_in_sphere(rng);
        }
        p.pending = lambert ? 1u : 2u;
        p.p_att = texv;
        p.p_n = rec.n;
    }
    p.ray.o = rec.p;
    p.ray.d = dir;
    p.ray.inside = inside;
    p.depth++;
    return true;
}

// A sample counts only if it is finite (main.cpp:163-165).  In the reference's nested
// evaluation a non-finite weight (e.g. 0/0 when scattering_pdf == pdf_v == 0) poisons
// the whole sample even if nothing is emitted further down the path (0 * NaN = NaN), so
// the throughput is checked together with the radiance.
MRT_HD bool path_sample_finite(const Path &p) {
    return is_finite(p.L.x) && is_finite(p.L.y) && is_finite(p.L.z) && is_finite(p.T.x) && is_finite(p.T.y) && is_finite(p.T.z);
}

}  // namespace mrt
