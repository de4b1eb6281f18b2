/* TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.
 *
 * mrt_oracle.c -- plain-C CPU restatement of the reference's per-pixel path-tracing bounce loop
 * (Maraneshi/MiniRayTracer).  It is independent of the product code: it shares no source with
 * miniraytracer_b200/, keeps the reference's RECURSIVE structure (virtual hit() -> switch on a tag,
 * recursive trace()), and takes its scene from the canonical text dump that the reference itself prints
 * (`oracle/_ref/mrt_ref dump-scene`, format in ref_harness.cpp), so the only things it restates are the
 * functions of SURVEY.md section 8(a).  Every function cites the reference file:line it follows.
 *
 * Pinning: tests/test_oracle_restatement.py checks it against oracle/_ref/mrt_ref (the reference's own
 * code) on all nine scenes: identical trace() counts, identical finite-sample counts, radiance equal up to
 * nothing at all (the same float operations in the same order; compiled with -ffp-contract=off).
 *
 * Conventions shared with the reference harness: one PCG32 stream per (pixel, sample),
 * pcg32_srandom(seed, (y*W+x)*N+s); argument lists that draw random numbers are evaluated left to right;
 * sinf cosf atan2f asinf logf powf(.,5) are the correctly rounded values (double evaluation rounded once,
 * see miniraytracer_b200/csrc/mrt_libm.h for the rationale).
 *
 * usage: mrt_oracle -dump scene.txt -width W -height H -samples N -depth D -seed X [-s0 a -s1 b] [-x0 a -x1 b -y0 c -y1 d]
 *                   [-sky 0|1] [-image earthmap.ppm] [-threads T] -out acc.bin
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ Vec3 (vec3.h) */
typedef struct { float x, y, z; } V3;
static V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
static V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static V3 mulv(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static V3 muls(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static V3 divs(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
static float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }          /* vec3.h:245-248 */
static V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } /* vec3.h:250-266 */
static V3 normalize(V3 a) { return divs(a, sqrtf(dot(a, a))); }                     /* vec3.h:133-139 */
static float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
#define PI_F 3.14159265358979323846f

/* canonical libm (see header) */
static float cr_sinf(float x) { return (float) sin((double) x); }
static float cr_cosf(float x) { return (float) cos((double) x); }
static float cr_logf(float x) { return (float) log((double) x); }
static float cr_atan2f(float y, float x) { return (float) atan2((double) y, (double) x); }
static float cr_asinf(float x) { return (float) asin((double) x); }
static float cr_pow5f(float x) { double d = x, d2 = d * d, d4 = d2 * d2; return (float) (d4 * d); }

/* ------------------------------------------------------------------ PCG32 (pcg.cpp:13-62) */
typedef struct { uint64_t state, inc; } Rng;
static uint32_t pcg32(Rng *r) {
    uint64_t old = r->state;
    r->state = old * 6364136223846793005ULL + r->inc;
    uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
}
static void pcg32_seed(Rng *r, uint64_t initstate, uint64_t initseq) {
    r->state = 0u; r->inc = (initseq << 1u) | 1u; pcg32(r); r->state += initstate; pcg32(r);
}
static float randf(Rng *r) { union { float f; uint32_t u; } a; a.u = 0x3f800000u | (pcg32(r) & 0x007FFFFFu); return a.f - 1.0f; }
static V3 random_in_sphere(Rng *r) {   /* pcg.cpp:70-77 */
    V3 p;
    do { float x = randf(r), y = randf(r), z = randf(r); p = v3(2.0f * x - 1.0f, 2.0f * y - 1.0f, 2.0f * z - 1.0f); } while (dot(p, p) >= 1.0f);
    return p;
}
static V3 random_in_disk(Rng *r) {     /* pcg.cpp:112-119 */
    V3 p;
    do { float x = randf(r), y = randf(r); p = v3(2.0f * x - 1.0f, 2.0f * y - 1.0f, 0.0f); } while (dot(p, p) >= 1.0f);
    return p;
}
static V3 random_cosine_direction(Rng *r) {   /* pcg.cpp:87-95 */
    float r1 = randf(r), r2 = randf(r);
    float z = sqrtf(1 - r2);
    float phi = 2 * PI_F * r1;
    float x = cr_cosf(phi) * 2 * sqrtf(r2);
    float y = cr_sinf(phi) * 2 * sqrtf(r2);
    return v3(x, y, z);
}
static V3 random_towards_sphere(float radius, float dist_sq, Rng *r) {   /* pcg.cpp:125-133 */
    float r1 = randf(r), r2 = randf(r);
    float z = 1 + r2 * (sqrtf(1 - radius * radius / dist_sq) - 1);
    float phi = 2 * PI_F * r1;
    float x = cr_cosf(phi) * sqrtf(1 - z * z);
    float y = cr_sinf(phi) * sqrtf(1 - z * z);
    return v3(x, y, z);
}

/* ------------------------------------------------------------------ scene structures */
enum { T_COLOR, T_CHECKER, T_PERLIN, T_IMAGE };
typedef struct Tex { int kind; V3 color; struct Tex *even, *odd; float scale; int w, h; const uint8_t *data; } Tex;
enum { M_LAMBERTIAN, M_ISOTROPIC, M_METAL, M_DIELECTRIC, M_LIGHT };
typedef struct Mat { int kind; Tex *tex; float param; } Mat;
typedef struct { V3 min, max; } Aabb;
enum { O_SPHERE, O_XY, O_XZ, O_YZ, O_BOX, O_LIST, O_BVH, O_TRANSLATE, O_ROTATE, O_VOLUME, O_POD };
typedef struct { V3 m, u, v, mn, un, vn; } Tri;
typedef struct { Aabb box; uint32_t left, off, cnt; uint32_t order; } PodNode;
typedef struct Obj {
    int kind;
    V3 c0, c1; float t0, t1, radius; int moving;          /* sphere.h:11-17 */
    float a0, a1, b0, b1, k, sign;                         /* rect.h */
    Mat *mat;
    Aabb box; int has_box;
    struct Obj **children; size_t n;                       /* object_list */
    struct Obj *left, *right; uint32_t order;              /* bvh_node */
    struct Obj *child;                                     /* box list / translate / rotate_y / volume boundary */
    V3 offset; float sin_t, cos_t, density;
    PodNode *nodes; Tri *tris; uint32_t n_nodes, n_tris;   /* pod_bvh<triangle> */
} Obj;
typedef struct { V3 origin, u, v, w, llcorner, horz, vert; float lens_radius, time0, time1; } Camera;
typedef struct { Obj *objects, *biased; Camera cam; int sky; } Scene;

typedef struct { V3 o, d; float time; int inside; uint32_t mask; } Ray;   /* ray.h:7-16 */
typedef struct { float t; V3 p, n; float u, v; Mat *mat; } HitRec;       /* scene_object.h:10-17 */

static Ray make_ray(V3 o, V3 dir, float time, int inside) {   /* ray.h:20-52 */
    Ray r; r.o = o; r.d = normalize(dir); r.time = time; r.inside = inside;
    union { float f; uint32_t u; } X, Y, Z; X.f = r.d.x; Y.f = r.d.y; Z.f = r.d.z;
    r.mask = 1u << ((Z.u >> 31) | ((Y.u >> 31) << 1) | ((X.u >> 31) << 2));
    return r;
}
static V3 ray_eval(const Ray *r, float t) { return add(r->o, muls(r->d, t)); }

/* aabb::hit, aabb.h:45-76: SSE min/max return the second operand if either is NaN; strict > */
static int aabb_hit(const Aabb *b, const Ray *r, float tmin, float tmax) {
    float ix = 1.0f / r->d.x, iy = 1.0f / r->d.y, iz = 1.0f / r->d.z;
    float t0x = (b->min.x - r->o.x) * ix, t1x = (b->max.x - r->o.x) * ix;
    float t0y = (b->min.y - r->o.y) * iy, t1y = (b->max.y - r->o.y) * iy;
    float t0z = (b->min.z - r->o.z) * iz, t1z = (b->max.z - r->o.z) * iz;
    float t;
    if (ix < 0.0f) { t = t0x; t0x = t1x; t1x = t; }
    if (iy < 0.0f) { t = t0y; t0y = t1y; t1y = t; }
    if (iz < 0.0f) { t = t0z; t0z = t1z; t1z = t; }
    float a0 = (t0x > t0z) ? t0x : t0z, a1 = (t0y > tmin) ? t0y : tmin, lo = (a0 > a1) ? a0 : a1;
    float b0 = (t1x < t1z) ? t1x : t1z, b1 = (t1y < tmax) ? t1y : tmax, hi = (b0 < b1) ? b0 : b1;
    return hi > lo;
}

/* ------------------------------------------------------------------ Perlin (texture.cpp:68-203) */
static float ranvec[256][3];
static int perm[3][256];
static void perlin_init(void) {
    Rng g = {11350390909718046443ULL, 6305599193148252115ULL};   /* raw G_rng, pcg.cpp:40 */
    for (int i = 0; i < 256; i++) { V3 v = random_in_sphere(&g); ranvec[i][0] = v.x; ranvec[i][1] = v.y; ranvec[i][2] = v.z; }
    for (int a = 0; a < 3; a++) {
        for (int i = 0; i < 256; i++) perm[a][i] = i;
        for (int i = 255; i > 0; i--) { int target = (int) (randf(&g) * (i + 1)); int tmp = perm[a][i]; perm[a][i] = perm[a][target]; perm[a][target] = tmp; }
    }
}
static float perlin_noise(V3 p) {   /* texture.cpp:114-151 + perlin_interp 68-105 */
    float u = p.x - floorf(p.x), v = p.y - floorf(p.y), w = p.z - floorf(p.z);
    int i = (int) floorf(p.x), j = (int) floorf(p.y), k = (int) floorf(p.z);
    float uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w);
    float acc = 0;
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++) {
                const float *c = ranvec[perm[0][(i + di) & 255] ^ perm[1][(j + dj) & 255] ^ perm[2][(k + dk) & 255]];
                V3 wt = v3(u - di, v - dj, w - dk);
                float ax = di ? uu : (1 - uu), ay = dj ? vv : (1 - vv), az = dk ? ww : (1 - ww);
                acc += ax * ay * az * dot(v3(c[0], c[1], c[2]), wt);
            }
    return acc;
}
static float perlin_turbulence(V3 p) {   /* texture.cpp:153-165, depth 7 */
    float acc = 0, weight = 1.0f;
    for (int i = 0; i < 7; i++) { acc += weight * perlin_noise(p); weight *= 0.5f; p = muls(p, 2); }
    return fabsf(acc);
}

static V3 tex_sample(const Tex *t, float u, float v, V3 p) {
    switch (t->kind) {
    case T_COLOR: return t->color;                                                    /* texture.h:18-20 */
    case T_CHECKER: {                                                                 /* texture.cpp:7-13 */
        float sines = cr_sinf(t->scale * p.x) * cr_sinf(t->scale * p.y) * cr_sinf(t->scale * p.z);
        return (sines < 0) ? tex_sample(t->odd, u, v, p) : tex_sample(t->even, u, v, p);
    }
    case T_PERLIN: return muls(v3(1, 1, 1), perlin_turbulence(muls(p, t->scale)));    /* texture.h:56-59 */
    default: {                                                                        /* texture.cpp:207-225 */
        int i = (int) (u * t->w), j = (int) ((1 - v) * t->h);
        i = i < 0 ? 0 : (i > t->w - 1 ? t->w - 1 : i);
        j = j < 0 ? 0 : (j > t->h - 1 ? t->h - 1 : j);
        const uint8_t *px = t->data + ((size_t) i + (size_t) t->w * j) * 3;
        return muls(v3((float) px[0], (float) px[1], (float) px[2]), 1.0f / 255.0f);
    }
    }
}

/* ------------------------------------------------------------------ hit() per object class */
static int obj_hit(const Obj *o, const Ray *r, float tmin, float tmax, HitRec *rec, Rng *rng);

static V3 sphere_center(const Obj *s, float time) {   /* sphere.h:24-31 */
    if (s->moving) return add(s->c0, muls(sub(s->c1, s->c0), (time - s->t0) / (s->t1 - s->t0)));
    return s->c0;
}
static int sphere_hit(const Obj *s, const Ray *r, float tmin, float tmax, HitRec *rec) {   /* sphere.cpp:13-46 */
    rec->mat = s->mat;
    V3 cen = sphere_center(s, r->time);
    V3 oc = sub(r->o, cen);
    float b = dot(oc, r->d), c = dot(oc, oc) - s->radius * s->radius, disc = b * b - c;
    if (disc > 0) {
        float t = (-b - sqrtf(disc));
        int ok = (t < tmax && t > tmin);
        if (!ok && r->inside) { t = (-b + sqrtf(disc)); ok = (t < tmax && t > tmin); }
        if (ok) {
            rec->t = t; rec->p = ray_eval(r, t); rec->n = divs(sub(rec->p, cen), s->radius);
            float phi = cr_atan2f(rec->n.z, rec->n.x), theta = cr_asinf(rec->n.y);   /* get_sphere_uv, sphere.cpp:6-11 */
            rec->u = 0.5f - phi * (1.0f / (2.0f * PI_F));
            rec->v = 0.5f + theta * (1.0f / PI_F);
            return 1;
        }
    }
    return 0;
}
static int rect_hit(const Obj *q, const Ray *r, float tmin, float tmax, HitRec *rec) {   /* rect.cpp:24-45,69-90,130-152 */
    float ok_, dk, oa, da, ob, db;
    if (q->kind == O_XY)      { ok_ = r->o.z; dk = r->d.z; oa = r->o.x; da = r->d.x; ob = r->o.y; db = r->d.y; }
    else if (q->kind == O_XZ) { ok_ = r->o.y; dk = r->d.y; oa = r->o.x; da = r->d.x; ob = r->o.z; db = r->d.z; }
    else                      { ok_ = r->o.x; dk = r->d.x; oa = r->o.y; da = r->d.y; ob = r->o.z; db = r->d.z; }
    if (dk * q->sign > 0.0f) return 0;
    float t = (q->k - ok_) / dk;
    if (t < tmin || t > tmax) return 0;
    float a = oa + t * da, b = ob + t * db;
    if (a < q->a0 || a > q->a1 || b < q->b0 || b > q->b1) return 0;
    rec->u = (a - q->a0) / (q->a1 - q->a0);
    rec->v = (b - q->b0) / (q->b1 - q->b0);
    rec->t = t; rec->mat = q->mat; rec->p = ray_eval(r, t);
    rec->n = q->kind == O_XY ? v3(0, 0, q->sign) : (q->kind == O_XZ ? v3(0, q->sign, 0) : v3(q->sign, 0, 0));
    return 1;
}
static int tri_hit(const Tri *tr, Mat *mat, const Ray *r, float tmin, float tmax, HitRec *rec) {   /* triangle.cpp:222-266 */
    V3 pvec = cross(r->d, tr->v);
    float det = dot(tr->u, pvec), sign = 1.0f;
    if (r->inside) { sign = det < 0.0f ? -1.0f : 1.0f; det = sign * det; }
    if (det < 0.00001f) return 0;
    V3 tvec = sub(r->o, tr->m);
    float uu = dot(tvec, pvec) * sign;
    V3 qvec = cross(tvec, tr->u);
    float vv = dot(r->d, qvec) * sign;
    if ((uu < 0) | (uu > det) | (vv < 0) | ((uu + vv) > det)) return 0;
    float invDet = 1 / det;
    float t = dot(tr->v, qvec) * invDet * sign;
    if ((t < tmin) | (t > tmax)) return 0;
    uu *= invDet; vv *= invDet;
    rec->t = t; rec->p = ray_eval(r, t);
    rec->n = normalize(add(add(muls(tr->mn, 1 - uu - vv), muls(tr->un, uu)), muls(tr->vn, vv)));
    rec->u = uu; rec->v = vv; rec->mat = mat;
    return 1;
}
static int pod_hit(const Obj *o, uint32_t ni, const Ray *r, float tmin, float tmax, HitRec *rec) {   /* triangle.h:171-213 */
    const PodNode *node = &o->nodes[ni];
    if (!aabb_hit(&node->box, r, tmin, tmax)) return 0;
    int has_hit = 0;
    if (node->cnt) {
        for (uint32_t i = 0; i < node->cnt; i++)
            if (tri_hit(&o->tris[node->off + i], o->mat, r, tmin, tmax, rec)) { has_hit = 1; tmax = rec->t; }
        return has_hit;
    }
    uint32_t closer, farther;
    if (node->order & r->mask) { closer = node->left; farther = node->left + 1; }
    else { closer = node->left + 1; farther = node->left; }
    if (pod_hit(o, closer, r, tmin, tmax, rec)) return 1;
    return pod_hit(o, farther, r, tmin, tmax, rec);
}
static int obj_hit(const Obj *o, const Ray *r, float tmin, float tmax, HitRec *rec, Rng *rng) {
    switch (o->kind) {
    case O_SPHERE: return sphere_hit(o, r, tmin, tmax, rec);
    case O_XY: case O_XZ: case O_YZ: return rect_hit(o, r, tmin, tmax, rec);
    case O_BOX: return obj_hit(o->child, r, tmin, tmax, rec, rng);   /* box.h:23-25 */
    case O_LIST: {   /* object_list::hit, scene_object.h:79-103 */
        if (!o->has_box || aabb_hit(&o->box, r, tmin, tmax)) {
            HitRec cur; int hit = 0; float closest = tmax;
            for (size_t i = 0; i < o->n; i++)
                if (obj_hit(o->children[i], r, tmin, closest, &cur, rng)) { hit = 1; closest = cur.t; *rec = cur; }
            return hit;
        }
        return 0;
    }
    case O_BVH: {    /* bvh_node::hit, scene_object.h:208-244 */
        if (!aabb_hit(&o->box, r, tmin, tmax)) return 0;
        const Obj *closer = (o->order & r->mask) ? o->left : o->right;
        const Obj *farther = (o->order & r->mask) ? o->right : o->left;
        if (obj_hit(closer, r, tmin, tmax, rec, rng)) return 1;
        return obj_hit(farther, r, tmin, tmax, rec, rng);
    }
    case O_TRANSLATE: {   /* scene_object.cpp:9-18 */
        Ray moved = make_ray(sub(r->o, o->offset), r->d, r->time, 0);
        if (obj_hit(o->child, &moved, tmin, tmax, rec, rng)) { rec->p = add(rec->p, o->offset); return 1; }
        return 0;
    }
    case O_ROTATE: {      /* scene_object.cpp:70-98 */
        if (o->has_box && !aabb_hit(&o->box, r, tmin, tmax)) return 0;
        V3 origin = r->o, dir = r->d;
        origin.x = o->cos_t * r->o.x - o->sin_t * r->o.z;
        origin.z = o->cos_t * r->o.z + o->sin_t * r->o.x;
        dir.x = o->cos_t * r->d.x - o->sin_t * r->d.z;
        dir.z = o->cos_t * r->d.z + o->sin_t * r->d.x;
        Ray rot = make_ray(origin, dir, r->time, 0);
        if (obj_hit(o->child, &rot, tmin, tmax, rec, rng)) {
            V3 p = rec->p, n = rec->n;
            p.x = o->cos_t * rec->p.x + o->sin_t * rec->p.z;
            p.z = o->cos_t * rec->p.z - o->sin_t * rec->p.x;
            n.x = o->cos_t * rec->n.x + o->sin_t * rec->n.z;
            n.z = o->cos_t * rec->n.z - o->sin_t * rec->n.x;
            rec->p = p; rec->n = n;
            return 1;
        }
        return 0;
    }
    case O_VOLUME: {      /* constant_volume::hit, volumes.cpp:5-36 */
        HitRec rec1, rec2;
        if (obj_hit(o->child, r, -FLT_MAX, FLT_MAX, &rec1, rng)) {
            if (obj_hit(o->child, r, rec1.t + 0.0001f, FLT_MAX, &rec2, rng)) {
                if (rec1.t < tmin) rec1.t = tmin;
                if (rec2.t > tmax) rec2.t = tmax;
                if (rec1.t >= rec2.t) return 0;
                if (rec1.t < 0) rec1.t = 0;
                float inside_dist = (rec2.t - rec1.t);
                float hit_dist = -(1 / o->density) * cr_logf(randf(rng));
                if (hit_dist < inside_dist) {
                    rec->t = rec1.t + hit_dist; rec->p = ray_eval(r, rec->t); rec->n = v3(1, 0, 0); rec->mat = o->mat;
                    return 1;
                }
            }
        }
        return 0;
    }
    default: return pod_hit(o, 0, r, tmin, tmax, rec);
    }
}

/* ------------------------------------------------------------------ light pdfs (scene_object.h:64-77) */
static float obj_pdf_value(const Obj *o, V3 origin, V3 dir, float time, Rng *rng) {
    if (o->kind == O_LIST) {
        float sum = 0;
        for (size_t i = 0; i < o->n; i++) sum += obj_pdf_value(o->children[i], origin, dir, time, rng);
        return sum / o->n;
    }
    HitRec rec;
    if (o->kind == O_XZ) {       /* rect.cpp:92-102 */
        Ray r = make_ray(origin, dir, 0.0f, 0);
        if (rect_hit(o, &r, 0.001f, FLT_MAX, &rec)) {
            float area = (o->a1 - o->a0) * (o->b1 - o->b0);
            float dist_sq = rec.t * rec.t;
            float cosine = fabsf(dot(dir, rec.n));
            return dist_sq / (cosine * area);
        }
        return 0;
    }
    if (o->kind == O_SPHERE) {   /* sphere.cpp:63-72 */
        Ray r = make_ray(origin, dir, time, 0);
        if (sphere_hit(o, &r, 0.001f, FLT_MAX, &rec)) {
            V3 d = sub(sphere_center(o, time), origin);
            float cos_theta_max = sqrtf(1 - o->radius * o->radius / dot(d, d));
            float solid_angle = 2 * PI_F * (1 - cos_theta_max);
            return 1 / solid_angle;
        }
        return 0;
    }
    return 0;   /* scene_object::pdf_value default, scene_object.h:24-26 */
}
typedef struct { V3 u, v, w; } Onb;
static Onb make_onb(V3 n) {   /* onb.h:19-23 */
    Onb o; o.w = n;
    V3 a = (fabsf(n.x) > 0.9f) ? v3(0, 1, 0) : v3(1, 0, 0);
    o.v = normalize(cross(o.w, a));
    o.u = cross(o.w, o.v);
    return o;
}
static V3 onb_local(const Onb *o, V3 a) { return add(add(muls(o->u, a.x), muls(o->v, a.y)), muls(o->w, a.z)); }
static V3 obj_pdf_generate(const Obj *o, V3 origin, float time, Rng *rng) {
    if (o->kind == O_LIST) { int i = (int) (randf(rng) * o->n); return obj_pdf_generate(o->children[i], origin, time, rng); }
    if (o->kind == O_XZ) {       /* rect.cpp:104-107, x then z */
        float rx = randf(rng), rz = randf(rng);
        return sub(v3(o->a0 + rx * (o->a1 - o->a0), o->k, o->b0 + rz * (o->b1 - o->b0)), origin);
    }
    if (o->kind == O_SPHERE) {   /* sphere.cpp:74-79 */
        V3 dir = sub(sphere_center(o, time), origin);
        float dist_sq = dot(dir, dir);
        Onb uvw = make_onb(normalize(dir));
        return onb_local(&uvw, random_towards_sphere(o->radius, dist_sq, rng));
    }
    return v3(1, 0, 0);
}

/* ------------------------------------------------------------------ trace (main.cpp:66-118) */
static uint32_t g_max_bounces = 32;
typedef struct { uint64_t rays; } Stats;
static V3 trace(const Scene *sc, const Ray *r, uint32_t depth, Rng *rng, Stats *stats) {
    stats->rays++;
    HitRec rec;
    if (obj_hit(sc->objects, r, 0.001f, FLT_MAX, &rec, rng)) {
        const Mat *m = rec.mat;
        V3 emitted = v3(0, 0, 0);
        if (m->kind == M_LIGHT && dot(rec.n, r->d) < 0.0f) emitted = muls(tex_sample(m->tex, rec.u, rec.v, rec.p), m->param);   /* material.h:190-199 */
        if (!(depth < g_max_bounces) || m->kind == M_LIGHT) return emitted;
        if (m->kind == M_METAL) {   /* material.h:84-98 */
            V3 reflected = sub(r->d, muls(rec.n, 2.0f * dot(r->d, rec.n)));
            V3 fuzz = muls(random_in_sphere(rng), 1 - m->param);
            Ray spec = make_ray(rec.p, add(reflected, fuzz), r->time, 0);
            V3 att = tex_sample(m->tex, rec.u, rec.v, rec.p);
            return mulv(att, trace(sc, &spec, depth + 1, rng, stats));
        }
        if (m->kind == M_DIELECTRIC) {   /* material.h:106-175 */
            float ref_index = m->param, ni_over_nt, cosI = -dot(r->d, rec.n);
            V3 fn;
            if (cosI < 0) { fn = v3(-rec.n.x, -rec.n.y, -rec.n.z); ni_over_nt = ref_index; }
            else { fn = rec.n; ni_over_nt = 1.0f / ref_index; }
            V3 reflected = sub(r->d, muls(rec.n, 2.0f * dot(r->d, rec.n)));
            float ncosI = dot(r->d, fn);   /* refract(), vec3.h:185-198 */
            float sinT2 = (ni_over_nt * ni_over_nt) * (1.0f - ncosI * ncosI);
            Ray spec;
            if (sinT2 <= 1.0f) {
                float cosT = sqrtf(1.0f - sinT2);
                V3 refracted = add(muls(r->d, ni_over_nt), muls(fn, ni_over_nt * (-ncosI) - cosT));
                float cosine_schlick = (cosI < 0) ? sqrtf(1.0f - ni_over_nt * ni_over_nt * (1.0f - cosI * cosI)) : cosI;
                float r0 = (1 - ref_index) / (1 + ref_index);
                r0 = r0 * r0;
                float reflect_prob = r0 + (1 - r0) * cr_pow5f((1 - cosine_schlick));
                if (randf(rng) < reflect_prob) {
                    spec = make_ray(rec.p, reflected, r->time, r->inside);
                } else {
                    int inside = r->inside;
                    if (cosI < 0) { inside--; if (inside < 0) inside = 0; } else inside++;
                    spec = make_ray(rec.p, refracted, r->time, inside);
                }
            } else {
                spec = make_ray(rec.p, reflected, r->time, r->inside);
            }
            return trace(sc, &spec, depth + 1, rng, stats);   /* attenuation (1,1,1) */
        }
        /* lambertian (material.h:40-53) / isotropic (material.h:64-73) + pdf sampling, main.cpp:84-102 */
        V3 att = tex_sample(m->tex, rec.u, rec.v, rec.p);
        int lambert = (m->kind == M_LAMBERTIAN);
        Onb uvw; if (lambert) uvw = make_onb(rec.n);
        V3 dir;
        int use_light = 0;
        if (sc->biased) use_light = randf(rng) < 0.5f;   /* mix_pdf::generate, pdf.h:74-79 */
        if (use_light) dir = obj_pdf_generate(sc->biased, rec.p, r->time, rng);
        else if (lambert) dir = onb_local(&uvw, random_cosine_direction(rng));
        else dir = random_in_sphere(rng);
        Ray scattered = make_ray(rec.p, dir, r->time, 0);
        float mat_pdf;
        if (lambert) { float cosine = dot(scattered.d, uvw.w); mat_pdf = (cosine > 0) ? cosine / PI_F : 0.0f; }
        else mat_pdf = 1 / (2 * PI_F);
        float pdf_v = mat_pdf;
        if (sc->biased) pdf_v = 0.5f * (obj_pdf_value(sc->biased, rec.p, scattered.d, r->time, rng) + mat_pdf);
        float spdf;
        if (lambert) { float cosine = dot(rec.n, scattered.d); spdf = (cosine < 0) ? 0.0f : cosine * (1.0f / PI_F); }
        else spdf = 1.0f / (2.0f * PI_F);
        V3 li = trace(sc, &scattered, depth + 1, rng, stats);
        return add(emitted, divs(mulv(muls(att, spdf), li), pdf_v));   /* main.cpp:102 */
    }
    if (!sc->sky) return v3(0, 0, 0);   /* main.cpp:110-111 */
    float t = 0.5f * (r->d.y + 1.0f);
    return add(v3(1.0f - t, 1.0f - t, 1.0f - t), muls(v3(0.5f, 0.7f, 1.0f), t));
}

static Ray camera_get_ray(const Camera *c, float s, float t, Rng *rng) {   /* camera.h:38-45 */
    V3 rd = muls(random_in_disk(rng), c->lens_radius);
    V3 offset = add(muls(c->u, rd.x), muls(c->v, rd.y));
    float time = c->time0 + (c->time1 - c->time0) * randf(rng);
    V3 dir = sub(sub(add(add(c->llcorner, muls(c->horz, s)), muls(c->vert, t)), c->origin), offset);
    return make_ray(add(c->origin, offset), dir, time, 0);
}

/* ------------------------------------------------------------------ scene dump parser */
static char *g_lines[1 << 20]; static int g_nlines, g_pos;
static uint8_t *g_image; static int g_img_w, g_img_h;
static float hexf(const char *s) { union { float f; uint32_t u; } a; a.u = (uint32_t) strtoul(s, NULL, 16); return a.f; }
static const char *field(const char *line, const char *name) {
    size_t n = strlen(name);
    for (const char *p = line; (p = strstr(p, name)); p += n)
        if ((p == line || p[-1] == ' ' || p[-1] == '[' || p[-1] == '{') && p[n] == '=') return p + n + 1;
    fprintf(stderr, "missing field %s in: %.80s\n", name, line); exit(1);
}
static float ff(const char *line, const char *name) { return hexf(field(line, name)); }
static V3 fv(const char *line, const char *name) { const char *p = field(line, name); return v3(hexf(p), hexf(p + 9), hexf(p + 18)); }
static int fi(const char *line, const char *name) { return atoi(field(line, name)); }
static Tex *parse_tex(const char **pp) {
    const char *p = *pp; Tex *t = calloc(1, sizeof(Tex));
    if (!strncmp(p, "[color", 6)) { t->kind = T_COLOR; t->color = fv(p, "c"); p = strchr(p, ']') + 1; }
    else if (!strncmp(p, "[checker", 8)) {
        t->kind = T_CHECKER; t->scale = ff(p, "scale");
        p = strstr(p, "even=") + 5; t->even = parse_tex(&p);
        p = strstr(p, "odd=") + 4; t->odd = parse_tex(&p);
        p = strchr(p, ']') + 1;
    } else if (!strncmp(p, "[perlin", 7)) { t->kind = T_PERLIN; t->scale = ff(p, "scale"); p = strchr(p, ']') + 1; }
    else if (!strncmp(p, "[image", 6)) {
        t->kind = T_IMAGE; t->w = fi(p, "w"); t->h = fi(p, "h"); t->data = g_image;
        if (!g_image || g_img_w != t->w || g_img_h != t->h) { fprintf(stderr, "image texture needs -image <ppm %dx%d>\n", t->w, t->h); exit(1); }
        p = strchr(p, ']') + 1;
    } else { fprintf(stderr, "bad texture: %.40s\n", p); exit(1); }
    *pp = p; return t;
}
static Mat *parse_mat(const char *line) {
    const char *p = strstr(line, "mat={");
    if (!p) { fprintf(stderr, "no material in: %.80s\n", line); exit(1); }
    p += 4; Mat *m = calloc(1, sizeof(Mat));
    const char *tx = strstr(p, "tex=");
    if (!strncmp(p, "{lambertian", 11)) m->kind = M_LAMBERTIAN;
    else if (!strncmp(p, "{isotropic", 10)) m->kind = M_ISOTROPIC;
    else if (!strncmp(p, "{metal", 6)) { m->kind = M_METAL; m->param = ff(p, "gloss"); }
    else if (!strncmp(p, "{dielectric", 11)) { m->kind = M_DIELECTRIC; m->param = ff(p, "idx"); tx = NULL; }
    else if (!strncmp(p, "{light", 6)) { m->kind = M_LIGHT; m->param = ff(p, "scale"); }
    else { fprintf(stderr, "bad material: %.40s\n", p); exit(1); }
    if (tx) { tx += 4; m->tex = parse_tex(&tx); }
    return m;
}
static Obj *parse_obj(void) {
    char *line = g_lines[g_pos++];
    while (*line == ' ') line++;
    Obj *o = calloc(1, sizeof(Obj));
    if (!strncmp(line, "sphere", 6)) {
        o->kind = O_SPHERE; o->c0 = fv(line, "c0"); o->c1 = fv(line, "c1"); o->t0 = ff(line, "t0"); o->t1 = ff(line, "t1");
        o->moving = fi(line, "moving"); o->radius = ff(line, "r"); o->mat = parse_mat(line);
    } else if (!strncmp(line, "xy_rect", 7) || !strncmp(line, "xz_rect", 7) || !strncmp(line, "yz_rect", 7)) {
        o->kind = line[1] == 'y' ? O_XY : (line[0] == 'x' ? O_XZ : O_YZ);
        o->a0 = ff(line, "a0"); o->a1 = ff(line, "a1"); o->b0 = ff(line, "b0"); o->b1 = ff(line, "b1"); o->k = ff(line, "k");
        o->sign = ff(line, "sign"); o->mat = parse_mat(line);
    } else if (!strncmp(line, "box", 3)) {
        o->kind = O_BOX; o->box.min = fv(line, "min"); o->box.max = fv(line, "max"); o->child = parse_obj();
    } else if (!strncmp(line, "list", 4)) {
        o->kind = O_LIST; o->n = (size_t) fi(line, "n"); o->has_box = fi(line, "hasBox");
        if (o->has_box) { o->box.min = fv(line, "min"); o->box.max = fv(line, "max"); }
        o->children = calloc(o->n, sizeof(Obj *));
        for (size_t i = 0; i < o->n; i++) o->children[i] = parse_obj();
    } else if (!strncmp(line, "bvh", 3)) {
        o->kind = O_BVH; o->order = (uint32_t) strtoul(field(line, "order"), NULL, 16);
        o->box.min = fv(line, "min"); o->box.max = fv(line, "max");
        o->left = parse_obj(); o->right = parse_obj();
    } else if (!strncmp(line, "translate", 9)) {
        o->kind = O_TRANSLATE; o->offset = fv(line, "offset"); o->child = parse_obj();
    } else if (!strncmp(line, "rotate_y", 8)) {
        o->kind = O_ROTATE; o->sin_t = ff(line, "sin"); o->cos_t = ff(line, "cos"); o->has_box = fi(line, "hasBox");
        o->box.min = fv(line, "min"); o->box.max = fv(line, "max"); o->child = parse_obj();
    } else if (!strncmp(line, "volume", 6)) {
        o->kind = O_VOLUME; o->density = ff(line, "density"); o->mat = parse_mat(line); o->child = parse_obj();
    } else if (!strncmp(line, "podbvh", 6)) {
        o->kind = O_POD; o->n_tris = (uint32_t) fi(line, "prims"); o->n_nodes = (uint32_t) fi(line, "nodes");
        o->nodes = calloc(o->n_nodes, sizeof(PodNode)); o->tris = calloc(o->n_tris, sizeof(Tri));
        for (uint32_t i = 0; i < o->n_nodes; i++) {
            char *l = g_lines[g_pos++]; PodNode *n = &o->nodes[i];
            n->left = (uint32_t) fi(l, "left"); n->off = (uint32_t) fi(l, "off"); n->cnt = (uint32_t) fi(l, "cnt");
            n->order = (uint32_t) strtoul(field(l, "order"), NULL, 16); n->box.min = fv(l, "min"); n->box.max = fv(l, "max");
        }
        for (uint32_t i = 0; i < o->n_tris; i++) {
            char *l = g_lines[g_pos++]; Tri *t = &o->tris[i];
            t->m = fv(l, "m"); t->u = fv(l, "u"); t->v = fv(l, "v"); t->mn = fv(l, "mn"); t->un = fv(l, "un"); t->vn = fv(l, "vn");
            if (i == 0) o->mat = parse_mat(l);
        }
    } else { fprintf(stderr, "unknown object line: %.60s\n", line); exit(1); }
    return o;
}
static void parse_scene(const char *path, Scene *sc) {
    FILE *f = fopen(path, "r");
    if (!f) { perror(path); exit(1); }
    char *buf = NULL; size_t cap = 0; ssize_t len;
    while ((len = getline(&buf, &cap, f)) > 0) { if (buf[len - 1] == '\n') buf[len - 1] = 0; g_lines[g_nlines++] = strdup(buf); }
    fclose(f); free(buf);
    const char *c = g_lines[0];
    Camera *cam = &sc->cam;
    cam->origin = fv(c, "origin"); cam->u = fv(c, "u"); cam->v = fv(c, "v"); cam->w = fv(c, "w"); cam->llcorner = fv(c, "llcorner");
    cam->horz = fv(c, "horz"); cam->vert = fv(c, "vert"); cam->lens_radius = ff(c, "lens_radius"); cam->time0 = ff(c, "time0"); cam->time1 = ff(c, "time1");
    g_pos = 2;   /* line 1 = "objects" */
    sc->objects = parse_obj();
    if (!strncmp(g_lines[g_pos], "biased none", 11)) sc->biased = NULL;
    else { g_pos++; sc->biased = parse_obj(); }
}

/* ------------------------------------------------------------------ render driver */
typedef struct { char magic[8]; uint32_t width, height, samples, s0, s1, depth, scene, threads; uint64_t seed, rays; double seconds; } FileHeader;
static Scene g_scene; static uint32_t W = 500, H = 500, N = 16, S0 = 0, S1 = 0, SQ = 4, X0 = 0, X1 = 0, Y0 = 0, Y1 = 0;   /* [X0,X1)x[Y0,Y1): crop window */ static uint64_t SEED = 11350390909718046443ULL;
static float *g_acc; static uint32_t g_next_row; static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER; static uint64_t g_rays;
static void *worker(void *arg) {
    (void) arg; Stats st = {0};
    for (;;) {
        pthread_mutex_lock(&g_lock); uint32_t y = g_next_row++; pthread_mutex_unlock(&g_lock);
        if (y >= Y1) break;
        for (uint32_t x = X0; x < X1; x++) {
            V3 color = v3(0, 0, 0); uint32_t cnt = 0;
            for (uint32_t s = S0; s < S1; s++) {
                Rng rng; pcg32_seed(&rng, SEED, ((uint64_t) y * W + x) * N + s);
                uint32_t i = s / SQ, j = s % SQ;   /* regular grid, main.cpp:319-332 */
                float sx = (i + 0.5f) / (float) SQ, sy = (j + 0.5f) / (float) SQ;
                float u = (x + sx) / (float) W, v = (y + sy) / (float) H;   /* main.cpp:156-157 */
                Ray r = camera_get_ray(&g_scene.cam, u, v, &rng);
                V3 c = trace(&g_scene, &r, 0, &rng, &st);
                if (isfinite(c.x) && isfinite(c.y) && isfinite(c.z)) { color = add(color, c); cnt++; }   /* main.cpp:163-165 */
            }
            float *o = &g_acc[((size_t) (y - Y0) * (X1 - X0) + (x - X0)) * 4];
            o[0] = color.x; o[1] = color.y; o[2] = color.z; o[3] = (float) cnt;
        }
    }
    pthread_mutex_lock(&g_lock); g_rays += st.rays; pthread_mutex_unlock(&g_lock);
    return NULL;
}
static const char *arg(int argc, char **argv, const char *name, const char *def) {
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return argv[i + 1];
    return def;
}
int main(int argc, char **argv) {
    const char *dump = arg(argc, argv, "-dump", NULL), *out = arg(argc, argv, "-out", NULL), *image = arg(argc, argv, "-image", NULL);
    if (!dump) { fprintf(stderr, "usage: %s -dump scene.txt [-image earthmap.ppm] -width W -height H -samples N -depth D -out acc.bin\n", argv[0]); return 2; }
    W = (uint32_t) atoi(arg(argc, argv, "-width", "500")); H = (uint32_t) atoi(arg(argc, argv, "-height", "500"));
    uint32_t spp = (uint32_t) atoi(arg(argc, argv, "-samples", "16"));
    g_max_bounces = (uint32_t) atoi(arg(argc, argv, "-depth", "32"));
    SEED = strtoull(arg(argc, argv, "-seed", "11350390909718046443"), NULL, 0);
    int threads = atoi(arg(argc, argv, "-threads", "8"));
    SQ = (uint32_t) sqrtf((float) spp); N = SQ * SQ;
    S0 = (uint32_t) atoi(arg(argc, argv, "-s0", "0")); S1 = (uint32_t) atoi(arg(argc, argv, "-s1", "0"));
    if (S1 == 0 || S1 > N) S1 = N;
    X0 = (uint32_t) atoi(arg(argc, argv, "-x0", "0")); X1 = (uint32_t) atoi(arg(argc, argv, "-x1", "0"));
    Y0 = (uint32_t) atoi(arg(argc, argv, "-y0", "0")); Y1 = (uint32_t) atoi(arg(argc, argv, "-y1", "0"));
    if (X1 == 0 || X1 > W) X1 = W;
    if (Y1 == 0 || Y1 > H) Y1 = H;
    if (X0 >= X1 || Y0 >= Y1) { fprintf(stderr, "bad crop window\n"); return 2; }
    g_next_row = Y0;
    if (image) {
        FILE *f = fopen(image, "rb"); int maxv;
        if (!f || fscanf(f, "P6 %d %d %d", &g_img_w, &g_img_h, &maxv) != 3) { fprintf(stderr, "bad ppm %s\n", image); return 1; }
        fgetc(f); g_image = malloc((size_t) g_img_w * g_img_h * 3);
        if (fread(g_image, 3, (size_t) g_img_w * g_img_h, f) != (size_t) g_img_w * g_img_h) { fprintf(stderr, "short ppm\n"); return 1; }
        fclose(f);
    }
    perlin_init();
    parse_scene(dump, &g_scene);
    g_scene.sky = atoi(arg(argc, argv, "-sky", "0"));
    g_acc = calloc((size_t) (X1 - X0) * (Y1 - Y0) * 4, sizeof(float));
    pthread_t th[256]; if (threads > 256) threads = 256; if (threads < 1) threads = 1;
    for (int i = 0; i < threads; i++) pthread_create(&th[i], NULL, worker, NULL);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    printf("{\"mode\":\"restatement\",\"rays\":%llu,\"paths\":%llu}\n", (unsigned long long) g_rays, (unsigned long long) (X1 - X0) * (Y1 - Y0) * (S1 - S0));
    if (out) {
        FileHeader h; memset(&h, 0, sizeof(h)); memcpy(h.magic, "MRTACC1", 8);
        h.width = X1 - X0; h.height = Y1 - Y0; h.samples = N; h.s0 = S0; h.s1 = S1; h.depth = g_max_bounces; h.threads = (uint32_t) threads; h.seed = SEED; h.rays = g_rays;
        FILE *f = fopen(out, "wb"); if (!f) { perror(out); return 1; }
        fwrite(&h, sizeof(h), 1, f); fwrite(g_acc, sizeof(float) * 4, (size_t) (X1 - X0) * (Y1 - Y0), f); fclose(f);
    }
    return 0;
}
