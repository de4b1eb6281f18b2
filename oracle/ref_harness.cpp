// TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.
//
// Harness around the UNMODIFIED arithmetic of the reference renderer
// (/root/reference, patched only as listed in patch_reference.py).  It pulls in
// the reference's main.cpp with `main` renamed so that trace() (main.cpp:66),
// the file-static parameters and the linear back buffer are reachable, and
// adds what the reference lacks:
//   * a per-(pixel, sample) RNG stream: Init_Thread_RNG(seed, (y*W+x)*N+s)
//     (pcg.cpp:44) before every camera::get_ray (camera.h:38) + trace call,
//     so that results are independent of thread->tile scheduling
//     (the reference seeds per worker thread, main.cpp:143);
//   * a raw dump of the accumulator: float4 per pixel = (sum of finite
//     radiance samples, finite-sample count);
//   * a canonical text dump of the constructed scene graph (scene.cpp) used by
//     the scene-parity test of the new host-side scene builder;
//   * known-answer vectors for the PCG32 generator and the Perlin tables;
//   * "stock" mode: the reference's own main() (own threading, own per-thread
//     RNG, own clock) run headless, for the CPU baseline.
//
// usage:
//   mrt_ref render -scene S -width W -height H -samples N -depth D -seed X
//                  [-s0 a -s1 b] [-x0 a -x1 b -y0 c -y1 d] [-threads T] [-maxlum L] [-lights all] [-extra triangles] [-draw2 1] -out file.bin
//                  (crop window: only those pixels of the W x H frame are traced; stream ids and u,v stay the
//                   full frame's, so BASELINE.json's full-size configurations can be spot-checked in seconds)
//   mrt_ref stock  <reference command line>  [-dump file.bin] [-dumpargb file.u32]   (final linear frame / its tone map)
//   mrt_ref dump-scene -scene S -width W -height H -out file.txt
//   mrt_ref kat
//   mrt_ref dump-image out.ppm          (decodes ../earthmap.jpg via the
//                                        reference's vendored stb_image)
// Built twice by build_ref.sh: mrt_ref (parity: -ffp-contract=off, portable -march, canonical libm interposed at link time)
// and mrt_ref_native / mrt_ref_v3 (-DMRT_REF_TIMING_ONLY: render + stock only, the reference's OWN build flags
// clang/clang_build_linux.sh:23-29 = -O3 -march=native -fno-exceptions -fno-rtti, host libm) for the CPU baseline.
// The process must be started with cwd = <assets>/run (the reference opens
// "../earthmap.jpg" and "../obj/*.obj", scene.cpp:139,503,509).
#include <atomic>
#include <limits>
#include <memory>
#include <thread>
#include <vector>
#include <string>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define main ref_main
#include "main.cpp"
#undef main

#ifndef MRT_REF_TIMING_ONLY
#include "stb_image.h"
#endif

extern bool MRT_headless_quiet;
const char *MRT_headless_last_title();

// globals defined (non-static) in texture.cpp:200-203
extern Vec3 *rv;
extern int *px;
extern int *py;
extern int *pz;

static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

#ifndef MRT_REF_TIMING_ONLY   // the timing-only build (reference's own flags incl. -fno-rtti, host libm) has no scene dump
// ---------------------------------------------------------------- scene dump
static void ind(FILE *f, int d) { for (int i = 0; i < d; i++) fputc(' ', f); }
static void pv(FILE *f, const char *name, const Vec3 &v) {
    fprintf(f, " %s=%08x,%08x,%08x", name, fbits(v.x), fbits(v.y), fbits(v.z));
}
static void pf(FILE *f, const char *name, float v) { fprintf(f, " %s=%08x", name, fbits(v)); }

static uint32_t crc_bytes(const uint8 *d, size_t n) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) { h ^= d[i]; h *= 16777619u; }
    return h;
}

static void dump_tex(FILE *f, const texture *t) {
    if (auto c = dynamic_cast<const color_tex *>(t)) {
        fprintf(f, "[color"); pv(f, "c", c->color); fprintf(f, "]");
    } else if (auto c = dynamic_cast<const checker_tex *>(t)) {
        fprintf(f, "[checker"); pf(f, "scale", c->scale);
        fprintf(f, " even="); dump_tex(f, c->even);
        fprintf(f, " odd="); dump_tex(f, c->odd); fprintf(f, "]");
    } else if (auto c = dynamic_cast<const perlin_tex *>(t)) {
        fprintf(f, "[perlin"); pf(f, "scale", c->scale); fprintf(f, "]");
    } else if (auto c = dynamic_cast<const image_tex *>(t)) {
        fprintf(f, "[image w=%d h=%d fnv=%08x]", c->width, c->height,
                crc_bytes(c->data, (size_t) c->width * c->height * 3));
    } else {
        fprintf(f, "[unknown-tex]");
    }
}

static void dump_mat(FILE *f, const material *m) {
    if (auto c = dynamic_cast<const lambertian *>(m)) {
        fprintf(f, "{lambertian tex="); dump_tex(f, c->albedo); fprintf(f, "}");
    } else if (auto c = dynamic_cast<const isotropic *>(m)) {
        fprintf(f, "{isotropic tex="); dump_tex(f, c->albedo); fprintf(f, "}");
    } else if (auto c = dynamic_cast<const metal *>(m)) {
        fprintf(f, "{metal"); pf(f, "gloss", c->gloss); fprintf(f, " tex="); dump_tex(f, c->albedo); fprintf(f, "}");
    } else if (auto c = dynamic_cast<const dielectric *>(m)) {
        fprintf(f, "{dielectric"); pf(f, "idx", c->ref_index); fprintf(f, "}");
    } else if (auto c = dynamic_cast<const diffuse_light *>(m)) {
        fprintf(f, "{light"); pf(f, "scale", c->scale); fprintf(f, " tex="); dump_tex(f, c->emissive); fprintf(f, "}");
    } else {
        fprintf(f, "{unknown-mat}");
    }
}

template <typename T> static void dump_list_children(FILE *f, const object_list<T> *l, int d);
static void dump_obj(FILE *f, const scene_object *o, int d);

static void dump_sphere(FILE *f, const sphere *s, int d) {
    ind(f, d); fprintf(f, "sphere");
    pv(f, "c0", s->center0); pv(f, "c1", s->center1); pf(f, "t0", s->time0); pf(f, "t1", s->time1);
    fprintf(f, " moving=%d", (int) s->isMoving); pf(f, "r", s->radius);
    fprintf(f, " mat="); dump_mat(f, s->mat_ptr); fprintf(f, "\n");
}

static void dump_box(FILE *f, const box *b, int d) {
    ind(f, d); fprintf(f, "box"); pv(f, "min", b->min); pv(f, "max", b->max); fprintf(f, "\n");
    dump_obj(f, b->rect_list, d + 1);
}

template <typename T> static void dump_typed(FILE *f, const T *o, int d);
template <> void dump_typed<sphere>(FILE *f, const sphere *o, int d) { dump_sphere(f, o, d); }
template <> void dump_typed<box>(FILE *f, const box *o, int d) { dump_box(f, o, d); }
template <> void dump_typed<scene_object>(FILE *f, const scene_object *o, int d) { dump_obj(f, o, d); }

template <typename T> static void dump_list(FILE *f, const object_list<T> *l, int d) {
    ind(f, d); fprintf(f, "list n=%zu hasBox=%d", l->count, (int) l->hasBox);
    if (l->hasBox) { pv(f, "min", l->box.min); pv(f, "max", l->box.max); }
    fprintf(f, "\n");
    for (size_t i = 0; i < l->count; i++) dump_typed<T>(f, l->list[i], d + 1);
}

template <typename T> static void dump_bvh(FILE *f, const bvh_node<T> *b, int d) {
    ind(f, d); fprintf(f, "bvh order=%02x same=%d", (unsigned) b->node_order, (int) (b->left == b->right));
    pv(f, "min", b->box.min); pv(f, "max", b->box.max); fprintf(f, "\n");
    dump_obj(f, b->left, d + 1);
    dump_obj(f, b->right, d + 1);
}

static void dump_podbvh(FILE *f, const pod_bvh<triangle> *b, int d) {
    ind(f, d); fprintf(f, "podbvh prims=%u nodes=%u root=%u\n", b->prim_count, b->node_count, b->root_node);
    for (uint32 i = 0; i < b->node_count; i++) {
        const pod_bvh_node &n = b->nodes[i];
        ind(f, d + 1); fprintf(f, "node %u left=%u off=%u cnt=%u order=%02x", i,
                               n.prim_count ? 0u : n.left, n.prim_offset, n.prim_count,
                               n.prim_count ? 0u : (unsigned) n.node_order);
        pv(f, "min", n.box.min); pv(f, "max", n.box.max); fprintf(f, "\n");
    }
    for (uint32 i = 0; i < b->prim_count; i++) {
        const triangle &t = b->prims[i];
        ind(f, d + 1); fprintf(f, "tri %u", i);
        pv(f, "m", t.m); pv(f, "u", t.u); pv(f, "v", t.v); pv(f, "mn", t.mn); pv(f, "un", t.un); pv(f, "vn", t.vn);
        if (i == 0) { fprintf(f, " mat="); dump_mat(f, t.mat_ptr); }
        fprintf(f, "\n");
    }
}

static void dump_obj(FILE *f, const scene_object *o, int d) {
    if (auto s = dynamic_cast<const sphere *>(o)) {
        dump_sphere(f, s, d);
    } else if (auto r = dynamic_cast<const xy_rect *>(o)) {
        ind(f, d); fprintf(f, "xy_rect"); pf(f, "a0", r->x0); pf(f, "a1", r->x1); pf(f, "b0", r->y0); pf(f, "b1", r->y1);
        pf(f, "k", r->z); pf(f, "sign", r->normal_sign); fprintf(f, " mat="); dump_mat(f, r->mat_ptr); fprintf(f, "\n");
    } else if (auto r = dynamic_cast<const xz_rect *>(o)) {
        ind(f, d); fprintf(f, "xz_rect"); pf(f, "a0", r->x0); pf(f, "a1", r->x1); pf(f, "b0", r->z0); pf(f, "b1", r->z1);
        pf(f, "k", r->y); pf(f, "sign", r->normal_sign); fprintf(f, " mat="); dump_mat(f, r->mat_ptr); fprintf(f, "\n");
    } else if (auto r = dynamic_cast<const yz_rect *>(o)) {
        ind(f, d); fprintf(f, "yz_rect"); pf(f, "a0", r->y0); pf(f, "a1", r->y1); pf(f, "b0", r->z0); pf(f, "b1", r->z1);
        pf(f, "k", r->x); pf(f, "sign", r->normal_sign); fprintf(f, " mat="); dump_mat(f, r->mat_ptr); fprintf(f, "\n");
    } else if (auto b = dynamic_cast<const box *>(o)) {
        dump_box(f, b, d);
    } else if (auto l = dynamic_cast<const object_list<scene_object> *>(o)) {
        dump_list(f, l, d);
    } else if (auto l = dynamic_cast<const object_list<sphere> *>(o)) {
        dump_list(f, l, d);
    } else if (auto l = dynamic_cast<const object_list<box> *>(o)) {
        dump_list(f, l, d);
    } else if (auto b = dynamic_cast<const bvh_node<sphere> *>(o)) {
        dump_bvh(f, b, d);
    } else if (auto b = dynamic_cast<const bvh_node<box> *>(o)) {
        dump_bvh(f, b, d);
    } else if (auto b = dynamic_cast<const bvh_node<scene_object> *>(o)) {
        dump_bvh(f, b, d);
    } else if (auto t = dynamic_cast<const translate *>(o)) {
        ind(f, d); fprintf(f, "translate"); pv(f, "offset", t->offset); fprintf(f, "\n");
        dump_obj(f, t->obj, d + 1);
    } else if (auto r = dynamic_cast<const rotate_y *>(o)) {
        ind(f, d); fprintf(f, "rotate_y"); pf(f, "sin", r->sin_theta); pf(f, "cos", r->cos_theta);
        fprintf(f, " hasBox=%d", (int) r->hasBox); pv(f, "min", r->bbox.min); pv(f, "max", r->bbox.max); fprintf(f, "\n");
        dump_obj(f, r->obj, d + 1);
    } else if (auto v = dynamic_cast<const constant_volume *>(o)) {
        ind(f, d); fprintf(f, "volume"); pf(f, "density", v->density); fprintf(f, " mat="); dump_mat(f, v->phase_function); fprintf(f, "\n");
        dump_obj(f, v->boundary, d + 1);
    } else if (auto p = dynamic_cast<const pod_bvh<triangle> *>(o)) {
        dump_podbvh(f, p, d);
    } else if (auto t = dynamic_cast<const triangle_scene_object *>(o)) {
        ind(f, d); fprintf(f, "triangle_object");
        pv(f, "m", t->m); pv(f, "u", t->u); pv(f, "v", t->v); pv(f, "mn", t->mn); pv(f, "un", t->un); pv(f, "vn", t->vn);
        fprintf(f, " mat="); dump_mat(f, t->mat_ptr); fprintf(f, "\n");
    } else {
        ind(f, d); fprintf(f, "unknown-object\n");
    }
}

static void dump_scene(FILE *f, const scene &sc) {
    const camera *c = sc.camera;
    fprintf(f, "camera"); pv(f, "origin", c->origin); pv(f, "u", c->u); pv(f, "v", c->v); pv(f, "w", c->w);
    pv(f, "llcorner", c->llcorner); pv(f, "horz", c->horz); pv(f, "vert", c->vert);
    pf(f, "lens_radius", c->lens_radius); pf(f, "time0", c->time0); pf(f, "time1", c->time1); fprintf(f, "\n");
    fprintf(f, "objects\n");
    dump_obj(f, sc.objects, 1);
    if (sc.biased_objects) {
        fprintf(f, "biased\n");
        dump_obj(f, sc.biased_objects, 1);
    } else {
        fprintf(f, "biased none\n");
    }
}

#endif   // MRT_REF_TIMING_ONLY

// ------------------------------------------------------------------- helpers
static const char *argval(int argc, char **argv, const char *name, const char *def) {
    for (int i = 2; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return argv[i + 1];
    return def;
}

static scene build_scene(uint32 sceneSelect, uint32 W, uint32 H) {
    // main.cpp:302-309
    Init_Thread_RNG(11350390909718046443uLL, 6305599193148252115uLL);
    return select_scene((scenes) sceneSelect, float(W) / float(H));
}

// "-extra triangles" (Cornell box only): appends two triangle_scene_objects (triangle.cpp:5-175 -- a class of the reference that
// none of its scenes instantiates) to the scene's object list, one through each constructor, so that the class has an oracle.
// The reference's object_list is rebuilt with the longer array (its constructor also recomputes the list's box).
static void add_extra_triangles(scene &sc) {
    auto *old = (object_list<scene_object> *) sc.objects;
    const size_t n = old->count;
    scene_object **list = new scene_object *[n + 2];
    for (size_t i = 0; i < n; i++) list[i] = old->list[i];
    material *red = new lambertian(new color_tex(Vec3(0.65f, 0.055f, 0.06f)));
    material *alu = new metal(new color_tex(Vec3(0.8f, 0.85f, 0.88f)), 0.9f);
    list[n] = new triangle_scene_object(Vec3(100, 300, 250), Vec3(400, 320, 300), Vec3(250, 520, 420), red);
    list[n + 1] = new triangle_scene_object(Vec3(420, 60, 120), Vec3(520, 60, 260), Vec3(470, 260, 180), Vec3(0, 0, -1), Vec3(-0.6f, 0, -0.8f),
                                            Vec3(0, 0.6f, -0.8f), alu);
    sc.objects = new object_list<scene_object>(list, n + 2, 0.0f, 1.0f);
}

struct FileHeader {
    char magic[8];        // "MRTACC1\0"
    uint32_t width, height, samples, s0, s1, depth, scene, threads;
    uint64_t seed, rays;
    double seconds;
};

static int write_acc(const char *path, const FileHeader &h, const float *acc) {
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); return 1; }
    fwrite(&h, sizeof(h), 1, f);
    fwrite(acc, sizeof(float) * 4, (size_t) h.width * h.height, f);
    fclose(f);
    return 0;
}

// -------------------------------------------------------------------- render
static int cmd_render(int argc, char **argv) {
    MRT_Params p;
    uint32 W = strtoul(argval(argc, argv, "-width", "500"), 0, 0);
    uint32 H = strtoul(argval(argc, argv, "-height", "500"), 0, 0);
    uint32 spp = strtoul(argval(argc, argv, "-samples", "16"), 0, 0);
    p.windowWidth = p.bufferWidth = W;
    p.windowHeight = p.bufferHeight = H;
    p.samplesPerPixel = spp;
    p.maxBounces = strtoul(argval(argc, argv, "-depth", "32"), 0, 0);
    p.sceneSelect = strtoul(argval(argc, argv, "-scene", "0"), 0, 0);
    p.maxLuminance = strtof(argval(argc, argv, "-maxlum", "1000"), 0);
    uint64 seed = strtoull(argval(argc, argv, "-seed", "11350390909718046443"), 0, 0);
    uint32 nthreads = strtoul(argval(argc, argv, "-threads", "0"), 0, 0);
    if (!nthreads) nthreads = std::thread::hardware_concurrency();
    const char *out = argval(argc, argv, "-out", nullptr);
    *getParams() = p;

    scene sc = build_scene(p.sceneSelect, W, H);
    if (!strcmp(argval(argc, argv, "-extra", "none"), "triangles") && p.sceneSelect == 5) add_extra_triangles(sc);
    if (!strcmp(argval(argc, argv, "-lights", "ref"), "all") && sc.biased_objects) {
        // The Cornell box and the final scene allocate a light list of TWO objects (ceiling light, glass sphere) but pass
        // count 1 (scene.cpp:326-329, 456-459).  "-lights all" uses both, so that sphere::pdf_value / pdf_generate
        // (sphere.cpp:63-79) and random_towards_sphere (pcg.cpp:125-133) -- dead code in all nine stock scenes -- run.
        if (p.sceneSelect == 5 || p.sceneSelect == 7) ((object_list<scene_object> *) sc.biased_objects)->count = 2;
    }

    // regular sample grid, main.cpp:319-332
    uint32 sq = (uint32) MRT::sqrt((float) spp);
    uint32 N = sq * sq;
    std::vector<vec2> sd(N);
    for (uint32 i = 0; i < sq; i++)
        for (uint32 j = 0; j < sq; j++) {
            sd[i * sq + j].x = (i + 0.5f) / (float) sq;
            sd[i * sq + j].y = (j + 0.5f) / (float) sq;
        }
    uint32 s0 = strtoul(argval(argc, argv, "-s0", "0"), 0, 0);
    uint32 s1 = strtoul(argval(argc, argv, "-s1", "0"), 0, 0);
    if (s1 == 0 || s1 > N) s1 = N;

    uint32 x0 = strtoul(argval(argc, argv, "-x0", "0"), 0, 0), x1 = strtoul(argval(argc, argv, "-x1", "0"), 0, 0);
    uint32 y0 = strtoul(argval(argc, argv, "-y0", "0"), 0, 0), y1 = strtoul(argval(argc, argv, "-y1", "0"), 0, 0);
    if (!x1 || x1 > W) x1 = W;
    if (!y1 || y1 > H) y1 = H;
    if (x0 >= x1 || y0 >= y1) { fprintf(stderr, "bad crop window\n"); return 2; }
    const uint32 CW = x1 - x0, CH = y1 - y0;

    std::vector<float> acc((size_t) CW * CH * 4, 0.0f);
    std::atomic<uint32> nextRow(y0);
    G_rayCounter = 0;
    // -draw2 1: the pixel update of the reference's default worker, draw2 (main.cpp:214-231), instead of the plain sum:
    // running mean over the samples in order, a non-finite sample replaced by the mean so far (0 for the first), and the
    // luminance clamp applied after EVERY sample, feeding back into the mean.  Output: rgb = final mean, w = 1.
    const bool draw2_mode = atoi(argval(argc, argv, "-draw2", "0")) != 0;

    uint64 t0 = MRT_GetTime();
    auto worker = [&]() {
        for (;;) {
            uint32 y = nextRow.fetch_add(1);
            if (y >= y1) break;
            for (uint32 x = x0; x < x1; x++) {
                Vec3 color(0, 0, 0);
                uint32 cnt = 0;
                for (uint32 s = s0; s < s1 && !draw2_mode; s++) {
                    // one private PCG32 stream per (pixel, sample)
                    Init_Thread_RNG(seed, ((uint64) y * W + x) * N + s);
                    float u = (x + sd[s].x) / (float) W;   // main.cpp:156-157
                    float v = (y + sd[s].y) / (float) H;
                    ray r = sc.camera->get_ray(u, v);
                    Vec3 sample = trace(r, *sc.objects, sc.biased_objects, 0);
                    if (isfinite(sample.r) && isfinite(sample.g) && isfinite(sample.b)) {
                        color += sample;
                        cnt++;
                    }
                }
                float *o = &acc[((size_t) (y - y0) * CW + (x - x0)) * 4];
                if (draw2_mode) {
                    Vec3 mean(0.0f);
                    for (uint32 s = s0; s < s1; s++) {
                        Init_Thread_RNG(seed, ((uint64) y * W + x) * N + s);
                        float u = (x + sd[s].x) / (float) W;
                        float v = (y + sd[s].y) / (float) H;
                        ray r = sc.camera->get_ray(u, v);
                        Vec3 c = trace(r, *sc.objects, sc.biased_objects, 0);
                        const uint32 sampleCount = s - s0;
                        if (!isfinite(c.r) || !isfinite(c.g) || !isfinite(c.b)) c = sampleCount > 0 ? mean : Vec3(0.0f);   // main.cpp:214-219
                        if (sampleCount > 0) c = mean + (c - mean) * (1.0f / (sampleCount + 1.0f));                         // main.cpp:221-224
                        float lum = luminance(c);
                        if (lum > p.maxLuminance) c = c * (p.maxLuminance / lum);                                           // main.cpp:226-229
                        mean = c;
                    }
                    o[0] = mean.r; o[1] = mean.g; o[2] = mean.b; o[3] = 1.0f;
                    continue;
                }
                o[0] = color.r; o[1] = color.g; o[2] = color.b; o[3] = (float) cnt;
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32 i = 0; i < nthreads; i++) th.emplace_back(worker);
    for (auto &t : th) t.join();
    double secs = (MRT_GetTime() - t0) / 1e9;

    uint64 rays = G_rayCounter;
    double paths = (double) CW * CH * (s1 - s0);
    printf("{\"mode\":\"render\",\"scene\":%u,\"width\":%u,\"height\":%u,\"samples\":%u,\"s0\":%u,\"s1\":%u,"
           "\"depth\":%u,\"threads\":%u,\"seconds\":%.6f,\"rays\":%llu,\"paths\":%.0f,"
           "\"mrays_per_s\":%.4f,\"mpaths_per_s\":%.4f}\n",
           p.sceneSelect, W, H, N, s0, s1, p.maxBounces, nthreads, secs, (unsigned long long) rays, paths,
           rays / secs * 1e-6, paths / secs * 1e-6);

    if (out) {
        FileHeader h;
        memset(&h, 0, sizeof(h));
        memcpy(h.magic, "MRTACC1", 8);
        h.width = CW; h.height = CH; h.samples = N; h.s0 = s0; h.s1 = s1; h.depth = p.maxBounces;
        h.scene = p.sceneSelect; h.threads = nthreads; h.seed = seed; h.rays = rays; h.seconds = secs;
        return write_acc(out, h, acc.data());
    }
    return 0;
}

// --------------------------------------------------------------------- stock
static int cmd_stock(int argc, char **argv) {
    const char *dump = nullptr, *dump_argb = nullptr;
    std::vector<char *> args;
    args.push_back(argv[0]);
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "-dump") && i + 1 < argc) { dump = argv[++i]; continue; }
        if (!strcmp(argv[i], "-dumpargb") && i + 1 < argc) { dump_argb = argv[++i]; continue; }
        args.push_back(argv[i]);
    }
    MRT_headless_quiet = true;
    uint64 t0 = MRT_GetTime();
    ref_main((int) args.size(), args.data());
    double wall = (MRT_GetTime() - t0) / 1e9;
    MRT_Params *p = getParams();
    // title = "MiniRayTracer - Scene: Xms - Trace: %.2fs - %.3f Mrays/s | %.3f us/ray" (main.cpp:403)
    const char *title = MRT_headless_last_title();
    float trace_s = 0, mrays = 0;
    const char *t = strstr(title, "Trace: ");
    if (t) sscanf(t, "Trace: %fs - %f Mrays/s", &trace_s, &mrays);
    uint32 sq = (uint32) MRT::sqrt((float) p->samplesPerPixel);
    double paths = (double) p->bufferWidth * p->bufferHeight * sq * sq;
    printf("{\"mode\":\"stock\",\"scene\":%u,\"width\":%u,\"height\":%u,\"samples\":%u,\"depth\":%u,"
           "\"threads\":%u,\"threading_mode\":%u,\"trace_seconds\":%.3f,\"wall_seconds\":%.3f,\"rays\":%llu,"
           "\"paths\":%.0f,\"mrays_per_s\":%.4f,\"mpaths_per_s\":%.4f}\n",
           p->sceneSelect, p->bufferWidth, p->bufferHeight, sq * sq, p->maxBounces, p->numThreads,
           p->threadingMode, trace_s, wall, (unsigned long long) (size_t) G_rayCounter, paths,
           trace_s > 0 ? (size_t) G_rayCounter / trace_s * 1e-6 : 0.0, trace_s > 0 ? paths / trace_s * 1e-6 : 0.0);
    if (dump_argb) {   // G_backBuffer: the reference's own tone map of its final frame (main.cpp:416-444), raw uint32 ARGB
        FILE *f = fopen(dump_argb, "wb");
        if (!f) { perror(dump_argb); return 1; }
        fwrite(G_backBuffer, sizeof(uint32), (size_t) p->bufferWidth * p->bufferHeight, f);
        fclose(f);
    }
    if (dump) {
        uint32 W = p->bufferWidth, H = p->bufferHeight;
        std::vector<float> acc((size_t) W * H * 4);
        for (size_t i = 0; i < (size_t) W * H; i++) {
            acc[i * 4 + 0] = G_linearBackBuffer[i].r;
            acc[i * 4 + 1] = G_linearBackBuffer[i].g;
            acc[i * 4 + 2] = G_linearBackBuffer[i].b;
            acc[i * 4 + 3] = 1.0f; // already a mean (main.cpp:168)
        }
        FileHeader h;
        memset(&h, 0, sizeof(h));
        memcpy(h.magic, "MRTACC1", 8);
        h.width = W; h.height = H; h.samples = sq * sq; h.s0 = 0; h.s1 = sq * sq; h.depth = p->maxBounces;
        h.scene = p->sceneSelect; h.threads = p->numThreads; h.rays = G_rayCounter; h.seconds = trace_s;
        return write_acc(dump, h, acc.data());
    }
    return 0;
}

#ifndef MRT_REF_TIMING_ONLY
// ---------------------------------------------------------------- dump-scene
static int cmd_dump_scene(int argc, char **argv) {
    uint32 W = strtoul(argval(argc, argv, "-width", "500"), 0, 0);
    uint32 H = strtoul(argval(argc, argv, "-height", "500"), 0, 0);
    uint32 sel = strtoul(argval(argc, argv, "-scene", "0"), 0, 0);
    const char *out = argval(argc, argv, "-out", nullptr);
    MRT_headless_quiet = true;
    scene sc = build_scene(sel, W, H);
    if (!strcmp(argval(argc, argv, "-extra", "none"), "triangles") && sel == 5) add_extra_triangles(sc);
    if (!strcmp(argval(argc, argv, "-lights", "ref"), "all") && sc.biased_objects && (sel == 5 || sel == 7))
        ((object_list<scene_object> *) sc.biased_objects)->count = 2;   // see cmd_render
    FILE *f = out ? fopen(out, "w") : stdout;
    if (!f) { perror(out); return 1; }
    dump_scene(f, sc);
    if (out) fclose(f);
    return 0;
}

// ----------------------------------------------------------------------- kat
static int cmd_kat() {
    // PCG32 known answers (pcg.cpp:13-62)
    Init_Thread_RNG(42, 54);
    printf("pcg32 seed=42 seq=54 rand32:");
    for (int i = 0; i < 6; i++) printf(" %08x", rand32());
    printf("\n");
    Init_Thread_RNG(11350390909718046443uLL, 6305599193148252115uLL);
    printf("pcg32 mainseed rand32:");
    for (int i = 0; i < 4; i++) printf(" %08x", rand32());
    printf("\n");
    Init_Thread_RNG(11350390909718046443uLL, 6305599193148252115uLL);
    printf("pcg32 mainseed randf:");
    for (int i = 0; i < 4; i++) printf(" %08x", fbits(randf()));
    printf("\n");
    Init_Thread_RNG(42, 54);
    Vec3 s = random_in_sphere();
    printf("random_in_sphere seed=42,54: %08x %08x %08x\n", fbits(s.x), fbits(s.y), fbits(s.z));
    Init_Thread_RNG(42, 54);
    Vec3 dk = random_in_disk();
    printf("random_in_disk seed=42,54: %08x %08x %08x\n", fbits(dk.x), fbits(dk.y), fbits(dk.z));
    Init_Thread_RNG(42, 54);
    Vec3 c = random_cosine_direction();
    printf("random_cosine_direction seed=42,54: %08x %08x %08x\n", fbits(c.x), fbits(c.y), fbits(c.z));
    // Perlin tables (texture.cpp:167-203), built at static-init time from G_rng (pcg.cpp:40)
    printf("perlin ranvec:");
    for (int i = 0; i < 256; i++) printf(" %08x,%08x,%08x", fbits(rv[i].x), fbits(rv[i].y), fbits(rv[i].z));
    printf("\nperlin perm_x:");
    for (int i = 0; i < 256; i++) printf(" %d", px[i]);
    printf("\nperlin perm_y:");
    for (int i = 0; i < 256; i++) printf(" %d", py[i]);
    printf("\nperlin perm_z:");
    for (int i = 0; i < 256; i++) printf(" %d", pz[i]);
    printf("\n");
    // a few texture / noise samples
    perlin_noise pn;
    for (int i = 0; i < 8; i++) {
        Vec3 p(0.37f * i - 1.3f, 1.91f * i + 0.2f, -0.77f * i + 3.1f);
        printf("perlin p=%08x,%08x,%08x noise=%08x turb=%08x\n", fbits(p.x), fbits(p.y), fbits(p.z),
               fbits(pn.noise(p)), fbits(pn.turbulence(p)));
    }
    return 0;
}

static int cmd_dump_image(int argc, char **argv) {
    if (argc < 3) return 2;
    int w, h, ch;
    uint8 *pixels = stbi_load("../earthmap.jpg", &w, &h, &ch, 3);
    if (!pixels) { fprintf(stderr, "cannot decode ../earthmap.jpg\n"); return 1; }
    FILE *f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 1; }
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    fwrite(pixels, 3, (size_t) w * h, f);
    fclose(f);
    printf("wrote %s %dx%d fnv=%08x\n", argv[2], w, h, crc_bytes(pixels, (size_t) w * h * 3));
    return 0;
}

#endif   // MRT_REF_TIMING_ONLY

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s render|stock|dump-scene|kat|dump-image ...\n", argv[0]);
        return 2;
    }
    if (!strcmp(argv[1], "render")) return cmd_render(argc, argv);
    if (!strcmp(argv[1], "stock")) return cmd_stock(argc, argv);
#ifndef MRT_REF_TIMING_ONLY
    if (!strcmp(argv[1], "dump-scene")) return cmd_dump_scene(argc, argv);
    if (!strcmp(argv[1], "kat")) return cmd_kat();
    if (!strcmp(argv[1], "dump-image")) return cmd_dump_image(argc, argv);
#endif
    fprintf(stderr, "unknown command %s\n", argv[1]);
    return 2;
}
