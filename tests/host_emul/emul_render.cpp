// TEST-ONLY logic emulation: compiles the device tracer core (trace_core.h) with
// g++ and runs it on the flattened scene, so that traversal / shading logic can
// be checked against the reference oracle without a GPU.  This binary is built
// only by the test-suite; the shipped library has no CPU execution path.
//
// usage: emul_render -scene S -width W -height H -samples N -depth D -seed X
//                    [-s0 a -s1 b] [-x0 a -x1 b -y0 c -y1 d] [-threads T] -assets DIR -out file.bin [-counters] [-nocull]
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define MRT_COUNT_OPS 1   // algorithmic operation counters (trace_core.h: MRT_OP) -- this test binary only
#include "scene_graph.h"
#include "trace_core.h"

using namespace mrt;
namespace mrt { thread_local OpCounts mrt_ops; }

static const char *argval(int argc, char **argv, const char *name, const char *def) {
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return argv[i + 1];
    return def;
}
static bool argflag(int argc, char **argv, const char *name) {
    for (int i = 1; i < argc; i++) if (!strcmp(argv[i], name)) return true;
    return false;
}

struct FileHeader {
    char magic[8];
    uint32_t width, height, samples, s0, s1, depth, scene, threads;
    uint64_t seed, rays;
    double seconds;
};

int main(int argc, char **argv) {
    uint32_t W = strtoul(argval(argc, argv, "-width", "500"), 0, 0);
    uint32_t H = strtoul(argval(argc, argv, "-height", "500"), 0, 0);
    uint32_t spp = strtoul(argval(argc, argv, "-samples", "16"), 0, 0);
    uint32_t depth = strtoul(argval(argc, argv, "-depth", "32"), 0, 0);
    uint32_t scene = strtoul(argval(argc, argv, "-scene", "0"), 0, 0);
    uint64_t seed = strtoull(argval(argc, argv, "-seed", "11350390909718046443"), 0, 0);
    uint32_t nthreads = strtoul(argval(argc, argv, "-threads", "0"), 0, 0);
    if (!nthreads) nthreads = std::thread::hardware_concurrency();
    std::string assets = argval(argc, argv, "-assets", "assets");
    const char *out = argval(argc, argv, "-out", nullptr);
    bool want_counters = argflag(argc, argv, "-counters");
    const bool lite = argflag(argc, argv, "-specialised");   // run with the scene's own feature mask instead of MRT_FEAT_ALL
    FlattenOptions fopt;
    fopt.cull_boxes = !argflag(argc, argv, "-nocull");       // A/B of the translate cull boxes (must not change a bit)

    SceneGraph g;
    if (!build_scene(g, scene, float(W) / float(H), assets)) { fprintf(stderr, "scene: %s\n", g.error.c_str()); return 1; }
    FlatScene fs;
    if (!flatten_scene(g, &fs, fopt)) { fprintf(stderr, "flatten: %s\n", fs.error.c_str()); return 1; }
    const MrtSceneDesc &d = fs.desc;
    const uint32_t feat = lite ? (d.features ? d.features : MRT_FEAT_ALL) : MRT_FEAT_ALL;
    SceneView sv;
    sv.sphere = d.sphere; sv.rect = d.rect; sv.list = d.list; sv.bvh = d.bvh; sv.node2 = d.node2; sv.trileaf = d.trileaf; sv.tri = d.tri; sv.trin = d.trin;
    sv.xlate = d.xlate; sv.rot = d.rot; sv.vol = d.vol; sv.mat = d.mat; sv.tex = d.tex; sv.perlin_vec = d.perlin_vec;
    sv.child = d.child; sv.lights = d.lights; sv.perlin_perm = d.perlin_perm; sv.image = d.image;
    sv.root = d.root; sv.n_lights = d.n_lights; sv.sky = d.sky; sv.cam = d.camera;

    uint32_t sq = (uint32_t) sqrtf((float) spp);
    uint32_t N = sq * sq;
    uint32_t s0 = strtoul(argval(argc, argv, "-s0", "0"), 0, 0);
    uint32_t s1 = strtoul(argval(argc, argv, "-s1", "0"), 0, 0);
    if (s1 == 0 || s1 > N) s1 = N;

    // crop window: pixels [x0,x1) x [y0,y1) of the W x H frame (stream ids and u,v stay those of the full frame)
    uint32_t x0 = strtoul(argval(argc, argv, "-x0", "0"), 0, 0), x1 = strtoul(argval(argc, argv, "-x1", "0"), 0, 0);
    uint32_t y0 = strtoul(argval(argc, argv, "-y0", "0"), 0, 0), y1 = strtoul(argval(argc, argv, "-y1", "0"), 0, 0);
    if (!x1) x1 = W;
    if (!y1) y1 = H;
    const uint32_t CW = x1 - x0, CH = y1 - y0;
    std::vector<float> acc((size_t) CW * CH * 4, 0.0f);
    std::atomic<uint32_t> nextRow(y0);
    std::atomic<unsigned long long> rays(0);
    Counters total;
    memset(&total, 0, sizeof(total));
    std::vector<Counters> per_thread(nthreads);
    std::vector<OpCounts> ops_thread(nthreads);
    auto worker = [&](uint32_t tid) {
        std::vector<uint32_t> stack_mem(d.stack_words + 8);
        Counters cnt;
        memset(&cnt, 0, sizeof(cnt));
        for (;;) {
            uint32_t y = nextRow.fetch_add(1);
            if (y >= y1) break;
            for (uint32_t x = x0; x < x1; x++) {
                V3 color = v3(0, 0, 0);
                uint32_t n_ok = 0;
                for (uint32_t s = s0; s < s1; s++) {
                    Rng rng;
                    Path p;
                    path_begin(sv, p, rng, x, y, s, sq, W, H, seed);
                    for (;;) {
                        path_advance(feat, sv, p);
                        cnt.rays++;
                        Hit rec;
                        Stack st;
                        st.base = stack_mem.data(); st.stride = 1; st.sp = 0;
                        bool hit;
                        hit = intersect(feat, sv, p.ray, 0.001f, FLT_MAX, rec, rng, st, want_counters ? &cnt : nullptr);
                        if (st.sp != 0) { fprintf(stderr, "stack imbalance\n"); abort(); }
                        if (!path_shade(feat, sv, p, hit, rec, depth, rng)) break;
                    }
                    if (path_sample_finite(p)) {
                        color = color + p.L;
                        n_ok++;
                    }
                }
                float *o = &acc[((size_t) (y - y0) * CW + (x - x0)) * 4];
                o[0] = color.x; o[1] = color.y; o[2] = color.z; o[3] = (float) n_ok;
            }
        }
        per_thread[tid] = cnt;
        ops_thread[tid] = mrt_ops;
    };
    std::vector<std::thread> th;
    for (uint32_t i = 0; i < nthreads; i++) th.emplace_back(worker, i);
    for (auto &t : th) t.join();
    for (auto &c : per_thread) {
        total.rays += c.rays; total.aabb += c.aabb; total.sphere += c.sphere; total.rect += c.rect;
        total.tri += c.tri; total.vol += c.vol; total.xform += c.xform;
    }
    OpCounts ops;
    memset(&ops, 0, sizeof(ops));
    for (auto &o : ops_thread) {
        const unsigned long long *src = reinterpret_cast<const unsigned long long *>(&o);
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(&ops);
        for (size_t i = 0; i < sizeof(OpCounts) / sizeof(unsigned long long); i++) dst[i] += src[i];
    }
    printf("{\"mode\":\"emul\",\"scene\":%u,\"rays\":%llu,\"aabb\":%llu,\"sphere\":%llu,\"rect\":%llu,\"tri\":%llu,\"vol\":%llu,"
           "\"xform\":%llu,\"stack_words\":%u,\"paths\":%llu,\"ray_ctor\":%llu,\"rng\":%llu,\"sphere_hit\":%llu,\"sphere_moving\":%llu,"
           "\"rect_hit\":%llu,\"tri_hit\":%llu,\"translate\":%llu,\"rotate\":%llu,\"lambert\":%llu,\"metal\":%llu,\"dielectric\":%llu,"
           "\"isotropic\":%llu,\"lightpdf\":%llu,\"perlin\":%llu,\"image\":%llu,\"checker\":%llu,\"sky\":%llu}\n",
           scene, total.rays, total.aabb, total.sphere, total.rect, total.tri, total.vol, total.xform, d.stack_words, ops.paths, ops.ray_ctor, ops.rng,
           ops.sphere_hit, ops.sphere_moving, ops.rect_hit, ops.tri_hit, ops.translate, ops.rotate, ops.lambert, ops.metal, ops.dielectric,
           ops.isotropic, ops.lightpdf, ops.perlin, ops.image, ops.checker, ops.sky);
    if (out) {
        FileHeader h;
        memset(&h, 0, sizeof(h));
        memcpy(h.magic, "MRTACC1", 8);
        h.width = CW; h.height = CH; h.samples = N; h.s0 = s0; h.s1 = s1; h.depth = depth; h.scene = scene; h.threads = nthreads;
        h.seed = seed; h.rays = total.rays;
        FILE *f = fopen(out, "wb");
        if (!f) { perror(out); return 1; }
        fwrite(&h, sizeof(h), 1, f);
        fwrite(acc.data(), sizeof(float) * 4, (size_t) CW * CH, f);
        fclose(f);
    }
    return 0;
}
