// TEST-ONLY: runs the mode-B render kernel (render_kernels.cuh: render_pixel_binned) on the CPU through the SIMT shim
// of emul_warp.h, one block of four warps, and writes the accumulator like emul_render.  Lets the not-gpu test-suite
// check the warp scheduling logic -- ticket queue, bins, path pool, sample staging, chunk epilogue -- against the oracle.
//
// usage: emul_binned -scene S -width W -height H -samples N [-depth D] [-seed X] [-s0 a -s1 b] [-chunk K] [-bins B]
//                    [-x0 a -x1 b -y0 c -y1 d] [-mode B|W|P] -assets DIR -out file.bin        (W / P: the pool-less kernels of the same header)
#include "emul_warp.h"

#include <cmath>
#include <string>

#include "scene_graph.h"
#include "render_kernels.cuh"
#include "schedule.h"

using namespace mrt;

namespace mrt { void set_error(const std::string &) {} }

static const char *argval(int argc, char **argv, const char *name, const char *def) {
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return argv[i + 1];
    return def;
}
struct FileHeader {
    char magic[8];
    uint32_t width, height, samples, s0, s1, depth, scene, threads;
    uint64_t seed, rays;
    double seconds;
};

static RenderArgs g_args;
static char g_mode = 'B';
static bool g_coop = false;
static bool g_seq = false;   // launches with <= kStageBlock samples per pixel use the SEQ instantiation (render_kernel.cu)
static bool g_trees = true;   // scenes without BVH trees run a mask without MRT_FEAT_TREES, like the product's specialised variants
template <uint32_t FEAT>
static void entry_b() {
    if (g_seq) { if (g_coop) render_pixel_binned<FEAT, 6, true, true>(g_args); else render_pixel_binned<FEAT, 6, false, true>(g_args); }
    else if (g_coop) render_pixel_binned<FEAT, 6, true, false>(g_args);
    else render_pixel_binned<FEAT, 6, false, false>(g_args);
}
static void entry(void *) {
    if (g_mode == 'W') render_pixel_per_warp<MRT_FEAT_ALL, 6>(g_args);
    else if (g_mode == 'P') render_pixel_per_lane<MRT_FEAT_ALL, 6>(g_args);
    else if (g_trees) entry_b<MRT_FEAT_ALL>();
    else entry_b<(MRT_FEAT_ALL & ~(MRT_FEAT_TREES | MRT_FEAT_TRIS))>();
}

int main(int argc, char **argv) {
    uint32_t W = strtoul(argval(argc, argv, "-width", "16"), 0, 0), H = strtoul(argval(argc, argv, "-height", "9"), 0, 0);
    uint32_t spp = strtoul(argval(argc, argv, "-samples", "16"), 0, 0), depth = strtoul(argval(argc, argv, "-depth", "32"), 0, 0);
    uint32_t scene = strtoul(argval(argc, argv, "-scene", "5"), 0, 0);
    uint64_t seed = strtoull(argval(argc, argv, "-seed", "11350390909718046443"), 0, 0);
    uint32_t chunk = strtoul(argval(argc, argv, "-chunk", "0"), 0, 0), bins = strtoul(argval(argc, argv, "-bins", "2"), 0, 0);
    g_mode = argval(argc, argv, "-mode", "B")[0];
    g_coop = atoi(argval(argc, argv, "-coop", "0")) != 0;   // warp-cooperative tree traversal (coop_tree.cuh)
    std::string assets = argval(argc, argv, "-assets", "assets");
    const char *out = argval(argc, argv, "-out", nullptr);

    SceneGraph g;
    if (!build_scene(g, scene, float(W) / float(H), assets)) { fprintf(stderr, "scene: %s\n", g.error.c_str()); return 1; }
    FlatScene fs;
    if (!flatten_scene(g, &fs)) { fprintf(stderr, "flatten: %s\n", fs.error.c_str()); return 1; }
    const MrtSceneDesc &d = fs.desc;
    RenderArgs &a = g_args;
    memset(&a, 0, sizeof(a));
    SceneView &sv = a.sc;
    sv.sphere = d.sphere; sv.rect = d.rect; sv.list = d.list; sv.bvh = d.bvh; sv.node2 = d.node2; sv.trileaf = d.trileaf; sv.tri = d.tri; sv.trin = d.trin;
    sv.xlate = d.xlate; sv.rot = d.rot; sv.vol = d.vol; sv.mat = d.mat; sv.tex = d.tex; sv.perlin_vec = d.perlin_vec;
    sv.child = d.child; sv.lights = d.lights; sv.perlin_perm = d.perlin_perm; sv.image = d.image;
    sv.root = d.root; sv.n_lights = d.n_lights; sv.sky = d.sky; sv.cam = d.camera;

    uint32_t sq = (uint32_t) sqrtf((float) spp);
    uint32_t N = sq * sq;
    uint32_t s0 = strtoul(argval(argc, argv, "-s0", "0"), 0, 0), s1 = strtoul(argval(argc, argv, "-s1", "0"), 0, 0);
    if (s1 == 0 || s1 > N) s1 = N;
    uint32_t x0 = strtoul(argval(argc, argv, "-x0", "0"), 0, 0), x1 = strtoul(argval(argc, argv, "-x1", "0"), 0, 0);
    uint32_t y0 = strtoul(argval(argc, argv, "-y0", "0"), 0, 0), y1 = strtoul(argval(argc, argv, "-y1", "0"), 0, 0);
    if (!x1) x1 = W;
    if (!y1) y1 = H;
    const uint32_t CW = x1 - x0, CH = y1 - y0;
    const uint32_t ns = s1 - s0, n_pixels = CW * CH;
    a.crop_x0 = x0; a.crop_y0 = y0; a.crop_w = CW; a.n_pixels = n_pixels;
    a.width = W; a.height = H; a.sqrt_n = sq; a.s_begin = s0; a.s_end = s1; a.max_bounces = depth; a.seed = seed;
    g_seq = (s1 - s0) <= kStageBlock;
    g_trees = d.n_node2 != 0;
    a.accumulate = 0;
    a.stack_words = d.stack_words ? d.stack_words : 64;
    if (g_coop) {
        if (!d.stack_words_coop) { fprintf(stderr, "the scene's trees do not qualify for the cooperative traversal\n"); return 1; }
        a.stack_words = d.stack_words_coop;
    }
    uint32_t K = chunk ? chunk : (256u / ns ? 256u / ns : 1u);   // small chunks: several tasks per warp even on a tiny frame
    if ((uint64_t) K * ns > kMaxStageItems) K = kMaxStageItems / ns;
    if (K < 1) K = 1;
    if (g_mode == 'W') { if (ns < 32) { fprintf(stderr, "mode W needs >= 32 samples\n"); return 1; } K = chunk ? chunk : 2u; }
    if (g_mode == 'P') K = chunk ? ((chunk + 31u) & ~31u) : 32u;
    a.pixels_per_task = K;
    a.n_tasks = (n_pixels + K - 1) / K;
    const uint32_t plan_warps = strtoul(argval(argc, argv, "-plan", "0"), 0, 0);
    if (plan_warps && g_mode == 'B') {
        // -plan W: the shipped planner (schedule.h, as called by mrt_gpu_render_async) for W resident warps
        MrtTuning tn; memset(&tn, 0, sizeof(tn));
        tn.chunk_pixels = chunk;
        tn.chunk_paths = strtoul(argval(argc, argv, "-chunkpaths", "0"), 0, 0);
        const BinnedPlan pl = plan_binned_schedule(n_pixels, ns, plan_warps, d.n_node2 != 0, g_coop, tn, kMaxStageItems);
        K = pl.K;
        a.pixels_per_task = K;
        for (int r = 0; r < 3; r++) { a.sched_task0[r] = pl.task0[r]; a.sched_pix0[r] = pl.pix0[r]; a.sched_k[r] = pl.k[r]; }
        a.sched_pix0[3] = pl.pix0[3];
        a.n_tasks = pl.n_tasks;
    } else {   // ad hoc: -tail P = the last P pixels in chunks of K/4 and K/16
        uint32_t tail = strtoul(argval(argc, argv, "-tail", "0"), 0, 0);
        if (tail > n_pixels) tail = n_pixels;
        const uint32_t k1 = K / 4u ? K / 4u : 1u, k2 = K / 16u ? K / 16u : 1u;
        const uint32_t p1 = n_pixels - tail, p2 = p1 + tail * 2u / 3u;
        const uint32_t t1 = (p1 + K - 1u) / K, t2 = t1 + (p2 - p1 + k1 - 1u) / k1;
        a.sched_task0[0] = 0; a.sched_task0[1] = t1; a.sched_task0[2] = t2;
        a.sched_pix0[0] = 0; a.sched_pix0[1] = p1; a.sched_pix0[2] = p2; a.sched_pix0[3] = n_pixels;
        a.sched_k[0] = K; a.sched_k[1] = k1; a.sched_k[2] = k2;
        if (g_mode == 'B') a.n_tasks = t2 + (n_pixels - p2 + k2 - 1u) / k2;
    }
    // bins: classifier boxes are a grouping heuristic, any box will do -- the first rotate_y's bounds if the scene has one
    a.n_cls_boxes = 0; a.cls_pending = 0;
    if (bins >= 2 && d.n_rot) {
        a.n_cls_boxes = 1;
        const MrtF4 &b0 = d.rot[0], &b1 = d.rot[1];
        a.cls_box[0][0] = b0.x; a.cls_box[0][1] = b0.y; a.cls_box[0][2] = b0.z; a.cls_box[0][3] = b1.x; a.cls_box[0][4] = b1.y; a.cls_box[0][5] = b1.z;
        if (bins >= 3) a.cls_pending = 1;
    }
    a.n_bins = 1u << (a.n_cls_boxes + a.cls_pending);
    a.coop_leaf_batch = strtoul(argval(argc, argv, "-leafbatch", "16"), 0, 0);
    std::vector<float4> acc(n_pixels, make_float4(0, 0, 0, 0));
    a.acc = acc.data();
    unsigned int ticket = 0;
    unsigned long long counters[12] = {0};
    int cancel = 0;
    a.ticket = &ticket; a.counters = counters; a.cancel = &cancel; a.order = nullptr;
    const uint32_t warps = kWarpsPerBlock;
    std::vector<uint32_t> pool((size_t) warps * kPoolCap * kStateWords, 0u);
    a.stage_items = K * ns;
    std::vector<float4> stage((size_t) warps * a.stage_items, make_float4(0, 0, 0, 0));
    a.pool = pool.data(); a.stage = stage.data();
    const size_t smem_words = (size_t) warps * a.stack_words * 32u + ((size_t) warps * (a.n_bins + 1u) * kPoolCap + 3u) / 4u +
                              (g_mode == 'W' ? (size_t) warps * K * 32u * 4u : 0u) + (g_coop ? (size_t) warps * kCoopWords : 0u);
    if (smem_words > sizeof(smem_stack) / 4) { fprintf(stderr, "shared memory emulation too small\n"); return 1; }

    emul::run_block(entry, nullptr, kBlock, 0);

    if (out) {
        FILE *f = fopen(out, "wb");
        if (!f) { fprintf(stderr, "cannot write %s\n", out); return 1; }
        FileHeader h;
        memset(&h, 0, sizeof(h));
        memcpy(h.magic, "MRTACC1", 8);
        h.width = CW; h.height = CH; h.samples = N; h.s0 = s0; h.s1 = s1; h.depth = depth; h.scene = scene; h.threads = 1;
        h.seed = seed; h.rays = counters[0]; h.seconds = 0;
        fwrite(&h, sizeof(h), 1, f);
        fwrite(acc.data(), sizeof(float4), acc.size(), f);
        fclose(f);
    }
    printf("{\"rays\": %llu, \"iterations\": %llu, \"nonfinite\": %llu, \"collectives\": %llu, \"tasks\": %u, \"bins\": %u, "
           "\"coop_node_steps\": %llu, \"coop_node_items\": %llu, \"coop_leaf_steps\": %llu, \"coop_leaf_items\": %llu}\n", counters[0], counters[1], counters[2],
           emul::g_collectives, a.n_tasks, a.n_bins, counters[4], counters[5], counters[6], counters[7]);
    return 0;
}
