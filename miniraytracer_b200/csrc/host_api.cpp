// Host half of the C ABI (include/mrt_gpu.h): error string, command-line
// parameters (cmdline_parser.cpp restated: same options, defaults, range checks
// and warnings) and scene construction / flattening.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <string>

#include "mrt_gpu.h"
#include "scene_graph.h"

namespace mrt {
static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
}  // namespace mrt

using namespace mrt;

extern "C" const char *mrt_last_error(void) { return g_error.c_str(); }

// ------------------------------------------------------------------ parameters
extern "C" void mrt_params_default(MrtParams *p) {   // cmdline_parser.h:5-18
    memset(p, 0, sizeof(*p));
    p->window_width = 500;
    p->window_height = 500;
    p->buffer_width = 500;
    p->buffer_height = 500;
    p->samples_per_pixel = 128;
    p->tile_size = 32;
    p->num_threads = 0;
    p->max_bounces = 32;
    p->scene_select = 8;   // SCENE_TRIANGLES
    p->threading_mode = 1;
    p->max_luminance = 1000;
    p->delay = 0;
    p->num_gpus = 1;
    p->seed = 11350390909718046443ull;   // main.cpp:302 initstate
    strcpy(p->asset_dir, "assets");
}

namespace {
template <typename T> T read_value(char *arg);
template <> float read_value<float>(char *arg) { return strtof(arg, nullptr); }
template <> uint32_t read_value<uint32_t>(char *arg) { return (uint32_t) strtoul(arg, nullptr, 0); }
template <> uint64_t read_value<uint64_t>(char *arg) { return (uint64_t) strtoull(arg, nullptr, 0); }

// cmdline_parser.cpp:40-62
template <typename T>
int read_parameter(int argc, char **argv, const char *parameter, T *res, T min = std::numeric_limits<T>::min(),
                   T max = std::numeric_limits<T>::max()) {
    for (int i = 1; i < argc; i++) {
        if (strcmp(parameter, argv[i]) == 0) {
            if ((i + 1) == argc) {
                std::cout << "Warning: Missing value for parameter '" << parameter << "'." << std::endl;
                return 0;
            }
            T p = read_value<T>(argv[i + 1]);
            if ((p < min) || (p > max)) {
                std::cout << "Warning: Invalid value for parameter '" << parameter << "', must be in [" << min << ", " << max << "]."
                          << std::endl;
                return 0;
            }
            *res = p;
            return i;
        }
    }
    return 0;
}
int check_parameter(int argc, char **argv, const char *parameter) {   // cmdline_parser.cpp:64-71
    for (int i = 1; i < argc; i++)
        if (strcmp(parameter, argv[i]) == 0) return i;
    return 0;
}
int read_string(int argc, char **argv, const char *parameter, char *dst, size_t cap) {
    int i = check_parameter(argc, argv, parameter);
    if (!i) return 0;
    if (i + 1 == argc) {
        std::cout << "Warning: Missing value for parameter '" << parameter << "'." << std::endl;
        return 0;
    }
    strncpy(dst, argv[i + 1], cap - 1);
    dst[cap - 1] = 0;
    return i;
}
void print_help() {   // cmdline_parser.cpp:107-122 (+ the new options)
    printf("\n"
           "PARAMETERS:\n"
           "  -width    \t<value>\t\tWindow width\n"
           "  -height   \t<value>\t\tWindow height\n"
           "  -samples  \t<value>\t\tSamples per pixel\n"
           "  -depth    \t<value>\t\tMaximum bounce depth per primary ray\n"
           "  -maxlum   \t<value>\t\tClamp maximum luminance (introduces bias)\n"
           "  -threads  \t<value>\t\tNumber of execution threads (0 selects maximum hardware threads)\n"
           "  -tilesize \t<value>\t\tSize of image tiles (threads operate on tiles)\n"
           "  -mode     \t[0, 1]\t\tThreading/queue mode (0 for sequential, 1 for dynamic sampling)\n"
           "  -scene    \t[0, %i]\t\tSelect the scene\n"
           "  -delay    \t\t\tDelay start until keypress\n"
           "  -gpus     \t<value>\t\tNumber of GPUs (samples per pixel are split across them)\n"
           "  -seed     \t<value>\t\tPCG32 initstate of the per-(pixel, sample) streams\n"
           "  -assets   \t<dir>\t\tDirectory with earthmap.ppm and obj/\n"
           "  -out      \t<file>\t\tWrite the image (.ppm tone-mapped, .pfm linear)\n",
           8);
}
}  // namespace

extern "C" int mrt_params_parse(int argc, char **argv, MrtParams *out) {
    if (check_parameter(argc, argv, "-help") || check_parameter(argc, argv, "--help") || check_parameter(argc, argv, "-?")) {
        print_help();
        return 1;
    }
    MrtParams p;
    mrt_params_default(&p);
    if (read_parameter<uint32_t>(argc, argv, "-width", &p.window_width, 1u)) p.buffer_width = p.window_width;
    if (read_parameter<uint32_t>(argc, argv, "-height", &p.window_height, 1u)) p.buffer_height = p.window_height;
    read_parameter<uint32_t>(argc, argv, "-samples", &p.samples_per_pixel, 1u);
    read_parameter<uint32_t>(argc, argv, "-tilesize", &p.tile_size, 1u);
    read_parameter<uint32_t>(argc, argv, "-threads", &p.num_threads);
    read_parameter<uint32_t>(argc, argv, "-depth", &p.max_bounces);
    read_parameter<uint32_t>(argc, argv, "-scene", &p.scene_select, 0u, 8u);
    read_parameter<uint32_t>(argc, argv, "-mode", &p.threading_mode, 0u, 1u);
    read_parameter<float>(argc, argv, "-maxlum", &p.max_luminance);
    if (check_parameter(argc, argv, "-delay")) p.delay = 1;
    read_parameter<uint32_t>(argc, argv, "-gpus", &p.num_gpus, 1u, 64u);
    read_parameter<uint64_t>(argc, argv, "-seed", &p.seed);
    read_string(argc, argv, "-assets", p.asset_dir, sizeof(p.asset_dir));
    read_string(argc, argv, "-out", p.out_path, sizeof(p.out_path));
    *out = p;
    return 0;
}

// ----------------------------------------------------------------------- scene
struct MrtHostScene {
    SceneGraph graph;
    FlatScene flat;
    bool has_graph = true;
};

extern "C" int mrt_scene_create(uint32_t scene, float aspect, const char *asset_dir, MrtHostScene **out) {
    if (!out) { set_error("mrt_scene_create: null argument"); return MRT_E_INVALID; }
    *out = nullptr;
    if ((scene & 0xFFu) > 8 || (scene & ~(0xFFu | MRT_SCENE_ALL_LIGHTS))) { set_error("mrt_scene_create: scene must be in [0, 8] (| MRT_SCENE_ALL_LIGHTS)"); return MRT_E_INVALID; }
    MrtHostScene *s = new (std::nothrow) MrtHostScene();
    if (!s) { set_error("out of memory"); return MRT_E_INVALID; }
    if (!build_scene(s->graph, scene, aspect, asset_dir ? asset_dir : "assets")) {
        set_error("scene construction failed: " + s->graph.error);
        delete s;
        return MRT_E_SCENE;
    }
    if (!flatten_scene(s->graph, &s->flat)) {
        set_error("scene flattening failed: " + s->flat.error);
        delete s;
        return MRT_E_SCENE;
    }
    *out = s;
    return MRT_OK;
}

extern "C" const MrtSceneDesc *mrt_scene_desc(const MrtHostScene *s) { return s ? &s->flat.desc : nullptr; }

extern "C" int mrt_scene_dump(const MrtHostScene *s, const char *path) {
    if (!s || !path) { set_error("mrt_scene_dump: null argument"); return MRT_E_INVALID; }
    if (!s->has_graph) { set_error("mrt_scene_dump: scene was loaded from a flattened file (no graph)"); return MRT_E_STATE; }
    FILE *f = fopen(path, "w");
    if (!f) { set_error(std::string("cannot open ") + path); return MRT_E_INVALID; }
    dump_scene(s->graph, f);
    fclose(f);
    return MRT_OK;
}

// ---------------------------------------------------------------- scene file
namespace {
const char kMagic[8] = {'M', 'R', 'T', 'S', 'C', 'N', '1', 0};
template <typename T> bool put(FILE *f, const std::vector<T> &v) {
    uint64_t n = v.size();
    return fwrite(&n, sizeof(n), 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == n);
}
template <typename T> bool get(FILE *f, std::vector<T> &v) {
    uint64_t n = 0;
    if (fread(&n, sizeof(n), 1, f) != 1 || n > (1ull << 34) / sizeof(T)) return false;
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}
void rebind(FlatScene &o) {   // pointers of the description follow the vectors
    MrtSceneDesc &d = o.desc;
    d.lights = o.lights.data();
    d.sphere = o.sphere.data(); d.rect = o.rect.data(); d.list = o.list.data(); d.child = o.child.data();
    d.bvh = o.bvh.data(); d.node2 = o.node2.data(); d.trileaf = o.trileaf.data(); d.tri = o.tri.data(); d.trin = o.trin.data();
    d.xlate = o.xlate.data(); d.rot = o.rot.data(); d.vol = o.vol.data(); d.mat = o.mat.data(); d.tex = o.tex.data();
    d.perlin_vec = o.perlin_vec.empty() ? nullptr : o.perlin_vec.data();
    d.perlin_perm = o.perlin_perm.empty() ? nullptr : o.perlin_perm.data();
    d.image = o.image.empty() ? nullptr : o.image.data();
}
}  // namespace

extern "C" int mrt_scene_save(const MrtHostScene *s, const char *path) {
    if (!s || !path) { set_error("mrt_scene_save: null argument"); return MRT_E_INVALID; }
    FILE *f = fopen(path, "wb");
    if (!f) { set_error(std::string("cannot open ") + path); return MRT_E_INVALID; }
    const FlatScene &o = s->flat;
    MrtSceneDesc d = o.desc;   // scalar part; pointers are meaningless in the file
    bool ok = fwrite(kMagic, 8, 1, f) == 1 && fwrite(&d, sizeof(d), 1, f) == 1;
    ok = ok && put(f, o.sphere) && put(f, o.rect) && put(f, o.list) && put(f, o.child) && put(f, o.bvh) && put(f, o.node2) &&
         put(f, o.trileaf) && put(f, o.tri) && put(f, o.trin) && put(f, o.xlate) && put(f, o.rot) && put(f, o.vol) && put(f, o.mat) &&
         put(f, o.tex) && put(f, o.perlin_vec) && put(f, o.perlin_perm) && put(f, o.image) && put(f, o.lights);
    fclose(f);
    if (!ok) { set_error(std::string("short write to ") + path); return MRT_E_INVALID; }
    return MRT_OK;
}

extern "C" int mrt_scene_load(const char *path, MrtHostScene **out) {
    if (!path || !out) { set_error("mrt_scene_load: null argument"); return MRT_E_INVALID; }
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) { set_error(std::string("cannot open ") + path); return MRT_E_SCENE; }
    MrtHostScene *s = new (std::nothrow) MrtHostScene();
    if (!s) { fclose(f); set_error("out of memory"); return MRT_E_INVALID; }
    s->has_graph = false;
    FlatScene &o = s->flat;
    char magic[8];
    bool ok = fread(magic, 8, 1, f) == 1 && memcmp(magic, kMagic, 8) == 0 && fread(&o.desc, sizeof(o.desc), 1, f) == 1;
    ok = ok && get(f, o.sphere) && get(f, o.rect) && get(f, o.list) && get(f, o.child) && get(f, o.bvh) && get(f, o.node2) &&
         get(f, o.trileaf) && get(f, o.tri) && get(f, o.trin) && get(f, o.xlate) && get(f, o.rot) && get(f, o.vol) && get(f, o.mat) &&
         get(f, o.tex) && get(f, o.perlin_vec) && get(f, o.perlin_perm) && get(f, o.image) && get(f, o.lights);
    fclose(f);
    const MrtSceneDesc &d = o.desc;
    ok = ok && o.sphere.size() == (size_t) d.n_sphere * 3 && o.rect.size() == (size_t) d.n_rect * 2 && o.list.size() == (size_t) d.n_list * 2 &&
         o.child.size() == d.n_child && o.bvh.size() == (size_t) d.n_bvh * 2 && o.node2.size() == (size_t) d.n_node2 * 4 &&
         o.trileaf.size() == (size_t) d.n_trileaf * 2 && o.tri.size() == (size_t) d.n_tri * 3 && o.trin.size() == o.tri.size() &&
         o.xlate.size() == (size_t) d.n_xlate * 3 && o.rot.size() == (size_t) d.n_rot * 3 && o.vol.size() == d.n_vol && o.mat.size() == d.n_mat &&
         o.tex.size() == d.n_tex && o.image.size() == d.n_image_bytes && o.lights.size() == d.n_lights;
    if (!ok) { delete s; set_error(std::string("not a valid MRTSCN1 file: ") + path); return MRT_E_SCENE; }
    rebind(o);
    *out = s;
    return MRT_OK;
}

extern "C" void mrt_scene_free(MrtHostScene *s) { delete s; }
