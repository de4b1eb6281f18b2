// Host half of the C ABI (include/mrt_gpu.h): error string, command-line
// parameters (cmdline_parser.cpp restated: same options, defaults, range checks
// and warnings) and scene construction / flattening.
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <new>
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

// Option table: one descriptor per command-line option -- the reference's options (cmdline_parser.cpp:90-104: same names,
// same defaults via mrt_params_default, same accepted ranges) and the four this front end adds.  One pass over argv; like the
// reference, the FIRST occurrence of an option counts, a value outside the range or a missing value leaves the default in
// place and prints a warning, and arguments that are not options are ignored.
namespace {
enum class OptType { U32, U64, F32, Flag, Text };
struct OptDesc {
    const char *name;
    OptType type;
    size_t offset;        // field of MrtParams
    double lo, hi;        // accepted range (numbers)
    size_t text_cap;      // Text: capacity of the destination
    const char *operand;  // help: what follows the option
    const char *help;
};
#define MRT_OPT_FIELD(f) offsetof(MrtParams, f)
constexpr double kU32Max = 4294967295.0;
const OptDesc kOptions[] = {
    {"-width", OptType::U32, MRT_OPT_FIELD(window_width), 1, kU32Max, 0, "<pixels>", "image width"},
    {"-height", OptType::U32, MRT_OPT_FIELD(window_height), 1, kU32Max, 0, "<pixels>", "image height"},
    {"-samples", OptType::U32, MRT_OPT_FIELD(samples_per_pixel), 1, kU32Max, 0, "<n>", "samples per pixel (rounded down to a square number)"},
    {"-depth", OptType::U32, MRT_OPT_FIELD(max_bounces), 0, kU32Max, 0, "<n>", "bounce limit of a path"},
    {"-maxlum", OptType::F32, MRT_OPT_FIELD(max_luminance), 1.17549435e-38 /* numeric_limits<float>::min(), cmdline_parser.cpp:41 */, 3.4028234e38, 0, "<x>", "luminance above which a pixel is scaled down (biased)"},
    {"-threads", OptType::U32, MRT_OPT_FIELD(num_threads), 0, kU32Max, 0, "<n>", "CPU worker threads of the reference; accepted, unused by the GPU renderer"},
    {"-tilesize", OptType::U32, MRT_OPT_FIELD(tile_size), 1, kU32Max, 0, "<pixels>", "tile edge of the reference's work queue; accepted, unused"},
    {"-mode", OptType::U32, MRT_OPT_FIELD(threading_mode), 0, 1, 0, "0|1", "0: all samples in one launch, 1: progressive passes"},
    {"-scene", OptType::U32, MRT_OPT_FIELD(scene_select), 0, 8, 0, "0..8", "which of the nine scenes"},
    {"-delay", OptType::Flag, MRT_OPT_FIELD(delay), 0, 0, 0, "", "wait for a key press before starting (reference option; no effect headless)"},
    {"-gpus", OptType::U32, MRT_OPT_FIELD(num_gpus), 1, 64, 0, "<n>", "GPUs to split the samples per pixel across"},
    {"-seed", OptType::U64, MRT_OPT_FIELD(seed), 0, 1.8446744073709552e19, 0, "<n>", "PCG32 initstate of the per-(pixel, sample) streams"},
    {"-assets", OptType::Text, MRT_OPT_FIELD(asset_dir), 0, 0, sizeof(MrtParams::asset_dir), "<dir>", "directory holding earthmap.ppm and obj/"},
    {"-out", OptType::Text, MRT_OPT_FIELD(out_path), 0, 0, sizeof(MrtParams::out_path), "<file>", "image file: .ppm (tone mapped) or .pfm (linear)"},
};
constexpr size_t kNumOptions = sizeof(kOptions) / sizeof(kOptions[0]);

const OptDesc *find_option(const char *arg, size_t *index) {
    for (size_t i = 0; i < kNumOptions; i++)
        if (!strcmp(arg, kOptions[i].name)) { *index = i; return &kOptions[i]; }
    return nullptr;
}
bool wants_help(const char *arg) { return !strcmp(arg, "-help") || !strcmp(arg, "--help") || !strcmp(arg, "-?"); }

void print_usage() {
    printf("usage: mrt_b200 [options]\n");
    for (size_t i = 0; i < kNumOptions; i++) {
        const OptDesc &o = kOptions[i];
        printf("  %-10s %-9s %s", o.name, o.operand, o.help);
        if (o.type == OptType::U32 && o.hi < kU32Max) printf(" [%.0f, %.0f]", o.lo, o.hi);
        printf("\n");
    }
}

// Stores the operand of option `o` into `p`; false (with a message) if it is not acceptable.
bool store_operand(const OptDesc &o, const char *text, MrtParams *p) {
    char *field = reinterpret_cast<char *>(p) + o.offset;
    double as_number = 0;
    switch (o.type) {
    case OptType::U32: { const uint32_t v = (uint32_t) strtoul(text, nullptr, 0); as_number = v; if (as_number >= o.lo && as_number <= o.hi) { memcpy(field, &v, sizeof(v)); return true; } break; }
    case OptType::U64: { const uint64_t v = (uint64_t) strtoull(text, nullptr, 0); memcpy(field, &v, sizeof(v)); return true; }
    case OptType::F32: { const float v = strtof(text, nullptr); as_number = v; if (as_number >= o.lo && as_number <= o.hi) { memcpy(field, &v, sizeof(v)); return true; } break; }
    case OptType::Text: snprintf(field, o.text_cap, "%s", text); return true;
    case OptType::Flag: return true;
    }
    printf("warning: %s %s is outside [%.9g, %.9g]; keeping the default\n", o.name, text, o.lo, o.hi);
    return false;
}
}  // namespace

extern "C" int mrt_params_parse(int argc, char **argv, MrtParams *out) {
    for (int i = 1; i < argc; i++)
        if (wants_help(argv[i])) { print_usage(); return 1; }
    MrtParams p;
    mrt_params_default(&p);
    bool seen[kNumOptions] = {};
    for (int i = 1; i < argc; i++) {
        size_t k = 0;
        const OptDesc *o = find_option(argv[i], &k);
        if (!o) continue;                      // not an option: ignored, as in the reference
        const bool first = !seen[k];
        seen[k] = true;
        if (o->type == OptType::Flag) {
            if (first) { const uint32_t one = 1; memcpy(reinterpret_cast<char *>(&p) + o->offset, &one, sizeof(one)); }
            continue;
        }
        if (i + 1 >= argc) { if (first) printf("warning: %s needs a value; keeping the default\n", o->name); break; }
        i++;                                   // the operand is consumed even when a repeated option is ignored
        if (first) store_operand(*o, argv[i], &p);
    }
    // the reference renders into a buffer of the window's size (cmdline_parser.cpp:92-93)
    p.buffer_width = p.window_width;
    p.buffer_height = p.window_height;
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
    if ((scene & 0xFFu) > 8 || (scene & ~(0xFFu | MRT_SCENE_ALL_LIGHTS | MRT_SCENE_EXTRA_TRIANGLES))) { set_error("mrt_scene_create: scene must be in [0, 8] (| MRT_SCENE_ALL_LIGHTS)"); return MRT_E_INVALID; }
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
// "MRTSCN2": magic, sizeof(MrtSceneDesc) of the writer, the description with every pointer field zeroed, then each table as
// (uint64 count, raw records).  A file is checked structurally before it is accepted (validate_scene_desc): table sizes against
// the counts, every reference inside its table, terminated child runs, Perlin tables of 256 / 768 entries, stack depth.
namespace {
const char kMagic[8] = {'M', 'R', 'T', 'S', 'C', 'N', '2', 0};
template <typename T> bool put(FILE *f, const std::vector<T> &v) {
    uint64_t n = v.size();
    return fwrite(&n, sizeof(n), 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == n);
}
template <typename T> bool get(FILE *f, std::vector<T> &v, uint64_t max_count) {
    uint64_t n = 0;
    if (fread(&n, sizeof(n), 1, f) != 1 || n > max_count) return false;
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}
void clear_pointers(MrtSceneDesc &d) {
    d.lights = nullptr; d.sphere = d.rect = d.list = d.bvh = d.node2 = d.tri = d.trin = d.xlate = d.rot = d.vol = d.mat = d.tex = d.perlin_vec = nullptr;
    d.child = d.trileaf = nullptr; d.perlin_perm = nullptr; d.image = nullptr;
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
bool read_scene_file(FILE *f, FlatScene &o) {
    char magic[8];
    uint64_t desc_size = 0;
    if (fread(magic, 8, 1, f) != 1 || memcmp(magic, kMagic, 8) != 0) return false;
    if (fread(&desc_size, sizeof(desc_size), 1, f) != 1 || desc_size != sizeof(MrtSceneDesc)) return false;
    if (fread(&o.desc, sizeof(o.desc), 1, f) != 1) return false;
    clear_pointers(o.desc);
    const MrtSceneDesc &d = o.desc;
    // every table is read with its count bounded by what the description announces (no attacker-sized allocations)
    return get(f, o.sphere, (uint64_t) d.n_sphere * 3) && get(f, o.rect, (uint64_t) d.n_rect * 2) && get(f, o.list, (uint64_t) d.n_list * 2) &&
           get(f, o.child, d.n_child) && get(f, o.bvh, (uint64_t) d.n_bvh * 2) && get(f, o.node2, (uint64_t) d.n_node2 * 4) &&
           get(f, o.trileaf, (uint64_t) d.n_trileaf * 2) && get(f, o.tri, (uint64_t) d.n_tri * 3) && get(f, o.trin, (uint64_t) d.n_tri * 3) &&
           get(f, o.xlate, (uint64_t) d.n_xlate * 3) && get(f, o.rot, (uint64_t) d.n_rot * 3) && get(f, o.vol, d.n_vol) && get(f, o.mat, d.n_mat) &&
           get(f, o.tex, d.n_tex) && get(f, o.perlin_vec, 256) && get(f, o.perlin_perm, 768) && get(f, o.image, d.n_image_bytes) &&
           get(f, o.lights, d.n_lights);
}
}  // namespace

extern "C" int mrt_scene_save(const MrtHostScene *s, const char *path) {
    if (!s || !path) { set_error("mrt_scene_save: null argument"); return MRT_E_INVALID; }
    FILE *f = fopen(path, "wb");
    if (!f) { set_error(std::string("cannot open ") + path); return MRT_E_INVALID; }
    const FlatScene &o = s->flat;
    MrtSceneDesc d = o.desc;   // scalar part; host pointers are meaningless in a file and are not written
    clear_pointers(d);
    const uint64_t desc_size = sizeof(MrtSceneDesc);
    bool ok = fwrite(kMagic, 8, 1, f) == 1 && fwrite(&desc_size, sizeof(desc_size), 1, f) == 1 && fwrite(&d, sizeof(d), 1, f) == 1;
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
    bool ok = false;
    std::string why = "not a valid MRTSCN2 file";
    try {   // nothing may throw across the C ABI (a table count the machine cannot allocate)
        ok = read_scene_file(f, o);
    } catch (const std::bad_alloc &) {
        ok = false;
        why = "out of memory";
    }
    fclose(f);
    const MrtSceneDesc &d = o.desc;
    ok = ok && o.sphere.size() == (size_t) d.n_sphere * 3 && o.rect.size() == (size_t) d.n_rect * 2 && o.list.size() == (size_t) d.n_list * 2 &&
         o.child.size() == d.n_child && o.bvh.size() == (size_t) d.n_bvh * 2 && o.node2.size() == (size_t) d.n_node2 * 4 &&
         o.trileaf.size() == (size_t) d.n_trileaf * 2 && o.tri.size() == (size_t) d.n_tri * 3 && o.trin.size() == o.tri.size() &&
         o.xlate.size() == (size_t) d.n_xlate * 3 && o.rot.size() == (size_t) d.n_rot * 3 && o.vol.size() == d.n_vol && o.mat.size() == d.n_mat &&
         o.tex.size() == d.n_tex && o.image.size() == d.n_image_bytes && o.lights.size() == d.n_lights &&
         (o.perlin_vec.empty() || o.perlin_vec.size() == 256) && (o.perlin_perm.empty() || o.perlin_perm.size() == 768) &&
         o.perlin_vec.empty() == o.perlin_perm.empty();
    if (ok) {
        rebind(o);
        std::string detail;
        ok = validate_scene_desc(o.desc, &detail, nullptr);
        if (!ok) why = detail;
    }
    if (!ok) { delete s; set_error(why + ": " + path); return MRT_E_SCENE; }
    *out = s;
    return MRT_OK;
}

extern "C" void mrt_scene_free(MrtHostScene *s) { delete s; }
