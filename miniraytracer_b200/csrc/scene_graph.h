// Host-side scene description: a tagged node graph that mirrors the reference's
// scene_object / material / texture class hierarchy (scene_object.h, sphere.h,
// rect.h, box.h, triangle.h, volumes.h, material.h, texture.h) without virtual
// dispatch, plus the constructors' build-time logic (bounding boxes, the
// median-split bvh_node build, the pod_bvh build, rotate_y's box) restated so
// that the resulting topology is identical to the reference's -- the traversal
// result depends on it.
//
// Everything here is host-only C++20 and is compiled with -ffp-contract=off.
#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "mrt_gpu.h"

namespace mrt {

struct H3 {   // Vec3 with the reference's operation order (vec3.h)
    float x = 0, y = 0, z = 0;
    H3() = default;
    H3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float operator[](size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline H3 operator+(H3 a, H3 b) { return H3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline H3 operator-(H3 a, H3 b) { return H3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline H3 operator*(H3 a, float s) { return H3(a.x * s, a.y * s, a.z * s); }
inline H3 operator*(float s, H3 a) { return H3(a.x * s, a.y * s, a.z * s); }
inline H3 operator/(H3 a, float s) { return H3(a.x / s, a.y / s, a.z / s); }
inline float hdot(H3 a, H3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline H3 hcross(H3 a, H3 b) { return H3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
float hlength(H3 a);
H3 hnormalize(H3 a);
inline H3 hmin(H3 a, H3 b) { return H3(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z); }   // _mm_min_ps
inline H3 hmax(H3 a, H3 b) { return H3(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z); }   // _mm_max_ps

struct Aabb { H3 min, max; };

// PCG32 (pcg.cpp:13-62), host copy used for scene construction and the Perlin tables
struct HostRng {
    uint64_t state = 0, inc = 0;
    void seed(uint64_t initstate, uint64_t initseq);
    uint32_t next();
    float randf();
    H3 in_sphere();
};

enum class NodeKind : uint8_t { Sphere, RectXY, RectXZ, RectYZ, Box, List, Bvh, Translate, RotateY, Volume, PodBvh, Triangle };
enum class MatKind : uint8_t { Lambertian, Isotropic, Metal, Dielectric, Light };
enum class TexKind : uint8_t { Color, Checker, Perlin, Image };

struct Texture {
    TexKind kind = TexKind::Color;
    H3 color;
    int even = -1, odd = -1;
    float scale = 1;
    int image = -1;
};
struct Material {
    MatKind kind = MatKind::Lambertian;
    int tex = -1;
    float param = 0;   // metal gloss / dielectric ref_index / light scale
};
struct Image {
    int width = 0, height = 0;
    std::vector<uint8_t> rgb;
};
struct Triangle {   // triangle.h:13-22 (m, u = b - a, v = c - a)
    H3 m, u, v, mn, un, vn;
};
struct PodNode {    // triangle.h:46-56
    Aabb box;
    uint32_t left = 0, prim_offset = 0, prim_count = 0;
    uint8_t order = 0;
};
struct Mesh {
    std::vector<Triangle> tris;   // in pod_bvh order after the build
    std::vector<PodNode> nodes;
    int mat = -1;
};

struct Node {
    NodeKind kind = NodeKind::Sphere;
    // sphere
    H3 c0, c1;
    float t0 = 0, t1 = 0, radius = 0;
    bool moving = false;
    // rect: in-plane ranges a0..a1, b0..b1, plane k, normal sign
    float a0 = 0, a1 = 0, b0 = 0, b1 = 0, k = 0, sign = 1;
    int mat = -1;
    // box / list / bvh / rotate_y
    Aabb box;
    bool has_box = false;
    std::vector<int> children;   // list
    int left = -1, right = -1;   // bvh
    uint8_t order = 0;
    int child = -1;              // box (its rect list), translate, rotate_y, volume boundary
    H3 offset;                   // translate
    float sin_theta = 0, cos_theta = 0;
    float density = 0;           // volume (mat = isotropic phase function)
    int mesh = -1;               // pod_bvh
    Triangle tri;                // triangle_scene_object (triangle.cpp:5-175): a lone triangle as a scene object
};

struct Camera {   // camera.h
    H3 origin, u, v, w, llcorner, horz, vert;
    float lens_radius = 0, time0 = 0, time1 = 0;
    Camera() = default;
    Camera(H3 pos, H3 lookat, H3 up, float vfov, float aspect, float aperture, float focus_dist, float t0, float t1);
};

struct SceneGraph {
    std::vector<Node> nodes;
    std::vector<Material> mats;
    std::vector<Texture> texs;
    std::vector<Image> images;
    std::vector<Mesh> meshes;
    Camera camera;
    int objects = -1;   // scene.objects
    int biased = -1;    // scene.biased_objects (a List) or -1
    bool sky = false;   // sceneSelect < SCENE_CORNELL_BOX (main.cpp:110)
    bool uses_perlin = false;
    std::string error;

    // textures / materials
    int color_tex(H3 c);
    int checker_tex(int even, int odd, float scale);
    int perlin_tex(float scale);
    int image_tex(int image);
    int lambertian(int tex);
    int isotropic(int tex);
    int metal(int tex, float gloss);
    int dielectric(float ref_index);
    int diffuse_light(int tex, float scale = 1.0f);

    // objects (constructors of the reference classes)
    int sphere(H3 c0, float r, int mat, H3 c1 = H3(0, 0, 0), float t0 = 0.0f, float t1 = 0.0f);
    int triangle(H3 a, H3 b, H3 c, int mat);                                   // triangle.cpp:5-17 (face normal)
    int triangle(H3 a, H3 b, H3 c, H3 an, H3 bn, H3 cn, int mat);             // triangle.cpp:19-35 (vertex normals)
    int xy_rect(float x0, float x1, float y0, float y1, float z, int mat);
    int xz_rect(float x0, float x1, float z0, float z1, float y, int mat);
    int yz_rect(float y0, float y1, float z0, float z1, float x, int mat);
    int box(H3 min, H3 max, int mat);
    int list(const std::vector<int> &items, float time0, float time1);
    int bvh(std::vector<int> &items, size_t begin, size_t n, float time0, float time1);   // sorts items[begin..begin+n) in place
    int translate(int obj, H3 offset);
    int rotate_y(int obj, float angle_deg);
    int volume(int boundary, float density, int albedo_tex);
    int pod_bvh(std::vector<Triangle> &&tris, int mat);

    bool bounding_box(int node, float t0, float t1, Aabb *out) const;
};

// Perlin tables (texture.cpp:167-203), generated from the pre-seeded global generator (pcg.cpp:40)
struct PerlinTables {
    float ranvec[256][3];
    int32_t perm[3][256];
};
const PerlinTables &perlin_tables();

// scenes (scene.cpp) -- `scene` uses the reference's enum values (scene.h:6-17)
bool build_scene(SceneGraph &g, uint32_t scene_and_flags, float aspect, const std::string &asset_dir);   // flags: MRT_SCENE_ALL_LIGHTS (mrt_gpu.h)

// obj_loader.cpp restated; M4 is the reference's column-major Mat4 (mat4.h), c[col][row]
struct M4 {
    float c[4][4];
    static M4 identity();
    static M4 scale(float s);          // mat4.cpp:164-169
    static M4 rotate_y(float radians); // mat4.cpp:237-244
    static M4 invert(const M4 &m);     // mat4.cpp:60-125 (the cofactor path that is compiled)
    H3 mul_col(H3 v) const;            // Mat4 * Vec3, mat4.h:44-53
    H3 mul_row(H3 v) const;            // Vec3 * Mat4 (row vector), mat4.h:68-103
};
bool read_obj(const std::string &path, bool flip, const M4 &scale, H3 translate, const M4 &rotate,
              std::vector<Triangle> *out);

// canonical text dump (same format as oracle/ref_harness.cpp dump-scene)
void dump_scene(const SceneGraph &g, FILE *f);

// flattener: graph -> mrt_types.h tables
struct FlatScene {
    std::vector<MrtF4> sphere, rect, list, bvh, node2, tri, trin, xlate, rot, vol, mat, tex, perlin_vec;
    std::vector<uint32_t> child, lights, trileaf;
    std::vector<int32_t> perlin_perm;
    std::vector<uint8_t> image;
    MrtSceneDesc desc;
    std::string error;
};
struct FlattenOptions {
    bool cull_boxes = true;   // conservative cull boxes on translate nodes (trace_core.h: cull_miss); off = A/B for the parity test
};
bool flatten_scene(const SceneGraph &g, FlatScene *out, const FlattenOptions &opt = FlattenOptions());
bool coop_trees_supported(const MrtSceneDesc &d);
bool validate_scene_desc(const MrtSceneDesc &d, std::string *err, uint32_t *stack_words_needed);   // structural check (scene files, caller-built descriptions)   // do the BVH trees qualify for the warp-cooperative traversal (coop_tree.cuh)?

}  // namespace mrt
