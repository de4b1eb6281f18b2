// Host scene graph: constructors, bounding boxes and the two BVH builders of the
// reference, restated over tagged nodes (see scene_graph.h).
#include "scene_graph.h"
#include "mrt_libm.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>

namespace mrt {

float hlength(H3 a) { return sqrtf(hdot(a, a)); }            // Vec4::length, vec3.h:128-131
H3 hnormalize(H3 a) { return a / hlength(a); }              // vec3.h:133-139

// ------------------------------------------------------------------ PCG32
void HostRng::seed(uint64_t initstate, uint64_t initseq) {   // pcg.cpp:28-35
    state = 0u;
    inc = (initseq << 1u) | 1u;
    next();
    state += initstate;
    next();
}
uint32_t HostRng::next() {                                   // pcg.cpp:13-26
    uint64_t old = state;
    state = old * 6364136223846793005ULL + inc;
    uint32_t xorshifted = (uint32_t) (((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t) (old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
}
float HostRng::randf() {                                     // pcg.cpp:53-62
    uint32_t bits = 0x3f800000u | (next() & 0x007FFFFFu);
    float f;
    memcpy(&f, &bits, 4);
    return f - 1.0f;
}
H3 HostRng::in_sphere() {                                    // pcg.cpp:70-77, draws in x, y, z order
    H3 p;
    do {
        float rx = randf(), ry = randf(), rz = randf();
        p = H3(2.0f * rx - 1.0f, 2.0f * ry - 1.0f, 2.0f * rz - 1.0f);
    } while (hdot(p, p) >= 1.0f);
    return p;
}

// Perlin tables: texture.cpp:167-203, from the raw (un-seeded) G_rng state of pcg.cpp:40
const PerlinTables &perlin_tables() {
    static PerlinTables t;
    static bool init = false;
    if (!init) {
        HostRng g;
        g.state = 11350390909718046443uLL;
        g.inc = 6305599193148252115uLL;
        for (int i = 0; i < 256; i++) {
            H3 v = g.in_sphere();
            t.ranvec[i][0] = v.x; t.ranvec[i][1] = v.y; t.ranvec[i][2] = v.z;
        }
        for (int a = 0; a < 3; a++) {
            for (int i = 0; i < 256; i++) t.perm[a][i] = i;
            for (int i = 255; i > 0; i--) {
                int target = int(g.randf() * (i + 1));
                std::swap(t.perm[a][i], t.perm[a][target]);
            }
        }
        init = true;
    }
    return t;
}

// ----------------------------------------------------------------- camera
#define H_PI_F 3.14159265358979323846f
#define H_RAD(a) ((a) * (H_PI_F / 180.0f))

Camera::Camera(H3 pos, H3 lookat, H3 up, float vfov, float aspect, float aperture, float focus_dist, float t0, float t1) {
    // camera.h:16-36
    time0 = t0;
    time1 = t1;
    float theta = H_RAD(vfov);
    // the reference folds this tanf at compile time (correctly rounded); evaluate it the same way
    float height = 2.0f * (float) tan((double) (theta / 2));
    float width = aspect * height;
    origin = pos;
    w = hnormalize(pos - lookat);
    u = hnormalize(hcross(up, w));
    v = hcross(w, u);
    lens_radius = aperture / 2.0f;
    horz = focus_dist * width * u;
    vert = focus_dist * height * v;
    llcorner = origin - 0.5f * horz - 0.5f * vert - focus_dist * w;
}

// ------------------------------------------------------ textures / materials
int SceneGraph::color_tex(H3 c) { Texture t; t.kind = TexKind::Color; t.color = c; texs.push_back(t); return (int) texs.size() - 1; }
int SceneGraph::checker_tex(int even, int odd, float scale) {
    Texture t; t.kind = TexKind::Checker; t.even = even; t.odd = odd; t.scale = scale; texs.push_back(t); return (int) texs.size() - 1;
}
int SceneGraph::perlin_tex(float scale) {
    Texture t; t.kind = TexKind::Perlin; t.scale = scale; texs.push_back(t); uses_perlin = true; return (int) texs.size() - 1;
}
int SceneGraph::image_tex(int image) { Texture t; t.kind = TexKind::Image; t.image = image; texs.push_back(t); return (int) texs.size() - 1; }
int SceneGraph::lambertian(int tex) { Material m; m.kind = MatKind::Lambertian; m.tex = tex; mats.push_back(m); return (int) mats.size() - 1; }
int SceneGraph::isotropic(int tex) { Material m; m.kind = MatKind::Isotropic; m.tex = tex; mats.push_back(m); return (int) mats.size() - 1; }
int SceneGraph::metal(int tex, float gloss) {
    Material m; m.kind = MatKind::Metal; m.tex = tex; m.param = std::min(gloss, 1.0f);   // material.h:86-88
    mats.push_back(m); return (int) mats.size() - 1;
}
int SceneGraph::dielectric(float ref_index) { Material m; m.kind = MatKind::Dielectric; m.param = ref_index; mats.push_back(m); return (int) mats.size() - 1; }
int SceneGraph::diffuse_light(int tex, float scale) { Material m; m.kind = MatKind::Light; m.tex = tex; m.param = scale; mats.push_back(m); return (int) mats.size() - 1; }

// ------------------------------------------------------------------ objects
int SceneGraph::sphere(H3 c0, float r, int mat, H3 c1, float t0, float t1) {   // sphere.h:19-23
    Node n; n.kind = NodeKind::Sphere; n.c0 = c0; n.c1 = c1; n.t0 = t0; n.t1 = t1; n.radius = r; n.mat = mat;
    n.moving = (t1 - t0) > std::numeric_limits<float>::epsilon();
    nodes.push_back(n); return (int) nodes.size() - 1;
}
static int make_rect(SceneGraph &g, NodeKind kind, float a0, float a1, float b0, float b1, float k, int mat) {
    // rect.cpp:6-22,51-67,112-128: argument order decides the normal sign
    Node n; n.kind = kind; n.k = k; n.mat = mat; n.sign = 1;
    if (a0 > a1) { n.sign *= -1; std::swap(a0, a1); }
    if (b0 > b1) { n.sign *= -1; std::swap(b0, b1); }
    n.a0 = a0; n.a1 = a1; n.b0 = b0; n.b1 = b1;
    g.nodes.push_back(n); return (int) g.nodes.size() - 1;
}
int SceneGraph::xy_rect(float x0, float x1, float y0, float y1, float z, int mat) { return make_rect(*this, NodeKind::RectXY, x0, x1, y0, y1, z, mat); }
int SceneGraph::xz_rect(float x0, float x1, float z0, float z1, float y, int mat) { return make_rect(*this, NodeKind::RectXZ, x0, x1, z0, z1, y, mat); }
int SceneGraph::triangle(H3 a, H3 b, H3 c, int mat) {   // triangle.cpp:5-17: m, u = b - a, v = c - a, one face normal
    Node n;
    n.kind = NodeKind::Triangle;
    n.mat = mat;
    n.tri.m = a; n.tri.u = b - a; n.tri.v = c - a;
    n.tri.mn = n.tri.un = n.tri.vn = hnormalize(hcross(n.tri.u, n.tri.v));
    nodes.push_back(n);
    return (int) nodes.size() - 1;
}
int SceneGraph::triangle(H3 a, H3 b, H3 c, H3 an, H3 bn, H3 cn, int mat) {   // triangle.cpp:19-35
    Node n;
    n.kind = NodeKind::Triangle;
    n.mat = mat;
    n.tri.m = a; n.tri.u = b - a; n.tri.v = c - a;
    n.tri.mn = an; n.tri.un = bn; n.tri.vn = cn;
    nodes.push_back(n);
    return (int) nodes.size() - 1;
}
int SceneGraph::yz_rect(float y0, float y1, float z0, float z1, float x, int mat) { return make_rect(*this, NodeKind::RectYZ, y0, y1, z0, z1, x, mat); }

int SceneGraph::box(H3 mn, H3 mx, int mat) {   // box.h:12-21
    std::vector<int> l(6);
    l[0] = xy_rect(mn.x, mx.x, mn.y, mx.y, mx.z, mat);
    l[1] = xy_rect(mx.x, mn.x, mn.y, mx.y, mn.z, mat);
    l[2] = xz_rect(mn.x, mx.x, mn.z, mx.z, mx.y, mat);
    l[3] = xz_rect(mx.x, mn.x, mn.z, mx.z, mn.y, mat);
    l[4] = yz_rect(mn.y, mx.y, mn.z, mx.z, mx.x, mat);
    l[5] = yz_rect(mx.y, mn.y, mn.z, mx.z, mn.x, mat);
    int rl = list(l, 0, 0);
    Node n; n.kind = NodeKind::Box; n.box.min = mn; n.box.max = mx; n.child = rl;
    nodes.push_back(n); return (int) nodes.size() - 1;
}

static H3 sphere_center(const Node &s, float time) {   // sphere.h:24-31
    if (s.moving) return s.c0 + ((time - s.t0) / (s.t1 - s.t0)) * (s.c1 - s.c0);
    return s.c0;
}

bool SceneGraph::bounding_box(int id, float t0, float t1, Aabb *out) const {
    const Node &n = nodes[id];
    switch (n.kind) {
    case NodeKind::Sphere: {   // sphere.cpp:48-61
        float abs_r = fabsf(n.radius);
        H3 r(abs_r, abs_r, abs_r);
        H3 c0 = sphere_center(n, t0), c1 = sphere_center(n, t1);
        Aabb bb0{c0 - r, c0 + r}, bb1{c1 - r, c1 + r};
        out->min = hmin(bb0.min, bb1.min);   // surrounding_box, aabb.h:108-110
        out->max = hmax(bb0.max, bb1.max);
        return true;
    }
    case NodeKind::RectXY: *out = Aabb{H3(n.a0, n.b0, n.k - 0.0001f), H3(n.a1, n.b1, n.k + 0.0001f)}; return true;   // rect.h:18-21
    case NodeKind::RectXZ: *out = Aabb{H3(n.a0, n.k - 0.0001f, n.b0), H3(n.a1, n.k + 0.0001f, n.b1)}; return true;   // rect.h:38-41
    case NodeKind::RectYZ: *out = Aabb{H3(n.k - 0.0001f, n.a0, n.b0), H3(n.k + 0.0001f, n.a1, n.b1)}; return true;   // rect.h:61-64
    case NodeKind::Box: *out = n.box; return true;
    case NodeKind::List: if (!n.has_box) return false; *out = n.box; return true;
    case NodeKind::Bvh: *out = n.box; return true;
    case NodeKind::Translate: {   // scene_object.cpp:20-27
        if (!bounding_box(n.child, t0, t1, out)) return false;
        *out = Aabb{out->min + n.offset, out->max + n.offset};
        return true;
    }
    case NodeKind::RotateY: *out = n.box; return n.has_box;
    case NodeKind::Volume: return bounding_box(n.child, t0, t1, out);
    case NodeKind::PodBvh: *out = meshes[n.mesh].nodes[0].box; return true;
    case NodeKind::Triangle: {   // triangle.cpp:37-45: vmin / vmax of the three corners
        const H3 a = n.tri.m, b = n.tri.m + n.tri.u, c = n.tri.m + n.tri.v;
        *out = Aabb{hmin(hmin(a, b), c), hmax(hmax(a, b), c)};
        return true;
    }
    }
    return false;
}

int SceneGraph::list(const std::vector<int> &items, float time0, float time1) {   // scene_object.h:105-131
    Node n; n.kind = NodeKind::List; n.children = items;
    H3 minbb(FLT_MAX, FLT_MAX, FLT_MAX), maxbb(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    n.has_box = true;
    for (int c : items) {
        Aabb cur;
        if (bounding_box(c, time0, time1, &cur)) {
            minbb = hmin(minbb, cur.min);
            maxbb = hmax(maxbb, cur.max);
        } else {
            n.has_box = false;
            break;
        }
    }
    if (n.has_box) n.box = Aabb{minbb, maxbb};
    nodes.push_back(n); return (int) nodes.size() - 1;
}

static size_t max_dim(H3 a) {   // vec3.h:316-324
    bool v01 = (a.x > a.y), v02 = (a.x > a.z), v12 = (a.y > a.z);
    return v01 ? (v02 ? 0 : 2) : (v12 ? 1 : 2);
}

// node_order LUT, scene_object.h:154-205 / triangle.h:293-330
static uint8_t node_order(const Aabb &lbox, const Aabb &rbox) {
    H3 C0 = (lbox.max + lbox.min) * 0.5f;
    H3 C1 = (rbox.max + rbox.min) * 0.5f;
    H3 d = C0 - C1;
    uint8_t code = 0;
    int bit = 7;
    for (int sx = 0; sx < 2; sx++)
        for (int sy = 0; sy < 2; sy++)
            for (int sz = 0; sz < 2; sz++) {   // PPP, PPN, PNP, PNN, NPP, NPN, NNP, NNN
                H3 dir = hnormalize(H3(sx ? -1.0f : 1.0f, sy ? -1.0f : 1.0f, sz ? -1.0f : 1.0f));
                bool b = hdot(d, dir) < 0.0f;
                if (!b) code |= (uint8_t) (1u << bit);
                bit--;
            }
    return code;
}

int SceneGraph::bvh(std::vector<int> &items, size_t begin, size_t n, float time0, float time1) {   // scene_object.h:282-319
    Node node; node.kind = NodeKind::Bvh;
    {   // temporary object_list for the box of all objects
        H3 minbb(FLT_MAX, FLT_MAX, FLT_MAX), maxbb(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        for (size_t i = 0; i < n; i++) {
            Aabb cur;
            if (!bounding_box(items[begin + i], time0, time1, &cur)) { error = "no bounding box in bvh_node constructor"; return -1; }
            minbb = hmin(minbb, cur.min);
            maxbb = hmax(maxbb, cur.max);
        }
        node.box = Aabb{minbb, maxbb};
    }
    H3 dim = node.box.max - node.box.min;
    size_t axis = max_dim(dim);
    // qsort (glibc: stable merge sort) on bounding_box(0,0).min[axis]  (scene_object.h:246-267,295)
    std::vector<std::pair<float, int>> keyed(n);
    for (size_t i = 0; i < n; i++) {
        Aabb b;
        bounding_box(items[begin + i], 0, 0, &b);
        keyed[i] = {b.min[axis], items[begin + i]};
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<float, int> &a, const std::pair<float, int> &b) {
        return (a.first - b.first) < 0.0f;
    });
    for (size_t i = 0; i < n; i++) items[begin + i] = keyed[i].second;

    if (n == 1) {
        node.left = node.right = items[begin];
    } else if (n == 2) {
        node.left = items[begin];
        node.right = items[begin + 1];
    } else if (n < 11) {
        std::vector<int> l(items.begin() + begin, items.begin() + begin + n / 2);
        std::vector<int> r(items.begin() + begin + n / 2, items.begin() + begin + n);
        node.left = list(l, time0, time1);
        node.right = list(r, time0, time1);
    } else {
        node.left = bvh(items, begin, n / 2, time0, time1);
        node.right = bvh(items, begin + n / 2, n - n / 2, time0, time1);
    }
    Aabb lb, rb;
    bounding_box(node.left, 0, 1, &lb);
    bounding_box(node.right, 0, 1, &rb);
    node.order = node_order(lb, rb);
    nodes.push_back(node); return (int) nodes.size() - 1;
}

int SceneGraph::translate(int obj, H3 offset) {
    Node n; n.kind = NodeKind::Translate; n.child = obj; n.offset = offset;
    nodes.push_back(n); return (int) nodes.size() - 1;
}

int SceneGraph::rotate_y(int obj, float angle) {   // scene_object.cpp:33-68
    Node n; n.kind = NodeKind::RotateY; n.child = obj;
    float radians = H_RAD(angle);
    n.sin_theta = cr_sinf(radians);   // canonical libm, mrt_libm.h
    n.cos_theta = cr_cosf(radians);
    Aabb bbox;
    n.has_box = bounding_box(obj, 0, 1, &bbox);
    if (!n.has_box) {
        n.box = Aabb{H3(1, 1, 1), H3(-1, -1, -1)};
    } else {
        H3 minbb(FLT_MAX, FLT_MAX, FLT_MAX), maxbb(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++)
                for (int k = 0; k < 2; k++) {
                    float x = i * bbox.max.x + (1 - i) * bbox.min.x;
                    float y = j * bbox.max.y + (1 - j) * bbox.min.y;
                    float z = k * bbox.max.z + (1 - k) * bbox.min.z;
                    float newx = n.cos_theta * x + n.sin_theta * z;
                    float newz = n.cos_theta * z - n.sin_theta * x;
                    H3 testvec(newx, y, newz);
                    minbb = hmin(minbb, testvec);
                    maxbb = hmax(maxbb, testvec);
                }
        n.box = Aabb{minbb, maxbb};
    }
    nodes.push_back(n); return (int) nodes.size() - 1;
}

int SceneGraph::volume(int boundary, float density, int albedo_tex) {   // volumes.h:14-16
    Node n; n.kind = NodeKind::Volume; n.child = boundary; n.density = density; n.mat = isotropic(albedo_tex);
    nodes.push_back(n); return (int) nodes.size() - 1;
}

// ------------------------------------------------------------------ pod_bvh
namespace {
struct PodBuilder {   // triangle.h:77-168
    std::vector<Triangle> &prims;
    std::vector<H3> centroids;
    std::vector<PodNode> nodes;
    uint32_t node_count = 0;

    explicit PodBuilder(std::vector<Triangle> &p) : prims(p) {}

    void update_node_box(uint32_t ni) {
        PodNode &node = nodes[ni];
        const float maxf = std::numeric_limits<float>::max();
        const float minf = std::numeric_limits<float>::min();   // smallest positive (triangle.h:160)
        node.box = Aabb{H3(maxf, maxf, maxf), H3(minf, minf, minf)};
        for (size_t i = 0; i < node.prim_count; i++) {
            const Triangle &t = prims[node.prim_offset + i];   // triangle.h:33-41
            node.box.min = hmin(node.box.min, t.m);
            node.box.min = hmin(node.box.min, t.m + t.u);
            node.box.min = hmin(node.box.min, t.m + t.v);
            node.box.max = hmax(node.box.max, t.m);
            node.box.max = hmax(node.box.max, t.m + t.u);
            node.box.max = hmax(node.box.max, t.m + t.v);
        }
    }
    void subdivide(uint32_t ni) {
        if (nodes[ni].prim_count <= 2) return;
        H3 extent2 = nodes[ni].box.max - nodes[ni].box.min;
        int axis = 0;
        if (extent2.y > extent2.x) axis = 1;
        if (extent2.z > extent2[axis]) axis = 2;
        float splitPos = nodes[ni].box.min[axis] + extent2[axis] * 0.5f;
        int i = (int) nodes[ni].prim_offset;
        int j = i + (int) nodes[ni].prim_count - 1;
        while (i <= j) {
            if (centroids[i][axis] < splitPos) i++;
            else {
                std::swap(prims[i], prims[j]);
                std::swap(centroids[i], centroids[j]);
                j--;
            }
        }
        int left_count = i - (int) nodes[ni].prim_offset;
        if (left_count == 0 || left_count == (int) nodes[ni].prim_count) return;
        uint32_t l = node_count++, r = node_count++;
        nodes[l].prim_offset = nodes[ni].prim_offset;
        nodes[l].prim_count = (uint32_t) left_count;
        nodes[r].prim_offset = (uint32_t) i;
        nodes[r].prim_count = nodes[ni].prim_count - (uint32_t) left_count;
        update_node_box(l);
        update_node_box(r);
        nodes[ni].left = l;
        nodes[ni].prim_count = 0;
        nodes[ni].order = node_order(nodes[l].box, nodes[r].box);
        subdivide(l);
        subdivide(r);
    }
};
}  // namespace

int SceneGraph::pod_bvh(std::vector<Triangle> &&tris, int mat) {
    Mesh mesh;
    mesh.tris = std::move(tris);
    mesh.mat = mat;
    size_t n = mesh.tris.size();
    PodBuilder b(mesh.tris);
    b.nodes.resize(n * 2 - 1);
    b.centroids.resize(n);
    for (size_t i = 0; i < n; i++) {   // triangle::get_centroid, triangle.h:30-32
        const Triangle &t = mesh.tris[i];
        b.centroids[i] = (t.m + (t.m + t.u) + (t.m + t.v)) * (1.0f / 3.0f);
    }
    b.node_count = 1;
    b.nodes[0].left = 0;
    b.nodes[0].prim_offset = 0;
    b.nodes[0].prim_count = (uint32_t) n;
    b.update_node_box(0);
    b.subdivide(0);
    b.nodes.resize(b.node_count);
    mesh.nodes = std::move(b.nodes);
    meshes.push_back(std::move(mesh));
    Node node; node.kind = NodeKind::PodBvh; node.mesh = (int) meshes.size() - 1; node.mat = mat;
    nodes.push_back(node); return (int) nodes.size() - 1;
}

// --------------------------------------------------------------------- Mat4
M4 M4::identity() { M4 m{}; for (int i = 0; i < 4; i++) m.c[i][i] = 1; return m; }
M4 M4::scale(float s) { M4 m{}; m.c[0][0] = s; m.c[1][1] = s; m.c[2][2] = s; m.c[3][3] = 1; return m; }
M4 M4::rotate_y(float radians) {
    float s = cr_sinf(radians), c = cr_cosf(radians);
    M4 m{};
    // Mat4(c,0,s,0, 0,1,0,0, -s,0,c,0, 0,0,0,1) given row-major, stored column-major
    m.c[0][0] = c;  m.c[1][0] = 0; m.c[2][0] = s; m.c[3][0] = 0;
    m.c[0][1] = 0;  m.c[1][1] = 1; m.c[2][1] = 0; m.c[3][1] = 0;
    m.c[0][2] = -s; m.c[1][2] = 0; m.c[2][2] = c; m.c[3][2] = 0;
    m.c[0][3] = 0;  m.c[1][3] = 0; m.c[2][3] = 0; m.c[3][3] = 1;
    return m;
}
H3 M4::mul_col(H3 v) const {   // (vx*c0 + vy*c1) + vz*c2
    return H3((v.x * c[0][0] + v.y * c[1][0]) + v.z * c[2][0],
              (v.x * c[0][1] + v.y * c[1][1]) + v.z * c[2][1],
              (v.x * c[0][2] + v.y * c[1][2]) + v.z * c[2][2]);
}
H3 M4::mul_row(H3 v) const {   // res[i] = (v.x*ci.x + v.z*ci.z) + (v.y*ci.y + v.w*ci.w), v.w = 0
    float r[3];
    for (int i = 0; i < 3; i++) r[i] = (v.x * c[i][0] + v.z * c[i][2]) + (v.y * c[i][1] + 0.0f * c[i][3]);
    return H3(r[0], r[1], r[2]);
}
M4 M4::invert(const M4 &m) {
    // Lane-by-lane restatement of the SSE cofactor inverse (mat4.cpp:63-124).
    const float(*c)[4] = m.c;
    auto F = [&](int A, int B, float out[4]) {
        float v = c[2][A] * c[3][B] - c[3][A] * c[2][B];
        out[0] = v; out[1] = v;
        out[2] = c[1][A] * c[3][B] - c[3][A] * c[1][B];
        out[3] = c[1][A] * c[2][B] - c[2][A] * c[1][B];
    };
    float f1[4], f2[4], f3[4], f4[4], f5[4], f6[4];
    F(2, 3, f1); F(1, 3, f2); F(1, 2, f3); F(0, 3, f4); F(0, 2, f5); F(0, 1, f6);
    float v[4][4];   // v[k] = [c1[k], c0[k], c0[k], c0[k]]
    for (int k = 0; k < 4; k++) { v[k][0] = c[1][k]; v[k][1] = c[0][k]; v[k][2] = c[0][k]; v[k][3] = c[0][k]; }
    float i1[4], i2[4], i3[4], i4[4];
    for (int l = 0; l < 4; l++) {
        i1[l] = (v[1][l] * f1[l] - v[2][l] * f2[l]) + v[3][l] * f3[l];
        i2[l] = (v[0][l] * f1[l] - v[2][l] * f4[l]) + v[3][l] * f5[l];
        i3[l] = (v[0][l] * f2[l] - v[1][l] * f4[l]) + v[3][l] * f6[l];
        i4[l] = (v[0][l] * f3[l] - v[1][l] * f5[l]) + v[3][l] * f6[l];
    }
    // sign masks: s1 flips lanes 1,3; s2 flips lanes 0,2
    i1[1] = -i1[1]; i1[3] = -i1[3]; i3[1] = -i3[1]; i3[3] = -i3[3];
    i2[0] = -i2[0]; i2[2] = -i2[2]; i4[0] = -i4[0]; i4[2] = -i4[2];
    float d0 = c[0][0] * i1[0], d1 = c[0][1] * i2[0], d2 = c[0][2] * i3[0], d3 = c[0][3] * i4[0];
    float e[4] = {d0 + d2, d1 + d3, d2 + d0, d3 + d1};
    float det[4] = {e[0] + e[1], e[1] + e[0], e[2] + e[1], e[3] + e[0]};
    M4 r;
    for (int l = 0; l < 4; l++) {
        float inv = 1.0f / det[l];
        r.c[0][l] = i1[l] * inv;
        r.c[1][l] = i2[l] * inv;
        r.c[2][l] = i3[l] * inv;
        r.c[3][l] = i4[l] * inv;
    }
    return r;
}

}  // namespace mrt
