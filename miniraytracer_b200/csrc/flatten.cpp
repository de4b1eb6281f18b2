// Flattener: host scene graph -> the typed structure-of-arrays tables of
// mrt_types.h.  Topology, child order and the node_order bytes are preserved
// exactly (the reference's traversal result depends on them).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "scene_graph.h"

namespace mrt {

namespace {
inline float ubits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t ubits_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline MrtF4 f4(float x, float y, float z, float w) { MrtF4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

struct Flattener {
    const SceneGraph &g;
    FlatScene &o;
    std::map<int, uint32_t> memo;        // graph node -> typed ref
    std::map<int, uint32_t> mat_memo;    // graph material -> flat material index
    std::map<int, uint32_t> tex_memo;
    std::vector<uint64_t> image_offset;
    FlattenOptions opt;
    bool ok = true;

    Flattener(const SceneGraph &g_, FlatScene &o_, const FlattenOptions &opt_) : g(g_), o(o_), opt(opt_) {}

    void fail(const std::string &msg) { if (ok) o.error = msg; ok = false; }

    bool tex_needs_uv(int t) const {
        const Texture &x = g.texs[t];
        if (x.kind == TexKind::Image) return true;
        if (x.kind == TexKind::Checker) return tex_needs_uv(x.even) || tex_needs_uv(x.odd);
        return false;
    }
    uint32_t tex(int t) {
        auto it = tex_memo.find(t);
        if (it != tex_memo.end()) return it->second;
        const Texture &x = g.texs[t];
        uint32_t even = 0, odd = 0;
        if (x.kind == TexKind::Checker) { even = tex(x.even); odd = tex(x.odd); }
        uint32_t id = (uint32_t) o.tex.size();
        switch (x.kind) {
        case TexKind::Color: o.tex.push_back(f4(ubits(MRT_X_COLOR), x.color.x, x.color.y, x.color.z)); break;
        case TexKind::Checker: o.tex.push_back(f4(ubits(MRT_X_CHECKER), ubits(even), ubits(odd), x.scale)); break;
        case TexKind::Perlin: o.tex.push_back(f4(ubits(MRT_X_PERLIN), x.scale, 0, 0)); break;
        case TexKind::Image: {
            const Image &im = g.images[x.image];
            if (image_offset[x.image] > 0xFFFFFFFFull) fail("image table larger than 4 GiB");
            o.tex.push_back(f4(ubits(MRT_X_IMAGE), ubits((uint32_t) im.width), ubits((uint32_t) im.height), ubits((uint32_t) image_offset[x.image])));
            break;
        }
        }
        tex_memo[t] = id;
        return id;
    }
    uint32_t mat(int m) {
        auto it = mat_memo.find(m);
        if (it != mat_memo.end()) return it->second;
        const Material &x = g.mats[m];
        uint32_t kind = 0, t = 0;
        bool uv = false;
        switch (x.kind) {
        case MatKind::Lambertian: kind = MRT_M_LAMBERTIAN; break;
        case MatKind::Isotropic: kind = MRT_M_ISOTROPIC; break;
        case MatKind::Metal: kind = MRT_M_METAL; break;
        case MatKind::Dielectric: kind = MRT_M_DIELECTRIC; break;
        case MatKind::Light: kind = MRT_M_LIGHT; break;
        }
        if (x.tex >= 0) { t = tex(x.tex); uv = tex_needs_uv(x.tex); }
        uint32_t id = (uint32_t) o.mat.size();
        o.mat.push_back(f4(ubits(kind | (uv ? MRT_MAT_NEEDS_UV : 0u)), ubits(t), x.param, 0));
        mat_memo[m] = id;
        return id;
    }

    bool contains_volume(int id) const {
        const Node &n = g.nodes[id];
        switch (n.kind) {
        case NodeKind::Volume: return true;
        case NodeKind::List: for (int c : n.children) if (contains_volume(c)) return true; return false;
        case NodeKind::Bvh: return contains_volume(n.left) || contains_volume(n.right);
        case NodeKind::Box: case NodeKind::Translate: case NodeKind::RotateY: return contains_volume(n.child);
        default: return false;
        }
    }

    // The device traversal tests primitives only inside object_list loops (one copy of each hit test).
    // A primitive referenced directly by a bvh_node (n <= 2 leaves, scene_object.h:297-303), a transform or
    // a volume boundary is therefore wrapped into a box-less one-element list -- same result: the list
    // passes (tmin, tmax) through and reports its only child's hit (scene_object.h:79-103).
    std::map<uint32_t, uint32_t> wrap_memo;
    static bool is_prim(uint32_t ref) { return MRT_REF_TYPE(ref) <= MRT_T_RECT_YZ || MRT_REF_TYPE(ref) == MRT_T_TRI; }
    uint32_t wrap(uint32_t ref) {
        if (!is_prim(ref)) return ref;
        auto it = wrap_memo.find(ref);
        if (it != wrap_memo.end()) return it->second;
        uint32_t i = (uint32_t) o.list.size() / 2;
        uint32_t first = (uint32_t) o.child.size();
        o.child.push_back(ref);
        o.child.push_back(MRT_REF_END);
        o.list.push_back(f4(0, 0, 0, ubits(first)));
        o.list.push_back(f4(0, 0, 0, ubits(1u)));
        uint32_t w = MRT_REF(MRT_T_LIST, i);
        wrap_memo[ref] = w;
        return w;
    }
    bool node_is_prim(int id) const {
        NodeKind k = g.nodes[id].kind;
        return k == NodeKind::Sphere || k == NodeKind::RectXY || k == NodeKind::RectXZ || k == NodeKind::RectYZ || k == NodeKind::Triangle;
    }

    // ---- wide BVH nodes (children's boxes stored in the parent, see mrt_types.h)
    std::map<int, uint32_t> list_nobox_memo;
    uint32_t list_nobox(int id) {   // header copy with hasBox = 0 sharing the child table of the original list
        auto it = list_nobox_memo.find(id);
        if (it != list_nobox_memo.end()) return it->second;
        uint32_t orig = node(id);
        uint32_t oi = MRT_REF_INDEX(orig);
        MrtF4 h0 = o.list[2 * oi], h1 = o.list[2 * oi + 1];
        uint32_t w1; memcpy(&w1, &h1.w, 4);
        uint32_t i = (uint32_t) o.list.size() / 2;
        o.list.push_back(h0);
        o.list.push_back(f4(h1.x, h1.y, h1.z, ubits(w1 & 0x7FFFFFFFu)));
        uint32_t r = MRT_REF(MRT_T_LIST, i);
        list_nobox_memo[id] = r;
        return r;
    }
    // child of a bvh_node: the ref to visit once the child's own box (if it has one) has been passed
    uint32_t bvh_child(int id, Aabb *box, bool *has_box) {
        const Node &c = g.nodes[id];
        switch (c.kind) {
        case NodeKind::Bvh: *box = c.box; *has_box = true; return bvh_inner(id);
        case NodeKind::List: *box = c.box; *has_box = c.has_box; return c.has_box ? list_nobox(id) : node(id);
        case NodeKind::Box: return bvh_child(c.child, box, has_box);   // box::hit == its rect list's hit (box.h:23-25)
        default: *has_box = false; *box = Aabb(); return wrap(node(id));   // spheres, transforms, ...: no box test of their own here
        }
    }
    // nodes are laid out in depth-first PRE-order (a node right before its first subtree): a traversal that
    // descends keeps touching neighbouring 64-byte records (L1 hit rate of the triangle scene was 69 % with
    // post-order placement)
    uint32_t reserve_node2() {
        uint32_t i = (uint32_t) o.node2.size() / 4;
        for (int k = 0; k < 4; k++) o.node2.push_back(f4(0, 0, 0, 0));
        return i;
    }
    void fill_node2(uint32_t i, const Aabb &lb, bool lhas, uint32_t l, const Aabb &rb, bool rhas, uint32_t r, uint8_t order) {
        o.node2[4 * i + 0] = f4(lb.min.x, lb.min.y, lb.min.z, ubits(l | ((uint32_t) (order & 15u) << 28)));
        o.node2[4 * i + 1] = f4(lb.max.x, lb.max.y, lb.max.z, ubits(r | ((uint32_t) (order >> 4) << 28)));
        // flags: bit 0 / 1 = the left / right child has a box to test; bits 2-3 / 4-5 = kind of the left / right child for the
        // warp-cooperative traversal (coop_tree.cuh: 0 inner node, 1 object_list leaf, 2 triangle leaf)
        auto kind = [](uint32_t ref) { return MRT_REF_TYPE(ref) == MRT_T_NODE2 ? 0u : (MRT_REF_TYPE(ref) == MRT_T_TRILEAF ? 2u : 1u); };
        o.node2[4 * i + 2] = f4(rb.min.x, rb.min.y, rb.min.z, ubits((lhas ? 1u : 0u) | (rhas ? 2u : 0u) | (kind(l) << 2) | (kind(r) << 4)));
        o.node2[4 * i + 3] = f4(rb.max.x, rb.max.y, rb.max.z, 0);
    }
    std::map<int, uint32_t> inner_memo;
    uint32_t bvh_inner(int id) {
        auto it = inner_memo.find(id);
        if (it != inner_memo.end()) return it->second;
        const Node &n = g.nodes[id];
        Aabb lb, rb;
        bool lhas = false, rhas = false;
        uint32_t i = reserve_node2();
        uint32_t ref = MRT_REF(MRT_T_NODE2, i);
        inner_memo[id] = ref;
        uint32_t l = bvh_child(n.left, &lb, &lhas), r = bvh_child(n.right, &rb, &rhas);
        fill_node2(i, lb, lhas, l, rb, rhas, r, n.order);
        return ref;
    }
    uint32_t pod_child(const Mesh &mesh, uint32_t ni, uint32_t tri_base) {
        const PodNode &pn = mesh.nodes[ni];
        if (pn.prim_count) {
            uint32_t i = (uint32_t) o.trileaf.size() / 2;
            o.trileaf.push_back(tri_base + pn.prim_offset);
            o.trileaf.push_back(pn.prim_count);
            return MRT_REF(MRT_T_TRILEAF, i);
        }
        uint32_t i = reserve_node2();
        uint32_t l = pod_child(mesh, pn.left, tri_base), r = pod_child(mesh, pn.left + 1, tri_base);
        fill_node2(i, mesh.nodes[pn.left].box, true, l, mesh.nodes[pn.left + 1].box, true, r, pn.order);
        return MRT_REF(MRT_T_NODE2, i);
    }

    uint32_t node(int id) {
        auto it = memo.find(id);
        if (it != memo.end()) return it->second;
        const Node &n = g.nodes[id];
        uint32_t ref = MRT_REF_NONE;
        switch (n.kind) {
        case NodeKind::Sphere: {
            uint32_t i = (uint32_t) o.sphere.size() / 3;
            uint32_t m = mat(n.mat);
            o.sphere.push_back(f4(n.c0.x, n.c0.y, n.c0.z, n.radius));
            o.sphere.push_back(f4(n.c1.x, n.c1.y, n.c1.z, ubits(m | (n.moving ? 0x80000000u : 0u))));
            o.sphere.push_back(f4(n.t0, n.t1, 0, 0));
            ref = MRT_REF(MRT_T_SPHERE, i);
            break;
        }
        case NodeKind::RectXY:
        case NodeKind::RectXZ:
        case NodeKind::RectYZ: {
            uint32_t i = (uint32_t) o.rect.size() / 2;
            uint32_t m = mat(n.mat);
            o.rect.push_back(f4(n.a0, n.a1, n.b0, n.b1));
            o.rect.push_back(f4(n.k, n.sign, ubits(m), 0));
            uint32_t type = n.kind == NodeKind::RectXY ? MRT_T_RECT_XY : (n.kind == NodeKind::RectXZ ? MRT_T_RECT_XZ : MRT_T_RECT_YZ);
            ref = MRT_REF(type, i);
            break;
        }
        case NodeKind::Triangle: {   // triangle_scene_object: one record in the triangle tables, referenced as a primitive
            uint32_t i = (uint32_t) o.tri.size() / 3;
            uint32_t m = mat(n.mat);
            const Triangle &t = n.tri;
            o.tri.push_back(f4(t.m.x, t.m.y, t.m.z, ubits(m)));
            o.tri.push_back(f4(t.u.x, t.u.y, t.u.z, 0));
            o.tri.push_back(f4(t.v.x, t.v.y, t.v.z, 0));
            o.trin.push_back(f4(t.mn.x, t.mn.y, t.mn.z, 0));
            o.trin.push_back(f4(t.un.x, t.un.y, t.un.z, 0));
            o.trin.push_back(f4(t.vn.x, t.vn.y, t.vn.z, 0));
            ref = MRT_REF(MRT_T_TRI, i);
            break;
        }
        case NodeKind::Box:   // box::hit forwards to its rect list (box.h:23-25)
            ref = node(n.child);
            break;
        case NodeKind::List: {
            std::vector<uint32_t> refs;
            for (int c : n.children) refs.push_back(node(c));
            uint32_t i = (uint32_t) o.list.size() / 2;
            uint32_t first = (uint32_t) o.child.size();
            for (uint32_t r : refs) o.child.push_back(r);
            o.child.push_back(MRT_REF_END);
            if (n.children.size() > 0x7FFFFFFFu) fail("list too long");
            H3 mn = n.has_box ? n.box.min : H3(0, 0, 0), mx = n.has_box ? n.box.max : H3(0, 0, 0);
            o.list.push_back(f4(mn.x, mn.y, mn.z, ubits(first)));
            o.list.push_back(f4(mx.x, mx.y, mx.z, ubits((uint32_t) n.children.size() | (n.has_box ? 0x80000000u : 0u))));
            ref = MRT_REF(MRT_T_LIST, i);
            break;
        }
        case NodeKind::Bvh: {   // root header: the tree's own box test, then the wide nodes
            uint32_t root = bvh_inner(id);
            uint32_t i = (uint32_t) o.bvh.size() / 2;
            o.bvh.push_back(f4(n.box.min.x, n.box.min.y, n.box.min.z, ubits(root)));
            o.bvh.push_back(f4(n.box.max.x, n.box.max.y, n.box.max.z, 0));
            ref = MRT_REF(MRT_T_BVH, i);
            break;
        }
        case NodeKind::Translate: {
            uint32_t c = wrap(node(n.child));
            uint32_t i = (uint32_t) o.xlate.size();   // 3 records per translate
            o.xlate.push_back(f4(n.offset.x, n.offset.y, n.offset.z, ubits(c)));
            // cull box: the object's bounds in the parent frame (scene_object.cpp:20-27), inflated after the
            // whole scene is known (flatten_scene); w of the first record = 1 if there is one
            Aabb bb;
            const bool has = opt.cull_boxes && g.bounding_box(id, g.camera.time0, g.camera.time1, &bb);
            o.xlate.push_back(has ? f4(bb.min.x, bb.min.y, bb.min.z, ubits(1u)) : f4(0, 0, 0, ubits(0u)));
            o.xlate.push_back(has ? f4(bb.max.x, bb.max.y, bb.max.z, 0) : f4(0, 0, 0, 0));
            ref = MRT_REF(MRT_T_TRANSLATE, i / 3);
            break;
        }
        case NodeKind::RotateY: {
            uint32_t c = wrap(node(n.child));
            uint32_t i = (uint32_t) o.rot.size() / 3;
            o.rot.push_back(f4(n.box.min.x, n.box.min.y, n.box.min.z, ubits(c)));
            o.rot.push_back(f4(n.box.max.x, n.box.max.y, n.box.max.z, ubits(n.has_box ? 1u : 0u)));
            o.rot.push_back(f4(n.sin_theta, n.cos_theta, 0, 0));
            ref = MRT_REF(MRT_T_ROTATE_Y, i);
            break;
        }
        case NodeKind::Volume: {
            if (contains_volume(n.child)) fail("a constant_volume boundary must not contain another constant_volume");
            uint32_t b = wrap(node(n.child));
            uint32_t m = mat(n.mat);
            uint32_t i = (uint32_t) o.vol.size();
            o.vol.push_back(f4(ubits(b), n.density, ubits(m), 0));
            ref = MRT_REF(MRT_T_VOLUME, i);
            break;
        }
        case NodeKind::PodBvh: {
            const Mesh &mesh = g.meshes[n.mesh];
            uint32_t m = mat(mesh.mat);
            uint32_t tri_base = (uint32_t) o.tri.size() / 3;
            for (const Triangle &t : mesh.tris) {
                o.tri.push_back(f4(t.m.x, t.m.y, t.m.z, ubits(m)));
                o.tri.push_back(f4(t.u.x, t.u.y, t.u.z, 0));
                o.tri.push_back(f4(t.v.x, t.v.y, t.v.z, 0));
                o.trin.push_back(f4(t.mn.x, t.mn.y, t.mn.z, 0));
                o.trin.push_back(f4(t.un.x, t.un.y, t.un.z, 0));
                o.trin.push_back(f4(t.vn.x, t.vn.y, t.vn.z, 0));
            }
            uint32_t root = pod_child(mesh, 0, tri_base);
            uint32_t i = (uint32_t) o.bvh.size() / 2;
            const Aabb &rb = mesh.nodes[0].box;
            o.bvh.push_back(f4(rb.min.x, rb.min.y, rb.min.z, ubits(root)));
            o.bvh.push_back(f4(rb.max.x, rb.max.y, rb.max.z, 0));
            ref = MRT_REF(MRT_T_BVH, i);
            break;
        }
        }
        memo[id] = ref;
        return ref;
    }

    // worst-case number of 32-bit stack words `intersect` needs below this node (trace_core.h)
    uint32_t pod_depth(const Mesh &m, uint32_t ni) const {
        const PodNode &pn = m.nodes[ni];
        if (pn.prim_count) return 0;
        return 1 + std::max(pod_depth(m, pn.left), pod_depth(m, pn.left + 1));
    }
    // coop = true: BVH trees are traversed warp-cooperatively (coop_tree.cuh) and cost no per-lane stack frames
    uint32_t depth_w(int id, bool coop) const { return node_is_prim(id) ? 1u : depth(id, coop); }   // wrapped primitives cost one LIST frame
    uint32_t depth(int id, bool coop) const {
        const Node &n = g.nodes[id];
        switch (n.kind) {
        case NodeKind::List: {
            uint32_t d = 0;
            for (int c : n.children) d = std::max(d, depth(c, coop));
            return 1 + d;
        }
        case NodeKind::Box: return depth(n.child, coop);
        case NodeKind::Bvh: return coop ? 0u : 1 + std::max(depth_w(n.left, coop), depth_w(n.right, coop));
        case NodeKind::Translate:
        case NodeKind::RotateY: return 11 + depth_w(n.child, coop);
        case NodeKind::Volume: return 1 + depth_w(n.child, coop);
        case NodeKind::PodBvh: return coop ? 0u : pod_depth(g.meshes[n.mesh], 0);
        default: return 0;
        }
    }
};
}  // namespace

// Do the scene's BVH trees qualify for the warp-cooperative traversal (coop_tree.cuh)?  Every tree at most 31 levels deep
// (the rank has one bit per level), every leaf a triangle leaf or an object_list of spheres / rects / lists of those
// (what coop_leaf_hit evaluates), and no tree inside a constant_volume boundary (probe mode stays per-lane).
// Works on the flattened tables, so it also covers scenes loaded from a file.
bool coop_trees_supported(const MrtSceneDesc &d) {
    auto u = [](float f) { uint32_t v; memcpy(&v, &f, 4); return v; };
    auto list_ok = [&](uint32_t li, bool nested, auto &&self) -> bool {
        if (li >= d.n_list) return false;
        for (uint32_t ci = u(d.list[2 * li].w); ; ci++) {
            if (ci >= d.n_child) return false;
            const uint32_t c = d.child[ci], t = MRT_REF_TYPE(c);
            if (t == MRT_T_END) return true;
            if (t <= MRT_T_RECT_YZ) continue;
            if (t == MRT_T_LIST && !nested && self(MRT_REF_INDEX(c), true, self)) continue;
            return false;
        }
    };
    struct Item { uint32_t ref, depth; };
    std::vector<Item> todo;
    for (uint32_t b = 0; b < d.n_bvh; b++) todo.push_back(Item{u(d.bvh[2 * b].w), 1});
    size_t visited = 0;
    while (!todo.empty()) {
        Item it = todo.back(); todo.pop_back();
        if (++visited > (size_t) d.n_node2 * 2 + d.n_bvh + 16) return false;   // not a tree
        const uint32_t t = MRT_REF_TYPE(it.ref), i = MRT_REF_INDEX(it.ref);
        if (t == MRT_T_NODE2) {
            if (i >= d.n_node2 || it.depth > 31) return false;
            todo.push_back(Item{u(d.node2[4 * i].w) & 0x0FFFFFFFu, it.depth + 1});
            todo.push_back(Item{u(d.node2[4 * i + 1].w) & 0x0FFFFFFFu, it.depth + 1});
        } else if (t == MRT_T_TRILEAF) {
            if (i >= d.n_trileaf) return false;
        } else if (t == MRT_T_LIST) {
            if (!list_ok(i, false, list_ok)) return false;
        } else return false;
    }
    // volumes: the boundary subtree must not contain a tree
    std::vector<uint32_t> st;
    for (uint32_t v = 0; v < d.n_vol; v++) st.push_back(u(d.vol[v].x));
    visited = 0;
    while (!st.empty()) {
        const uint32_t r = st.back(); st.pop_back();
        if (++visited > (size_t) d.n_child + d.n_list + d.n_xlate + d.n_rot + d.n_vol + 16) return false;
        const uint32_t t = MRT_REF_TYPE(r), i = MRT_REF_INDEX(r);
        if (t == MRT_T_BVH || t == MRT_T_NODE2 || t == MRT_T_TRILEAF) return false;
        if (t == MRT_T_LIST && i < d.n_list) { for (uint32_t ci = u(d.list[2 * i].w); ci < d.n_child && MRT_REF_TYPE(d.child[ci]) != MRT_T_END; ci++) st.push_back(d.child[ci]); }
        else if (t == MRT_T_TRANSLATE && i < d.n_xlate) st.push_back(u(d.xlate[3 * i].w));
        else if (t == MRT_T_ROTATE_Y && i < d.n_rot) st.push_back(u(d.rot[3 * i].w));
        else if (t == MRT_T_VOLUME && i < d.n_vol) st.push_back(u(d.vol[i].x));
    }
    return d.n_bvh > 0;
}

// Structural check of a flattened scene that did not come straight out of flatten_scene (a scene file, a description built by
// a caller of the C ABI): every typed reference and index inside its table, every child run terminated, Perlin tables of the
// sizes the kernels read, image extents inside the image table, no reference cycles, and a traversal stack at least as deep
// as the scene needs.  On success *stack_words_needed (may be null) is the depth the flattener would have computed.
bool validate_scene_desc(const MrtSceneDesc &d, std::string *err, uint32_t *stack_words_needed) {
    auto u = [](float f) { uint32_t v; memcpy(&v, &f, 4); return v; };
    auto fail = [&](const char *what) { if (err) *err = what; return false; };
    if ((d.n_sphere && !d.sphere) || (d.n_rect && !d.rect) || (d.n_list && !d.list) || (d.n_child && !d.child) || (d.n_bvh && !d.bvh) ||
        (d.n_node2 && !d.node2) || (d.n_trileaf && !d.trileaf) || (d.n_tri && (!d.tri || !d.trin)) || (d.n_xlate && !d.xlate) ||
        (d.n_rot && !d.rot) || (d.n_vol && !d.vol) || (d.n_mat && !d.mat) || (d.n_tex && !d.tex) || (d.n_lights && !d.lights) ||
        (d.n_image_bytes && !d.image))
        return fail("a table pointer is null although its count is not");
    auto ref_ok = [&](uint32_t ref) {
        if (ref >> 28) return false;
        const uint32_t t = MRT_REF_TYPE(ref), i = MRT_REF_INDEX(ref);
        switch (t) {
        case MRT_T_SPHERE: return i < d.n_sphere;
        case MRT_T_RECT_XY: case MRT_T_RECT_XZ: case MRT_T_RECT_YZ: return i < d.n_rect;
        case MRT_T_LIST: return i < d.n_list;
        case MRT_T_BVH: return i < d.n_bvh;
        case MRT_T_NODE2: return i < d.n_node2;
        case MRT_T_TRANSLATE: return i < d.n_xlate;
        case MRT_T_ROTATE_Y: return i < d.n_rot;
        case MRT_T_VOLUME: return i < d.n_vol;
        case MRT_T_TRILEAF: return i < d.n_trileaf;
        case MRT_T_TRI: return i < d.n_tri;
        default: return false;
        }
    };
    // materials and textures
    for (uint32_t i = 0; i < d.n_tex; i++) {
        const MrtF4 &t = d.tex[i];
        const uint32_t kind = u(t.x);
        if (kind > MRT_X_IMAGE) return fail("texture kind out of range");
        if (kind == MRT_X_CHECKER && (u(t.y) >= d.n_tex || u(t.z) >= d.n_tex || u(t.y) == i || u(t.z) == i)) return fail("checker texture refers outside the texture table");
        if (kind == MRT_X_PERLIN && (!d.perlin_vec || !d.perlin_perm)) return fail("perlin texture without perlin tables");
        if (kind == MRT_X_IMAGE) {
            const uint64_t w = u(t.y), h = u(t.z), off = u(t.w);
            if (!w || !h || off + w * h * 3u > d.n_image_bytes) return fail("image texture outside the image table");
        }
    }
    for (uint32_t i = 0; i < d.n_tex; i++) {   // checker chains must end (no cycles): follow at most n_tex links
        uint32_t t = i, steps = 0;
        while (u(d.tex[t].x) == MRT_X_CHECKER) { t = u(d.tex[t].y); if (++steps > d.n_tex) return fail("checker textures form a cycle"); }
    }
    for (uint32_t i = 0; i < d.n_mat; i++) {
        if ((u(d.mat[i].x) & 0xFFu) > MRT_M_LIGHT) return fail("material kind out of range");
        if (u(d.mat[i].y) >= d.n_tex && (u(d.mat[i].x) & 0xFFu) != MRT_M_DIELECTRIC) return fail("material refers outside the texture table");
    }
    auto mat_ok = [&](uint32_t m) { return m < d.n_mat; };
    for (uint32_t i = 0; i < d.n_sphere; i++) if (!mat_ok(u(d.sphere[3 * i + 1].w) & 0x7FFFFFFFu)) return fail("sphere material out of range");
    for (uint32_t i = 0; i < d.n_rect; i++) if (!mat_ok(u(d.rect[2 * i + 1].z))) return fail("rect material out of range");
    for (uint32_t i = 0; i < d.n_tri; i++) if (!mat_ok(u(d.tri[3 * i].w))) return fail("triangle material out of range");
    for (uint32_t i = 0; i < d.n_vol; i++) if (!mat_ok(u(d.vol[i].z)) || !ref_ok(u(d.vol[i].x))) return fail("volume refers outside its tables");
    for (uint32_t i = 0; i < d.n_trileaf; i++) {
        const uint64_t first = d.trileaf[2 * i], cnt = d.trileaf[2 * i + 1];
        if (first + cnt > d.n_tri) return fail("triangle leaf outside the triangle table");
    }
    for (uint32_t i = 0; i < d.n_list; i++) {
        uint32_t ci = u(d.list[2 * i].w);
        for (;; ci++) {
            if (ci >= d.n_child) return fail("list children not terminated inside the child table");
            if (MRT_REF_TYPE(d.child[ci]) == MRT_T_END) break;
            if (!ref_ok(d.child[ci])) return fail("list child refers outside its table");
        }
    }
    for (uint32_t i = 0; i < d.n_bvh; i++) if (!ref_ok(u(d.bvh[2 * i].w))) return fail("tree root refers outside its table");
    for (uint32_t i = 0; i < d.n_node2; i++) {
        const uint32_t l = u(d.node2[4 * i].w) & 0x0FFFFFFFu, r = u(d.node2[4 * i + 1].w) & 0x0FFFFFFFu, fl = u(d.node2[4 * i + 2].w);
        if (!ref_ok(l) || !ref_ok(r)) return fail("tree node child refers outside its table");
        auto kind = [](uint32_t ref) { return MRT_REF_TYPE(ref) == MRT_T_NODE2 ? 0u : (MRT_REF_TYPE(ref) == MRT_T_TRILEAF ? 2u : 1u); };
        if (((fl >> 2) & 3u) != kind(l) || ((fl >> 4) & 3u) != kind(r)) return fail("tree node flags disagree with its children");
    }
    for (uint32_t i = 0; i < d.n_xlate; i++) if (!ref_ok(u(d.xlate[3 * i].w))) return fail("translate child refers outside its table");
    for (uint32_t i = 0; i < d.n_rot; i++) if (!ref_ok(u(d.rot[3 * i].w))) return fail("rotate_y child refers outside its table");
    for (uint32_t i = 0; i < d.n_lights; i++) {
        const uint32_t t = MRT_REF_TYPE(d.lights[i]);
        if (!ref_ok(d.lights[i]) || t > MRT_T_RECT_YZ) return fail("light list entry is not a primitive of the scene");
    }
    if (!ref_ok(d.root)) return fail("root refers outside its table");
    // depth of the traversal stack (mirrors Flattener::depth) with cycle / blow-up protection
    uint64_t visits = 0;
    const uint64_t visit_cap = 64ull * ((uint64_t) d.n_child + d.n_list + d.n_node2 + d.n_bvh + d.n_xlate + d.n_rot + d.n_vol + 16u);
    bool bad = false;
    bool coop = false;   // second pass: trees cost no per-lane stack (stack_words_coop)
    auto depth = [&](uint32_t ref, uint32_t level, auto &&self) -> uint32_t {
        if (bad) return 0;
        if (level > 512 || ++visits > visit_cap) { bad = true; return 0; }
        const uint32_t t = MRT_REF_TYPE(ref), i = MRT_REF_INDEX(ref);
        switch (t) {
        case MRT_T_LIST: {
            uint32_t m = 0;
            for (uint32_t ci = u(d.list[2 * i].w); MRT_REF_TYPE(d.child[ci]) != MRT_T_END; ci++)
                if (MRT_REF_TYPE(d.child[ci]) > MRT_T_RECT_YZ && MRT_REF_TYPE(d.child[ci]) != MRT_T_TRI) m = std::max(m, self(d.child[ci], level + 1, self));
            return 1 + m;
        }
        case MRT_T_BVH: return coop ? 0u : self(u(d.bvh[2 * i].w), level + 1, self);
        case MRT_T_NODE2: return 1 + std::max(self(u(d.node2[4 * i].w) & 0x0FFFFFFFu, level + 1, self), self(u(d.node2[4 * i + 1].w) & 0x0FFFFFFFu, level + 1, self));
        case MRT_T_TRANSLATE: return 11 + self(u(d.xlate[3 * i].w), level + 1, self);
        case MRT_T_ROTATE_Y: return 11 + self(u(d.rot[3 * i].w), level + 1, self);
        case MRT_T_VOLUME: return 1 + self(u(d.vol[i].x), level + 1, self);
        default: return 0;   // primitives (only legal as list children; elsewhere the traversal ignores them) and triangle leaves
        }
    };
    const uint32_t need = depth(d.root, 0, depth) + 2;
    if (bad) return fail("the object graph is cyclic or unreasonably deep");
    if (stack_words_needed) *stack_words_needed = need;
    if (d.stack_words && d.stack_words < need) return fail("stack_words is smaller than the scene's traversal depth");
    if (d.stack_words > 1024u || need > 1024u) return fail("traversal stack deeper than 1024 words");
    if (d.stack_words_coop) {
        if (!coop_trees_supported(d)) return fail("stack_words_coop set although the trees do not qualify");
        coop = true;
        visits = 0;
        if (d.stack_words_coop < depth(d.root, 0, depth) + 2 || d.stack_words_coop > 1024u) return fail("stack_words_coop is smaller than the scene needs");
    }
    return true;
}

bool flatten_scene(const SceneGraph &g, FlatScene *out, const FlattenOptions &opt) {
    FlatScene &o = *out;
    o = FlatScene();
    Flattener fl(g, o, opt);
    uint64_t off = 0;
    for (const Image &im : g.images) {
        fl.image_offset.push_back(off);
        o.image.insert(o.image.end(), im.rgb.begin(), im.rgb.end());
        off += im.rgb.size();
    }
    uint32_t root = fl.wrap(fl.node(g.objects));
    if (g.biased >= 0) {
        const Node &b = g.nodes[g.biased];
        if (b.kind != NodeKind::List) fl.fail("biased_objects must be an object_list");
        else for (int c : b.children) o.lights.push_back(fl.node(c));
    }
    if (g.uses_perlin) {
        const PerlinTables &pt = perlin_tables();
        for (int i = 0; i < 256; i++) o.perlin_vec.push_back(f4(pt.ranvec[i][0], pt.ranvec[i][1], pt.ranvec[i][2], 0));
        for (int a = 0; a < 3; a++) for (int i = 0; i < 256; i++) o.perlin_perm.push_back(pt.perm[a][i]);
    }
    const size_t lim = 0xFFFFFFu;
    if (o.sphere.size() / 3 > lim || o.rect.size() / 2 > lim || o.list.size() / 2 > lim || o.bvh.size() / 2 > lim ||
        o.node2.size() / 4 > lim || o.trileaf.size() / 2 > lim || o.tri.size() / 3 > lim || o.xlate.size() / 3 > lim || o.rot.size() / 3 > lim || o.vol.size() > lim)
        fl.fail("too many objects of one type for a 24-bit index");
    if (o.child.size() > 0x0FFFFFFFu) fl.fail("child table too large");
    if (!fl.ok) return false;

    {   // Inflate the translate cull boxes by 1e-3 of the scene scale: ~1000x the float32 rounding of anything the
        // exact path computes at these coordinates, so a ray that misses the inflated box cannot be reported as a hit
        // by the exact transformed-space test (trace_core.h: cull_miss).
        float scale = 1.0f;
        auto grow = [&](float v) { v = std::fabs(v); if (v < 1e30f && v > scale) scale = v; };
        for (size_t i = 0; i < o.sphere.size(); i += 3) { const MrtF4 &a = o.sphere[i], &b = o.sphere[i + 1]; grow(a.x); grow(a.y); grow(a.z); grow(b.x); grow(b.y); grow(b.z); grow(a.w); }
        for (size_t i = 0; i < o.rect.size(); i += 2) { const MrtF4 &a = o.rect[i]; grow(a.x); grow(a.y); grow(a.z); grow(a.w); grow(o.rect[i + 1].x); }
        for (size_t i = 0; i < o.bvh.size(); i++) { grow(o.bvh[i].x); grow(o.bvh[i].y); grow(o.bvh[i].z); }
        for (size_t i = 0; i < o.xlate.size(); i++) { grow(o.xlate[i].x); grow(o.xlate[i].y); grow(o.xlate[i].z); }
        grow(g.camera.origin.x); grow(g.camera.origin.y); grow(g.camera.origin.z);
        const float m = 1e-3f * scale;
        for (size_t i = 0; i < o.xlate.size(); i += 3) {
            if (!ubits_of(o.xlate[i + 1].w)) continue;
            o.xlate[i + 1].x -= m; o.xlate[i + 1].y -= m; o.xlate[i + 1].z -= m;
            o.xlate[i + 2].x += m; o.xlate[i + 2].y += m; o.xlate[i + 2].z += m;
            o.xlate[i + 2].w = m;
        }
    }

    MrtSceneDesc &d = o.desc;
    memset(&d, 0, sizeof(d));
    d.root = root;
    d.n_lights = (uint32_t) o.lights.size();
    d.lights = o.lights.data();
    d.sky = g.sky ? 1u : 0u;
    d.stack_words = fl.depth_w(g.objects, false) + 2;
    {   // feature mask: what a specialised kernel must be able to handle
        uint32_t f = 0;
        if (!o.node2.empty() || !o.bvh.empty() || !o.trileaf.empty()) f |= MRT_FEAT_TREES;
        if (!o.vol.empty()) f |= MRT_FEAT_VOLUMES;
        if (!o.xlate.empty() || !o.rot.empty()) f |= MRT_FEAT_XFORM;
        for (const Texture &t : g.texs) if (t.kind != TexKind::Color) f |= MRT_FEAT_TEX;
        for (const Material &m : g.mats) {
            if (m.kind == MatKind::Metal) f |= MRT_FEAT_METAL;
            if (m.kind == MatKind::Dielectric) f |= MRT_FEAT_DIELECTRIC;
            if (m.kind == MatKind::Isotropic) f |= MRT_FEAT_VOLUMES;
        }
        for (const Node &n : g.nodes) if (n.kind == NodeKind::Sphere && n.moving) f |= MRT_FEAT_MOVING;
        for (uint32_t l : o.lights) if (MRT_REF_TYPE(l) != MRT_T_RECT_XZ) f |= MRT_FEAT_LIGHT_SPHERE;
        if (!o.sphere.empty()) f |= MRT_FEAT_SPHERES;
        if (!o.trileaf.empty()) f |= MRT_FEAT_TRIS;
        for (const Node &n : g.nodes) if (n.kind == NodeKind::Triangle) f |= MRT_FEAT_TRI_OBJECT;
        for (size_t i = 0; i < o.node2.size(); i += 4) {   // kinds of the two children, see fill_node2
            const uint32_t fl = ubits_of(o.node2[i + 2].w);
            if (((fl >> 2) & 3u) == 1u || ((fl >> 4) & 3u) == 1u) f |= MRT_FEAT_LEAF_LISTS;
        }
        for (size_t i = 0; i < o.bvh.size(); i += 2) {
            const uint32_t t = MRT_REF_TYPE(ubits_of(o.bvh[i].w));
            if (t != MRT_T_NODE2 && t != MRT_T_TRILEAF) f |= MRT_FEAT_LEAF_LISTS;
        }
        d.features = f;
    }
    const Camera &c = g.camera;
    auto put = [](float *dst, H3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
    put(d.camera.origin, c.origin); put(d.camera.u, c.u); put(d.camera.v, c.v); put(d.camera.w, c.w);
    put(d.camera.llcorner, c.llcorner); put(d.camera.horz, c.horz); put(d.camera.vert, c.vert);
    d.camera.lens_radius = c.lens_radius; d.camera.time0 = c.time0; d.camera.time1 = c.time1;
    d.sphere = o.sphere.data(); d.n_sphere = (uint32_t) o.sphere.size() / 3;
    d.rect = o.rect.data();     d.n_rect = (uint32_t) o.rect.size() / 2;
    d.list = o.list.data();     d.n_list = (uint32_t) o.list.size() / 2;
    d.child = o.child.data();   d.n_child = (uint32_t) o.child.size();
    d.bvh = o.bvh.data();       d.n_bvh = (uint32_t) o.bvh.size() / 2;
    d.node2 = o.node2.data();   d.n_node2 = (uint32_t) o.node2.size() / 4;
    d.trileaf = o.trileaf.data(); d.n_trileaf = (uint32_t) o.trileaf.size() / 2;
    d.tri = o.tri.data();       d.n_tri = (uint32_t) o.tri.size() / 3;
    d.trin = o.trin.data();
    d.xlate = o.xlate.data();   d.n_xlate = (uint32_t) o.xlate.size() / 3;
    d.rot = o.rot.data();       d.n_rot = (uint32_t) o.rot.size() / 3;
    d.vol = o.vol.data();       d.n_vol = (uint32_t) o.vol.size();
    d.mat = o.mat.data();       d.n_mat = (uint32_t) o.mat.size();
    d.tex = o.tex.data();       d.n_tex = (uint32_t) o.tex.size();
    d.perlin_vec = o.perlin_vec.empty() ? nullptr : o.perlin_vec.data();
    d.perlin_perm = o.perlin_perm.empty() ? nullptr : o.perlin_perm.data();
    d.image = o.image.empty() ? nullptr : o.image.data();
    d.n_image_bytes = o.image.size();
    d.stack_words_coop = coop_trees_supported(d) ? fl.depth_w(g.objects, true) + 2 : 0u;
    return true;
}

}  // namespace mrt
