// Canonical text dump of a host scene graph.  The format is shared with
// oracle/ref_harness.cpp (dump-scene), which prints the reference's own scene
// graph; tests/test_scene_parity.py compares the two byte for byte.
#include <cstring>

#include "scene_graph.h"

namespace mrt {

static uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static void ind(FILE *f, int d) { for (int i = 0; i < d; i++) fputc(' ', f); }
static void pv(FILE *f, const char *name, H3 v) { fprintf(f, " %s=%08x,%08x,%08x", name, fbits(v.x), fbits(v.y), fbits(v.z)); }
static void pf(FILE *f, const char *name, float v) { fprintf(f, " %s=%08x", name, fbits(v)); }

static uint32_t fnv(const uint8_t *d, size_t n) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; i++) { h ^= d[i]; h *= 16777619u; }
    return h;
}

static void dump_tex(const SceneGraph &g, FILE *f, int t) {
    const Texture &x = g.texs[t];
    switch (x.kind) {
    case TexKind::Color: fprintf(f, "[color"); pv(f, "c", x.color); fprintf(f, "]"); break;
    case TexKind::Checker:
        fprintf(f, "[checker"); pf(f, "scale", x.scale);
        fprintf(f, " even="); dump_tex(g, f, x.even);
        fprintf(f, " odd="); dump_tex(g, f, x.odd); fprintf(f, "]");
        break;
    case TexKind::Perlin: fprintf(f, "[perlin"); pf(f, "scale", x.scale); fprintf(f, "]"); break;
    case TexKind::Image: {
        const Image &im = g.images[x.image];
        fprintf(f, "[image w=%d h=%d fnv=%08x]", im.width, im.height, fnv(im.rgb.data(), im.rgb.size()));
        break;
    }
    }
}

static void dump_mat(const SceneGraph &g, FILE *f, int m) {
    const Material &x = g.mats[m];
    switch (x.kind) {
    case MatKind::Lambertian: fprintf(f, "{lambertian tex="); dump_tex(g, f, x.tex); fprintf(f, "}"); break;
    case MatKind::Isotropic: fprintf(f, "{isotropic tex="); dump_tex(g, f, x.tex); fprintf(f, "}"); break;
    case MatKind::Metal: fprintf(f, "{metal"); pf(f, "gloss", x.param); fprintf(f, " tex="); dump_tex(g, f, x.tex); fprintf(f, "}"); break;
    case MatKind::Dielectric: fprintf(f, "{dielectric"); pf(f, "idx", x.param); fprintf(f, "}"); break;
    case MatKind::Light: fprintf(f, "{light"); pf(f, "scale", x.param); fprintf(f, " tex="); dump_tex(g, f, x.tex); fprintf(f, "}"); break;
    }
}

static void dump_obj(const SceneGraph &g, FILE *f, int id, int d) {
    const Node &n = g.nodes[id];
    switch (n.kind) {
    case NodeKind::Sphere:
        ind(f, d); fprintf(f, "sphere");
        pv(f, "c0", n.c0); pv(f, "c1", n.c1); pf(f, "t0", n.t0); pf(f, "t1", n.t1);
        fprintf(f, " moving=%d", (int) n.moving); pf(f, "r", n.radius);
        fprintf(f, " mat="); dump_mat(g, f, n.mat); fprintf(f, "\n");
        break;
    case NodeKind::RectXY:
    case NodeKind::RectXZ:
    case NodeKind::RectYZ:
        ind(f, d);
        fprintf(f, n.kind == NodeKind::RectXY ? "xy_rect" : (n.kind == NodeKind::RectXZ ? "xz_rect" : "yz_rect"));
        pf(f, "a0", n.a0); pf(f, "a1", n.a1); pf(f, "b0", n.b0); pf(f, "b1", n.b1); pf(f, "k", n.k); pf(f, "sign", n.sign);
        fprintf(f, " mat="); dump_mat(g, f, n.mat); fprintf(f, "\n");
        break;
    case NodeKind::Triangle:
        ind(f, d); fprintf(f, "triangle_object");
        pv(f, "m", n.tri.m); pv(f, "u", n.tri.u); pv(f, "v", n.tri.v); pv(f, "mn", n.tri.mn); pv(f, "un", n.tri.un); pv(f, "vn", n.tri.vn);
        fprintf(f, " mat="); dump_mat(g, f, n.mat); fprintf(f, "\n");
        break;
    case NodeKind::Box:
        ind(f, d); fprintf(f, "box"); pv(f, "min", n.box.min); pv(f, "max", n.box.max); fprintf(f, "\n");
        dump_obj(g, f, n.child, d + 1);
        break;
    case NodeKind::List:
        ind(f, d); fprintf(f, "list n=%zu hasBox=%d", n.children.size(), (int) n.has_box);
        if (n.has_box) { pv(f, "min", n.box.min); pv(f, "max", n.box.max); }
        fprintf(f, "\n");
        for (int c : n.children) dump_obj(g, f, c, d + 1);
        break;
    case NodeKind::Bvh:
        ind(f, d); fprintf(f, "bvh order=%02x same=%d", (unsigned) n.order, (int) (n.left == n.right));
        pv(f, "min", n.box.min); pv(f, "max", n.box.max); fprintf(f, "\n");
        dump_obj(g, f, n.left, d + 1);
        dump_obj(g, f, n.right, d + 1);
        break;
    case NodeKind::Translate:
        ind(f, d); fprintf(f, "translate"); pv(f, "offset", n.offset); fprintf(f, "\n");
        dump_obj(g, f, n.child, d + 1);
        break;
    case NodeKind::RotateY:
        ind(f, d); fprintf(f, "rotate_y"); pf(f, "sin", n.sin_theta); pf(f, "cos", n.cos_theta);
        fprintf(f, " hasBox=%d", (int) n.has_box); pv(f, "min", n.box.min); pv(f, "max", n.box.max); fprintf(f, "\n");
        dump_obj(g, f, n.child, d + 1);
        break;
    case NodeKind::Volume:
        ind(f, d); fprintf(f, "volume"); pf(f, "density", n.density); fprintf(f, " mat="); dump_mat(g, f, n.mat); fprintf(f, "\n");
        dump_obj(g, f, n.child, d + 1);
        break;
    case NodeKind::PodBvh: {
        const Mesh &m = g.meshes[n.mesh];
        ind(f, d); fprintf(f, "podbvh prims=%zu nodes=%zu root=0\n", m.tris.size(), m.nodes.size());
        for (size_t i = 0; i < m.nodes.size(); i++) {
            const PodNode &pn = m.nodes[i];
            ind(f, d + 1); fprintf(f, "node %zu left=%u off=%u cnt=%u order=%02x", i, pn.prim_count ? 0u : pn.left, pn.prim_offset,
                                   pn.prim_count, pn.prim_count ? 0u : (unsigned) pn.order);
            pv(f, "min", pn.box.min); pv(f, "max", pn.box.max); fprintf(f, "\n");
        }
        for (size_t i = 0; i < m.tris.size(); i++) {
            const Triangle &t = m.tris[i];
            ind(f, d + 1); fprintf(f, "tri %zu", i);
            pv(f, "m", t.m); pv(f, "u", t.u); pv(f, "v", t.v); pv(f, "mn", t.mn); pv(f, "un", t.un); pv(f, "vn", t.vn);
            if (i == 0) { fprintf(f, " mat="); dump_mat(g, f, m.mat); }
            fprintf(f, "\n");
        }
        break;
    }
    }
}

void dump_scene(const SceneGraph &g, FILE *f) {
    const Camera &c = g.camera;
    fprintf(f, "camera"); pv(f, "origin", c.origin); pv(f, "u", c.u); pv(f, "v", c.v); pv(f, "w", c.w);
    pv(f, "llcorner", c.llcorner); pv(f, "horz", c.horz); pv(f, "vert", c.vert);
    pf(f, "lens_radius", c.lens_radius); pf(f, "time0", c.time0); pf(f, "time1", c.time1); fprintf(f, "\n");
    fprintf(f, "objects\n");
    dump_obj(g, f, g.objects, 1);
    if (g.biased >= 0) {
        fprintf(f, "biased\n");
        dump_obj(g, f, g.biased, 1);
    } else {
        fprintf(f, "biased none\n");
    }
}

}  // namespace mrt
