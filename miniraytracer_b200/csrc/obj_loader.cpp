// Wavefront OBJ reader with the reference's accepted subset and transform order
// (obj_loader.cpp:14-162): `v`, `vn`, `f a b c` (no normals in the file so far)
// or `f a//an b//bn c//cn`; vertices are scaled, rotated, then translated;
// normals are multiplied as row vectors by the inverse rotation; flip swaps the
// first and last corner (and their normals).
#include <cstdio>
#include <cstring>

#include "scene_graph.h"

namespace mrt {

static Triangle make_triangle(H3 a, H3 b, H3 c) {   // triangle.cpp:178-192
    Triangle t;
    t.m = a;
    t.u = b - a;
    t.v = c - a;
    t.mn = t.un = t.vn = hnormalize(hcross(t.u, t.v));
    return t;
}
static Triangle make_triangle(H3 a, H3 b, H3 c, H3 an, H3 bn, H3 cn) {   // triangle.cpp:194-212
    Triangle t;
    t.m = a;
    t.u = b - a;
    t.v = c - a;
    t.mn = an;
    t.un = bn;
    t.vn = cn;
    return t;
}

static void skip_line(FILE *f) {
    char buf[64] = {0};
    do {
        if (!fgets(buf, sizeof(buf), f)) break;
    } while (buf[strlen(buf) - 1] != '\n' && !feof(f));
}

bool read_obj(const std::string &path, bool flip, const M4 &scale, H3 translate, const M4 &rotate, std::vector<Triangle> *out) {
    std::vector<H3> verts, norms;
    M4 invRot = M4::invert(rotate);
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return false;
    float x, y, z;
    while (!feof(f)) {
        char s = (char) getc(f);
        if (s == '#') {
            skip_line(f);
        } else if (s == '\n' || s == ' ' || s == '\t') {
        } else if (s == 'v') {
            s = (char) getc(f);
            if (s == ' ' || s == '\t') {
                if (fscanf(f, " %f %f %f ", &x, &y, &z) == 3) verts.push_back(H3(x, y, z));
                else break;
            } else if (s == 'n') {
                if (fscanf(f, " %f %f %f ", &x, &y, &z) == 3) norms.push_back(H3(x, y, z));
                else break;
            }
        } else if (s == 'f') {
            if (norms.empty()) {
                int ai, bi, ci;
                if (fscanf(f, " %i %i %i ", &ai, &bi, &ci) != 3) break;
                if (flip) std::swap(ai, ci);
                H3 a = scale.mul_col(verts[ai - 1]), b = scale.mul_col(verts[bi - 1]), c = scale.mul_col(verts[ci - 1]);
                a = rotate.mul_col(a); b = rotate.mul_col(b); c = rotate.mul_col(c);
                a = a + translate; b = b + translate; c = c + translate;
                out->push_back(make_triangle(a, b, c));
            } else {
                int ai, bi, ci, ani, bni, cni;
                if (fscanf(f, " %i//%i %i//%i %i//%i ", &ai, &ani, &bi, &bni, &ci, &cni) != 6) break;
                if (flip) { std::swap(ai, ci); std::swap(ani, cni); }
                H3 a = scale.mul_col(verts[ai - 1]), b = scale.mul_col(verts[bi - 1]), c = scale.mul_col(verts[ci - 1]);
                H3 an = invRot.mul_row(norms[ani - 1]), bn = invRot.mul_row(norms[bni - 1]), cn = invRot.mul_row(norms[cni - 1]);
                a = rotate.mul_col(a); b = rotate.mul_col(b); c = rotate.mul_col(c);
                a = a + translate; b = b + translate; c = c + translate;
                out->push_back(make_triangle(a, b, c, an, bn, cn));
            }
        } else {
            skip_line(f);
        }
    }
    fclose(f);
    return true;
}

}  // namespace mrt
