// Wavefront OBJ reader: a line tokenizer that accepts the subset the reference accepts (obj_loader.cpp:14-162) and bakes
// the same transforms in the same order -- `v`, `vn`, `f a b c` (no normals declared so far) or `f a//an b//bn c//cn`;
// vertices are scaled, rotated, then translated; normals are multiplied as row vectors by the inverse rotation; flip
// swaps the first and last corner (and their normals).  Scene parity (tests/test_scene_parity.py) compares the resulting
// triangles bit for bit with the reference's.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

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

// One line of the file, cut into whitespace-separated tokens (views into the line buffer).
struct Tokens {
    const char *tok[8];
    int n = 0;
    explicit Tokens(char *line) {
        for (char *p = line; *p && n < 8;) {
            while (*p == ' ' || *p == '\t' || *p == '\r') p++;
            if (!*p) break;
            tok[n++] = p;
            while (*p && *p != ' ' && *p != '\t' && *p != '\r') p++;
            if (*p) *p++ = 0;
        }
    }
};
static bool parse_floats(const Tokens &t, float out[3]) {
    if (t.n < 4) return false;
    for (int i = 0; i < 3; i++) {
        char *end = nullptr;
        out[i] = strtof(t.tok[1 + i], &end);
        if (end == t.tok[1 + i]) return false;
    }
    return true;
}
// corner "a" (positions only) or "a//an" (position // normal); the reference accepts nothing else (obj_loader.cpp:71-134)
static bool parse_corner(const char *s, bool with_normal, int *vi, int *ni) {
    char *end = nullptr;
    *vi = (int) strtol(s, &end, 0);
    if (end == s) return false;
    if (!with_normal) return *end == 0;
    if (end[0] != '/' || end[1] != '/') return false;
    const char *q = end + 2;
    *ni = (int) strtol(q, &end, 0);
    return end != q;
}

bool read_obj(const std::string &path, bool flip, const M4 &scale, H3 translate, const M4 &rotate, std::vector<Triangle> *out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string text;
    char chunk[1 << 16];
    for (size_t got; (got = fread(chunk, 1, sizeof(chunk), f)) > 0;) text.append(chunk, got);
    fclose(f);

    std::vector<H3> verts, norms;
    const M4 inv_rot = M4::invert(rotate);
    // vertex: scale, rotate, translate (obj_loader.cpp:80-90); normal: row vector times the inverse rotation (:112-118)
    auto place = [&](int index) { return rotate.mul_col(scale.mul_col(verts[(size_t) index - 1])) + translate; };
    size_t pos = 0;
    while (pos < text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::string line = text.substr(pos, eol - pos);
        pos = eol + 1;
        Tokens t(line.data());
        if (t.n == 0 || t.tok[0][0] == '#') continue;
        float v[3];
        if (!strcmp(t.tok[0], "v")) {
            if (!parse_floats(t, v)) break;          // malformed: the reference stops reading here
            verts.push_back(H3(v[0], v[1], v[2]));
        } else if (!strcmp(t.tok[0], "vn")) {
            if (!parse_floats(t, v)) break;
            norms.push_back(H3(v[0], v[1], v[2]));
        } else if (!strcmp(t.tok[0], "f")) {
            // the face format is decided by whether the file has declared normals SO FAR (obj_loader.cpp:71); only the
            // first three corners of a face are used
            const bool with_normal = !norms.empty();
            int vi[3], ni[3] = {0, 0, 0};
            bool ok = t.n >= 4;
            for (int c = 0; ok && c < 3; c++) ok = parse_corner(t.tok[1 + c], with_normal, &vi[c], &ni[c]);
            if (!ok) break;
            if (flip) { std::swap(vi[0], vi[2]); std::swap(ni[0], ni[2]); }
            const H3 a = place(vi[0]), b = place(vi[1]), c = place(vi[2]);
            if (with_normal)
                out->push_back(make_triangle(a, b, c, inv_rot.mul_row(norms[(size_t) ni[0] - 1]), inv_rot.mul_row(norms[(size_t) ni[1] - 1]),
                                             inv_rot.mul_row(norms[(size_t) ni[2] - 1])));
            else
                out->push_back(make_triangle(a, b, c));
        }
        // anything else (vt, g, usemtl, s, ...) is skipped like a comment
    }
    return true;
}

}  // namespace mrt
