#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle) -- not part of the shipped product.

Copies the reference's first-party C++ sources from /root/reference into a
scratch directory (never into the repo) and applies the mechanical patches that
let g++ build them headless.  Every patch is asserted to match exactly once (or
the stated count) so a changed reference fails loudly instead of silently.

Patch list:
  P1 mrt_math.h:66   '#error INSERT LZCNT INTRINSIC HERE' -> __builtin_clz
  P2 platform_*.cpp  dropped; oracle/ref_platform_headless.cpp supplies platform.h
  P3 onb.h:9-16      anonymous union of ctor'd members -> plain members
  P4 mat4.h:10-13    anonymous struct alias c0..c3 removed, uses rewritten to c[i]
  P5 triangle.h      + #include <cstring> (memcpy at :85)
  P6 build flags     -D__cdecl= -D__stdcall= -fpermissive (see build_ref.sh)
  P7 scene.cpp:509   "teapot3_no_vt.obj" is mis-cased on case-sensitive file
                     systems; the asset is staged under the lower-case name.
  P8 evaluation order: the reference's documented compiler (clang) evaluates
     function arguments left->right, g++ right->left.  The RNG-consuming
     argument lists are sequenced explicitly left->right so that the oracle is
     compiler independent: pcg.cpp:73,115  rect.cpp:105
     scene.cpp:78,86,90,156,164,169,450.
  P9 access only: `public:` added to pod_bvh (triangle.h:58) so the scene dump can read it.
  P10 access only: `public:` added to triangle_scene_object (triangle.h:326) for the same reason.
No arithmetic is changed by any source patch.  Separately, at LINK time, the six libm functions the path
calls are bound to correctly rounded versions (oracle/cr_libm.cpp; rationale in
miniraytracer_b200/csrc/mrt_libm.h); MRT_ORACLE_LIBM=host restores the host libm's own.
"""
import os
import re
import shutil
import sys

SRC = sys.argv[1]
DST = sys.argv[2]

FILES = [f for f in os.listdir(SRC)
         if (f.endswith(".cpp") or f.endswith(".h")) and not f.startswith("platform_")]
os.makedirs(DST, exist_ok=True)
for f in FILES:
    shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))


def patch(fname, old, new, count=1, regex=False):
    p = os.path.join(DST, fname)
    s = open(p, encoding="utf-8", errors="surrogateescape").read()
    if regex:
        s2, n = re.subn(old, new, s)
    else:
        n = s.count(old)
        s2 = s.replace(old, new)
    if count is not None and n != count:
        raise SystemExit(f"patch failed: {fname}: expected {count} matches of {old!r}, got {n}")
    open(p, "w", encoding="utf-8", errors="surrogateescape").write(s2)


# P1
patch("mrt_math.h", "#error INSERT LZCNT INTRINSIC HERE", "uint32 i = (uint32) __builtin_clz(v);")

# P3
patch("onb.h",
      re.compile(r"union \{\s*struct \{\s*Vec3 u;\s*Vec3 v;\s*Vec3 w;\s*\};\s*Vec3 axis\[3\];\s*\};"),
      "Vec3 u; Vec3 v; Vec3 w;", regex=True)

# P4
patch("mat4.h", re.compile(r"struct \{\s*Vec4 c0, c1, c2, c3;\s*\};"), "", regex=True)
patch("mat4.h", ") : c0(c0), c1(c1), c2(c2), c3(c3) {}",
      ") { c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3; }", count=2)
for f in ("mat4.h", "mat4.cpp"):
    patch(f, re.compile(r"\b(m|res)\.c([0-3])\b"), r"\1.c[\2]", count=None, regex=True)
    patch(f, re.compile(r"(?<![\w.\]])c([0-3])\.(m|x)\b"), r"c[\1].\2", count=None, regex=True)

# P5
patch("triangle.h", '#include "scene_object.h"', '#include <cstring>\n#include "scene_object.h"')

# P9 access only: pod_bvh keeps its arrays private (triangle.h:58-65); the scene
# dump of the harness needs to read them.
patch("triangle.h", "class pod_bvh final : public scene_object {",
      "class pod_bvh final : public scene_object {\npublic:")
# P10 access only: the same for triangle_scene_object's members (triangle.h:326-334)
patch("triangle.h", "class triangle_scene_object final : public scene_object {",
      "class triangle_scene_object final : public scene_object {\npublic:")

# P8 -- explicit left->right sequencing of RNG draws
patch("pcg.cpp",
      "p = 2.0f * Vec3(randf(rng), randf(rng), randf(rng)) - Vec3(1, 1, 1);",
      "{ float rx_ = randf(rng); float ry_ = randf(rng); float rz_ = randf(rng); "
      "p = 2.0f * Vec3(rx_, ry_, rz_) - Vec3(1, 1, 1); }")
patch("pcg.cpp",
      "p = 2.0f * Vec3(randf(rng), randf(rng), 0) - Vec3(1, 1, 0);",
      "{ float rx_ = randf(rng); float ry_ = randf(rng); "
      "p = 2.0f * Vec3(rx_, ry_, 0) - Vec3(1, 1, 0); }")
patch("rect.cpp",
      "Vec3 rand = Vec3(x0 + randf() * (x1 - x0), y, z0 + randf() * (z1 - z0));",
      "float rx_ = randf(); float rz_ = randf(); "
      "Vec3 rand = Vec3(x0 + rx_ * (x1 - x0), y, z0 + rz_ * (z1 - z0));")
patch("scene.cpp",
      "Vec3 center(a + 0.9f * randf(), 0.2f, b + 0.9f * randf());",
      "float cx_ = randf(); float cz_ = randf(); "
      "Vec3 center(a + 0.9f * cx_, 0.2f, b + 0.9f * cz_);", count=2)
patch("scene.cpp",
      "mat = new lambertian(new color_tex(Vec3(randf()*randf(), randf()*randf(), randf()*randf())));",
      "{ float r0_ = randf(); float r1_ = randf(); float r2_ = randf(); float r3_ = randf(); "
      "float r4_ = randf(); float r5_ = randf(); "
      "mat = new lambertian(new color_tex(Vec3(r0_*r1_, r2_*r3_, r4_*r5_))); }", count=2)
patch("scene.cpp",
      "mat = new metal(new color_tex(0.5f * Vec3(1 + randf(), 1 + randf(), 1 + randf())), randf());",
      "{ float r0_ = randf(); float r1_ = randf(); float r2_ = randf(); float r3_ = randf(); "
      "mat = new metal(new color_tex(0.5f * Vec3(1 + r0_, 1 + r1_, 1 + r2_)), r3_); }", count=2)
patch("scene.cpp",
      "spherelist[i] = new sphere(Vec3(165 * randf(), 165 * randf(), 165 * randf()), 10, white);",
      "{ float r0_ = randf(); float r1_ = randf(); float r2_ = randf(); "
      "spherelist[i] = new sphere(Vec3(165 * r0_, 165 * r1_, 165 * r2_), 10, white); }")

print(f"patched {len(FILES)} files into {DST}")
