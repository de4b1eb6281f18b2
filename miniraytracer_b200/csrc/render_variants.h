// Kernel variants specialised by scene-feature mask (MRT_FEAT_*).  One translation unit per mask
// (render_variant_<name>.cu) instantiates the two megakernels of render_kernels.cuh for it; the launcher
// takes the first variant whose mask covers the scene's.
#pragma once
#include <stdint.h>

#include "mrt_types.h"

namespace mrt {

// exact masks of the two Cornell configurations (no metal; C3 has no dielectric either)
#define MRT_VARIANT_CORNELL (MRT_FEAT_XFORM | MRT_FEAT_DIELECTRIC | MRT_FEAT_SPHERES)
#define MRT_VARIANT_CORNELL_VOL (MRT_FEAT_XFORM | MRT_FEAT_VOLUMES)
// Cornell-style scenes: lists, rects, spheres, transforms; lambertian / metal / dielectric / light; colour textures
#define MRT_VARIANT_LISTS (MRT_FEAT_XFORM | MRT_FEAT_METAL | MRT_FEAT_DIELECTRIC | MRT_FEAT_SPHERES)
// ... plus constant-density volumes
#define MRT_VARIANT_LISTS_VOL (MRT_VARIANT_LISTS | MRT_FEAT_VOLUMES)
// triangle meshes in a Cornell box: no spheres, no transforms, volumes, procedural / image textures
#define MRT_VARIANT_TREES (MRT_FEAT_TREES | MRT_FEAT_TRIS | MRT_FEAT_METAL | MRT_FEAT_DIELECTRIC)
// sphere BVH (object_list leaves) with procedural textures and motion blur ("In One Weekend").  MRT_FEAT_TRIS stays in although
// these scenes have no triangles: without the triangle-leaf case nvcc 12.9 lays the per-lane traversal loop out differently
// and the kernel is 57 % SLOWER on scene 0 (measured, profiles/r2_notes.md) -- a code-generation accident, kept on the good side
#define MRT_VARIANT_TREES_TEX (MRT_FEAT_TREES | MRT_FEAT_TRIS | MRT_FEAT_LEAF_LISTS | MRT_FEAT_SPHERES | MRT_FEAT_METAL | MRT_FEAT_DIELECTRIC | MRT_FEAT_TEX | MRT_FEAT_MOVING)

// everything the nine stock scenes can contain (no triangle_scene_object: no stock scene has one)
#define MRT_VARIANT_STOCK (MRT_FEAT_ALL & ~MRT_FEAT_TRI_OBJECT)

const void *variant_cornell(int kind, int minb);
const void *variant_cornell_vol(int kind, int minb);
const void *variant_lists(int kind, int minb);
const void *variant_lists_vol(int kind, int minb);
const void *variant_trees(int kind, int minb);
const void *variant_trees_tex(int kind, int minb);
const void *variant_all(int kind, int minb);
const void *variant_full(int kind, int minb);

struct Variant {
    uint32_t mask;
    const void *(*get)(int kind, int minb);
    const char *name;
};
inline const Variant *pick_variant(uint32_t scene_features) {
    static const Variant table[] = {
        {MRT_VARIANT_CORNELL, variant_cornell, "cornell"},
        {MRT_VARIANT_CORNELL_VOL, variant_cornell_vol, "cornell+volumes"},
        {MRT_VARIANT_LISTS, variant_lists, "lists"},
        {MRT_VARIANT_LISTS_VOL, variant_lists_vol, "lists+volumes"},
        {MRT_VARIANT_TREES, variant_trees, "trees"},
        {MRT_VARIANT_TREES_TEX, variant_trees_tex, "trees+textures"},
        {MRT_VARIANT_STOCK, variant_all, "all stock features"},
        {MRT_FEAT_ALL, variant_full, "all"},
    };
    if (scene_features == 0) scene_features = MRT_FEAT_ALL;   // unknown: keep everything
    for (const Variant &v : table)
        if ((scene_features & ~v.mask) == 0) return &v;
    return &table[sizeof(table) / sizeof(table[0]) - 1];
}

}  // namespace mrt
