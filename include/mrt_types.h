// Flattened GPU scene layout shared by the host flattener, the CUDA kernels and
// the C ABI (include/mrt_gpu.h re-exports the POD description).
//
// The reference scene is a pointer graph of virtual scene_object's
// (scene_object.h:20-31).  Here every object lives in a typed structure-of-
// arrays table of 16-byte records (read with one coalesced/vectorised 128-bit
// load each) and is addressed by a 28-bit typed reference:
//
//      ref = (type << 24) | index          type: 4 bits, index: 24 bits
//
// Tables (all `MrtF4`, 16 B aligned):
//   sphere[3*i+0] = (c0.x, c0.y, c0.z, radius)                 sphere.h:11-17
//   sphere[3*i+1] = (c1.x, c1.y, c1.z, bits(mat | moving<<31))
//   sphere[3*i+2] = (time0, time1, 0, 0)
//   rect[2*i+0]   = (a0, a1, b0, b1)                            rect.h:6-11
//   rect[2*i+1]   = (k, normal_sign, bits(mat), 0)     axis is in the ref type
//   list[2*i+0]   = (box.min.xyz, bits(first_child))            scene_object.h:37-44
//   list[2*i+1]   = (box.max.xyz, bits(count | hasBox<<31))
//                   children = child[first_child .. +count], then MRT_REF_END
//   Both of the reference's BVHs -- bvh_node<T> (scene_object.h:138-244) and pod_bvh<triangle>
//   (triangle.h:46-213) -- have the same traversal rule (test the node's own box; visit the closer child
//   by node_order & dirMask; return on its first hit; else the farther child; tmin/tmax never change
//   inside one tree).  They are flattened into ONE node format in which a node carries its CHILDREN's
//   boxes, so one visit decides both children (the box test is a pure function of the ray and the
//   unchanged tmin/tmax, so testing a child's box at its parent gives the same answer):
//   bvh[2*i+0]    = (root box.min.xyz, bits(root child ref))        root header: the tree's own box test
//   bvh[2*i+1]    = (root box.max.xyz, 0)
//   node2[4*i+0]  = (left  box.min.xyz, bits(left ref  | (order & 15) << 28))
//   node2[4*i+1]  = (left  box.max.xyz, bits(right ref | (order >> 4) << 28))
//   node2[4*i+2]  = (right box.min.xyz, bits(flags))  flags bit0/bit1: left/right child has a box to test; bits 2-3 / 4-5: kind of
//                   the left / right child (0 inner node, 1 object_list, 2 triangle leaf) for the warp-cooperative traversal
//   node2[4*i+3]  = (right box.max.xyz, 0)
//                   child refs: NODE2 (inner), LIST (header copy with hasBox = 0: its box is the one stored
//                   here), TRILEAF, or any other object (no box flag -> visited unconditionally)
//   trileaf[2*i]  = first triangle, trileaf[2*i+1] = triangle count      (pod_bvh leaf, triangle.h:179-187)
//   tri[3*i+0..2] = (m.xyz, bits(mat)), (u.xyz, 0), (v.xyz, 0)         triangle.h:13-22
//   trin[3*i+0..2]= (mn.xyz,0), (un.xyz,0), (vn.xyz,0)
//   xlate[3*i+0]  = (offset.xyz, bits(child))                          scene_object.h:325-333
//   xlate[3*i+1]  = (cull box min.xyz, bits(hasCullBox))   bounds of the translated object in the parent frame,
//   xlate[3*i+2]  = (cull box max.xyz, slack)              inflated by slack = 1e-3 x scene scale: a ray that
//                   misses it skips the node (conservative; the reference tests nothing here)
//   rot[3*i+0]    = (bbox.min.xyz, bits(child))                        scene_object.h:340-355
//   rot[3*i+1]    = (bbox.max.xyz, bits(hasBox))
//   rot[3*i+2]    = (sin_theta, cos_theta, 0, 0)
//   vol[i]        = (bits(boundary), density, bits(mat), 0)            volumes.h:8-12
//   mat[i]        = (bits(kind | needs_uv<<8), bits(tex), param, 0)    material.h
//   tex[i]        = COLOR  (bits(kind), r, g, b)                       texture.h
//                   CHECKER(bits(kind), bits(even), bits(odd), scale)
//                   PERLIN (bits(kind), scale, 0, 0)
//                   IMAGE  (bits(kind), bits(width), bits(height), bits(byte offset into image[]))
//   perlin_vec[256] = (ranvec.xyz, 0)   perlin_perm[3*256] = perm_x,perm_y,perm_z (texture.cpp:107-112)
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
#if defined(__GNUC__) || defined(__CUDACC__)
    __attribute__((aligned(16)))
#endif
    MrtF4 {
    float x, y, z, w;
} MrtF4;

enum MrtRefType {
    MRT_T_SPHERE = 0,
    MRT_T_RECT_XY = 1,
    MRT_T_RECT_XZ = 2,
    MRT_T_RECT_YZ = 3,
    MRT_T_LIST = 4,
    MRT_T_BVH = 5,
    MRT_T_NODE2 = 6,
    MRT_T_TRANSLATE = 7,
    MRT_T_ROTATE_Y = 8,
    MRT_T_VOLUME = 9,
    MRT_T_TRILEAF = 10,
    MRT_T_TRI = 11, /* triangle_scene_object: one triangle of tri[] / trin[] as a list child (triangle.cpp:5-175) */
    MRT_T_END = 15 /* list terminator */
};
#define MRT_REF(type, index) ((uint32_t) (((uint32_t) (type) << 24) | ((uint32_t) (index) & 0xFFFFFFu)))
#define MRT_REF_TYPE(ref) (((ref) >> 24) & 15u)
#define MRT_REF_INDEX(ref) ((ref) & 0xFFFFFFu)
#define MRT_REF_END MRT_REF(MRT_T_END, 0)
#define MRT_REF_NONE 0xFFFFFFFFu

enum MrtMatKind {
    MRT_M_LAMBERTIAN = 0,
    MRT_M_ISOTROPIC = 1,
    MRT_M_METAL = 2,
    MRT_M_DIELECTRIC = 3,
    MRT_M_LIGHT = 4
};
#define MRT_MAT_NEEDS_UV 0x100u

/* Scene features (MrtSceneDesc.features): which object classes / materials / textures occur.  The renderer
   picks a kernel specialised for a superset of this mask. */
#define MRT_FEAT_TREES 1u      /* bvh_node / pod_bvh trees, triangles */
#define MRT_FEAT_VOLUMES 2u    /* constant_volume (+ isotropic phase function) */
#define MRT_FEAT_XFORM 4u      /* translate / rotate_y */
#define MRT_FEAT_TEX 8u        /* checker / perlin / image textures (and the uv they need) */
#define MRT_FEAT_METAL 16u
#define MRT_FEAT_DIELECTRIC 32u
#define MRT_FEAT_MOVING 64u    /* moving spheres */
#define MRT_FEAT_LIGHT_SPHERE 128u /* the light list holds something other than xz_rects */
#define MRT_FEAT_SPHERES 256u     /* the scene has spheres */
#define MRT_FEAT_TRIS 512u        /* triangle leaves (pod_bvh) */
#define MRT_FEAT_LEAF_LISTS 1024u /* some tree leaf is an object_list (bvh_node leaves) */
#define MRT_FEAT_TRI_OBJECT 2048u /* lone triangles as scene objects (triangle_scene_object) */
#define MRT_FEAT_ALL 4095u

enum MrtTexKind { MRT_X_COLOR = 0, MRT_X_CHECKER = 1, MRT_X_PERLIN = 2, MRT_X_IMAGE = 3 };

/* camera.h:8-14 */
typedef struct MrtCamera {
    float origin[3];
    float u[3], v[3], w[3];
    float llcorner[3];
    float horz[3];
    float vert[3];
    float lens_radius;
    float time0, time1;
} MrtCamera;

/* The flattened scene.  All pointers are HOST pointers when handed to
   mrt_gpu_scene_upload(); the library owns its device copies. */
typedef struct MrtSceneDesc {
    uint32_t root;            /* typed ref of scene.objects (scene.h:20) */
    uint32_t n_lights;        /* scene.biased_objects: object_list count (0 = nullptr) */
    const uint32_t *lights;   /* typed refs (sphere or xz_rect have a pdf; others evaluate to 0) */
    uint32_t sky;             /* 1: sky gradient on miss (sceneSelect < SCENE_CORNELL_BOX, main.cpp:110) */
    uint32_t stack_words;     /* worst-case traversal stack depth in 32-bit words (computed by the flattener) */
    uint32_t features;        /* MRT_FEAT_* mask of what the scene contains (0 is treated as MRT_FEAT_ALL) */
    uint32_t stack_words_coop; /* the same when BVH trees are traversed warp-cooperatively (the per-lane stack then holds no tree
                                  frames); 0 = the scene's trees do not qualify (see coop_tree.cuh), per-lane traversal only */
    MrtCamera camera;

    const MrtF4 *sphere;  uint32_t n_sphere;
    const MrtF4 *rect;    uint32_t n_rect;
    const MrtF4 *list;    uint32_t n_list;
    const uint32_t *child; uint32_t n_child;
    const MrtF4 *bvh;     uint32_t n_bvh;
    const MrtF4 *node2;   uint32_t n_node2;
    const uint32_t *trileaf; uint32_t n_trileaf;
    const MrtF4 *tri;     uint32_t n_tri;
    const MrtF4 *trin;
    const MrtF4 *xlate;   uint32_t n_xlate;
    const MrtF4 *rot;     uint32_t n_rot;
    const MrtF4 *vol;     uint32_t n_vol;
    const MrtF4 *mat;     uint32_t n_mat;
    const MrtF4 *tex;     uint32_t n_tex;
    const MrtF4 *perlin_vec;      /* 256 entries or NULL */
    const int32_t *perlin_perm;   /* 768 entries or NULL */
    const uint8_t *image;  uint64_t n_image_bytes;   /* all RGB8 images, concatenated */
} MrtSceneDesc;

/* One render call = all pixels x samples [sample_begin, sample_end). */
typedef struct MrtRenderParams {
    uint32_t width, height;
    uint32_t samples;         /* N = floor(sqrt(spp))^2 (main.cpp:319-320); the stream id uses this N */
    uint32_t sample_begin, sample_end;
    uint32_t max_bounces;     /* MRT_Params::maxBounces (cmdline_parser.h:13) */
    uint64_t seed;            /* PCG32 initstate; the stream (initseq) is (y*W+x)*N+s */
    float max_luminance;      /* applied by finalize only (main.cpp:170-173) */
    uint32_t flags;           /* MRT_RENDER_* */
    /* Crop window: only pixels [crop_x0, crop_x1) x [crop_y0, crop_y1) of the width x height frame are rendered and
       the accumulator holds (crop_x1-crop_x0) x (crop_y1-crop_y0) pixels, row-major.  Sub-pixel positions (u, v) and
       PCG32 stream ids stay those of the FULL frame (main.cpp:156-157), so a window is bit-identical to the same
       pixels of a full render -- used for tile sharding and for parity checks at full-size configurations.
       All four zero = the whole frame. */
    uint32_t crop_x0, crop_y0, crop_x1, crop_y1;
} MrtRenderParams;

#define MRT_RENDER_ACCUMULATE 1u /* add to the accumulator instead of overwriting it */
#define MRT_RENDER_CONTINUE 2u   /* this launch continues the previous one(s): ray / path statistics and the kernel-time window of
                                    mrt_gpu_stats keep running instead of restarting */

#ifdef __cplusplus
}
#endif
