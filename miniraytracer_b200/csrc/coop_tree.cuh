// Warp-cooperative traversal of the reference's BVH trees (bvh_node<T>, scene_object.h:208-244; pod_bvh<triangle>,
// triangle.h:171-213).
//
// Why.  Per-lane depth-first traversal leaves most lanes of a warp idle inside a tree: the trip counts are heavy-
// tailed (a ray that grazes a mesh visits every node whose box it crosses, a ray that hits stops at its first leaf),
// and node visits, leaf tests and stack pops are different instruction streams (ncu, round 1: 5.8 of 32 lanes per
// instruction on the triangle scene, 10 on the final scene, 14 on the sphere scene).
//
// What makes sharing legal.  Inside one tree (tmin, tmax) is constant -- the reference never tightens tmax between
// children and returns on the FIRST child, front to back by node_order & dirMask, that reports a hit
// (scene_object.h:224-239, triangle.h:188-210).  So (i) every subtree's own answer is a pure function of the ray,
// independent of what was visited before, and (ii) the tree's answer is the hit of the leaf that comes first in the
// ray's depth-first order among all leaves that report a hit.  The order is made explicit as a RANK: the path from
// the root, one bit per level (0 = the child the ray visits first), most significant bit first, closed by a marker
// bit.  For two different leaves the integer order of the ranks is their depth-first order; a pending node whose rank
// is above the best hit's rank lies entirely behind that leaf and is dropped.
//
// How.  When lanes of a warp stand at tree roots (isect_run<true> suspends there), the warp traverses all those trees
// together: the rays are stored in shared memory, work items (ray, node, rank) live on a per-warp stack and
// (ray, leaf, rank) items in a per-warp queue, also in shared memory.  A NODE STEP pops up to 32 node items -- of
// whichever rays -- and every lane tests the two child boxes of its node with the reference's exact slab test
// (aabb_hit) and pushes the children that are hit, the farther one first.  A LEAF STEP runs when a batch of leaves is
// queued (or nothing else is left): object_list leaves are evaluated one per lane in the reference's order
// (scene_object.h:88-95), triangle leaves likewise (triangle.h:179-187); a leaf that hits posts its rank with an atomic
// minimum and the winner stores its hit record for the ray's owner lane.  All hit
// tests are the same functions, on the same operands, as in the per-lane traversal, so the result is bit-identical;
// the extra work is speculative (nodes behind a hit that was not known yet) and only occupies lanes that would idle.
//
// Bounded memory: the node stack is processed 32 items at a time while it holds <= kCoopThresh items, one item at a
// time (plain depth-first order, which nets at most one item per level) above that; it can never exceed
// kCoopThresh + 32 + kCoopMaxDepth items.
#pragma once
#include "trace_core.h"

namespace mrt {

constexpr uint32_t kCoopMaxDepth = 31;      // rank bits; deeper trees use the per-lane traversal (checked at upload)
constexpr uint32_t kCoopThresh = 224;
constexpr uint32_t kCoopNodeCap = kCoopThresh + 32 + kCoopMaxDepth + 1;   // 288 items
constexpr uint32_t kCoopLeafCap = 31 + 64 + 1;                            // < 32 waiting + one node step's pushes
constexpr uint32_t kCoopRayWords = 13;
constexpr uint32_t kCoopRecWords = 10;
// shared memory per warp, in 32-bit words (8-byte aligned)
constexpr uint32_t kCoopWords = kCoopRayWords * 32u + 32u + kCoopRecWords * 32u + 2u * kCoopNodeCap + 2u * kCoopLeafCap;
constexpr uint32_t kRankNone = 0xFFFFFFFFu;

// Work item: x = table index (24 bits) | kind << 24 | job lane << 26, y = rank.  The kind of a tree node's children is
// precomputed by the flattener (node2 flags bits 2-5): 0 = inner node, 1 = object_list leaf, 2 = triangle leaf.
#define MRT_COOP_KIND_NODE 0u
#define MRT_COOP_KIND_LIST 1u
#define MRT_COOP_KIND_TRILEAF 2u

struct CoopArea {   // views into one warp's shared-memory area
    uint32_t *ray;    // [kCoopRayWords][32]: word k of job lane j at [k * 32 + j]
    uint32_t *best;   // [32] rank of the first (depth-first) leaf that reported a hit so far, kRankNone = none yet
    uint32_t *rec;    // [kCoopRecWords][32] hit record of that leaf (t, p, n, u, v, mat)
    uint2 *nodes;     // node stack
    uint2 *leaves;    // leaf queue
    MRT_HD void bind(uint32_t *base) {
        ray = base;
        best = base + kCoopRayWords * 32u;
        rec = best + 32u;
        nodes = reinterpret_cast<uint2 *>(rec + kCoopRecWords * 32u);
        leaves = nodes + kCoopNodeCap;
    }
};

struct CoopStats { uint32_t node_steps, node_items, leaf_steps, leaf_items; };   // per warp and chunk; flushed by the kernel

// The slab test of aabb_hit for a ray whose reciprocal direction is finite in all three components (no +-0 direction
// component): then no product is NaN, the SSE select rule of aabb.h:69-76 (second operand on NaN) never fires and
// min / max are the plain ones; picking the near / far plane by the sign of inv before the multiply yields the very
// values aabb_hit has after its swap.  Same boolean, ~24 instead of ~40 instructions.  n? = inv.? < 0.
MRT_HD bool aabb_hit_finite(const MrtF4 &bmin, const MrtF4 &bmax, V3 o, V3 inv, bool nx, bool ny, bool nz, float tmin, float tmax) {
    const float t0x = ((nx ? bmax.x : bmin.x) - o.x) * inv.x, t1x = ((nx ? bmin.x : bmax.x) - o.x) * inv.x;
    const float t0y = ((ny ? bmax.y : bmin.y) - o.y) * inv.y, t1y = ((ny ? bmin.y : bmax.y) - o.y) * inv.y;
    const float t0z = ((nz ? bmax.z : bmin.z) - o.z) * inv.z, t1z = ((nz ? bmin.z : bmax.z) - o.z) * inv.z;
    const float lo = fmaxf(fmaxf(t0x, t0z), fmaxf(t0y, tmin));
    const float hi = fminf(fminf(t1x, t1z), fminf(t1y, tmax));
    return hi > lo;
}

// Closest hit among the primitives of one object_list leaf, in the reference's order, with the full hit record: an
// object_list of spheres / rects, or of boxes (= object_list with a box of six rects, box.h:12-25): own box test where
// the list has one, children in order with shrinking closest (scene_object.h:83-97).  Anything else inside a tree leaf
// makes the scene fall back to the per-lane traversal (coop_trees_supported()).
MRT_FN bool coop_list_leaf_hit(const uint32_t feat, const SceneView &sc, uint32_t idx, const Ray &r, float tmin, float tmax, Hit &rec) {
    bool found = false;
    // LIST header: the copy referenced from a tree node has hasBox = 0 (its box was tested at the parent)
    MrtF4 l0 = ld4(sc.list, 2 * idx), l1 = ld4(sc.list, 2 * idx + 1);
    if (f2u(l1.w) >> 31) { if (!aabb_hit(l0, l1, r, tmin, tmax)) return false; }
    uint32_t ci = f2u(l0.w);
    uint32_t outer_ci = 0;      // position to come back to after a nested list (0 = not inside one)
    for (;;) {
        const uint32_t c = ldu(sc.child, ci);
        const uint32_t ctype = MRT_REF_TYPE(c);
        if (ctype == MRT_T_END) {
            if (!outer_ci) break;
            ci = outer_ci; outer_ci = 0;
            continue;
        }
        ci++;
        if (MRT_HAS(feat, MRT_FEAT_SPHERES) && ctype == MRT_T_SPHERE) {
            if (hit_sphere(feat, sc, MRT_REF_INDEX(c), r, tmin, tmax, true, rec)) { found = true; tmax = rec.t; }
        } else if (ctype <= MRT_T_RECT_YZ) {
            if (hit_rect(feat, sc, ctype - MRT_T_RECT_XY, MRT_REF_INDEX(c), r, tmin, tmax, true, rec)) { found = true; tmax = rec.t; }
        } else {   // nested object_list of primitives (a box): its own box test with the current closest
            const uint32_t li = MRT_REF_INDEX(c);
            MrtF4 n0 = ld4(sc.list, 2 * li), n1 = ld4(sc.list, 2 * li + 1);
            if ((f2u(n1.w) >> 31) && !aabb_hit(n0, n1, r, tmin, tmax)) continue;
            outer_ci = ci;
            ci = f2u(n0.w);
        }
    }
    return found;
}

MRT_HD void coop_prefetch(const void *p) {
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void) p;
#endif
}
MRT_HD uint32_t coop_kind_of_ref(uint32_t ref) {
    const uint32_t t = MRT_REF_TYPE(ref);
    return t == MRT_T_NODE2 ? MRT_COOP_KIND_NODE : (t == MRT_T_TRILEAF ? MRT_COOP_KIND_TRILEAF : MRT_COOP_KIND_LIST);
}

#if defined(__CUDACC__) || defined(MRT_EMUL_WARP)
// Traverses the trees of all lanes with has_job together.  root = the tree's root child (bvh[2i].w).  On return, for
// job lanes: true + rec (full record) if the tree reports a hit.  Must be called by all 32 lanes.
__device__ __forceinline__ bool coop_traverse(const uint32_t feat, const SceneView &sc, CoopArea &ca, bool has_job, uint32_t root, const Ray &ray,
                                              float tmin, float tmax, Hit &rec, CoopStats &stats, const uint32_t leaf_batch) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t gt_mask = ~(lt_mask | (1u << lane));
    // publish the rays
    if (has_job) {
        uint32_t *q = ca.ray + lane;
        q[0 * 32] = f2u(ray.o.x); q[1 * 32] = f2u(ray.o.y); q[2 * 32] = f2u(ray.o.z);
        q[3 * 32] = f2u(ray.d.x); q[4 * 32] = f2u(ray.d.y); q[5 * 32] = f2u(ray.d.z);
        q[6 * 32] = f2u(ray.inv.x); q[7 * 32] = f2u(ray.inv.y); q[8 * 32] = f2u(ray.inv.z);
        q[9 * 32] = f2u(tmin); q[10 * 32] = f2u(tmax); q[11 * 32] = f2u(ray.time);
        const bool fin = is_finite(ray.inv.x) && is_finite(ray.inv.y) && is_finite(ray.inv.z);
        q[12 * 32] = ray.mask | (fin ? 0x100u : 0u) | ((uint32_t) ray.inside << 16);
        ca.best[lane] = kRankNone;
    }
    // root items
    uint32_t n_nodes, n_leaves;
    {
        const uint32_t kind = coop_kind_of_ref(root);
        const bool root_is_node = has_job && kind == MRT_COOP_KIND_NODE;
        const bool root_is_leaf = has_job && kind != MRT_COOP_KIND_NODE;
        const uint32_t mn = __ballot_sync(0xFFFFFFFFu, root_is_node), ml = __ballot_sync(0xFFFFFFFFu, root_is_leaf);
        const uint2 it = make_uint2(MRT_REF_INDEX(root) | (kind << 24) | (lane << 26), 0x80000000u);
        if (root_is_node) ca.nodes[__popc(mn & lt_mask)] = it;
        if (root_is_leaf) ca.leaves[__popc(ml & lt_mask)] = it;
        n_nodes = __popc(mn); n_leaves = __popc(ml);
    }
    __syncwarp();

    while (n_nodes | n_leaves) {
        if (n_leaves >= leaf_batch || n_nodes == 0u) {
            // ------------------------------------------------------------ leaf step
            const uint32_t k = min(32u, n_leaves);
            n_leaves -= k;
            bool hit = false, valid = false;
            uint32_t j = 0, rank = 0, kind = 0, idx = 0;
            Hit lrec;
            if (lane < k) {
                const uint2 it = ca.leaves[n_leaves + lane];
                j = it.x >> 26; rank = it.y; kind = (it.x >> 24) & 3u; idx = it.x & 0xFFFFFFu;
                valid = rank < ca.best[j];   // else: a leaf that comes earlier in this ray's order has already reported a hit
            }
            stats.leaf_steps++; stats.leaf_items += k;
            if (MRT_HAS(feat, MRT_FEAT_TRIS)) {
                // triangle leaf: all triangles in order, the closest shrinks after each hit (triangle.h:179-187).  (Dealing the
                // (ray, triangle) pairs of a step out over the lanes instead -- leaves hold 1..17 triangles -- was measured
                // slower: prefix scan, owner table, 64-bit shared atomics and the second evaluation of the winner cost more than
                // the idle lanes of this loop; profiles/r2_notes.md)
                if (valid && (kind == MRT_COOP_KIND_TRILEAF || !MRT_HAS(feat, MRT_FEAT_LEAF_LISTS))) {
                    const uint32_t *q = ca.ray + j;
                    Ray r;
                    r.o = v3(u2f(q[0 * 32]), u2f(q[1 * 32]), u2f(q[2 * 32]));
                    r.d = v3(u2f(q[3 * 32]), u2f(q[4 * 32]), u2f(q[5 * 32]));
                    r.inside = (int) (q[12 * 32] >> 16);
                    const float tmn = u2f(q[9 * 32]);
                    float tmx = u2f(q[10 * 32]);
                    const uint32_t first = ldu(sc.trileaf, 2 * idx), cnt = ldu(sc.trileaf, 2 * idx + 1);
                    for (uint32_t i = 0; i < cnt; i++) if (hit_triangle(sc, first + i, r, tmn, tmx, true, lrec)) { hit = true; tmx = lrec.t; }
                }
            }
            if (MRT_HAS(feat, MRT_FEAT_LEAF_LISTS)) {
                if (valid && (kind == MRT_COOP_KIND_LIST || !MRT_HAS(feat, MRT_FEAT_TRIS))) {
                    const uint32_t *q = ca.ray + j;
                    Ray r;
                    r.o = v3(u2f(q[0 * 32]), u2f(q[1 * 32]), u2f(q[2 * 32]));
                    r.d = v3(u2f(q[3 * 32]), u2f(q[4 * 32]), u2f(q[5 * 32]));
                    r.inv = v3(u2f(q[6 * 32]), u2f(q[7 * 32]), u2f(q[8 * 32]));
                    r.time = u2f(q[11 * 32]);
                    const uint32_t fl = q[12 * 32];
                    r.mask = fl & 0xFFu; r.inside = (int) (fl >> 16);
                    hit = coop_list_leaf_hit(feat, sc, idx, r, u2f(q[9 * 32]), u2f(q[10 * 32]), lrec);
                }
            }
            if (hit) atomicMin(&ca.best[j], rank);
            __syncwarp();
            if (hit && ca.best[j] == rank) {   // ranks of different leaves differ: one winner per ray
                uint32_t *q = ca.rec + j;
                q[0 * 32] = f2u(lrec.t);
                q[1 * 32] = f2u(lrec.p.x); q[2 * 32] = f2u(lrec.p.y); q[3 * 32] = f2u(lrec.p.z);
                q[4 * 32] = f2u(lrec.n.x); q[5 * 32] = f2u(lrec.n.y); q[6 * 32] = f2u(lrec.n.z);
                q[7 * 32] = f2u(lrec.u); q[8 * 32] = f2u(lrec.v); q[9 * 32] = lrec.mat;
            }
            __syncwarp();
            continue;
        }
        // ---------------------------------------------------------------- node step
        const uint32_t k = (n_nodes <= kCoopThresh) ? min(32u, n_nodes) : 1u;
        n_nodes -= k;
        // children to push: A = the one the ray visits second, B = first (B ends up on top); each to the node stack or the leaf queue
        uint32_t a_x = 0, b_x = 0, rank = 0;
        bool a_ok = false, b_ok = false;
        if (lane < k) {
            const uint2 it = ca.nodes[n_nodes + (k - 1u - lane)];   // lane 0 takes the top of the stack
            const uint32_t jj = it.x & 0xFC000000u, j = it.x >> 26;
            rank = it.y;
            if (rank < ca.best[j]) {
                const uint32_t ni = it.x & 0xFFFFFFu;
                const uint32_t *q = ca.ray + j;
                const V3 o = v3(u2f(q[0 * 32]), u2f(q[1 * 32]), u2f(q[2 * 32]));
                const V3 inv = v3(u2f(q[6 * 32]), u2f(q[7 * 32]), u2f(q[8 * 32]));
                const float tmn = u2f(q[9 * 32]), tmx = u2f(q[10 * 32]);
                const uint32_t fl = q[12 * 32];
                MrtF4 n0 = ld4(sc.node2, 4 * ni), n1 = ld4(sc.node2, 4 * ni + 1);
                MrtF4 n2 = ld4(sc.node2, 4 * ni + 2), n3 = ld4(sc.node2, 4 * ni + 3);
                const uint32_t w0 = f2u(n0.w), w1 = f2u(n1.w), flags = f2u(n2.w);
                bool hl, hr;
                if (fl & 0x100u) {
                    const bool nx = inv.x < 0.0f, ny = inv.y < 0.0f, nz = inv.z < 0.0f;
                    hl = aabb_hit_finite(n0, n1, o, inv, nx, ny, nz, tmn, tmx);
                    hr = aabb_hit_finite(n2, n3, o, inv, nx, ny, nz, tmn, tmx);
                } else {
                    Ray r;
                    r.o = o; r.inv = inv;
                    hl = aabb_hit(n0, n1, r, tmn, tmx);
                    hr = aabb_hit(n2, n3, r, tmn, tmx);
                }
                hl = hl || !(flags & 1u);
                hr = hr || !(flags & 2u);
                const uint32_t order = (w0 >> 28) | ((w1 >> 28) << 4);
                const bool lfirst = (order & fl & 0xFFu) != 0;   // scene_object.h:224-231
                const uint32_t left_x = (w0 & 0xFFFFFFu) | (((flags >> 2) & 3u) << 24) | jj;
                const uint32_t right_x = (w1 & 0xFFFFFFu) | (((flags >> 4) & 3u) << 24) | jj;
                b_x = lfirst ? left_x : right_x; b_ok = lfirst ? hl : hr;
                a_x = lfirst ? right_x : left_x; a_ok = lfirst ? hr : hl;
            }
        }
        stats.node_steps++; stats.node_items += k;
        // ranks of the children: the path bit (0 = visited first) replaces the marker, the marker moves one down
        const uint32_t m = rank & (0u - rank);
        const uint32_t a_rank = rank | (m >> 1), b_rank = (rank ^ m) | (m >> 1);
        const bool a_node = a_ok && !(a_x & 0x03000000u), b_node = b_ok && !(b_x & 0x03000000u);
        {
            const uint32_t ma = __ballot_sync(0xFFFFFFFFu, a_node), mb = __ballot_sync(0xFFFFFFFFu, b_node);
            // lane 31's children lowest, lane 0's on top, a lane's own first-visited child above its second
            const uint32_t base = n_nodes + __popc(ma & gt_mask) + __popc(mb & gt_mask);
            if (a_node) ca.nodes[base] = make_uint2(a_x, a_rank);
            if (b_node) { ca.nodes[base + (a_node ? 1u : 0u)] = make_uint2(b_x, b_rank); coop_prefetch(sc.node2 + 4u * (b_x & 0xFFFFFFu)); }   // most likely popped by the next step
            n_nodes += __popc(ma) + __popc(mb);
        }
        {
            const bool a_leaf = a_ok && !a_node, b_leaf = b_ok && !b_node;
            const uint32_t ma = __ballot_sync(0xFFFFFFFFu, a_leaf), mb = __ballot_sync(0xFFFFFFFFu, b_leaf);
            if (ma | mb) {
                const uint32_t base = n_leaves + __popc(ma & gt_mask) + __popc(mb & gt_mask);
                if (a_leaf) ca.leaves[base] = make_uint2(a_x, a_rank);
                if (b_leaf) ca.leaves[base + (a_leaf ? 1u : 0u)] = make_uint2(b_x, b_rank);
                n_leaves += __popc(ma) + __popc(mb);
            }
        }
        __syncwarp();
    }
    // the owner lane takes the record of the winning leaf
    bool ret = false;
    if (has_job && ca.best[lane] != kRankNone) {
        const uint32_t *q = ca.rec + lane;
        rec.t = u2f(q[0 * 32]);
        rec.p = v3(u2f(q[1 * 32]), u2f(q[2 * 32]), u2f(q[3 * 32]));
        rec.n = v3(u2f(q[4 * 32]), u2f(q[5 * 32]), u2f(q[6 * 32]));
        rec.u = u2f(q[7 * 32]); rec.v = u2f(q[8 * 32]); rec.mat = q[9 * 32];
        ret = true;
    }
    __syncwarp();   // the area is reused by the next round
    return ret;
}
#endif

}  // namespace mrt
