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
// (aabb_hit) and pushes the children that are hit, the farther one first.  A LEAF STEP runs when 32 leaves are
// queued (or nothing else is left): every lane evaluates one leaf in the reference's order (closest hit among its
// primitives: triangle.h:179-187, scene_object.h:88-95) and posts its rank with an atomic minimum; the winner records
// its primitive.  At the end the owner lane re-evaluates the winning primitive to get the full hit record.  All hit
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
// shared memory per warp, in 32-bit words (8-byte aligned)
constexpr uint32_t kCoopWords = kCoopRayWords * 32u + 32u + 32u + 2u * kCoopNodeCap + 2u * kCoopLeafCap;
constexpr uint32_t kRankNone = 0xFFFFFFFFu;

struct CoopArea {   // views into one warp's shared-memory area
    uint32_t *ray;    // [kCoopRayWords][32]: word k of job lane j at [k * 32 + j]
    uint32_t *best;   // [32] rank of the first (depth-first) leaf that reported a hit, kRankNone = none yet
    uint32_t *prim;   // [32] typed ref of that leaf's closest primitive (sphere / rect / triangle index)
    uint2 *nodes;     // node stack: x = node2 index | job lane << 24, y = rank
    uint2 *leaves;    // leaf queue: x = index | is_trileaf << 24 | job lane << 25, y = rank
    MRT_HD void bind(uint32_t *base) {
        ray = base;
        best = base + kCoopRayWords * 32u;
        prim = best + 32u;
        nodes = reinterpret_cast<uint2 *>(prim + 32u);
        leaves = nodes + kCoopNodeCap;
    }
};

struct CoopStats { unsigned long long node_steps, node_items, leaf_steps, leaf_items; };

// MRT_T_TRI: pseudo ref type of a winning triangle (index into tri[]); only used inside this file
#define MRT_T_TRI_WIN 14u

// Closest hit among the primitives of one tree leaf, in the reference's order; returns the winning primitive.
//   TRILEAF: all triangles in order, tmax shrinks after each hit (triangle.h:179-187)
//   LIST   : object_list of spheres / rects, or of boxes (= object_list with a box of six rects, box.h:12-25):
//            own box test where the list has one, children in order with shrinking closest (scene_object.h:83-97)
// Anything else inside a tree leaf makes the scene fall back to the per-lane traversal (coop_supported()).
MRT_FN bool coop_leaf_hit(const uint32_t feat, const SceneView &sc, uint32_t leaf_is_tri, uint32_t idx, const Ray &r, float tmin, float tmax, uint32_t *win) {
    bool found = false;
    Hit rec;
    if (leaf_is_tri) {
        const uint32_t first = ldu(sc.trileaf, 2 * idx), count = ldu(sc.trileaf, 2 * idx + 1);
        for (uint32_t i = 0; i < count; i++) {
            if (hit_triangle(sc, first + i, r, tmin, tmax, false, rec)) { found = true; tmax = rec.t; *win = MRT_REF(MRT_T_TRI_WIN, first + i); }
        }
        return found;
    }
    // LIST header: the copy referenced from a tree node has hasBox = 0 (its box was tested at the parent)
    {
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
            if (ctype == MRT_T_SPHERE) {
                if (hit_sphere(feat, sc, MRT_REF_INDEX(c), r, tmin, tmax, false, rec)) { found = true; tmax = rec.t; *win = c; }
            } else if (ctype <= MRT_T_RECT_YZ) {
                if (hit_rect(feat, sc, ctype - MRT_T_RECT_XY, MRT_REF_INDEX(c), r, tmin, tmax, false, rec)) { found = true; tmax = rec.t; *win = c; }
            } else {   // nested object_list of primitives (a box): its own box test with the current closest
                const uint32_t li = MRT_REF_INDEX(c);
                MrtF4 n0 = ld4(sc.list, 2 * li), n1 = ld4(sc.list, 2 * li + 1);
                if ((f2u(n1.w) >> 31) && !aabb_hit(n0, n1, r, tmin, tmax)) continue;
                outer_ci = ci;
                ci = f2u(n0.w);
            }
        }
    }
    return found;
}

// Full hit record of the winning primitive: the same test once more with the tree's (tmin, tmax) -- the outcome of a
// primitive test does not depend on tmax except through the final range check, which the winner has passed.
MRT_FN void coop_finish_hit(const uint32_t feat, const SceneView &sc, uint32_t win, const Ray &r, float tmin, float tmax, Hit &rec) {
    const uint32_t type = MRT_REF_TYPE(win), idx = MRT_REF_INDEX(win);
    if (type == MRT_T_TRI_WIN) hit_triangle(sc, idx, r, tmin, tmax, true, rec);
    else if (type == MRT_T_SPHERE) hit_sphere(feat, sc, idx, r, tmin, tmax, true, rec);
    else hit_rect(feat, sc, type - MRT_T_RECT_XY, idx, r, tmin, tmax, true, rec);
}

#if defined(__CUDACC__) || defined(MRT_EMUL_WARP)
// Traverses the trees of all lanes with has_job together.  root = the tree's root child (bvh[2i].w).  On return, for
// job lanes: true + rec (full record) if the tree reports a hit.  Must be called by all 32 lanes.
__device__ __forceinline__ bool coop_traverse(const uint32_t feat, const SceneView &sc, CoopArea &ca, bool has_job, uint32_t root, const Ray &ray,
                                              float tmin, float tmax, Hit &rec, CoopStats &stats) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // publish the rays
    if (has_job) {
        uint32_t *q = ca.ray + lane;
        q[0 * 32] = f2u(ray.o.x); q[1 * 32] = f2u(ray.o.y); q[2 * 32] = f2u(ray.o.z);
        q[3 * 32] = f2u(ray.d.x); q[4 * 32] = f2u(ray.d.y); q[5 * 32] = f2u(ray.d.z);
        q[6 * 32] = f2u(ray.inv.x); q[7 * 32] = f2u(ray.inv.y); q[8 * 32] = f2u(ray.inv.z);
        q[9 * 32] = f2u(tmin); q[10 * 32] = f2u(tmax); q[11 * 32] = f2u(ray.time);
        q[12 * 32] = ray.mask | ((uint32_t) ray.inside << 16);
        ca.best[lane] = kRankNone;
    }
    // root items
    uint32_t n_nodes, n_leaves;
    {
        const bool root_is_node = has_job && MRT_REF_TYPE(root) == MRT_T_NODE2;
        const bool root_is_leaf = has_job && !root_is_node;
        const uint32_t mn = __ballot_sync(0xFFFFFFFFu, root_is_node), ml = __ballot_sync(0xFFFFFFFFu, root_is_leaf);
        if (root_is_node) ca.nodes[__popc(mn & lt_mask)] = make_uint2(MRT_REF_INDEX(root) | (lane << 24), 0x80000000u);
        if (root_is_leaf) ca.leaves[__popc(ml & lt_mask)] = make_uint2(MRT_REF_INDEX(root) | ((MRT_REF_TYPE(root) == MRT_T_TRILEAF ? 1u : 0u) << 24) | (lane << 25), 0x80000000u);
        n_nodes = __popc(mn); n_leaves = __popc(ml);
    }
    __syncwarp();

    while (n_nodes | n_leaves) {
        if (n_leaves >= 32u || n_nodes == 0u) {
            // ------------------------------------------------------------ leaf step
            const uint32_t k = min(32u, n_leaves);
            n_leaves -= k;
            bool hit = false;
            uint32_t j = 0, rank = 0, win = 0;
            if (lane < k) {
                const uint2 it = ca.leaves[n_leaves + lane];
                j = it.x >> 25; rank = it.y;
                if (rank < ca.best[j]) {   // else: a leaf that comes earlier in this ray's order has already reported a hit
                    const uint32_t *q = ca.ray + j;
                    Ray r;
                    r.o = v3(u2f(q[0 * 32]), u2f(q[1 * 32]), u2f(q[2 * 32]));
                    r.d = v3(u2f(q[3 * 32]), u2f(q[4 * 32]), u2f(q[5 * 32]));
                    r.inv = v3(u2f(q[6 * 32]), u2f(q[7 * 32]), u2f(q[8 * 32]));
                    r.time = u2f(q[11 * 32]);
                    const uint32_t fl = q[12 * 32];
                    r.mask = fl & 0xFFFFu; r.inside = (int) (fl >> 16);
                    hit = coop_leaf_hit(feat, sc, (it.x >> 24) & 1u, it.x & 0xFFFFFFu, r, u2f(q[9 * 32]), u2f(q[10 * 32]), &win);
                    if (hit) atomicMin(&ca.best[j], rank);
                }
            }
            stats.leaf_steps++; stats.leaf_items += k;
            __syncwarp();
            if (hit && ca.best[j] == rank) ca.prim[j] = win;   // ranks of different leaves differ: one winner per ray
            __syncwarp();
            continue;
        }
        // ---------------------------------------------------------------- node step
        const uint32_t k = (n_nodes <= kCoopThresh) ? min(32u, n_nodes) : 1u;
        n_nodes -= k;
        // children to push: A = the one the ray visits second, B = first (B ends up on top); each to the node stack or the leaf queue
        uint32_t a_x = 0, b_x = 0, a_rank = 0, b_rank = 0;
        bool a_node = false, a_leaf = false, b_node = false, b_leaf = false;
        if (lane < k) {
            const uint2 it = ca.nodes[n_nodes + (k - 1u - lane)];   // lane 0 takes the top of the stack
            const uint32_t j = it.x >> 24, rank = it.y;
            if (rank < ca.best[j]) {
                const uint32_t ni = it.x & 0xFFFFFFu;
                const uint32_t *q = ca.ray + j;
                Ray r;
                r.o = v3(u2f(q[0 * 32]), u2f(q[1 * 32]), u2f(q[2 * 32]));
                r.inv = v3(u2f(q[6 * 32]), u2f(q[7 * 32]), u2f(q[8 * 32]));
                const float tmn = u2f(q[9 * 32]), tmx = u2f(q[10 * 32]);
                const uint32_t mask = q[12 * 32] & 0xFFFFu;
                MrtF4 n0 = ld4(sc.node2, 4 * ni), n1 = ld4(sc.node2, 4 * ni + 1);
                MrtF4 n2 = ld4(sc.node2, 4 * ni + 2), n3 = ld4(sc.node2, 4 * ni + 3);
                const uint32_t w0 = f2u(n0.w), w1 = f2u(n1.w), flags = f2u(n2.w);
                const uint32_t order = (w0 >> 28) | ((w1 >> 28) << 4);
                const uint32_t left = w0 & 0x0FFFFFFFu, right = w1 & 0x0FFFFFFFu;
                const bool hl = !(flags & 1u) || aabb_hit(n0, n1, r, tmn, tmx);
                const bool hr = !(flags & 2u) || aabb_hit(n2, n3, r, tmn, tmx);
                const bool lfirst = (order & mask) != 0;   // scene_object.h:224-231
                const bool h_first = lfirst ? hl : hr, h_second = lfirst ? hr : hl;
                const uint32_t first = lfirst ? left : right, second = lfirst ? right : left;
                const uint32_t m = rank & (0u - rank);      // marker bit; children: path bit 0 / 1 at its place, marker one lower
                if (h_first) {
                    b_rank = (rank & ~m) | (m >> 1);
                    if (MRT_REF_TYPE(first) == MRT_T_NODE2) { b_node = true; b_x = MRT_REF_INDEX(first) | (j << 24); }
                    else { b_leaf = true; b_x = MRT_REF_INDEX(first) | ((MRT_REF_TYPE(first) == MRT_T_TRILEAF ? 1u : 0u) << 24) | (j << 25); }
                }
                if (h_second) {
                    a_rank = rank | (m >> 1);
                    if (MRT_REF_TYPE(second) == MRT_T_NODE2) { a_node = true; a_x = MRT_REF_INDEX(second) | (j << 24); }
                    else { a_leaf = true; a_x = MRT_REF_INDEX(second) | ((MRT_REF_TYPE(second) == MRT_T_TRILEAF ? 1u : 0u) << 24) | (j << 25); }
                }
            }
        }
        stats.node_steps++; stats.node_items += k;
        {
            const uint32_t ma = __ballot_sync(0xFFFFFFFFu, a_node), mb = __ballot_sync(0xFFFFFFFFu, b_node);
            // lane 31's children lowest, lane 0's on top, a lane's own first-visited child above its second
            const uint32_t gt_mask = ~(lt_mask | (1u << lane));
            const uint32_t base = n_nodes + __popc(ma & gt_mask) + __popc(mb & gt_mask);
            if (a_node) ca.nodes[base] = make_uint2(a_x, a_rank);
            if (b_node) ca.nodes[base + (a_node ? 1u : 0u)] = make_uint2(b_x, b_rank);
            n_nodes += __popc(ma) + __popc(mb);
        }
        {
            const uint32_t ma = __ballot_sync(0xFFFFFFFFu, a_leaf), mb = __ballot_sync(0xFFFFFFFFu, b_leaf);
            if (ma | mb) {
                const uint32_t gt_mask = ~(lt_mask | (1u << lane));
                const uint32_t base = n_leaves + __popc(ma & gt_mask) + __popc(mb & gt_mask);
                if (a_leaf) ca.leaves[base] = make_uint2(a_x, a_rank);
                if (b_leaf) ca.leaves[base + (a_leaf ? 1u : 0u)] = make_uint2(b_x, b_rank);
                n_leaves += __popc(ma) + __popc(mb);
            }
        }
        __syncwarp();
    }
    // the owner lane completes the hit record of its winning primitive
    bool ret = false;
    if (has_job && ca.best[lane] != kRankNone) {
        coop_finish_hit(feat, sc, ca.prim[lane], ray, tmin, tmax, rec);
        ret = true;
    }
    __syncwarp();   // the area is reused by the next round
    return ret;
}
#endif

}  // namespace mrt
