// Render megakernels, templated on the scene-feature mask (MRT_FEAT_*, see trace_core.h) and on the
// launch-bounds variant.  Each feature mask is instantiated in its own translation unit
// (render_variant_*.cu) so that the variants compile in parallel.
//
// Work decomposition (replaces work_queue.cpp + the draw() loop nest, main.cpp:138-188):
//   * persistent warps: the grid is sized to the SM count x resident blocks and every warp pulls work from
//     one global atomic ticket counter (one atomic per warp task, issued by lane 0, broadcast by shuffle);
//   * a warp task is a chunk of (pixel, sample) items that the 32 lanes drain cooperatively.  Mode B (default)
//     regroups the live paths between segments through a per-warp pool and bins; in modes W / P a lane keeps
//     its path and is regenerated from the item stream at the next warp-converged point (ballot + popc
//     prefix), so lanes do not idle while the longest path of a batch finishes;
//   * mode B hands the frame out in chunks of decreasing size (guided self-scheduling: big chunks first, the last stretch in
//     chunks of 1/4 and 1/16 of the size), so a launch does not end with warps waiting for a big chunk;
//   * per-thread traversal stacks live in shared memory, interleaved by lane (word k of lane l at
//     [k*32 + l]) so pushes/pops are bank-conflict free;
//   * BVH trees are traversed per lane or -- mode B, big trees -- warp-cooperatively (coop_tree.cuh).
// Every accumulator element has exactly one writer: no float atomics, results are reproducible run to run.
#pragma once
#include "gpu_internal.h"
#include "coop_tree.cuh"

namespace mrt {

constexpr uint32_t kPoolCap = 128;     // path slots per warp (slot ids are bytes)
constexpr uint32_t kStateWords = 21;   // words per parked path, see park_path()
constexpr uint32_t kMaxClsBoxes = 3;
constexpr uint32_t kMaxBins = 16;
constexpr uint32_t kMaxStageItems = 8192;   // item index has 14 bits in the parked state

struct RenderArgs {
    SceneView sc;
    uint32_t width, height, sqrt_n, s_begin, s_end, max_bounces;   // width x height = the FULL frame (sub-pixel positions, stream ids)
    uint32_t crop_x0, crop_y0, crop_w, n_pixels;                   // rendered window: accumulator pixel i = (crop_x0 + i % crop_w, crop_y0 + i / crop_w)
    uint64_t seed;
    uint32_t accumulate;
    uint32_t stack_words;
    uint32_t n_tasks, pixels_per_task;
    // mode B, guided self-scheduling: the ticket queue hands out the frame in three runs of tasks with decreasing chunk size
    // (pixels_per_task = the largest): run r starts at task sched_task0[r] / pixel sched_pix0[r] and uses chunks of sched_k[r]
    // pixels; sched_pix0[3] = n_pixels.  Small last chunks keep the warps from idling while the last big chunk finishes.
    uint32_t sched_task0[3], sched_pix0[4], sched_k[3];
    float4 *acc;
    unsigned int *ticket;             // global task counter
    unsigned long long *counters;     // [0] rays [1] warp iterations [2] nonfinite [4..7] cooperative traversal: node steps, node items, leaf steps, leaf items
    const uint32_t *order;            // work order: item i of the queue is pixel order[i] (Morton tiles), or NULL = row-major
    const volatile int *cancel;       // device flag, written by mrt_gpu_cancel through a side stream
    // binned mode (render_pixel_binned): per-warp path pool in global memory (L2 resident) and the ray classifier
    uint32_t *pool;                   // [warp][kStateWords][kPoolCap]
    float4 *stage;                    // [warp][stage_items]: finished samples of the warp's current chunk, by item index
    uint32_t stage_items;             // pixels_per_task * samples of this launch (<= kMaxStageItems)
    uint32_t n_bins;                  // 1 << (n_cls_boxes + cls_pending)
    uint32_t n_cls_boxes, cls_pending;
    uint32_t coop_leaf_batch;         // cooperative tree traversal: a leaf step runs once this many leaves are queued (1..32)
    float cls_box[kMaxClsBoxes][6];   // world-space boxes (min xyz, max xyz) of the root list's composite children
};

constexpr int kWarpsPerBlock = kBlock / 32;
constexpr uint32_t kStageBlock = 128;   // mode B: launches with up to this many samples per pixel add them up sequentially (SEQ, see render_pixel_binned)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// One segment of a live path: traverse, shade.  Returns true while the path continues.
template <uint32_t FEAT>
__device__ __forceinline__ bool path_step(const RenderArgs &a, Path &p, Rng &rng, Stack &st) {
    Hit rec;
    path_advance(FEAT, a.sc, p);   // normalise the pending direction, apply the deferred diffuse weight
    bool hit = intersect(FEAT, a.sc, p.ray, 0.001f, FLT_MAX, rec, rng, st, nullptr);
    return path_shade(FEAT, a.sc, p, hit, rec, a.max_bounces, rng);
}

// The same with the warp-cooperative tree traversal (coop_tree.cuh); called by ALL lanes of the warp, active or not.
// The per-lane machine runs up to the root of a BVH tree and suspends; the warp then traverses the trees of all suspended
// lanes together; the lanes resume with the tree's outcome (a scene may hold several trees: one round per tree).
template <uint32_t FEAT>
__device__ __forceinline__ bool path_step_coop(const RenderArgs &a, Path &p, Rng &rng, Stack &st, bool active, CoopArea &ca, CoopStats &cs) {
    Hit rec;
    Isect s;
    s.ret = false; s.cur = 0; s.tmin = 0; s.tmax = 0;
    if (active) {
        path_advance(FEAT, a.sc, p);
        isect_begin(s, a.sc, 0.001f, FLT_MAX, st);
    }
    bool done = !active, resume = false;
    for (;;) {   // one call site of the per-lane machine (instruction-cache footprint)
        if (!done) done = isect_run<true>(FEAT, a.sc, p.ray, s, rec, rng, st, resume, nullptr);
        if (!__any_sync(0xFFFFFFFFu, !done)) break;
        const bool job = !done;
        const bool h = coop_traverse(FEAT, a.sc, ca, job, s.cur, p.ray, s.tmin, s.tmax, rec, cs, a.coop_leaf_batch);
        if (job) {
            s.ret = h;
            if (h) s.tmax = rec.t;
            resume = true;
        }
    }
    if (!active) return false;
    return path_shade(FEAT, a.sc, p, s.ret, rec, a.max_bounces, rng);
}

// ------------------------------------------------------------------ mode W
// Warp task = a chunk of `pixels_per_task` consecutive pixels x all samples of this launch, handed out as
// one stream of items (pixel-major, sample-minor).  A lane keeps the running sum of the pixel it is
// working on; when its next item belongs to another pixel it parks that partial sum in its own column of
// a per-warp shared array part[k][lane] (one writer per element: a lane visits a pixel in one contiguous
// period because items are handed out in increasing order).  At the end of the chunk every pixel's 32
// partials are combined by a fixed-order shuffle tree.  The idle tail (lanes waiting for the last paths)
// is paid once per chunk instead of once per pixel.
template <uint32_t FEAT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) render_pixel_per_warp(const RenderArgs a) {
    extern __shared__ uint32_t smem_stack[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    Stack st;
    st.base = smem_stack + (size_t) warp * a.stack_words * 32u + lane;
    st.stride = 32u;
    st.sp = 0;
    const uint32_t K = a.pixels_per_task;
    float4 *part = reinterpret_cast<float4 *>(smem_stack + (size_t) kWarpsPerBlock * a.stack_words * 32u) + (size_t) warp * K * 32u + lane;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t n_pixels = a.n_pixels;
    const uint32_t ns = a.s_end - a.s_begin;
    unsigned long long rays = 0, nonfinite = 0, iters = 0;

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = (*a.cancel) ? 0xFFFFFFFFu : atomicAdd(a.ticket, 1u);
        task = __shfl_sync(0xFFFFFFFFu, task, 0);
        if (task >= a.n_tasks) break;
        const uint32_t pix0 = task * K;
        const uint32_t kp = min(K, n_pixels - pix0);   // pixels in this chunk
        const uint32_t n_items = kp * ns;
#pragma unroll 1
        for (uint32_t k = 0; k < kp; k++) part[k * 32u] = make_float4(0.f, 0.f, 0.f, 0.f);

        uint32_t next_i = 0;           // warp-uniform stream cursor
        uint32_t cur_k = 0xFFFFFFFFu;  // pixel slot of this lane's running sum
        bool alive = false;
        Path p;
        Rng rng;
        float sr = 0, sg = 0, sb = 0, sc = 0;
        for (;;) {
            // regenerate terminated lanes from the stream (warp-converged point)
            const uint32_t need = __ballot_sync(0xFFFFFFFFu, !alive);
            if (!alive) {
                const uint32_t i = next_i + __popc(need & lt_mask);
                if (i < n_items) {
                    const uint32_t k = i / ns, s = a.s_begin + (i - k * ns);
                    if (k != cur_k) {
                        if (cur_k != 0xFFFFFFFFu) part[cur_k * 32u] = make_float4(sr, sg, sb, sc);
                        sr = sg = sb = sc = 0;
                        cur_k = k;
                    }
                    const uint32_t pix = a.order ? __ldg(a.order + pix0 + k) : pix0 + k;
                    const uint32_t cy = pix / a.crop_w, x = a.crop_x0 + (pix - cy * a.crop_w), y = a.crop_y0 + cy;
                    path_begin(a.sc, p, rng, x, y, s, a.sqrt_n, a.width, a.height, a.seed);
                    alive = true;
                }
            }
            next_i = min(next_i + (uint32_t) __popc(need), n_items);
            if (!__any_sync(0xFFFFFFFFu, alive)) break;
            iters++;
            if (alive) {
                rays++;
                if (!path_step<FEAT>(a, p, rng, st)) {
                    if (path_sample_finite(p)) { sr += p.L.x; sg += p.L.y; sb += p.L.z; sc += 1.0f; }
                    else nonfinite++;
                    alive = false;
                }
            }
        }
        if (cur_k != 0xFFFFFFFFu) part[cur_k * 32u] = make_float4(sr, sg, sb, sc);
        __syncwarp();
#pragma unroll 1
        for (uint32_t k = 0; k < kp; k++) {
            float4 v = part[k * 32u];
            v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z); v.w = warp_sum(v.w);
            if (lane == 0) {
                const uint32_t pix = a.order ? __ldg(a.order + pix0 + k) : pix0 + k;
                if (a.accumulate) { float4 o = a.acc[pix]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                a.acc[pix] = v;
            }
        }
        __syncwarp();
    }
    // statistics: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) {
        rays += __shfl_xor_sync(0xFFFFFFFFu, rays, o);
        nonfinite += __shfl_xor_sync(0xFFFFFFFFu, nonfinite, o);
    }
    if (lane == 0) {
        atomicAdd(&a.counters[0], rays);
        atomicAdd(&a.counters[1], iters);   // warp iterations: rays / (32 * iters) = share of lanes with a live path
        atomicAdd(&a.counters[2], nonfinite);
    }
}

// ------------------------------------------------------------------ mode P
template <uint32_t FEAT, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) render_pixel_per_lane(const RenderArgs a) {
    extern __shared__ uint32_t smem_stack[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    Stack st;
    st.base = smem_stack + (size_t) warp * a.stack_words * 32u + lane;
    st.stride = 32u;
    st.sp = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t n_pixels = a.n_pixels;
    unsigned long long rays = 0, nonfinite = 0, iters = 0;

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = (*a.cancel) ? 0xFFFFFFFFu : atomicAdd(a.ticket, 1u);
        task = __shfl_sync(0xFFFFFFFFu, task, 0);
        if (task >= a.n_tasks) break;
        uint32_t next_p = task * a.pixels_per_task;
        const uint32_t end_p = min(next_p + a.pixels_per_task, n_pixels);

        bool has_pixel = false, alive = false;
        uint32_t pix = 0, x = 0, y = 0, s = 0;
        Path p;
        Rng rng;
        float sr = 0, sg = 0, sb = 0, sc = 0;
        for (;;) {
            const uint32_t need = __ballot_sync(0xFFFFFFFFu, !has_pixel);
            if (!has_pixel) {
                const uint32_t cand = next_p + __popc(need & lt_mask);
                if (cand < end_p) {
                    pix = a.order ? __ldg(a.order + cand) : cand;
                    y = pix / a.crop_w; x = a.crop_x0 + (pix - y * a.crop_w); y += a.crop_y0;
                    s = a.s_begin;
                    sr = sg = sb = sc = 0;
                    has_pixel = true;
                }
            }
            next_p = min(next_p + (uint32_t) __popc(need), end_p);
            if (!__any_sync(0xFFFFFFFFu, has_pixel)) break;
            iters++;
            if (has_pixel) {
                if (!alive) {
                    path_begin(a.sc, p, rng, x, y, s, a.sqrt_n, a.width, a.height, a.seed);
                    alive = true;
                }
                rays++;
                if (!path_step<FEAT>(a, p, rng, st)) {
                    if (path_sample_finite(p)) { sr += p.L.x; sg += p.L.y; sb += p.L.z; sc += 1.0f; }
                    else nonfinite++;
                    alive = false;
                    if (++s == a.s_end) {
                        float4 v = make_float4(sr, sg, sb, sc);
                        if (a.accumulate) { float4 o = a.acc[pix]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                        a.acc[pix] = v;
                        has_pixel = false;
                    }
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        rays += __shfl_xor_sync(0xFFFFFFFFu, rays, o);
        nonfinite += __shfl_xor_sync(0xFFFFFFFFu, nonfinite, o);
    }
    if (lane == 0) {
        atomicAdd(&a.counters[0], rays);
        atomicAdd(&a.counters[1], iters);   // warp iterations: rays / (32 * iters) = share of lanes with a live path
        atomicAdd(&a.counters[2], nonfinite);
    }
}

// ------------------------------------------------------------------ mode B (binned pool)
// Same warp task as mode W (a chunk of K pixels x all samples of the launch), but the paths of a chunk no
// longer stay with "their" lane.  Between two segments every live path is parked in a per-warp pool (global
// memory, L2 resident; 21 words per path) and its slot id is queued in one of n_bins per-warp bins chosen by
// a cheap classifier of the NEXT ray (which top-level composites its half-line crosses, whether a diffuse
// weight is pending).  Each iteration the warp
//   * pops 32 paths from the fullest bin if it holds 32            -> 32 lanes that will do similar things,
//   * else starts 32 new paths (consecutive samples of a pixel)    -> primary rays, fully coherent,
//   * else (chunk drained) pops what is left,
// runs ONE segment for them (path_step, single call site), writes finished samples to the chunk's staging array
// (global memory, one float4 per item) and parks the survivors again.  At the end of the chunk every pixel's
// samples are summed from the staging array in an order fixed by the launch's samples per pixel alone (SEQ: one after
// the other; else lane-strided + shuffle tree), so the result does not depend on the schedule at all;
// per-path arithmetic is untouched (same RNG stream, same operations).
// A live path carries no radiance (only lights and the sky emit, and both end the path), so L is not parked.
__device__ __forceinline__ void park_path(uint32_t *pool, uint32_t slot, const Path &p, const Rng &rng, uint32_t item) {
    uint32_t *q = pool + slot;
    __stcg(q + 0 * kPoolCap, f2u(p.ray.o.x)); __stcg(q + 1 * kPoolCap, f2u(p.ray.o.y)); __stcg(q + 2 * kPoolCap, f2u(p.ray.o.z));
    __stcg(q + 3 * kPoolCap, f2u(p.ray.d.x)); __stcg(q + 4 * kPoolCap, f2u(p.ray.d.y)); __stcg(q + 5 * kPoolCap, f2u(p.ray.d.z));
    __stcg(q + 6 * kPoolCap, f2u(p.T.x)); __stcg(q + 7 * kPoolCap, f2u(p.T.y)); __stcg(q + 8 * kPoolCap, f2u(p.T.z));
    __stcg(q + 9 * kPoolCap, f2u(p.p_att.x)); __stcg(q + 10 * kPoolCap, f2u(p.p_att.y)); __stcg(q + 11 * kPoolCap, f2u(p.p_att.z));
    __stcg(q + 12 * kPoolCap, f2u(p.p_n.x)); __stcg(q + 13 * kPoolCap, f2u(p.p_n.y)); __stcg(q + 14 * kPoolCap, f2u(p.p_n.z));
    __stcg(q + 15 * kPoolCap, (uint32_t) rng.state); __stcg(q + 16 * kPoolCap, (uint32_t) (rng.state >> 32));
    __stcg(q + 17 * kPoolCap, (uint32_t) rng.inc); __stcg(q + 18 * kPoolCap, (uint32_t) (rng.inc >> 32));
    __stcg(q + 19 * kPoolCap, f2u(p.ray.time));
    __stcg(q + 20 * kPoolCap, (p.depth & 0xFFu) | (p.pending << 8) | (((uint32_t) p.ray.inside & 0xFFu) << 10) | (item << 18));
}
__device__ __forceinline__ void unpark_path(const uint32_t *pool, uint32_t slot, Path &p, Rng &rng, uint32_t &item) {
    const uint32_t *q = pool + slot;
    p.ray.o = v3(u2f(__ldcg(q + 0 * kPoolCap)), u2f(__ldcg(q + 1 * kPoolCap)), u2f(__ldcg(q + 2 * kPoolCap)));
    p.ray.d = v3(u2f(__ldcg(q + 3 * kPoolCap)), u2f(__ldcg(q + 4 * kPoolCap)), u2f(__ldcg(q + 5 * kPoolCap)));
    p.T = v3(u2f(__ldcg(q + 6 * kPoolCap)), u2f(__ldcg(q + 7 * kPoolCap)), u2f(__ldcg(q + 8 * kPoolCap)));
    p.p_att = v3(u2f(__ldcg(q + 9 * kPoolCap)), u2f(__ldcg(q + 10 * kPoolCap)), u2f(__ldcg(q + 11 * kPoolCap)));
    p.p_n = v3(u2f(__ldcg(q + 12 * kPoolCap)), u2f(__ldcg(q + 13 * kPoolCap)), u2f(__ldcg(q + 14 * kPoolCap)));
    rng.state = (uint64_t) __ldcg(q + 15 * kPoolCap) | ((uint64_t) __ldcg(q + 16 * kPoolCap) << 32);
    rng.inc = (uint64_t) __ldcg(q + 17 * kPoolCap) | ((uint64_t) __ldcg(q + 18 * kPoolCap) << 32);
    p.ray.time = u2f(__ldcg(q + 19 * kPoolCap));
    const uint32_t f = __ldcg(q + 20 * kPoolCap);
    p.depth = f & 0xFFu;
    p.pending = (f >> 8) & 3u;
    p.ray.inside = (int) ((f >> 10) & 0xFFu);
    item = f >> 18;
    p.L = v3(0, 0, 0);
}

// Bin of the path's next ray.  Only a grouping heuristic (every lane still runs the exact algorithm), so
// fast reciprocals and the un-normalised direction are fine.
__device__ __forceinline__ uint32_t ray_class(const RenderArgs &a, const Path &p) {
    uint32_t c = 0;
    const float ix = __fdividef(1.0f, p.ray.d.x), iy = __fdividef(1.0f, p.ray.d.y), iz = __fdividef(1.0f, p.ray.d.z);
#pragma unroll 1
    for (uint32_t j = 0; j < a.n_cls_boxes; j++) {
        const float *b = a.cls_box[j];
        float t0 = (b[0] - p.ray.o.x) * ix, t1 = (b[3] - p.ray.o.x) * ix;
        float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
        t0 = (b[1] - p.ray.o.y) * iy; t1 = (b[4] - p.ray.o.y) * iy;
        lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
        t0 = (b[2] - p.ray.o.z) * iz; t1 = (b[5] - p.ray.o.z) * iz;
        lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
        if (hi >= fmaxf(lo, 0.0f)) c |= 1u << j;
    }
    if (a.cls_pending && p.pending) c |= 1u << a.n_cls_boxes;
    return c;
}

__device__ __forceinline__ unsigned long long warp_clock_ns() {
#if defined(__CUDA_ARCH__)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#else
    return 1ull;   // CPU build of the kernels (tests/host_emul)
#endif
}

template <uint32_t FEAT, int MINB, bool COOP, bool SEQ>
__global__ void __launch_bounds__(kBlock, MINB) render_pixel_binned(const RenderArgs a) {
    extern __shared__ uint32_t smem_stack[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    Stack st;
    st.base = smem_stack + (size_t) warp * a.stack_words * 32u + lane;
    st.stride = 32u;
    st.sp = 0;
    const uint32_t NB = a.n_bins;
    uint8_t *binq = reinterpret_cast<uint8_t *>(smem_stack + (size_t) kWarpsPerBlock * a.stack_words * 32u) + (size_t) warp * (NB + 1u) * kPoolCap;
    uint8_t *freeq = binq + (size_t) NB * kPoolCap;
    CoopArea ca;
    CoopStats cstats = {0, 0, 0, 0};
    unsigned long long cs_node_steps = 0, cs_node_items = 0, cs_leaf_steps = 0, cs_leaf_items = 0;
    if (COOP) ca.bind(smem_stack + (size_t) kWarpsPerBlock * a.stack_words * 32u + (size_t) kWarpsPerBlock * (NB + 1u) * (kPoolCap / 4u) + (size_t) warp * kCoopWords);
    uint32_t *pool = a.pool + ((size_t) blockIdx.x * kWarpsPerBlock + warp) * (size_t) (kPoolCap * kStateWords);
    // finished samples go to a per-warp staging array indexed by item (global memory, written once, read once) and
    // are summed per pixel at the end of the chunk in ITEM order -- so the result does not depend on which lane ran
    // which path, the chunk size does not depend on the samples per pixel, and no shared memory is needed for sums
    float4 *stage = a.stage + ((size_t) blockIdx.x * kWarpsPerBlock + warp) * (size_t) a.stage_items;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t ns = a.s_end - a.s_begin;
    // Statistics.  The 80-register kernels of the list scenes count rays and iterations per CHUNK, warp-uniform, in two 32-bit
    // registers flushed at the chunk end (C3 329 -> 317 ms: the segment loop is at the edge of its register budget); the 96-register
    // tree kernels keep three 64-bit per-lane counters for the whole launch (the per-chunk form cost them 1 %; profiles/r2_notes.md).
    constexpr bool kChunkCounters = (FEAT & MRT_FEAT_TREES) == 0;
    uint32_t rays = 0, iters = 0;                          // kChunkCounters
    unsigned long long rays64 = 0, nonfinite64 = 0, iters64 = 0;   // !kChunkCounters
    if (lane == 0) {   // when the warps were at work (MrtRenderStats.warp_time_sum_ns ...): nothing kept in a register across the launch
        const unsigned long long t_entry = warp_clock_ns();
        atomicMax(&a.counters[8], ~t_entry);
        atomicAdd(&a.counters[10], 0ull - t_entry);
    }

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = (*a.cancel) ? 0xFFFFFFFFu : atomicAdd(a.ticket, 1u);
        task = __shfl_sync(0xFFFFFFFFu, task, 0);
        if (task >= a.n_tasks) break;
        const bool r1 = task >= a.sched_task0[1], r2 = task >= a.sched_task0[2];
        const uint32_t run_task0 = r2 ? a.sched_task0[2] : (r1 ? a.sched_task0[1] : 0u);
        const uint32_t run_pix0 = r2 ? a.sched_pix0[2] : (r1 ? a.sched_pix0[1] : 0u);
        const uint32_t run_end = r2 ? a.sched_pix0[3] : (r1 ? a.sched_pix0[2] : a.sched_pix0[1]);
        const uint32_t run_k = r2 ? a.sched_k[2] : (r1 ? a.sched_k[1] : a.sched_k[0]);
        const uint32_t pix0 = run_pix0 + (task - run_task0) * run_k;
        const uint32_t kp = min(run_k, run_end - pix0);   // pixels in this chunk
        const uint32_t n_items = kp * ns;
#pragma unroll 1
        for (uint32_t i = lane; i < kPoolCap; i += 32u) freeq[i] = (uint8_t) i;
        uint32_t nfree = kPoolCap;   // warp-uniform
        uint32_t mycnt = 0;          // lane b holds the length of bin b
        uint32_t next_i = 0;         // warp-uniform stream cursor
        __syncwarp();

        for (;;) {
            // fullest bin
            uint32_t key = (lane < NB) ? ((mycnt << 8) | lane) : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(0xFFFFFFFFu, key, o));
            const uint32_t b = key & 0xFFu, c = key >> 8;
            uint32_t n;
            bool gen, drain = false;
            if (c >= 32u) { gen = false; n = 32u; }
            else if (next_i < n_items && nfree >= 32u) { gen = true; n = min(32u, n_items - next_i); }
            else if (c > 0u) { gen = false; drain = true; n = c; }
            else break;
            uint32_t slot = 0xFFu, k = 0;
            // Tree scenes have up to 8 bins and low-spp chunks: draining them one partial bin at a time left 14-20 %
            // of the lanes without a path (C4/C5).  List scenes (2-4 bins) lose 2 % to it instead -> compile-time.
            constexpr bool kMultiDrain = (FEAT & MRT_FEAT_TREES) != 0;
            if (kMultiDrain && drain && NB > 1u) {
                // no bin holds a full warp and there is nothing left to start (end of the chunk, or the pool is
                // full): fill the 32 lanes from several bins, in bin order
                n = 0;
#pragma unroll 1
                for (uint32_t b2 = 0; b2 < NB && n < 32u; b2++) {
                    const uint32_t cb = __shfl_sync(0xFFFFFFFFu, mycnt, b2);
                    const uint32_t take = min(cb, 32u - n);
                    if (lane >= n && lane < n + take) slot = binq[b2 * kPoolCap + (cb - take) + (lane - n)];
                    if (lane == b2) mycnt -= take;
                    n += take;
                }
            }
            const bool active = lane < n;
            Path p;
            Rng rng;
            if (gen) {
                if (active) {
                    k = next_i + lane;   // item index within the chunk
                    const uint32_t kk = k / ns;
                    const uint32_t s = a.s_begin + (k - kk * ns);
                    const uint32_t pix = a.order ? __ldg(a.order + pix0 + kk) : pix0 + kk;
                    const uint32_t cy = pix / a.crop_w, x = a.crop_x0 + (pix - cy * a.crop_w), y = a.crop_y0 + cy;
                    path_begin(a.sc, p, rng, x, y, s, a.sqrt_n, a.width, a.height, a.seed);
                }
                next_i += n;
            } else {
                if (!(kMultiDrain && drain && NB > 1u)) {
                    if (active) slot = binq[b * kPoolCap + (c - n) + lane];
                    if (lane == b) mycnt -= n;
                }
                if (active) unpark_path(pool, slot, p, rng, k);
            }
            bool cont = false;
            if constexpr (kChunkCounters) { iters++; rays += n; }   // n = lanes with a path in this step (warp-uniform)
            else { iters64++; if (active) rays64++; }
            if constexpr (COOP) cont = path_step_coop<FEAT>(a, p, rng, st, active, ca, cstats);
            else if (active) cont = path_step<FEAT>(a, p, rng, st);
            if (active && !cont) {
                const bool fin = path_sample_finite(p);
                __stcs(stage + k, fin ? make_float4(p.L.x, p.L.y, p.L.z, 1.0f) : make_float4(0.f, 0.f, 0.f, 0.f));
                if constexpr (kChunkCounters) { if (!fin) atomicAdd(&a.counters[2], 1ull); }   // rare
                else if (!fin) nonfinite64++;
            }
            // slots: finished paths return theirs, new survivors take one (never both in one iteration)
            const uint32_t m_free = __ballot_sync(0xFFFFFFFFu, active && !cont && slot != 0xFFu);
            const uint32_t m_alloc = __ballot_sync(0xFFFFFFFFu, cont && slot == 0xFFu);
            if (m_free) {
                if ((m_free >> lane) & 1u) freeq[nfree + __popc(m_free & lt_mask)] = (uint8_t) slot;
                nfree += __popc(m_free);
            }
            if (m_alloc) {
                if ((m_alloc >> lane) & 1u) slot = freeq[nfree - 1u - __popc(m_alloc & lt_mask)];
                nfree -= __popc(m_alloc);
            }
            uint32_t mybin = 0xFFFFFFFFu;
            if (cont) {
                mybin = (NB > 1u) ? ray_class(a, p) : 0u;
                park_path(pool, slot, p, rng, k);
            }
#pragma unroll 1
            for (uint32_t b2 = 0; b2 < NB; b2++) {
                const uint32_t m = __ballot_sync(0xFFFFFFFFu, mybin == b2);
                const uint32_t base = __shfl_sync(0xFFFFFFFFu, mycnt, b2);
                if (mybin == b2) binq[b2 * kPoolCap + base + __popc(m & lt_mask)] = (uint8_t) slot;
                if (lane == b2) mycnt += __popc(m);
            }
            __syncwarp();
        }
        if (COOP) {   // 32-bit per-chunk counters -> 64-bit totals
            cs_node_steps += cstats.node_steps; cs_node_items += cstats.node_items; cs_leaf_steps += cstats.leaf_steps; cs_leaf_items += cstats.leaf_items;
            cstats.node_steps = cstats.node_items = cstats.leaf_steps = cstats.leaf_items = 0u;
        }
        __syncwarp();
        __threadfence_block();   // this warp's staged samples (written by other lanes) are visible to every lane
        // Sum of a pixel's staged samples, in an order that depends on nothing but the samples per pixel of the launch.  Two forms,
        // two instantiations (SEQ), so that neither changes the register allocation of the other's segment loop:
        const unsigned long long t_sum0 = warp_clock_ns();
        if constexpr (SEQ) {
            // ns <= kStageBlock (the slice of a multi-GPU render, low-spp frames): lane = pixel, its samples added up one after the
            // other -- the reference's own order (main.cpp:154-166) -- with eight loads in flight per lane, 32 pixels per pass.  A
            // chunk costs 16 rounds of memory latency; lane-strided and one pixel at a time it cost 128 rounds, and the warp time spent
            // here went 3.7 % -> 1.4 % on a 128-sample slice of C2 (MrtRenderStats.stage_sum_ns; profiles/r2_notes.md).
#pragma unroll 1
            for (uint32_t k0 = 0; k0 < kp; k0 += 32u) {
                if (k0 + lane < kp) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 *src = stage + (k0 + lane) * ns;
#pragma unroll 1
                    for (uint32_t i = 0; i < ns; i += 8u) {
                        float4 q[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) q[j] = (i + j < ns) ? __ldcs(src + i + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < 8; j++) { v.x += q[j].x; v.y += q[j].y; v.z += q[j].z; v.w += q[j].w; }   // + 0.0f past the end: exact (v is never -0)
                    }
                    const uint32_t pix = a.order ? __ldg(a.order + pix0 + k0 + lane) : pix0 + k0 + lane;
                    if (a.accumulate) { float4 o = a.acc[pix]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                    a.acc[pix] = v;
                }
            }
        } else {
            // more samples per pixel: lane l adds items l, l+32, ... in order (eight loads in flight), then a fixed shuffle tree
#pragma unroll 1
            for (uint32_t k = 0; k < kp; k++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 *src = stage + k * ns;
                uint32_t i = lane;
#pragma unroll 1
                for (; i + 7u * 32u < ns; i += 8u * 32u) {
                    float4 q[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) q[j] = __ldcs(src + i + 32u * j);
#pragma unroll
                    for (int j = 0; j < 8; j++) { v.x += q[j].x; v.y += q[j].y; v.z += q[j].z; v.w += q[j].w; }
                }
#pragma unroll 1
                for (; i < ns; i += 32u) {
                    const float4 q = __ldcs(src + i);
                    v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                }
                v.x = warp_sum(v.x); v.y = warp_sum(v.y); v.z = warp_sum(v.z); v.w = warp_sum(v.w);
                if (lane == 0) {
                    const uint32_t pix = a.order ? __ldg(a.order + pix0 + k) : pix0 + k;
                    if (a.accumulate) { float4 o = a.acc[pix]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
                    a.acc[pix] = v;
                }
            }
        }
        if constexpr (kChunkCounters) {
            if (lane == 0) {
                atomicAdd(&a.counters[3], warp_clock_ns() - t_sum0);   // MrtRenderStats.stage_sum_ns
                atomicAdd(&a.counters[0], (unsigned long long) rays);
                atomicAdd(&a.counters[1], (unsigned long long) iters);
            }
            rays = 0; iters = 0;
        } else {
            if (lane == 0) atomicAdd(&a.counters[3], warp_clock_ns() - t_sum0);   // MrtRenderStats.stage_sum_ns
        }
        __syncwarp();
    }
    if constexpr (!kChunkCounters) {
        for (int o = 16; o > 0; o >>= 1) {
            rays64 += __shfl_xor_sync(0xFFFFFFFFu, rays64, o);
            nonfinite64 += __shfl_xor_sync(0xFFFFFFFFu, nonfinite64, o);
        }
    }
    if (lane == 0) {
        if constexpr (!kChunkCounters) {
            atomicAdd(&a.counters[0], rays64);
            atomicAdd(&a.counters[1], iters64);
            atomicAdd(&a.counters[2], nonfinite64);
        }
        if (COOP) {
            atomicAdd(&a.counters[4], cs_node_steps); atomicAdd(&a.counters[5], cs_node_items);
            atomicAdd(&a.counters[6], cs_leaf_steps); atomicAdd(&a.counters[7], cs_leaf_items);
        }
        const unsigned long long t_exit = warp_clock_ns();   // counters[10] = sum of (exit - entry), modulo 2^64
        atomicMax(&a.counters[9], t_exit); atomicAdd(&a.counters[10], t_exit); atomicMax(&a.counters[11], ~t_exit);
    }
}

// kernel entry for a feature mask (defined once per render_variant_*.cu); kind: 0 = pixel per lane,
// 1 = pixel per warp, 2 = pixel per warp with binned path pool, 3 = binned + warp-cooperative tree traversal (only
// instantiated for masks with MRT_FEAT_TREES); 4, 5 = 2, 3 for launches with at most kStageBlock samples per pixel (SEQ)
template <uint32_t FEAT, int MINB>
inline const void *variant_kernel_minb(int kind) {
    if constexpr ((FEAT & MRT_FEAT_TREES) != 0) {
        if (kind == 3) return (const void *) render_pixel_binned<FEAT, MINB, true, false>;
        if (kind == 5) return (const void *) render_pixel_binned<FEAT, MINB, true, true>;
    }
    if (kind == 4) return (const void *) render_pixel_binned<FEAT, MINB, false, true>;
    if (kind >= 2) return (const void *) render_pixel_binned<FEAT, MINB, false, false>;
    return kind ? (const void *) render_pixel_per_warp<FEAT, MINB> : (const void *) render_pixel_per_lane<FEAT, MINB>;
}
template <uint32_t FEAT>
inline const void *variant_kernel(int kind, int minb) {
    // small variants also come with tighter register budgets (7 / 8 resident blocks of 128 threads)
    if constexpr ((FEAT & (MRT_FEAT_TREES | MRT_FEAT_TEX)) == 0) {
        if (minb == 7) return variant_kernel_minb<FEAT, 7>(kind);
        if (minb == 8) return variant_kernel_minb<FEAT, 8>(kind);
    }
    return (minb == 5) ? variant_kernel_minb<FEAT, 5>(kind) : variant_kernel_minb<FEAT, 6>(kind);
}

}  // namespace mrt
