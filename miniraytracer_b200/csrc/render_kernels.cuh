// Render megakernels, templated on the scene-feature mask (MRT_FEAT_*, see trace_core.h) and on the
// launch-bounds variant.  Each feature mask is instantiated in its own translation unit
// (render_variant_*.cu) so that the variants compile in parallel.
//
// Work decomposition (replaces work_queue.cpp + the draw() loop nest, main.cpp:138-188):
//   * persistent warps: the grid is sized to the SM count x resident blocks and every warp pulls work from
//     one global atomic ticket counter (one atomic per warp task, issued by lane 0, broadcast by shuffle);
//   * a warp task is a POOL of (pixel, sample) items that the 32 lanes drain cooperatively: whenever a
//     lane's path terminates it is regenerated from the pool at the next warp-converged point (ballot +
//     popc prefix), so lanes do not idle while the longest path of a batch finishes;
//   * per-thread traversal stacks live in shared memory, interleaved by lane (word k of lane l at
//     [k*32 + l]) so pushes/pops are bank-conflict free.
// Every accumulator element has exactly one writer: no float atomics, results are reproducible run to run.
#pragma once
#include "gpu_internal.h"

namespace mrt {

struct RenderArgs {
    SceneView sc;
    uint32_t width, height, sqrt_n, s_begin, s_end, max_bounces;
    uint64_t seed;
    uint32_t accumulate;
    uint32_t stack_words;
    uint32_t n_tasks, pixels_per_task;
    float4 *acc;
    unsigned int *ticket;             // global task counter
    unsigned long long *counters;     // [0] rays [1] paths [2] nonfinite
    const uint32_t *order;            // work order: item i of the queue is pixel order[i] (Morton tiles), or NULL = row-major
    const volatile int *cancel;       // device flag, written by mrt_gpu_cancel through a side stream
};

constexpr int kWarpsPerBlock = kBlock / 32;

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
    const uint32_t n_pixels = a.width * a.height;
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
                    const uint32_t y = pix / a.width, x = pix - y * a.width;
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
    const uint32_t n_pixels = a.width * a.height;
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
                    y = pix / a.width; x = pix - y * a.width;
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

// kernel entry for a feature mask (defined once per render_variant_*.cu)
template <uint32_t FEAT>
inline const void *variant_kernel(bool pixel_per_warp, int minb) {
    // small variants also come with tighter register budgets (7 / 8 resident blocks of 128 threads)
    if constexpr ((FEAT & (MRT_FEAT_TREES | MRT_FEAT_TEX)) == 0) {
        if (minb == 7) return pixel_per_warp ? (const void *) render_pixel_per_warp<FEAT, 7> : (const void *) render_pixel_per_lane<FEAT, 7>;
        if (minb == 8) return pixel_per_warp ? (const void *) render_pixel_per_warp<FEAT, 8> : (const void *) render_pixel_per_lane<FEAT, 8>;
    }
    if (pixel_per_warp) return (minb == 5) ? (const void *) render_pixel_per_warp<FEAT, 5> : (const void *) render_pixel_per_warp<FEAT, 6>;
    return (minb == 5) ? (const void *) render_pixel_per_lane<FEAT, 5> : (const void *) render_pixel_per_lane<FEAT, 6>;
}

}  // namespace mrt
