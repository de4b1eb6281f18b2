// Wavefront renderer (the "optional wavefront ray compaction" of the design): the same per-path phases
// as the megakernel (trace_core.h: path_begin / path_advance / traversal / path_shade), but split into two
// kernels per bounce iteration over a resident pool of path slots:
//
//   wf_logic : one thread per slot -- shade the hit of the previous segment, start the slot's next sample
//              when its path ended, normalise the new ray (+ deferred diffuse weight) and store it;
//   wf_trav  : persistent warps running the INCREMENTAL traversal (trav_step); a lane whose ray has finished
//              is refilled from the slot pool at once, so lanes do not idle while the longest traversal of a
//              batch completes (ncu round 1: 14-23 % warp execution efficiency in the BVH scenes with one
//              ray per lane per segment).
//
// A slot owns a fixed (pixel, sub-stream) pair and walks its samples in increasing order, accumulating its
// own sum: no atomics on radiance, results are reproducible and, with one slot per pixel, the per-pixel sum
// has the reference's order.  Path state lives in HBM/L2 as float4 streams (coalesced: slot index == thread
// index in wf_logic); the slot pool is sized to stay L2 resident (~150 B per slot).
#include <cstdlib>

#include "gpu_internal.h"

namespace mrt {

#define WF_HAS_RAY (1u << 20)
#define WF_FINISHED (1u << 21)
#define WF_MISS 0xFFFFFFFFu

struct WaveState {
    float4 *R0, *R1, *R2;   // ray: (o, time) (d, inside) (inv, mask)
    float4 *H0, *H1;        // hit: (p, mat | WF_MISS) (n, u)
    float *H2;              //      v
    float4 *S0, *S1, *S4;   // (T, flags|depth) (L, sample index) (sum of finished samples, count)
    uint4 *G;               // PCG32 state, inc
};

struct WaveArgs {
    SceneView sc;
    WaveState st;
    uint32_t width, height, sqrt_n, s_begin, s_end, max_bounces;
    uint64_t seed;
    uint32_t pix0, n_slots, c;      // slot i <-> pixel pix0 + i / c, sub-stream i % c (samples s_begin + k, + c, ...)
    uint32_t has_volumes, stack_words, accumulate, first, count_live;
    float4 *acc;
    unsigned int *ctrl;             // [0] traversal cursor  [1] live slots (when count_live)
    unsigned long long *counters;   // [0] rays [1] traversal warp iterations [2] non-finite samples
};

__device__ __forceinline__ float4 f4(V3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ V3 xyz(float4 v) { return v3(v.x, v.y, v.z); }

// ------------------------------------------------------------------ logic
__global__ void __launch_bounds__(256) wf_logic(const WaveArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    bool live = false;
    uint32_t new_rays = 0, dropped = 0;
    if (i < a.n_slots) {
        const uint32_t pix = a.pix0 + i / a.c, k = i - (i / a.c) * a.c;
        Path p;
        Rng rng;
        uint32_t s;
        float4 sum;
        bool need_new = false, finished = false;
        if (a.first) {
            s = a.s_begin + k - a.c;   // "next sample" below lands on s_begin + k
            sum = make_float4(0.f, 0.f, 0.f, 0.f);
            rng.state = 0; rng.inc = 0;
            need_new = true;
        } else {
            const float4 s0 = a.st.S0[i];
            const uint32_t flags = __float_as_uint(s0.w);
            if (flags & WF_FINISHED) {
                finished = true;
            } else {
                const float4 s1 = a.st.S1[i];
                const uint4 g = a.st.G[i];
                sum = a.st.S4[i];
                s = __float_as_uint(s1.w);
                rng.state = ((uint64_t) g.y << 32) | g.x;
                rng.inc = ((uint64_t) g.w << 32) | g.z;
                const float4 r0 = a.st.R0[i], r1 = a.st.R1[i];
                p.ray.o = xyz(r0); p.ray.time = r0.w;
                p.ray.d = xyz(r1); p.ray.inside = (int) __float_as_uint(r1.w);
                p.T = xyz(s0); p.L = xyz(s1);
                p.depth = flags & 0xFFFFu;
                p.pending = 0;
                const float4 h0 = a.st.H0[i];
                Hit rec;
                const uint32_t mat = __float_as_uint(h0.w);
                const bool hit = mat != WF_MISS;
                if (hit) {
                    const float4 h1 = a.st.H1[i];
                    rec.p = xyz(h0); rec.n = xyz(h1); rec.u = h1.w; rec.v = a.st.H2[i]; rec.mat = mat; rec.t = 0.f;
                }
                if (!path_shade(MRT_FEAT_ALL, a.sc, p, hit, rec, a.max_bounces, rng)) {
                    if (path_sample_finite(p)) { sum.x += p.L.x; sum.y += p.L.y; sum.z += p.L.z; sum.w += 1.0f; }
                    else dropped++;
                    need_new = true;
                }
            }
        }
        if (!finished) {
            if (need_new) {
                s += a.c;
                if (s < a.s_end) {
                    const uint32_t y = pix / a.width, x = pix - y * a.width;
                    path_begin(a.sc, p, rng, x, y, s, a.sqrt_n, a.width, a.height, a.seed);
                } else {
                    finished = true;
                    a.st.S0[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(WF_FINISHED));
                    a.st.S4[i] = sum;
                }
            }
            if (!finished) {
                path_advance(MRT_FEAT_ALL, a.sc, p);
                new_rays = 1;
                live = true;
                a.st.R0[i] = f4(p.ray.o, p.ray.time);
                a.st.R1[i] = f4(p.ray.d, __uint_as_float((uint32_t) p.ray.inside));
                a.st.R2[i] = f4(p.ray.inv, __uint_as_float(p.ray.mask));
                a.st.S0[i] = f4(p.T, __uint_as_float((p.depth & 0xFFFFu) | WF_HAS_RAY));
                a.st.S1[i] = f4(p.L, __uint_as_float(s));
                if (need_new) a.st.S4[i] = sum;
                a.st.G[i] = make_uint4((uint32_t) rng.state, (uint32_t) (rng.state >> 32), (uint32_t) rng.inc, (uint32_t) (rng.inc >> 32));
            }
        }
    }
    // statistics, one atomic per warp
    const uint32_t nr = __popc(__ballot_sync(0xFFFFFFFFu, new_rays != 0));
    const uint32_t nd = __reduce_add_sync(0xFFFFFFFFu, dropped);
    const uint32_t nl = __popc(__ballot_sync(0xFFFFFFFFu, live));
    if (lane == 0) {
        if (nr) atomicAdd(&a.counters[0], (unsigned long long) nr);
        if (nd) atomicAdd(&a.counters[2], (unsigned long long) nd);
        if (a.count_live && nl) atomicAdd(&a.ctrl[1], nl);
    }
    if (i == 0) a.ctrl[0] = 0;   // rewind the traversal cursor for the wf_trav launch that follows
}

// -------------------------------------------------------------- traversal
constexpr int kTravBlock = 128;
constexpr int kTravSteps = 8;   // actions between two refills

template <int MINB>
__global__ void __launch_bounds__(kTravBlock, MINB) wf_trav(const WaveArgs a) {
    extern __shared__ uint32_t smem_stack[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    Stack st;
    st.base = smem_stack + (size_t) warp * a.stack_words * 32u + lane;
    st.stride = 32u;
    st.sp = 0;
    const uint32_t lt_mask = (1u << lane) - 1u;
    bool busy = false;
    uint32_t slot = 0;
    Ray ray;
    Trav tr;
    Hit rec;
    Rng rng;
    rng.state = 0; rng.inc = 0;
    uint32_t pool_next = 0, pool_end = 0;   // warp-uniform: slots [pool_next, pool_end) are ours to hand out
    bool exhausted = false;
    unsigned long long iters = 0;

    for (;;) {
        uint32_t idle = __ballot_sync(0xFFFFFFFFu, !busy);
        if (idle) {
            if (pool_next == pool_end && !exhausted) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&a.ctrl[0], 64u);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (base >= a.n_slots) exhausted = true;
                else { pool_next = base; pool_end = min(base + 64u, a.n_slots); }
            }
            if (pool_next != pool_end) {
                const uint32_t n = min((uint32_t) __popc(idle), pool_end - pool_next);
                if (!busy) {
                    const uint32_t rank = __popc(idle & lt_mask);
                    if (rank < n) {
                        const uint32_t my = pool_next + rank;
                        const uint32_t flags = __float_as_uint(a.st.S0[my].w);
                        if ((flags & WF_HAS_RAY) && !(flags & WF_FINISHED)) {
                            const float4 r0 = a.st.R0[my], r1 = a.st.R1[my], r2 = a.st.R2[my];
                            ray.o = xyz(r0); ray.time = r0.w;
                            ray.d = xyz(r1); ray.inside = (int) __float_as_uint(r1.w);
                            ray.inv = xyz(r2); ray.mask = __float_as_uint(r2.w);
                            if (a.has_volumes) {
                                const uint4 g = a.st.G[my];
                                rng.state = ((uint64_t) g.y << 32) | g.x;
                                rng.inc = ((uint64_t) g.w << 32) | g.z;
                            }
                            trav_begin(a.sc, tr, 0.001f, FLT_MAX, st);
                            slot = my;
                            busy = true;
                        }
                    }
                }
                pool_next += n;
            }
        }
        if (!__any_sync(0xFFFFFFFFu, busy)) {
            if (exhausted && pool_next == pool_end) break;
            continue;
        }
        iters++;
#pragma unroll 1
        for (int step = 0; step < kTravSteps; step++) {
            if (busy) {
                if (trav_active(tr, st)) {
                    trav_step(a.sc, tr, ray, rec, rng, st, nullptr);
                } else {
                    if (trav_hit(tr)) {
                        a.st.H0[slot] = f4(rec.p, __uint_as_float(rec.mat));
                        a.st.H1[slot] = f4(rec.n, rec.u);
                        a.st.H2[slot] = rec.v;
                    } else {
                        a.st.H0[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(WF_MISS));
                    }
                    if (a.has_volumes)
                        a.st.G[slot] = make_uint4((uint32_t) rng.state, (uint32_t) (rng.state >> 32), (uint32_t) rng.inc, (uint32_t) (rng.inc >> 32));
                    busy = false;
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) iters += __shfl_xor_sync(0xFFFFFFFFu, iters, o);
    if (lane == 0) atomicAdd(&a.counters[1], iters / 32ull);
}

// ------------------------------------------------------------------ flush
__global__ void wf_flush(const WaveArgs a, uint32_t n_pixels_tile) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_pixels_tile) return;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t k = 0; k < a.c; k++) {   // fixed order
        const float4 p = a.st.S4[j * a.c + k];
        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    const uint32_t pix = a.pix0 + j;
    if (a.accumulate) { const float4 o = a.acc[pix]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    a.acc[pix] = v;
}

}  // namespace mrt

using namespace mrt;

void mrt_wavefront_free(MrtScene *s) {
    if (s->wf_state) { cudaFree(s->wf_state); s->wf_state = nullptr; s->wf_state_bytes = 0; }
    if (s->wf_ctrl) { cudaFree(s->wf_ctrl); s->wf_ctrl = nullptr; }
}

int mrt_wavefront_render(MrtScene *s, const MrtRenderParams *p, float4 *acc, uint32_t sqrt_n) {
    const uint32_t n_pixels = p->width * p->height;
    const uint32_t ns = p->sample_end - p->sample_begin;
    uint32_t target = 512u * 1024u;
    if (const char *e = getenv("MRT_WF_SLOTS")) target = (uint32_t) atoi(e);
    if (target < 1024u) target = 1024u;
    uint32_t c = 1, tile_pixels = target;
    if (n_pixels < target) {
        tile_pixels = n_pixels;
        c = target / n_pixels;
        if (c > ns) c = ns;
        if (c > 64u) c = 64u;
        if (c < 1u) c = 1u;
    }
    const size_t B = (size_t) tile_pixels * c;
    const size_t per_slot = 9 * sizeof(float4) + sizeof(float) + sizeof(uint4);
    const size_t bytes = B * per_slot + 256;
    if (s->wf_state_bytes < bytes) {
        if (s->wf_state) { cudaFree(s->wf_state); s->wf_state = nullptr; s->wf_state_bytes = 0; }
        CUDA_TRY(cudaMalloc(&s->wf_state, bytes));
        s->wf_state_bytes = bytes;
    }
    if (!s->wf_ctrl) CUDA_TRY(cudaMalloc(&s->wf_ctrl, 4 * sizeof(unsigned int)));

    WaveArgs a;
    a.sc = s->view;
    char *base = (char *) s->wf_state;
    auto take = [&](size_t elem) { char *r = base; base += B * elem; return r; };
    a.st.R0 = (float4 *) take(16); a.st.R1 = (float4 *) take(16); a.st.R2 = (float4 *) take(16);
    a.st.H0 = (float4 *) take(16); a.st.H1 = (float4 *) take(16);
    a.st.S0 = (float4 *) take(16); a.st.S1 = (float4 *) take(16); a.st.S4 = (float4 *) take(16);
    a.st.G = (uint4 *) take(16);
    a.st.H2 = (float *) take(4);
    a.width = p->width; a.height = p->height; a.sqrt_n = sqrt_n;
    a.s_begin = p->sample_begin; a.s_end = p->sample_end; a.max_bounces = p->max_bounces;
    a.seed = p->seed;
    a.c = c;
    a.has_volumes = s->has_volumes;
    a.stack_words = s->stack_words;
    a.accumulate = (p->flags & MRT_RENDER_ACCUMULATE) ? 1u : 0u;
    a.acc = acc;
    a.ctrl = s->wf_ctrl;
    a.counters = s->counters;

    typedef void (*trav_t)(const WaveArgs);
    int minb = 8;
    if (const char *e = getenv("MRT_WF_MINB")) minb = atoi(e);
    trav_t trav = (minb <= 4) ? wf_trav<4> : (minb <= 6 ? wf_trav<6> : wf_trav<8>);
    const size_t smem = (size_t) (kTravBlock / 32) * s->stack_words * 32u * sizeof(uint32_t);
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(trav, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    int blocks_per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, trav, kTravBlock, smem));
    if (blocks_per_sm < 1) { set_error("wavefront traversal kernel does not fit on an SM"); return MRT_E_CUDA; }
    const uint32_t trav_grid = (uint32_t) s->sm_count * (uint32_t) blocks_per_sm;
    int poll_every = 16;
    if (const char *e = getenv("MRT_WF_POLL")) poll_every = atoi(e);
    if (poll_every < 1) poll_every = 1;

    unsigned int *live_host = (unsigned int *) s->poll_host;   // pinned
    for (uint32_t pix0 = 0; pix0 < n_pixels; pix0 += tile_pixels) {
        const uint32_t tp = (n_pixels - pix0 < tile_pixels) ? (n_pixels - pix0) : tile_pixels;
        a.pix0 = pix0;
        a.n_slots = tp * c;
        const uint32_t logic_grid = (a.n_slots + 255u) / 256u;
        a.first = 1; a.count_live = 0;
        wf_logic<<<logic_grid, 256, 0, s->stream>>>(a);
        a.first = 0;
        for (uint64_t it = 1;; it++) {
            trav<<<trav_grid, kTravBlock, smem, s->stream>>>(a);
            const bool poll = (it % (uint64_t) poll_every) == 0;
            a.count_live = poll ? 1u : 0u;
            if (poll) CUDA_TRY(cudaMemsetAsync(&s->wf_ctrl[1], 0, sizeof(unsigned int), s->stream));
            wf_logic<<<logic_grid, 256, 0, s->stream>>>(a);
            if (poll) {
                CUDA_TRY(cudaMemcpyAsync(live_host, &s->wf_ctrl[1], sizeof(unsigned int), cudaMemcpyDeviceToHost, s->stream));
                CUDA_TRY(cudaStreamSynchronize(s->stream));
                if (*live_host == 0) break;
            }
        }
        wf_flush<<<(tp + 255u) / 256u, 256, 0, s->stream>>>(a, tp);
        CUDA_TRY(cudaGetLastError());
    }
    return MRT_OK;
}
