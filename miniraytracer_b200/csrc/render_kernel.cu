// Device half of the C ABI (include/mrt_gpu.h): scene upload, launch configuration of the render kernels
// (render_kernels.cuh -- the work decomposition is described there), finalize / tone-map kernels, readback.
// The library reads no environment variables: scheduling knobs come in through mrt_gpu_set_tuning (MrtTuning).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "gpu_internal.h"
#include "scene_graph.h"
#include "render_kernels.cuh"
#include "render_variants.h"
#include "schedule.h"

namespace mrt {

// --------------------------------------------------------------- finalize
// color = sum / count ; luminance clamp (main.cpp:168-173, vec3.h:275-279)
__global__ void finalize_kernel(const float4 *acc, float4 *out, uint32_t n, float max_lum) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = acc[i];
    float r = 0, g = 0, b = 0;
    if (a.w > 0) { r = __fdiv_rn(a.x, a.w); g = __fdiv_rn(a.y, a.w); b = __fdiv_rn(a.z, a.w); }
    float lum = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.212655f), __fmul_rn(g, 0.715158f)), __fmul_rn(b, 0.072187f));
    if (lum > max_lum) {
        float k = __fdiv_rn(max_lum, lum);
        r = __fmul_rn(r, k); g = __fmul_rn(g, k); b = __fmul_rn(b, k);
    }
    out[i] = make_float4(r, g, b, a.w);
}

// --------------------------------------------------------------- draw2's pixel update
// The reference's default worker (draw2, main.cpp:193-243) renders sample-major passes and keeps a RUNNING MEAN per pixel:
// a non-finite sample is replaced by the mean so far (by 0 for the first sample), the mean advances by (c - mean) / (s + 1),
// and the luminance clamp is applied after every pass and feeds back into the mean (main.cpp:214-231).  `acc` holds ONE
// sample per pixel (a one-sample launch: xyz = the sample or 0, w = 1 if it was finite); `mean` is updated in place.
__global__ void running_mean_kernel(const float4 *acc, float4 *mean, uint32_t n, uint32_t pass, float max_lum) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = acc[i];
    const float4 m = mean[i];
    float r = a.x, g = a.y, b = a.z;
    if (!(a.w > 0)) { r = pass ? m.x : 0.0f; g = pass ? m.y : 0.0f; b = pass ? m.z : 0.0f; }
    if (pass) {
        const float k = __fdiv_rn(1.0f, __fadd_rn((float) pass, 1.0f));
        r = __fadd_rn(m.x, __fmul_rn(__fsub_rn(r, m.x), k));
        g = __fadd_rn(m.y, __fmul_rn(__fsub_rn(g, m.y), k));
        b = __fadd_rn(m.z, __fmul_rn(__fsub_rn(b, m.z), k));
    }
    const float lum = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.212655f), __fmul_rn(g, 0.715158f)), __fmul_rn(b, 0.072187f));
    if (lum > max_lum) {
        const float k = __fdiv_rn(max_lum, lum);
        r = __fmul_rn(r, k); g = __fmul_rn(g, k); b = __fmul_rn(b, k);
    }
    mean[i] = make_float4(r, g, b, (float) (pass + 1u));
}

// --------------------------------------------------------------- tone map
// Adaptive logarithmic mapping of the reference's preview loop (main.cpp:416-444)
__global__ void max_luminance_kernel(const float4 *img, uint32_t n, unsigned int *max_bits) {
    float m = 0.0f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 c = img[i];
        float lum = __fadd_rn(__fadd_rn(__fmul_rn(c.x, 0.212655f), __fmul_rn(c.y, 0.715158f)), __fmul_rn(c.z, 0.072187f));
        m = fmaxf(m, lum);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(max_bits, __float_as_uint(m));   // lum >= 0: uint order == float order
}
__global__ void tonemap_kernel(const float4 *img, uint32_t *argb, uint32_t n, const unsigned int *max_bits) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float L_dmax = 230.0f;
    const float bias = logf(0.7f) / logf(0.5f);
    float L_wmax = __uint_as_float(*max_bits);
    float invlogmax = 1.0f / log10f(L_wmax + 1.0f);
    float invmax = 1.0f / L_wmax;
    float4 c = img[i];
    float lum = __fadd_rn(__fadd_rn(__fmul_rn(c.x, 0.212655f), __fmul_rn(c.y, 0.715158f)), __fmul_rn(c.z, 0.072187f));
    float loglw = logf(lum + 1.0f);
    float lum_new = (L_dmax * 0.01f * invlogmax) * (loglw / logf(2 + powf(lum * invmax, bias) * 8));
    float d = lum + 0.00001f;
    float r = fminf((lum_new * c.x) / d, 1.0f) * 255.99f;
    float g = fminf((lum_new * c.y) / d, 1.0f) * 255.99f;
    float b = fminf((lum_new * c.z) / d, 1.0f) * 255.99f;
    argb[i] = ((uint32_t) r << 16) | ((uint32_t) g << 8) | (uint32_t) b;   // ARGB32, vec3.h:327-333
}

// --------------------------------------------------------------- multi-GPU sum + finalize over NVLink
// The path shards by samples per pixel: GPU k holds the accumulator (sum of finite radiance, count) of its sample slice for
// the whole frame.  This kernel is the exchange step as ONE pass over peer memory: the launching GPU owns a stripe
// of the pixels, reads that stripe from every GPU's accumulator (its own from HBM, the others' over NVLink / NVSwitch -- peer
// loads), adds them in GPU order (deterministic, the order a host loop would use), applies mean + luminance clamp
// (main.cpp:168-173) and stores the finished pixels straight into the root GPU's image (peer store).  Every byte crosses
// NVLink once; nothing goes through the host.
constexpr int kMaxReduceGpus = 16;
struct ReduceArgs {
    const float4 *acc[kMaxReduceGpus];
    int n;
};
__global__ void reduce_finalize_kernel(const ReduceArgs a, float4 *out, uint32_t begin, uint32_t end, float max_lum) {
    const uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    float4 s = a.acc[0][i];
#pragma unroll 1
    for (int k = 1; k < a.n; k++) {
        const float4 v = a.acc[k][i];
        s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y); s.z = __fadd_rn(s.z, v.z); s.w = __fadd_rn(s.w, v.w);
    }
    float r = 0, g = 0, b = 0;
    if (s.w > 0) { r = __fdiv_rn(s.x, s.w); g = __fdiv_rn(s.y, s.w); b = __fdiv_rn(s.z, s.w); }
    const float lum = __fadd_rn(__fadd_rn(__fmul_rn(r, 0.212655f), __fmul_rn(g, 0.715158f)), __fmul_rn(b, 0.072187f));
    if (lum > max_lum) {
        const float k = __fdiv_rn(max_lum, lum);
        r = __fmul_rn(r, k); g = __fmul_rn(g, k); b = __fmul_rn(b, k);
    }
    out[i] = make_float4(r, g, b, s.w);
}

}  // namespace mrt

// =========================================================================
//                          C ABI (device half)
// =========================================================================
using namespace mrt;

// Scene tables are packed into ONE device allocation and copied with ONE transfer (a scene is ~20 small tables; a
// cudaMalloc + cudaMemcpy each cost ~10 ms per upload).  The control block (ticket, counters, ...) lives in the same
// allocation.
struct Packer {
    struct Item { const void *host; size_t bytes; size_t offset; const void **dev; };
    std::vector<Item> items;
    size_t total = 0;
    size_t reserve(size_t bytes) { size_t o = total; total += (bytes + 255u) & ~(size_t) 255u; return o; }
    template <typename T>
    void add(const T *host, size_t count, const T **dev) {
        *dev = nullptr;
        if (!host || !count) return;
        Item it{host, count * sizeof(T), 0, (const void **) dev};
        it.offset = reserve(it.bytes);
        items.push_back(it);
    }
};

// Small pinned words (cancel flag staging, poll results): slots of one process-wide pinned block, because
// cudaHostAlloc / cudaFreeHost per scene cost milliseconds.
static std::mutex g_pin_mu;
static unsigned char *g_pin_block = nullptr;
static uint64_t g_pin_used = 0;   // bitmap of 64 slots x 64 bytes
static void *pinned_slot_acquire() {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (!g_pin_block && cudaHostAlloc((void **) &g_pin_block, 64 * 64, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); g_pin_block = nullptr; return nullptr; }
    for (int i = 0; i < 64; i++)
        if (!((g_pin_used >> i) & 1u)) { g_pin_used |= (uint64_t) 1 << i; memset(g_pin_block + 64 * i, 0, 64); return g_pin_block + 64 * i; }
    return nullptr;
}
static bool pinned_slot_release(void *p) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (!g_pin_block || (unsigned char *) p < g_pin_block || (unsigned char *) p >= g_pin_block + 64 * 64) return false;
    g_pin_used &= ~((uint64_t) 1 << (((unsigned char *) p - g_pin_block) / 64));
    return true;
}

// Device buffers (scene tables, accumulators, path pool, sample staging, ...), the poll stream and the timing events of a
// scene come from per-process caches and go back there when the scene is destroyed: a sequence of scenes / frames then pays
// neither cudaMalloc nor cudaFree (an implicit device-wide sync that was measured at 0.2 .. 60 ms per call, and made the
// end-to-end figure jitter by 10 %).  A request is served by a cached buffer of the same device whose capacity is between
// the request and twice the request (so a 1 KB table never pins the 230 MB staging array); at most 32 buffers / 2 GiB stay cached.
struct DevBuf { int device; void *ptr; size_t bytes; };
static std::mutex g_buf_mu;
static std::vector<DevBuf> g_buf_free;
static size_t g_buf_cached = 0;
static void *devbuf_acquire(int device, size_t bytes, size_t *got_bytes) {
    if (bytes < 256) bytes = 256;
    {
        std::lock_guard<std::mutex> lk(g_buf_mu);
        size_t best = g_buf_free.size();
        for (size_t i = 0; i < g_buf_free.size(); i++)
            if (g_buf_free[i].device == device && g_buf_free[i].bytes >= bytes && g_buf_free[i].bytes <= 2 * bytes &&
                (best == g_buf_free.size() || g_buf_free[i].bytes < g_buf_free[best].bytes)) best = i;
        if (best != g_buf_free.size()) {
            DevBuf b = g_buf_free[best];
            g_buf_free.erase(g_buf_free.begin() + best);
            g_buf_cached -= b.bytes;
            if (got_bytes) *got_bytes = b.bytes;
            return b.ptr;
        }
    }
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    if (got_bytes) *got_bytes = bytes;
    return p;
}
static void devbuf_release(int device, void *ptr, size_t bytes) {
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(g_buf_mu);
    if (g_buf_free.size() >= 32 || g_buf_cached + bytes > ((size_t) 2 << 30)) { cudaFree(ptr); return; }
    g_buf_free.push_back(DevBuf{device, ptr, bytes});
    g_buf_cached += bytes;
}
static uint32_t *pool_acquire(int device, size_t words, size_t *got_words) {
    size_t got = 0;
    uint32_t *p = (uint32_t *) devbuf_acquire(device, words * sizeof(uint32_t), &got);
    *got_words = got / sizeof(uint32_t);
    return p;
}
static void pool_release(int device, uint32_t *ptr, size_t words) { devbuf_release(device, ptr, words * sizeof(uint32_t)); }

// pinned host staging for the scene upload (cudaHostAlloc costs milliseconds): same caching policy, at most 8 buffers / 256 MiB
struct PinBuf { void *ptr; size_t bytes; };
static std::vector<PinBuf> g_pin_free;
static size_t g_pin_cached = 0;
static void *pinbuf_acquire(size_t bytes, size_t *got_bytes) {
    if (bytes < 4096) bytes = 4096;
    {
        std::lock_guard<std::mutex> lk(g_buf_mu);
        for (size_t i = 0; i < g_pin_free.size(); i++)
            if (g_pin_free[i].bytes >= bytes && g_pin_free[i].bytes <= 2 * bytes) {
                PinBuf b = g_pin_free[i];
                g_pin_free.erase(g_pin_free.begin() + i);
                g_pin_cached -= b.bytes;
                *got_bytes = b.bytes;
                return b.ptr;
            }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *got_bytes = bytes;
    return p;
}
static void pinbuf_release(void *ptr, size_t bytes) {
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(g_buf_mu);
    if (g_pin_free.size() >= 8 || g_pin_cached + bytes > ((size_t) 256 << 20)) { cudaFreeHost(ptr); return; }
    g_pin_free.push_back(PinBuf{ptr, bytes});
    g_pin_cached += bytes;
}

struct SyncObjs { int device; cudaStream_t stream; cudaEvent_t ev0, ev1, ev_up; };
static std::vector<SyncObjs> g_sync_free;
static bool syncobjs_acquire(int device, SyncObjs *out) {
    {
        std::lock_guard<std::mutex> lk(g_buf_mu);
        for (size_t i = 0; i < g_sync_free.size(); i++)
            if (g_sync_free[i].device == device) { *out = g_sync_free[i]; g_sync_free.erase(g_sync_free.begin() + i); return true; }
    }
    out->device = device; out->stream = nullptr; out->ev0 = out->ev1 = out->ev_up = nullptr;
    if (cudaStreamCreateWithFlags(&out->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&out->ev0) != cudaSuccess ||
        cudaEventCreate(&out->ev1) != cudaSuccess || cudaEventCreateWithFlags(&out->ev_up, cudaEventDisableTiming) != cudaSuccess) return false;
    return true;
}
static void syncobjs_release(const SyncObjs &o) {
    std::lock_guard<std::mutex> lk(g_buf_mu);
    if (g_sync_free.size() >= 32) {
        if (o.stream) cudaStreamDestroy(o.stream);
        if (o.ev0) cudaEventDestroy(o.ev0);
        if (o.ev1) cudaEventDestroy(o.ev1);
        if (o.ev_up) cudaEventDestroy(o.ev_up);
        return;
    }
    g_sync_free.push_back(o);
}

extern "C" int mrt_gpu_init(int device, MrtDeviceInfo *info) {
    // Everything asked of the driver here is cached per device: cudaGetDeviceProperties costs ~100 ms and the clock-rate attribute
    // 5 - 100 ms PER CALL (measured: it is answered by the GPU, not from a table), and a renderer is created per frame.
    static std::mutex mu;
    static MrtDeviceInfo cached[64];
    static bool have[64];
    static int n_devices = -1;
    std::lock_guard<std::mutex> lock(mu);
    if (n_devices < 0) {
        int n = 0;
        CUDA_TRY(cudaGetDeviceCount(&n));
        n_devices = n;
    }
    if (device < 0 || device >= n_devices || device >= 64) { set_error("mrt_gpu_init: no such CUDA device"); return MRT_E_INVALID; }
    CUDA_TRY(cudaSetDevice(device));
    if (!have[device]) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        MrtDeviceInfo &d = cached[device];
        memset(&d, 0, sizeof(d));
        d.device = device;
        d.sm_count = prop.multiProcessorCount;
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        d.clock_khz = khz;
        d.cc_major = prop.major;
        d.cc_minor = prop.minor;
        d.total_mem = prop.totalGlobalMem;
        strncpy(d.name, prop.name, sizeof(d.name) - 1);
        have[device] = true;
    }
    if (cached[device].cc_major != 10) {
        set_error("mrt_gpu_init: kernels are built for sm_100a only");
        return MRT_E_CUDA;
    }
    if (info) *info = cached[device];
    return MRT_OK;
}

extern "C" void mrt_gpu_destroy(MrtScene *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->rendered) cudaEventSynchronize(s->ev1);   // this scene's own last render (not whatever else was queued on the stream since)
    if (s->poll_stream) cudaStreamSynchronize(s->poll_stream);
    if (s->scene_base) devbuf_release(s->device, s->scene_base, s->scene_bytes);
    if (s->own_acc) devbuf_release(s->device, s->own_acc, s->own_acc_bytes);
    if (s->order_dev) devbuf_release(s->device, s->order_dev, s->order_bytes);
    if (s->pool_dev) pool_release(s->device, s->pool_dev, s->pool_words);
    if (s->stage_dev) pool_release(s->device, s->stage_dev, s->stage_words);
    if (s->final_buf) devbuf_release(s->device, s->final_buf, s->final_bytes);
    if (s->argb_buf) devbuf_release(s->device, s->argb_buf, s->argb_bytes);
    if (s->poll_host && !pinned_slot_release(s->poll_host)) cudaFreeHost(s->poll_host);   // control words live in the scene allocation
    if (s->upload_pinned) pinbuf_release(s->upload_pinned, s->upload_pinned_bytes);   // (the poll stream, which carried the copy, was synchronised above)
    if (s->poll_stream || s->ev0 || s->ev1) syncobjs_release(SyncObjs{s->device, s->poll_stream, s->ev0, s->ev1, s->ev_up});
    delete s;
}

static inline uint32_t bits_of(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// World-space bounds of a flattened object (host tables).  Only feeds the ray classifier of the binned
// renderer, which is a grouping heuristic -- so plain float arithmetic, no parity concerns.
static bool ref_bounds(const MrtSceneDesc *d, uint32_t ref, float lo[3], float hi[3], int depth = 0) {
    if (depth > 64) return false;
    const uint32_t type = MRT_REF_TYPE(ref), idx = MRT_REF_INDEX(ref);
    auto set = [&](const MrtF4 &a, const MrtF4 &b) { lo[0] = a.x; lo[1] = a.y; lo[2] = a.z; hi[0] = b.x; hi[1] = b.y; hi[2] = b.z; };
    auto merge = [&](const float l2[3], const float h2[3], bool first) {
        for (int i = 0; i < 3; i++) { lo[i] = first ? l2[i] : std::fmin(lo[i], l2[i]); hi[i] = first ? h2[i] : std::fmax(hi[i], h2[i]); }
    };
    switch (type) {
    case MRT_T_SPHERE: {
        const MrtF4 &c0 = d->sphere[3 * idx], &c1 = d->sphere[3 * idx + 1];
        const float r = std::fabs(c0.w);
        const bool moving = (bits_of(c1.w) >> 31) != 0;
        for (int i = 0; i < 3; i++) {
            const float a = (&c0.x)[i], b = moving ? (&c1.x)[i] : a;
            lo[i] = std::fmin(a, b) - r; hi[i] = std::fmax(a, b) + r;
        }
        return true;
    }
    case MRT_T_RECT_XY: case MRT_T_RECT_XZ: case MRT_T_RECT_YZ: {
        const MrtF4 &q0 = d->rect[2 * idx], &q1 = d->rect[2 * idx + 1];
        const int ka = type == MRT_T_RECT_XY ? 2 : type == MRT_T_RECT_XZ ? 1 : 0;
        const int aa = type == MRT_T_RECT_YZ ? 1 : 0, ba = type == MRT_T_RECT_XY ? 1 : 2;
        lo[ka] = q1.x - 1e-4f; hi[ka] = q1.x + 1e-4f;
        lo[aa] = q0.x; hi[aa] = q0.y; lo[ba] = q0.z; hi[ba] = q0.w;
        return true;
    }
    case MRT_T_LIST: {
        const MrtF4 &l0 = d->list[2 * idx], &l1 = d->list[2 * idx + 1];
        if (bits_of(l1.w) >> 31) { set(l0, l1); return true; }
        bool any = false;
        for (uint32_t ci = bits_of(l0.w); ci < d->n_child && MRT_REF_TYPE(d->child[ci]) != MRT_T_END; ci++) {
            float l2[3], h2[3];
            if (!ref_bounds(d, d->child[ci], l2, h2, depth + 1)) return false;
            merge(l2, h2, !any);
            any = true;
        }
        return any;
    }
    case MRT_T_BVH: set(d->bvh[2 * idx], d->bvh[2 * idx + 1]); return true;
    case MRT_T_NODE2: {
        float l2[3] = {d->node2[4 * idx].x, d->node2[4 * idx].y, d->node2[4 * idx].z}, h2[3] = {d->node2[4 * idx + 1].x, d->node2[4 * idx + 1].y, d->node2[4 * idx + 1].z};
        merge(l2, h2, true);
        float l3[3] = {d->node2[4 * idx + 2].x, d->node2[4 * idx + 2].y, d->node2[4 * idx + 2].z}, h3[3] = {d->node2[4 * idx + 3].x, d->node2[4 * idx + 3].y, d->node2[4 * idx + 3].z};
        merge(l3, h3, false);
        return true;
    }
    case MRT_T_TRANSLATE: {
        const MrtF4 &x = d->xlate[3 * idx];
        if (!ref_bounds(d, bits_of(x.w), lo, hi, depth + 1)) return false;
        lo[0] += x.x; lo[1] += x.y; lo[2] += x.z; hi[0] += x.x; hi[1] += x.y; hi[2] += x.z;
        return true;
    }
    case MRT_T_ROTATE_Y: {
        const MrtF4 &r0 = d->rot[3 * idx], &r1 = d->rot[3 * idx + 1], &r2 = d->rot[3 * idx + 2];
        if (bits_of(r1.w)) { set(r0, r1); return true; }
        float l2[3], h2[3];
        if (!ref_bounds(d, bits_of(r0.w), l2, h2, depth + 1)) return false;
        bool first = true;
        for (int c = 0; c < 8; c++) {   // object -> world: x' = cos x + sin z, z' = cos z - sin x (scene_object.cpp:86-93)
            const float x = (c & 1) ? h2[0] : l2[0], y = (c & 2) ? h2[1] : l2[1], z = (c & 4) ? h2[2] : l2[2];
            const float pw[3] = {r2.y * x + r2.x * z, y, r2.y * z - r2.x * x};
            merge(pw, pw, first);
            first = false;
        }
        return true;
    }
    case MRT_T_VOLUME: return ref_bounds(d, bits_of(d->vol[idx].x), lo, hi, depth + 1);
    default: return false;
    }
}

// Classifier boxes of the binned renderer: the composite children of the root list (boxes behind transforms,
// volumes, trees).  Up to kMaxClsBoxes of them, largest first.
static void find_classifier_boxes(const MrtSceneDesc *d, MrtScene *s) {
    s->n_cls_boxes = 0;
    if (MRT_REF_TYPE(d->root) != MRT_T_LIST) return;
    struct Cand { float vol; float lo[3], hi[3]; };
    std::vector<Cand> cands;
    const uint32_t first = bits_of(d->list[2 * MRT_REF_INDEX(d->root)].w);
    for (uint32_t ci = first; ci < d->n_child && MRT_REF_TYPE(d->child[ci]) != MRT_T_END; ci++) {
        const uint32_t t = MRT_REF_TYPE(d->child[ci]);
        if (t <= MRT_T_RECT_YZ) continue;
        Cand c;
        if (!ref_bounds(d, d->child[ci], c.lo, c.hi)) continue;
        c.vol = (c.hi[0] - c.lo[0]) * (c.hi[1] - c.lo[1]) * (c.hi[2] - c.lo[2]);
        cands.push_back(c);
    }
    std::stable_sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) { return a.vol > b.vol; });
    for (size_t i = 0; i < cands.size() && s->n_cls_boxes < kMaxClsBoxes; i++) {
        float *b = s->cls_box[s->n_cls_boxes++];
        for (int k = 0; k < 3; k++) { b[k] = cands[i].lo[k]; b[3 + k] = cands[i].hi[k]; }
    }
}

extern "C" int mrt_gpu_scene_upload(const MrtSceneDesc *d, MrtScene **out) {
    if (!d || !out) { set_error("mrt_gpu_scene_upload: null argument"); return MRT_E_INVALID; }
    *out = nullptr;
    {   // a description built by a caller (or read from a file) is checked before anything of it reaches the device
        std::string why;
        if (!validate_scene_desc(*d, &why, nullptr)) { set_error("mrt_gpu_scene_upload: invalid scene description: " + why); return MRT_E_SCENE; }
    }
    MrtScene *s = new (std::nothrow) MrtScene();
    if (!s) { set_error("out of memory"); return MRT_E_INVALID; }
    auto fail = [&](int code) { mrt_gpu_destroy(s); return code; };
    if (cudaGetDevice(&s->device) != cudaSuccess) { set_error("no CUDA device (mrt_gpu_init not called?)"); return fail(MRT_E_CUDA); }
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, s->device);
    SceneView &v = s->view;
    memset(&v, 0, sizeof(v));
    Packer pk;
    pk.add(d->sphere, (size_t) d->n_sphere * 3, &v.sphere);
    pk.add(d->rect, (size_t) d->n_rect * 2, &v.rect);
    pk.add(d->list, (size_t) d->n_list * 2, &v.list);
    pk.add(d->child, (size_t) d->n_child, &v.child);
    pk.add(d->bvh, (size_t) d->n_bvh * 2, &v.bvh);
    pk.add(d->node2, (size_t) d->n_node2 * 4, &v.node2);
    pk.add(d->trileaf, (size_t) d->n_trileaf * 2, &v.trileaf);
    pk.add(d->tri, (size_t) d->n_tri * 3, &v.tri);
    pk.add(d->trin, (size_t) d->n_tri * 3, &v.trin);
    pk.add(d->xlate, (size_t) d->n_xlate * 3, &v.xlate);
    pk.add(d->rot, (size_t) d->n_rot * 3, &v.rot);
    pk.add(d->vol, (size_t) d->n_vol, &v.vol);
    pk.add(d->mat, (size_t) d->n_mat, &v.mat);
    pk.add(d->tex, (size_t) d->n_tex, &v.tex);
    pk.add(d->perlin_vec, d->perlin_vec ? 256 : 0, &v.perlin_vec);
    pk.add(d->perlin_perm, d->perlin_perm ? 768 : 0, &v.perlin_perm);
    pk.add(d->image, (size_t) d->n_image_bytes, &v.image);
    pk.add(d->lights, (size_t) d->n_lights, &v.lights);
    const size_t ctrl_off = pk.reserve(256);   // counters[12] | ticket | max_bits | cancel, zero-initialised
    {
        SyncObjs so;
        const bool ok = syncobjs_acquire(s->device, &so);
        s->poll_stream = so.stream; s->ev0 = so.ev0; s->ev1 = so.ev1; s->ev_up = so.ev_up;
        if (!ok) { cudaGetLastError(); set_error("cudaStreamCreate / cudaEventCreate failed"); return fail(MRT_E_CUDA); }
    }
    {
        // The tables go through a pinned staging buffer and ONE asynchronous copy on the scene's own non-blocking stream; the first
        // render waits for it with an event.  (A plain cudaMemcpy runs on the legacy default stream and so waits for whatever
        // another scene is rendering on it: measured, it serialised the upload of frame k+1 behind the render of frame k.)
        unsigned char *staging = (unsigned char *) pinbuf_acquire(pk.total, &s->upload_pinned_bytes);
        if (!staging) { set_error("cudaHostAlloc upload staging failed"); return fail(MRT_E_CUDA); }
        s->upload_pinned = staging;
        memset(staging, 0, pk.total);
        for (const Packer::Item &it : pk.items) memcpy(staging + it.offset, it.host, it.bytes);
        unsigned char *base = (unsigned char *) devbuf_acquire(s->device, pk.total, &s->scene_bytes);
        if (!base) { set_error(std::string("cudaMalloc scene: ") + cudaGetErrorString(cudaGetLastError())); return fail(MRT_E_CUDA); }
        s->scene_base = base;
        if (cudaMemcpyAsync(base, staging, pk.total, cudaMemcpyHostToDevice, s->poll_stream) != cudaSuccess ||
            cudaEventRecord(s->ev_up, s->poll_stream) != cudaSuccess) { set_error(std::string("cudaMemcpyAsync scene: ") + cudaGetErrorString(cudaGetLastError())); return fail(MRT_E_CUDA); }
        s->upload_pending = true;
        for (const Packer::Item &it : pk.items) *it.dev = base + it.offset;
        s->counters = (unsigned long long *) (base + ctrl_off);          // 12 x 8 bytes
        s->ticket = (unsigned int *) (base + ctrl_off + 128);
        s->max_bits = (unsigned int *) (base + ctrl_off + 160);
        s->cancel_dev = (int *) (base + ctrl_off + 192);
    }
    v.root = d->root;
    v.n_lights = d->n_lights;
    v.sky = d->sky;
    v.cam = d->camera;
    s->stack_words = d->stack_words ? d->stack_words : 64;
    s->stack_words_coop = d->stack_words_coop;   // 0: trees do not qualify for the warp-cooperative traversal
    s->has_trees = d->n_node2 ? 1u : 0u;
    s->n_node2 = d->n_node2;
    s->features = d->features;
    find_classifier_boxes(d, s);

    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess) { set_error(std::string(what) + ": " + cudaGetErrorString(e)); return false; }
        return true;
    };
    {   // pinned words: poll_host[0..1] and the cancel staging word share one 64-byte slot
        void *slot = pinned_slot_acquire();
        if (!slot) {
            if (!cu(cudaHostAlloc(&slot, 64, cudaHostAllocDefault), "cudaHostAlloc control words")) return fail(MRT_E_CUDA);
            memset(slot, 0, 64);
        }
        s->poll_host = (unsigned long long *) slot;
        s->cancel_pinned = (int *) ((unsigned char *) slot + 32);
    }

    *out = s;
    return MRT_OK;
}

// The scene tables and the control block travel on the poll stream (mrt_gpu_scene_upload); whatever touches them on the render
// stream first is ordered after that copy by an event.
static cudaError_t order_after_upload(MrtScene *s) {
    if (!s->upload_pending) return cudaSuccess;
    s->upload_pending = false;
    return cudaStreamWaitEvent(s->stream, s->ev_up, 0);
}

extern "C" int mrt_gpu_set_tuning(MrtScene *s, const MrtTuning *t) {
    if (!s) { set_error("null scene"); return MRT_E_INVALID; }
    MrtTuning z;
    memset(&z, 0, sizeof(z));
    if (t) z = *t;
    if (z.mode > MRT_MODE_BINNED || z.bins > 3u || (z.min_blocks && (z.min_blocks < 5u || z.min_blocks > 8u)) || z.coop_trees > 2u || z.coop_leaf_batch > 32u) {
        set_error("mrt_gpu_set_tuning: value out of range");
        return MRT_E_INVALID;
    }
    s->tuning = z;
    return MRT_OK;
}

extern "C" int mrt_gpu_set_stream(MrtScene *s, void *cuda_stream) {
    if (!s) { set_error("null scene"); return MRT_E_INVALID; }
    s->stream = (cudaStream_t) cuda_stream;
    return MRT_OK;
}

extern "C" int mrt_gpu_bind_accumulator(MrtScene *s, void *device_ptr, uint32_t width, uint32_t height) {
    if (!s) { set_error("null scene"); return MRT_E_INVALID; }
    if (device_ptr && ((uintptr_t) device_ptr & 15u)) { set_error("accumulator must be 16-byte aligned"); return MRT_E_INVALID; }
    s->ext_acc = (float4 *) device_ptr;
    s->ext_w = width;
    s->ext_h = height;
    return MRT_OK;
}

// Work order of the ticket queue: pixels along a Z-curve (Morton order of (x, y)), so that the warps that are
// resident at the same time work on a compact block of the image -- their rays visit the same part of the
// scene, which keeps its nodes in L1/L2 (the reference orders its tiles by an INVERTED Hilbert curve,
// work_queue.cpp:84-127, to spread the threads; on a GPU locality wins).
static int ensure_order(MrtScene *s, uint32_t w, uint32_t h) {
    if (s->order_dev && s->order_w == w && s->order_h == h) return MRT_OK;
    if (s->order_dev) { CUDA_TRY(cudaStreamSynchronize(s->stream)); devbuf_release(s->device, s->order_dev, s->order_bytes); s->order_dev = nullptr; }
    std::vector<uint32_t> order;
    order.reserve((size_t) w * h);
    uint32_t side = 1;
    while (side < w || side < h) side <<= 1;
    auto compact = [](uint64_t v) {   // even bits of v
        v &= 0x5555555555555555ull;
        v = (v | (v >> 1)) & 0x3333333333333333ull;
        v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
        v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
        v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
        v = (v | (v >> 16)) & 0x00000000FFFFFFFFull;
        return (uint32_t) v;
    };
    for (uint64_t d = 0; d < (uint64_t) side * side; d++) {
        uint32_t x = compact(d), y = compact(d >> 1);
        if (x < w && y < h) order.push_back(y * w + x);
    }
    s->order_dev = (uint32_t *) devbuf_acquire(s->device, order.size() * sizeof(uint32_t), &s->order_bytes);
    if (!s->order_dev) { set_error("cudaMalloc pixel order"); return MRT_E_CUDA; }
    CUDA_TRY(cudaMemcpy(s->order_dev, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    s->order_w = w;
    s->order_h = h;
    return MRT_OK;
}

extern "C" int mrt_gpu_render_async(MrtScene *s, const MrtRenderParams *p) {
    if (!s || !p) { set_error("mrt_gpu_render_async: null argument"); return MRT_E_INVALID; }
    if (!p->width || !p->height || !p->samples || p->sample_begin >= p->sample_end || p->sample_end > p->samples) {
        set_error("mrt_gpu_render_async: bad image size or sample range");
        return MRT_E_INVALID;
    }
    uint32_t sq = (uint32_t) sqrtf((float) p->samples);
    while ((uint64_t) (sq + 1) * (sq + 1) <= p->samples) sq++;
    while ((uint64_t) sq * sq > p->samples) sq--;
    if (sq * sq != p->samples) { set_error("mrt_gpu_render_async: samples must be a perfect square (main.cpp:319-320)"); return MRT_E_INVALID; }
    // crop window (all zero = the whole frame): the accumulator has the window's size
    uint32_t cx0 = p->crop_x0, cy0 = p->crop_y0, cx1 = p->crop_x1, cy1 = p->crop_y1;
    if (!(cx0 | cy0 | cx1 | cy1)) { cx1 = p->width; cy1 = p->height; }
    if (cx0 >= cx1 || cy0 >= cy1 || cx1 > p->width || cy1 > p->height) { set_error("mrt_gpu_render_async: bad crop window"); return MRT_E_INVALID; }
    const uint32_t cw = cx1 - cx0, ch = cy1 - cy0;
    const uint64_t n_pixels64 = (uint64_t) cw * ch;
    if (n_pixels64 > 0x7FFFFFFFull) { set_error("image too large"); return MRT_E_INVALID; }
    const uint32_t n_pixels = (uint32_t) n_pixels64;
    CUDA_TRY(cudaSetDevice(s->device));

    float4 *acc = nullptr;
    if (s->ext_acc) {
        if (s->ext_w != cw || s->ext_h != ch) { set_error("bound accumulator has a different size"); return MRT_E_INVALID; }
        acc = s->ext_acc;
    } else {
        if (s->own_acc_pixels != n_pixels) {
            if (s->own_acc) { CUDA_TRY(cudaStreamSynchronize(s->stream)); devbuf_release(s->device, s->own_acc, s->own_acc_bytes); s->own_acc = nullptr; s->own_acc_pixels = 0; }
            s->own_acc = (float4 *) devbuf_acquire(s->device, (size_t) n_pixels * sizeof(float4), &s->own_acc_bytes);
            if (!s->own_acc) { set_error(std::string("cudaMalloc accumulator: ") + cudaGetErrorString(cudaGetLastError())); return MRT_E_CUDA; }
            s->own_acc_pixels = n_pixels;
            CUDA_TRY(cudaMemsetAsync(s->own_acc, 0, (size_t) n_pixels * sizeof(float4), s->stream));
        }
        acc = s->own_acc;
    }

    RenderArgs a;
    a.sc = s->view;
    a.width = p->width; a.height = p->height; a.sqrt_n = sq;
    a.crop_x0 = cx0; a.crop_y0 = cy0; a.crop_w = cw; a.n_pixels = n_pixels;
    a.s_begin = p->sample_begin; a.s_end = p->sample_end; a.max_bounces = p->max_bounces;
    a.seed = p->seed;
    a.accumulate = (p->flags & MRT_RENDER_ACCUMULATE) ? 1u : 0u;
    a.acc = acc;
    a.ticket = s->ticket;
    a.counters = s->counters;
    a.cancel = s->cancel_dev;
    a.order = nullptr;
    const MrtTuning &tn = s->tuning;
    if (tn.z_order) {
        int rc = ensure_order(s, cw, ch);
        if (rc) return rc;
        a.order = s->order_dev;
    }
    // kernel variant: specialised for the scene's feature mask; launch bounds by scene type (measured)
    const uint32_t ns = p->sample_end - p->sample_begin;
    // Mode B (default) takes any sample count that fits its staging array; a parked path keeps its depth (and its
    // count of nested dielectrics, which is bounded by the depth) in 8 bits.  Otherwise mode W for >= 32 samples per
    // launch, else mode P.  MrtTuning.mode forces a mode where it is applicable.
    const bool can_bin = ns <= kMaxStageItems && p->max_bounces <= 255u;
    bool binned = can_bin, mode_w = can_bin || ns >= 32;
    if (tn.mode == MRT_MODE_PER_LANE) { binned = false; mode_w = false; }
    else if (tn.mode == MRT_MODE_PER_WARP) { binned = false; mode_w = true; }
    else if (tn.mode == MRT_MODE_BINNED && !can_bin) { set_error("mrt_gpu_render_async: mode B needs <= 8192 samples per launch and <= 255 bounces"); return MRT_E_INVALID; }
    // measured (profiles/r1_notes.md): tree scenes want 96 registers; list scenes 64 registers / 8 blocks in modes
    // W and P, but 80 registers / 6 blocks in mode B (fewer resident warps thrash the instruction cache less)
    int minb = tn.min_blocks ? (int) tn.min_blocks : (s->has_trees ? 5 : (binned ? 6 : 8));
    const Variant *variant = tn.variant_all ? pick_variant(s->features | MRT_VARIANT_STOCK) : pick_variant(s->features);
    // BVH trees: warp-cooperative traversal (coop_tree.cuh) in mode B where the scene's trees qualify
    const bool coop_ok = binned && s->has_trees && s->stack_words_coop != 0u && (variant->mask & MRT_FEAT_TREES);
    if (tn.coop_trees == 2u && !coop_ok) { set_error("mrt_gpu_render_async: cooperative tree traversal needs mode B and qualifying trees"); return MRT_E_INVALID; }
    // default: only where it pays -- big trees (triangle meshes); the small sphere / box trees of scenes 0, 1, 7 are faster per lane
    // (profiles/r2_notes.md)
    const bool coop = coop_ok && (tn.coop_trees == 2u || (tn.coop_trees == 0u && s->n_node2 >= 1024u));
    const int seq = (binned && ns <= kStageBlock) ? 2 : 0;   // few samples per pixel: the instantiation that sums a pixel's samples sequentially
    const void *kernel = variant->get(coop ? 3 + seq : (binned ? 2 + seq : (mode_w ? 1 : 0)), minb);
    const uint32_t stack_words = coop ? s->stack_words_coop : s->stack_words;
    uint32_t n_bins = 1;
    a.pool = nullptr;
    a.stage = nullptr;
    a.stage_items = 0;
    a.n_cls_boxes = 0;
    a.cls_pending = 0;
    if (binned && tn.bins != 1u) {
        a.n_cls_boxes = s->n_cls_boxes;
        a.cls_pending = (tn.bins == 3u) ? 1u : 0u;
        memcpy(a.cls_box, s->cls_box, sizeof(a.cls_box));
        n_bins = 1u << (a.n_cls_boxes + a.cls_pending);
    }
    a.n_bins = n_bins;
    a.coop_leaf_batch = tn.coop_leaf_batch ? tn.coop_leaf_batch : 32u;
    const uint32_t threads = kBlock;
    const uint32_t warps_per_block = threads / 32u;
    // choose the task size so that every resident warp gets several tasks (load balance) while the idle
    // tail of a task stays small against its body
    uint32_t K = 1;
    BinnedPlan plan = {};
    size_t smem = 0;
    int blocks_per_sm = 0;
    uint32_t resident_warps = 0;
    auto occupancy = [&](uint32_t k) -> int {
        smem = (size_t) warps_per_block * stack_words * 32u * sizeof(uint32_t);
        if (coop) smem += (size_t) warps_per_block * kCoopWords * sizeof(uint32_t);
        if (mode_w && !binned) smem += (size_t) warps_per_block * k * 32u * sizeof(float4);
        if (binned) smem += (size_t) warps_per_block * (n_bins + 1u) * kPoolCap;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            if (e != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)); return MRT_E_CUDA; }
        }
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, (int) threads, smem);
        if (e != cudaSuccess) { set_error(std::string("cudaOccupancyMaxActiveBlocksPerMultiprocessor: ") + cudaGetErrorString(e)); return MRT_E_CUDA; }
        if (blocks_per_sm < 1) { set_error("render kernel does not fit on an SM (traversal stack too deep)"); return MRT_E_CUDA; }
        if (tn.blocks_per_sm && (int) tn.blocks_per_sm < blocks_per_sm) blocks_per_sm = (int) tn.blocks_per_sm;
        resident_warps = (uint32_t) s->sm_count * (uint32_t) blocks_per_sm * warps_per_block;
        return MRT_OK;
    };
    if (binned) {
        // chunk = warp task: 2048 ... 8192 paths whatever the samples per pixel (the sums live in a staging array in global
        // memory, not in shared memory), but at least 10 / 32 tasks per resident warp on small frames, at least 256 paths
        int rc = occupancy(0);
        if (rc) return rc;
        plan = plan_binned_schedule(n_pixels, ns, resident_warps, s->has_trees != 0u, coop, tn, kMaxStageItems);   // schedule.h
        K = plan.K;
    } else if (mode_w) {
        // chunk = pixels per warp task: small, so that the queue holds many short tasks (a task of 8 pixels x
        // 4096 samples keeps a warp busy for ~0.1 s and the warps that finish early idle at the end of the
        // launch: measured 2x slower on a 480x270 frame); the idle lanes at chunk ends cost 1-2 %
        // ... but a chunk should still hold ~4096 items (128 per lane) so that its own idle tail stays small
        uint32_t k_auto = (4096u + ns - 1u) / ns;
        if (k_auto > 8u) k_auto = 8u;
        if (k_auto < 1u) k_auto = 1u;
        K = tn.chunk_pixels ? tn.chunk_pixels : k_auto;
        for (;;) {
            int rc = occupancy(K);
            if (rc) return rc;
            if (K == 1 || tn.chunk_pixels) break;
            if (blocks_per_sm >= minb && n_pixels / K >= 32u * resident_warps) break;
            K >>= 1;
        }
    } else {
        int rc = occupancy(0);
        if (rc) return rc;
        K = tn.chunk_pixels ? tn.chunk_pixels : n_pixels / (8u * resident_warps);
        K = (K + 31u) & ~31u;
        if (K < 32u) K = 32u;
        if (K > 1024u) K = 1024u;
    }
    a.stack_words = stack_words;
    a.pixels_per_task = K;
    a.n_tasks = (n_pixels + K - 1) / K;
    if (binned) {
        for (int r = 0; r < 3; r++) { a.sched_task0[r] = plan.task0[r]; a.sched_pix0[r] = plan.pix0[r]; a.sched_k[r] = plan.k[r]; }
        a.sched_pix0[3] = plan.pix0[3];
        a.n_tasks = plan.n_tasks;
    }
    uint32_t grid = (uint32_t) s->sm_count * (uint32_t) blocks_per_sm;
    const uint32_t blocks_needed = (a.n_tasks + warps_per_block - 1) / warps_per_block;
    if (grid > blocks_needed) grid = blocks_needed;

    if (binned) {
        const size_t words = (size_t) grid * warps_per_block * kPoolCap * kStateWords;
        if (words > s->pool_words) {
            if (s->pool_dev) { CUDA_TRY(cudaStreamSynchronize(s->stream)); pool_release(s->device, s->pool_dev, s->pool_words); s->pool_dev = nullptr; s->pool_words = 0; }
            s->pool_dev = pool_acquire(s->device, words, &s->pool_words);
            if (!s->pool_dev) { set_error(std::string("cudaMalloc path pool: ") + cudaGetErrorString(cudaGetLastError())); return MRT_E_CUDA; }
        }
        a.pool = s->pool_dev;
        a.stage_items = K * ns;
        const size_t swords = (size_t) grid * warps_per_block * a.stage_items * 4u;
        if (swords > s->stage_words) {
            if (s->stage_dev) { CUDA_TRY(cudaStreamSynchronize(s->stream)); pool_release(s->device, s->stage_dev, s->stage_words); s->stage_dev = nullptr; s->stage_words = 0; }
            s->stage_dev = pool_acquire(s->device, swords, &s->stage_words);
            if (!s->stage_dev) { set_error(std::string("cudaMalloc sample staging: ") + cudaGetErrorString(cudaGetLastError())); return MRT_E_CUDA; }
        }
        a.stage = reinterpret_cast<float4 *>(s->stage_dev);
    }
    CUDA_TRY(order_after_upload(s));
    CUDA_TRY(cudaMemsetAsync(s->cancel_dev, 0, sizeof(int), s->stream));
    CUDA_TRY(cudaMemsetAsync(s->ticket, 0, sizeof(unsigned int), s->stream));
    s->final_is_running_mean = false;
    const bool cont = (p->flags & MRT_RENDER_CONTINUE) && s->rendered;
    if (!cont) {
        CUDA_TRY(cudaMemsetAsync(s->counters, 0, 12 * sizeof(unsigned long long), s->stream));
        CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
        s->stat_samples = 0;
    }
    s->stat_samples += ns;
    void *kargs[] = {(void *) &a};
    CUDA_TRY(cudaLaunchKernel(kernel, dim3(grid), dim3(threads), kargs, smem, s->stream));
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    s->rendered = true;
    s->last = *p;
    s->last_w = cw; s->last_h = ch;
    s->last_tasks = a.n_tasks;
    s->last_grid = grid;
    s->last_block = threads;
    s->last_smem = (uint32_t) smem;
    s->last_coop = coop;
    s->last_mode = binned ? MRT_MODE_BINNED : (mode_w ? MRT_MODE_PER_WARP : MRT_MODE_PER_LANE);
    s->last_acc = acc;
    return MRT_OK;
}

extern "C" int mrt_gpu_poll(MrtScene *s, float *pct_done, uint64_t *rays) {
    if (!s) { set_error("null scene"); return MRT_E_INVALID; }
    if (!s->rendered) { if (pct_done) *pct_done = 0; if (rays) *rays = 0; return MRT_OK; }
    CUDA_TRY(cudaSetDevice(s->device));
    if (cudaEventQuery(s->ev1) == cudaSuccess) {
        if (pct_done) *pct_done = 100.0f;
        if (rays) {
            CUDA_TRY(cudaMemcpyAsync(&s->poll_host[1], &s->counters[0], sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->poll_stream));
            CUDA_TRY(cudaStreamSynchronize(s->poll_stream));
            *rays = s->poll_host[1];
        }
        return MRT_OK;
    }
    s->poll_host[0] = 0;
    CUDA_TRY(cudaMemcpyAsync(&s->poll_host[0], s->ticket, sizeof(unsigned int), cudaMemcpyDeviceToHost, s->poll_stream));
    CUDA_TRY(cudaMemcpyAsync(&s->poll_host[1], &s->counters[0], sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->poll_stream));
    CUDA_TRY(cudaStreamSynchronize(s->poll_stream));
    // like work_queue_seq::getPercentDone (work_queue.cpp:142-149): tickets are taken at the start of work
    double done = (double) (unsigned int) s->poll_host[0] - (double) s->last_grid * (s->last_block / 32u);
    if (done < 0) done = 0;
    double pct = s->last_tasks ? done * 100.0 / s->last_tasks : 0.0;
    if (pct > 99.9) pct = 99.9;
    if (pct_done) *pct_done = (float) pct;
    if (rays) *rays = s->poll_host[1];   // warps add their counts when they retire
    return MRT_OK;
}

extern "C" int mrt_gpu_wait(MrtScene *s) {
    if (!s) { set_error("null scene"); return MRT_E_INVALID; }
    if (!s->rendered) return MRT_OK;
    CUDA_TRY(cudaSetDevice(s->device));
    CUDA_TRY(cudaEventSynchronize(s->ev1));
    return MRT_OK;
}

extern "C" int mrt_gpu_stats(MrtScene *s, MrtRenderStats *out) {
    if (!s || !out) { set_error("mrt_gpu_stats: null argument"); return MRT_E_INVALID; }
    if (!s->rendered) { set_error("mrt_gpu_stats: nothing rendered yet"); return MRT_E_STATE; }
    int rc = mrt_gpu_wait(s);
    if (rc) return rc;
    unsigned long long c[12];
    CUDA_TRY(cudaMemcpy(c, s->counters, sizeof(c), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof(*out));
    out->rays = c[0];
    out->paths = (uint64_t) s->last_w * s->last_h * s->stat_samples;
    out->nonfinite = c[2];
    out->warp_iterations = c[1];
    CUDA_TRY(cudaEventElapsedTime(&out->kernel_ms, s->ev0, s->ev1));
    out->grid = s->last_grid;
    out->block = s->last_block;
    out->smem_bytes = s->last_smem;
    out->mode = s->last_mode;
    out->coop_trees = s->last_coop ? 1u : 0u;
    out->coop_node_steps = c[4]; out->coop_node_items = c[5]; out->coop_leaf_steps = c[6]; out->coop_leaf_items = c[7];
    if (c[8]) {   // mode B: ~first entry | last exit | sum of (exit - entry) | ~first exit
        const unsigned long long t_first = ~c[8];
        out->warp_span_ns = c[9] - t_first;
        out->warp_time_sum_ns = c[10];
        out->first_exit_ns = ~c[11] - t_first;
        out->warps = s->last_grid * (s->last_block / 32u);
        out->stage_sum_ns = c[3];
    }
    return MRT_OK;
}

extern "C" int mrt_gpu_finalize_device(MrtScene *s, const void *acc_dev, void *out_dev, uint32_t width, uint32_t height, float max_luminance) {
    if (!s || !acc_dev || !out_dev) { set_error("mrt_gpu_finalize_device: null argument"); return MRT_E_INVALID; }
    CUDA_TRY(cudaSetDevice(s->device));
    uint32_t n = width * height;
    finalize_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>((const float4 *) acc_dev, (float4 *) out_dev, n, max_luminance);
    CUDA_TRY(cudaGetLastError());
    return MRT_OK;
}

static int ensure_final(MrtScene *s, size_t n) {
    if (s->final_pixels != n) {
        if (s->final_buf || s->argb_buf) CUDA_TRY(cudaStreamSynchronize(s->stream));
        if (s->final_buf) { devbuf_release(s->device, s->final_buf, s->final_bytes); s->final_buf = nullptr; }
        if (s->argb_buf) { devbuf_release(s->device, s->argb_buf, s->argb_bytes); s->argb_buf = nullptr; }
        s->final_pixels = 0;
        s->final_buf = (float4 *) devbuf_acquire(s->device, n * sizeof(float4), &s->final_bytes);
        s->argb_buf = (uint32_t *) devbuf_acquire(s->device, n * sizeof(uint32_t), &s->argb_bytes);
        if (!s->final_buf || !s->argb_buf) { set_error(std::string("cudaMalloc image buffers: ") + cudaGetErrorString(cudaGetLastError())); return MRT_E_CUDA; }
        s->final_pixels = n;
    }
    return MRT_OK;
}

extern "C" int mrt_gpu_readback(MrtScene *s, float *rgba_host, int finalize) {
    if (!s || !rgba_host) { set_error("mrt_gpu_readback: null argument"); return MRT_E_INVALID; }
    if (!s->rendered) { set_error("mrt_gpu_readback: nothing rendered yet"); return MRT_E_STATE; }
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t n = (size_t) s->last_w * s->last_h;
    const float4 *src = s->last_acc;
    if (finalize) {
        if (!s->final_is_running_mean) {
            int rc = ensure_final(s, n);
            if (rc) return rc;
            rc = mrt_gpu_finalize_device(s, s->last_acc, s->final_buf, s->last_w, s->last_h, s->last.max_luminance);
            if (rc) return rc;
        }
        src = s->final_buf;
    }
    CUDA_TRY(cudaMemcpyAsync(rgba_host, src, n * sizeof(float4), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return MRT_OK;
}

extern "C" int mrt_gpu_running_mean_update(MrtScene *s, const void *acc_dev, void *mean_dev, uint32_t width, uint32_t height, uint32_t pass,
                                           float max_luminance) {
    if (!s || !acc_dev || !mean_dev) { set_error("mrt_gpu_running_mean_update: null argument"); return MRT_E_INVALID; }
    CUDA_TRY(cudaSetDevice(s->device));
    const uint32_t n = width * height;
    running_mean_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>((const float4 *) acc_dev, (float4 *) mean_dev, n, pass, max_luminance);
    CUDA_TRY(cudaGetLastError());
    return MRT_OK;
}

// draw2 as a whole (main.cpp:193-243 with work_queue_dynamic, work_queue.cpp:158-166): `passes` one-sample launches
// (samples sample_begin .. sample_begin + passes - 1 of the grid), each followed by the running-mean update; the mean is
// left in the scene's image buffer and copied to rgba_host (w = number of passes).
extern "C" int mrt_gpu_render_running_mean(MrtScene *s, const MrtRenderParams *p, float *rgba_host) {
    if (!s || !p) { set_error("mrt_gpu_render_running_mean: null argument"); return MRT_E_INVALID; }
    if (s->ext_acc) { set_error("mrt_gpu_render_running_mean: unbind the external accumulator first"); return MRT_E_STATE; }
    uint32_t cx0 = p->crop_x0, cy0 = p->crop_y0, cx1 = p->crop_x1, cy1 = p->crop_y1;
    if (!(cx0 | cy0 | cx1 | cy1)) { cx1 = p->width; cy1 = p->height; }
    if (cx0 >= cx1 || cy0 >= cy1 || p->sample_begin >= p->sample_end) { set_error("mrt_gpu_render_running_mean: bad window or sample range"); return MRT_E_INVALID; }
    const uint32_t w = cx1 - cx0, h = cy1 - cy0;
    for (uint32_t pass = 0; p->sample_begin + pass < p->sample_end; pass++) {
        MrtRenderParams one = *p;
        one.sample_begin = p->sample_begin + pass;
        one.sample_end = one.sample_begin + 1u;
        one.flags = (p->flags & ~MRT_RENDER_ACCUMULATE) | (pass ? MRT_RENDER_CONTINUE : 0u);   // statistics run over all passes
        int rc = mrt_gpu_render_async(s, &one);
        if (rc) return rc;
        if (pass == 0) { rc = ensure_final(s, (size_t) w * h); if (rc) return rc; }
        rc = mrt_gpu_running_mean_update(s, s->last_acc, s->final_buf, w, h, pass, p->max_luminance);
        if (rc) return rc;
    }
    s->final_is_running_mean = true;   // readback(finalize) / tonemap now deliver this image (until the next plain render)
    if (rgba_host) {
        CUDA_TRY(cudaMemcpyAsync(rgba_host, s->final_buf, (size_t) w * h * sizeof(float4), cudaMemcpyDeviceToHost, s->stream));
        CUDA_TRY(cudaStreamSynchronize(s->stream));
    }
    return MRT_OK;
}

extern "C" int mrt_gpu_tonemap_device(MrtScene *s, const void *img_dev, void *argb_dev, uint32_t width, uint32_t height) {
    if (!s || !img_dev || !argb_dev) { set_error("mrt_gpu_tonemap_device: null argument"); return MRT_E_INVALID; }
    CUDA_TRY(cudaSetDevice(s->device));
    const uint32_t n = width * height;
    CUDA_TRY(order_after_upload(s));
    CUDA_TRY(cudaMemsetAsync(s->max_bits, 0, sizeof(unsigned int), s->stream));
    max_luminance_kernel<<<s->sm_count * 4, 256, 0, s->stream>>>((const float4 *) img_dev, n, s->max_bits);
    tonemap_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>((const float4 *) img_dev, (uint32_t *) argb_dev, n, s->max_bits);
    CUDA_TRY(cudaGetLastError());
    return MRT_OK;
}

extern "C" int mrt_gpu_tonemap(MrtScene *s, uint32_t *argb_host) {
    if (!s || !argb_host) { set_error("mrt_gpu_tonemap: null argument"); return MRT_E_INVALID; }
    if (!s->rendered) { set_error("mrt_gpu_tonemap: nothing rendered yet"); return MRT_E_STATE; }
    CUDA_TRY(cudaSetDevice(s->device));
    const size_t n = (size_t) s->last_w * s->last_h;
    int rc = MRT_OK;
    if (!s->final_is_running_mean) {
        rc = ensure_final(s, n);
        if (rc) return rc;
        rc = mrt_gpu_finalize_device(s, s->last_acc, s->final_buf, s->last_w, s->last_h, s->last.max_luminance);
        if (rc) return rc;
    }
    rc = mrt_gpu_tonemap_device(s, s->final_buf, s->argb_buf, s->last_w, s->last_h);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(argb_host, s->argb_buf, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return MRT_OK;
}

// Samples-per-pixel sharding inside ONE process (the C++ host, main_host.cpp -gpus N): sum the accumulators of n scenes that
// rendered the same frame on n GPUs, finalise, and leave the image on scenes[0]'s GPU -- see reduce_finalize_kernel.
extern "C" int mrt_gpu_reduce_finalize(MrtScene **scenes, int n, float max_luminance, float *rgba_host, uint32_t *argb_host) {
    if (!scenes || n < 1 || n > kMaxReduceGpus) { set_error("mrt_gpu_reduce_finalize: need 1..16 scenes"); return MRT_E_INVALID; }
    for (int k = 0; k < n; k++) {
        if (!scenes[k] || !scenes[k]->rendered) { set_error("mrt_gpu_reduce_finalize: every scene must have rendered"); return MRT_E_STATE; }
        if (scenes[k]->last_w != scenes[0]->last_w || scenes[k]->last_h != scenes[0]->last_h) { set_error("mrt_gpu_reduce_finalize: frame sizes differ"); return MRT_E_INVALID; }
    }
    MrtScene *root = scenes[0];
    const uint32_t W = root->last_w, H = root->last_h, P = W * H;
    // peer access between every pair of distinct devices (NVLink / NVSwitch on a B200 node)
    for (int a = 0; a < n; a++)
        for (int b = 0; b < n; b++) {
            const int da = scenes[a]->device, db = scenes[b]->device;
            if (da == db) continue;
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, da, db));
            if (!can) { set_error("mrt_gpu_reduce_finalize: no peer access between the GPUs (NVLink / P2P required)"); return MRT_E_CUDA; }
            CUDA_TRY(cudaSetDevice(da));
            cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return MRT_E_CUDA; }
        }
    CUDA_TRY(cudaSetDevice(root->device));
    int rc = ensure_final(root, P);
    if (rc) return rc;
    ReduceArgs ra;
    memset(&ra, 0, sizeof(ra));
    ra.n = n;
    for (int k = 0; k < n; k++) ra.acc[k] = scenes[k]->last_acc;
    std::vector<cudaEvent_t> done(n, nullptr);
    auto cleanup = [&]() { for (cudaEvent_t e : done) if (e) cudaEventDestroy(e); };
    for (int g = 0; g < n; g++) {
        MrtScene *sg = scenes[g];
        const uint32_t begin = (uint32_t) ((uint64_t) P * g / n), end = (uint32_t) ((uint64_t) P * (g + 1) / n);
        if (cudaSetDevice(sg->device) != cudaSuccess) { cleanup(); set_error("cudaSetDevice"); return MRT_E_CUDA; }
        for (int k = 0; k < n; k++)   // the stripe kernel reads every accumulator: wait for every render
            if (cudaStreamWaitEvent(sg->stream, scenes[k]->ev1, 0) != cudaSuccess) { cleanup(); set_error("cudaStreamWaitEvent"); return MRT_E_CUDA; }
        if (end > begin) reduce_finalize_kernel<<<(end - begin + 255) / 256, 256, 0, sg->stream>>>(ra, root->final_buf, begin, end, max_luminance);
        if (cudaGetLastError() != cudaSuccess || cudaEventCreateWithFlags(&done[g], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventRecord(done[g], sg->stream) != cudaSuccess) { cleanup(); set_error("mrt_gpu_reduce_finalize: launch failed"); return MRT_E_CUDA; }
    }
    CUDA_TRY(cudaSetDevice(root->device));
    for (int g = 0; g < n; g++) CUDA_TRY(cudaStreamWaitEvent(root->stream, done[g], 0));
    if (argb_host) {
        rc = mrt_gpu_tonemap_device(root, root->final_buf, root->argb_buf, W, H);
        if (rc) { cleanup(); return rc; }
        CUDA_TRY(cudaMemcpyAsync(argb_host, root->argb_buf, (size_t) P * sizeof(uint32_t), cudaMemcpyDeviceToHost, root->stream));
    }
    if (rgba_host) CUDA_TRY(cudaMemcpyAsync(rgba_host, root->final_buf, (size_t) P * sizeof(float4), cudaMemcpyDeviceToHost, root->stream));
    CUDA_TRY(cudaStreamSynchronize(root->stream));
    cleanup();
    return MRT_OK;
}

extern "C" int mrt_gpu_cancel(MrtScene *s) {
    if (!s) { set_error("null scene"); return MRT_E_INVALID; }
    if (s->cancel_dev && s->rendered) {
        CUDA_TRY(cudaSetDevice(s->device));
        *s->cancel_pinned = 1;
        CUDA_TRY(cudaMemcpyAsync(s->cancel_dev, s->cancel_pinned, sizeof(int), cudaMemcpyHostToDevice, s->poll_stream));
        CUDA_TRY(cudaStreamSynchronize(s->poll_stream));
    }
    return MRT_OK;
}
