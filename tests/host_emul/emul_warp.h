// TEST-ONLY SIMT shim: lets g++ compile render_kernels.cuh and run a render kernel on the CPU so that the
// warp-level scheduling logic (tickets, bins, path pool, sample staging) can be checked against the oracle without
// a GPU.  Every lane is a fiber (ucontext); the lanes of a warp run one after the other and switch at every warp
// collective (__shfl_sync, __ballot_sync, __syncwarp), which is exactly the lockstep the kernels rely on at those
// points.  One block at a time (the dynamic shared memory is one global array).  Nothing here ships.
#pragma once
#define MRT_EMUL_WARP 1
#include <cuda_runtime.h>   // vector types only; the CUDA function qualifiers are neutralised below
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __noinline__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __shared__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)

namespace emul {
struct Lane {
    ucontext_t ctx;
    std::vector<char> stack;
    uint32_t tid = 0, bid = 0;
    uint64_t seq = 0;      // collectives executed so far
    bool done = false;
};
struct Warp {
    uint64_t slots[2][32];
    Lane lanes[32];
};
inline ucontext_t g_sched;
inline Lane *g_lane = nullptr;
inline Warp *g_warp = nullptr;
inline unsigned long long g_collectives = 0;

// publish v, wait until every lane of the warp has published, return the buffer with all 32 values
inline const uint64_t *exchange(uint64_t v) {
    Lane *l = g_lane;
    Warp *w = g_warp;
    const uint64_t *buf = w->slots[l->seq & 1u];
    w->slots[l->seq & 1u][l->tid & 31u] = v;
    l->seq++;
    g_collectives++;
    swapcontext(&l->ctx, &g_sched);   // resumed by the scheduler after all lanes have arrived
    g_lane = l; g_warp = w;
    return buf;
}
struct Tid { uint32_t x, y, z; };
inline Tid tid() { return Tid{g_lane->tid, 0, 0}; }
inline Tid bid() { return Tid{g_lane->bid, 0, 0}; }

// run `fn(arg)` for one block of `threads` threads (multiple of 32)
template <typename F>
struct Launch { F *fn; };
inline void (*g_entry)(void *) = nullptr;
inline void *g_entry_arg = nullptr;
inline void trampoline() {
    g_entry(g_entry_arg);
    g_lane->done = true;
    swapcontext(&g_lane->ctx, &g_sched);
}
inline void run_block(void (*entry)(void *), void *arg, uint32_t threads, uint32_t block_id) {
    const uint32_t n_warps = threads / 32u;
    std::vector<Warp> *warps = new std::vector<Warp>(n_warps);
    g_entry = entry; g_entry_arg = arg;
    for (uint32_t w = 0; w < n_warps; w++)
        for (uint32_t l = 0; l < 32u; l++) {
            Lane &ln = (*warps)[w].lanes[l];
            ln.tid = w * 32u + l; ln.bid = block_id;
            ln.stack.resize(512 * 1024);
            getcontext(&ln.ctx);
            ln.ctx.uc_stack.ss_sp = ln.stack.data();
            ln.ctx.uc_stack.ss_size = ln.stack.size();
            ln.ctx.uc_link = nullptr;
            makecontext(&ln.ctx, trampoline, 0);
        }
    for (bool any = true; any;) {   // one round = every live lane runs to its next collective (or to the end)
        any = false;
        for (uint32_t w = 0; w < n_warps; w++) {
            uint64_t seq0 = ~0ull;
            uint32_t live = 0, finished = 0;
            for (uint32_t l = 0; l < 32u; l++) {
                Lane &ln = (*warps)[w].lanes[l];
                if (ln.done) { finished++; continue; }
                g_lane = &ln; g_warp = &(*warps)[w];
                swapcontext(&g_sched, &ln.ctx);
                if (ln.done) { finished++; continue; }
                live++;
                if (seq0 == ~0ull) seq0 = ln.seq;
                else if (seq0 != ln.seq) { fprintf(stderr, "emul: lanes of warp %u diverged at a collective\n", w); abort(); }
            }
            if (live && finished) { fprintf(stderr, "emul: some lanes of warp %u exited while others wait at a collective\n", w); abort(); }
            any = any || live;
        }
    }
    delete warps;
}
}  // namespace emul

#define threadIdx (emul::tid())
#define blockIdx (emul::bid())

namespace mrt {
inline uint32_t smem_stack[64 * 1024];   // the kernels' `extern __shared__ uint32_t smem_stack[]`
template <typename T> inline T emul_bits_to(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <typename T> inline uint64_t emul_bits_of(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T __shfl_sync(unsigned, T v, unsigned src) { return emul_bits_to<T>(emul::exchange(emul_bits_of(v))[src & 31u]); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const uint32_t lane = emul::g_lane->tid & 31u;
    const uint64_t *b = emul::exchange(emul_bits_of(v));
    return lane >= delta ? emul_bits_to<T>(b[lane - delta]) : v;
}
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask) {
    const uint32_t lane = emul::g_lane->tid & 31u;
    return emul_bits_to<T>(emul::exchange(emul_bits_of(v))[(lane ^ (uint32_t) mask) & 31u]);
}
inline unsigned __ballot_sync(unsigned, bool pred) {
    const uint64_t *b = emul::exchange(pred ? 1u : 0u);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (b[i] ? 1u : 0u) << i;
    return m;
}
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }
inline void __syncwarp() { emul::exchange(0); }
inline void __threadfence_block() {}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline unsigned atomicMin(unsigned *p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline void __stcg(T *p, T v) { *p = v; }
template <typename T> inline void __stcs(T *p, T v) { *p = v; }
inline float __fdividef(float a, float b) { return a / b; }
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
}  // namespace mrt
