// Host-side plan of the mode-B ticket queue (render_pixel_binned): how many pixels a warp task (chunk) holds and how the
// tasks tile the frame.  Plain C++ so that the test harness (tests/host_emul/emul_binned.cpp) runs the same code on the CPU.
#pragma once
#include <cstdint>

#include "mrt_gpu.h"

namespace mrt {

struct BinnedPlan {
    uint32_t K;            // pixels per big chunk
    uint32_t n_tasks;      // tickets
    // three runs of the queue: run r starts at task task0[r] / pixel pix0[r] and uses chunks of k[r] pixels; pix0[3] = n_pixels
    uint32_t task0[3], pix0[4], k[3];
};

// pixels [pix0, pix0 + kp) of ticket `task` -- the kernel's own mapping (render_pixel_binned), restated for the checker
inline void plan_task_pixels(const BinnedPlan &p, uint32_t task, uint32_t *pix0, uint32_t *kp) {
    const int r = task >= p.task0[2] ? 2 : (task >= p.task0[1] ? 1 : 0);
    *pix0 = p.pix0[r] + (task - p.task0[r]) * p.k[r];
    const uint32_t end = p.pix0[r + 1];
    *kp = p.k[r] < end - *pix0 ? p.k[r] : end - *pix0;
}

// n_pixels x ns paths on resident_warps persistent warps.  max_items = capacity of a warp's staging array.
inline BinnedPlan plan_binned_schedule(uint32_t n_pixels, uint32_t ns, uint32_t resident_warps, bool has_trees, bool coop,
                                       const MrtTuning &tn, uint32_t max_items) {
    BinnedPlan p;
    const uint64_t total = (uint64_t) n_pixels * ns;
    // >= 10 big tasks per resident warp for list scenes (chunks cost about the same; the end of the launch is balanced by the
    // small last chunks, below), >= 32 for scenes with trees, whose chunks differ several-fold in cost (mesh vs background
    // pixels: measured 176 -> 249 ms on a 960x540x256 frame of scene 7 with 10), >= 64 for the per-lane tree scenes (15-fold:
    // 178 -> 156 ms on that frame, warps at work 78 % -> 95 % of the launch, although fewer lanes hold a path: 93 % -> 88 %)
    uint64_t target = total / ((uint64_t) (resident_warps ? resident_warps : 1u) * (has_trees ? (coop ? 32u : 64u) : 10u));
    // big chunks: 4096 paths for the triangle meshes; 2048 for the per-lane tree scenes, whose chunks differ 15-fold in cost
    // (pixels on the glass / fog of scene 7): a heavy 4096-path chunk handed out late outlasts the whole guided tail (warps at
    // work 94 % -> 99.9 % of a 512-sample 1080p slice, 1111 -> 1065 ms; the triangle meshes lose 0.7 % with 2048); list scenes: as
    // big as the staging array allows (8192: lanes with a path 98.2 -> 99.1 %, C2 313.2 -> 311.5 ms).  profiles/r2_notes.md
    const uint64_t cap = tn.chunk_paths ? tn.chunk_paths : (has_trees ? (coop ? 4096u : 2048u) : max_items);
    if (target > cap) target = cap;
    if (target < 256u) target = 256u;
    uint32_t K = (uint32_t) (target / ns);
    if (tn.chunk_pixels) K = tn.chunk_pixels;
    if (K < 1u) K = 1u;
    if ((uint64_t) K * ns > max_items) K = max_items / ns;
    if (K > n_pixels) K = n_pixels;
    if (K < 1u) K = 1u;
    p.K = K;
    // Guided self-scheduling: big chunks first, the last stretch of the frame -- about two big chunks per resident warp --
    // in chunks of a quarter and then a sixteenth of that size, so that at the end of the launch a warp waits for a
    // small chunk, not a big one (with 18 big chunks per warp, the 1/8 slice of an 8-GPU run, the idle end was 3-5 %
    // of the launch; profiles/r2_notes.md).  MrtTuning.chunk_pixels = uniform chunks of that size.
    const uint32_t k1 = K / 4u ? K / 4u : 1u, k2 = K / 16u ? K / 16u : 1u;
    uint64_t tail = tn.chunk_pixels ? 0u : (uint64_t) (tn.tail_tasks ? tn.tail_tasks : 2u) * resident_warps * K;   // pixels handed out in small chunks
    if (tail > n_pixels / 4u) tail = n_pixels / 4u;
    if (k1 == K) tail = 0;
    const uint32_t p1 = n_pixels - (uint32_t) tail;                                 // run 0: [0, p1) in chunks of K
    const uint32_t p2 = (k2 == k1) ? n_pixels : p1 + (uint32_t) (tail * 2u / 3u);   // run 1: [p1, p2) in chunks of k1; run 2: the rest
    const uint32_t t1 = (p1 + K - 1u) / K, t2 = t1 + (p2 - p1 + k1 - 1u) / k1;
    p.task0[0] = 0; p.task0[1] = t1; p.task0[2] = t2;
    p.pix0[0] = 0; p.pix0[1] = p1; p.pix0[2] = p2; p.pix0[3] = n_pixels;
    p.k[0] = K; p.k[1] = k1; p.k[2] = k2;
    p.n_tasks = t2 + (n_pixels - p2 + k2 - 1u) / k2;
    return p;
}

}  // namespace mrt
