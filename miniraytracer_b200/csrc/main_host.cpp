// Command-line front end: the reference's main() (main.cpp:283-498) with the
// worker threads replaced by the GPU renderer behind the C ABI.  Same options
// (cmdline_parser.cpp:90-104) plus -gpus/-seed/-assets/-out; headless (the SDL
// preview of platform_linux.cpp is optional in the reference's design and SDL2
// is not available here), so the "window title" statistics go to stdout and the
// image goes to a file.  With -gpus N the samples per pixel are split across N
// devices of this process; the accumulators are summed, finalised and tone-mapped
// on the GPUs over NVLink peer memory (mrt_gpu_reduce_finalize) -- the
// torch.distributed/NCCL path of bench.py is the multi-process equivalent.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "mrt_gpu.h"

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int write_ppm(const char *path, const uint32_t *argb, uint32_t w, uint32_t h) {
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    fprintf(f, "P6\n%u %u\n255\n", w, h);
    for (uint32_t y = 0; y < h; y++) {
        const uint32_t *row = argb + (size_t) (h - 1 - y) * w;   // buffer is y-up (platform_linux.cpp:84 flips at display)
        for (uint32_t x = 0; x < w; x++) {
            unsigned char px[3] = {(unsigned char) (row[x] >> 16), (unsigned char) (row[x] >> 8), (unsigned char) row[x]};
            fwrite(px, 1, 3, f);
        }
    }
    fclose(f);
    return 0;
}

static int write_pfm(const char *path, const float *rgba, uint32_t w, uint32_t h) {
    FILE *f = fopen(path, "wb");
    if (!f) return 1;
    fprintf(f, "PF\n%u %u\n-1.0\n", w, h);   // little endian, bottom-to-top rows = our y-up order
    for (size_t i = 0; i < (size_t) w * h; i++) fwrite(rgba + i * 4, sizeof(float), 3, f);
    fclose(f);
    return 0;
}

int main(int argc, char **argv) {
    MrtParams p;
    if (mrt_params_parse(argc, argv, &p)) return 0;
    const uint32_t W = p.buffer_width, H = p.buffer_height;

    double t0 = now_s();
    MrtHostScene *hs = nullptr;
    if (mrt_scene_create(p.scene_select, float(W) / float(H), p.asset_dir, &hs)) {
        fprintf(stderr, "error: %s\n", mrt_last_error());
        return 1;
    }
    printf("MiniRayTracer - Scene: %.0fms\n", 1000.0 * (now_s() - t0));

    uint32_t sq = (uint32_t) sqrtf((float) p.samples_per_pixel);   // main.cpp:319-320
    uint32_t N = sq * sq;
    uint32_t G = p.num_gpus;
    if (G > N) G = N;

    std::vector<MrtScene *> scenes(G, nullptr);
    MrtDeviceInfo info;
    for (uint32_t g = 0; g < G; g++) {
        if (mrt_gpu_init((int) g, &info) || mrt_gpu_scene_upload(mrt_scene_desc(hs), &scenes[g])) {
            fprintf(stderr, "error: %s\n", mrt_last_error());
            return 1;
        }
    }
    double t1 = now_s();
    // -mode 0 (work_queue_seq, main.cpp:347-349): all samples of a GPU's slice in one launch.
    // -mode 1 (work_queue_dynamic + draw2, main.cpp:350-354,193-243): sample-major passes, the image refines
    // progressively; here each pass is one launch that ACCUMULATES a slice of the samples (sum + count, so the
    // running mean of draw2 is the finalised accumulator after every pass).
    // With one GPU -mode 1 is draw2 itself: one-sample passes and the running mean with its per-pass luminance clamp
    // (mrt_gpu_render_running_mean).  With several GPUs the samples are sharded, which a clamped running mean over ALL
    // samples cannot be; there the passes accumulate sums (identical unless a running mean crosses -maxlum on the way).
    const bool running_mean = (p.threading_mode == 1) && G == 1;
    uint32_t passes = (p.threading_mode == 1) ? 8u : 1u;
    if (passes > N / G) passes = (N / G) ? (N / G) : 1u;   // at least one sample per pass and GPU
    uint64_t rays = 0, paths = 0, dropped = 0;
    float kernel_ms = 0;
    auto collect = [&]() -> int {   // statistics of the launches in flight (blocks until they have finished)
        float ms_max = 0;
        for (uint32_t g = 0; g < G; g++) {
            MrtRenderStats st;
            mrt_gpu_init((int) g, nullptr);
            if (mrt_gpu_stats(scenes[g], &st)) { fprintf(stderr, "error: %s\n", mrt_last_error()); return 1; }
            rays += st.rays; paths += st.paths; dropped += st.nonfinite;
            if (st.kernel_ms > ms_max) ms_max = st.kernel_ms;
        }
        kernel_ms += ms_max;
        return 0;
    };
    if (running_mean) {
        MrtRenderParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.width = W; rp.height = H; rp.samples = N; rp.sample_begin = 0; rp.sample_end = N;
        rp.max_bounces = p.max_bounces; rp.seed = p.seed; rp.max_luminance = p.max_luminance;
        mrt_gpu_init(0, nullptr);
        if (mrt_gpu_render_running_mean(scenes[0], &rp, nullptr)) { fprintf(stderr, "error: %s\n", mrt_last_error()); return 1; }
        passes = 0;
    }
    for (uint32_t pass = 0; pass < passes; pass++) {
        if (pass && collect()) return 1;
        for (uint32_t g = 0; g < G; g++) {
            const uint32_t g0 = (uint32_t) ((uint64_t) N * g / G), g1 = (uint32_t) ((uint64_t) N * (g + 1) / G);
            MrtRenderParams rp;
            memset(&rp, 0, sizeof(rp));
            rp.width = W; rp.height = H; rp.samples = N;
            rp.sample_begin = g0 + (uint32_t) ((uint64_t) (g1 - g0) * pass / passes);
            rp.sample_end = g0 + (uint32_t) ((uint64_t) (g1 - g0) * (pass + 1) / passes);
            if (rp.sample_begin == rp.sample_end) continue;
            rp.max_bounces = p.max_bounces;
            rp.seed = p.seed;
            rp.max_luminance = p.max_luminance;
            rp.flags = pass ? MRT_RENDER_ACCUMULATE : 0u;
            mrt_gpu_init((int) g, nullptr);
            if (mrt_gpu_render_async(scenes[g], &rp)) { fprintf(stderr, "error: %s\n", mrt_last_error()); return 1; }
        }
        if (passes > 1) fprintf(stderr, "\rpass %u/%u", pass + 1, passes);
    }
    // poll loop (main.cpp:387-411)
    for (;;) {
        float pct_min = 100.0f;
        for (uint32_t g = 0; g < G; g++) {
            float pct = 0;
            mrt_gpu_init((int) g, nullptr);
            mrt_gpu_poll(scenes[g], &pct, nullptr);
            if (pct < pct_min) pct_min = pct;
        }
        if (pct_min >= 100.0f) break;
        double el = now_s() - t1;
        fprintf(stderr, "\rTrace: %.2fs (%.0f%%)", el, pct_min);
        std::this_thread::sleep_for(std::chrono::milliseconds(33));
    }
    if (collect()) return 1;
    double secs = now_s() - t1;
    fprintf(stderr, "\r");
    printf("Trace: %.2fs - %.3f Mrays/s | %.6f us/ray | %.3f Mpaths/s | kernel %.1f ms | %llu samples dropped (non-finite)\n", secs,
           rays * 1e-6 / secs, secs * 1e6 / (double) rays, paths * 1e-6 / secs, kernel_ms, (unsigned long long) dropped);

    if (p.out_path[0]) {
        // Sum of the GPUs' accumulators + mean + luminance clamp (main.cpp:168-173) + tone map (main.cpp:416-444): one call,
        // done on the GPUs over NVLink peer memory (mrt_gpu_reduce_finalize); the host only receives the finished image.
        const size_t len = strlen(p.out_path);
        const bool pfm = len > 4 && !strcmp(p.out_path + len - 4, ".pfm");
        std::vector<float> rgba(pfm ? (size_t) W * H * 4 : 0);
        std::vector<uint32_t> argb(pfm ? 0 : (size_t) W * H);
        int rc_img;
        if (running_mean) rc_img = pfm ? mrt_gpu_readback(scenes[0], rgba.data(), 1) : mrt_gpu_tonemap(scenes[0], argb.data());   // the running mean itself
        else rc_img = mrt_gpu_reduce_finalize(scenes.data(), (int) G, p.max_luminance, pfm ? rgba.data() : nullptr, pfm ? nullptr : argb.data());
        if (rc_img) {
            fprintf(stderr, "error: %s\n", mrt_last_error());
            return 1;
        }
        const int rc = pfm ? write_pfm(p.out_path, rgba.data(), W, H) : write_ppm(p.out_path, argb.data(), W, H);
        if (rc) { fprintf(stderr, "cannot write %s\n", p.out_path); return 1; }
        printf("wrote %s\n", p.out_path);
    }
    for (uint32_t g = 0; g < G; g++) { mrt_gpu_init((int) g, nullptr); mrt_gpu_destroy(scenes[g]); }
    mrt_scene_free(hs);
    return 0;
}
