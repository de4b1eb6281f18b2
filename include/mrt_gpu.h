/* mrt_gpu.h -- C ABI of the B200 renderer for MiniRayTracer's bounce loop.
 *
 * The reference has no plugin / FFI interface: its renderer is entered by
 * spawning worker threads on draw()/draw2() (main.cpp:138-243, spawned at
 * main.cpp:378-382) with a `drawArgs` (main.cpp:127-135) that points at the
 * scene graph (scene.h:19-23), the sample grid (main.cpp:319-332) and the
 * global parameters (cmdline_parser.h:5-18); the workers write
 * G_linearBackBuffer (main.cpp:58,175) and count rays in G_rayCounter
 * (main.cpp:55,68); the main thread polls work_queue::getPercentDone
 * (main.cpp:395).  This header is that seam as a C ABI: plain pointers and
 * sizes, opaque handles, int status codes, no C++ or torch types.
 *
 * All functions return MRT_OK (0) or a negative MRT_E_* code; the message is
 * available from mrt_last_error().  Nothing throws across the boundary (the
 * reference is built with -fno-exceptions, clang/clang_build_linux.sh:27).
 * All calls are expected from one host thread per scene handle.
 */
#ifndef MRT_GPU_H
#define MRT_GPU_H

#include <stddef.h>
#include <stdint.h>

#include "mrt_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MRT_OK 0
#define MRT_E_INVALID (-1)  /* bad argument */
#define MRT_E_CUDA (-2)     /* CUDA runtime error (no device, launch failure, ...) */
#define MRT_E_SCENE (-3)    /* scene could not be built / flattened (missing asset, limits) */
#define MRT_E_STATE (-4)    /* call order (e.g. readback before any render) */

const char *mrt_last_error(void);

/* ------------------------------------------------------------------ host side
 * Replaces: ParseArgv/getParams (cmdline_parser.cpp:74-104) and select_scene
 * (scene.cpp:27-49).  Host-only code: usable without a GPU. */

/* MRT_Params (cmdline_parser.h:5-18) plus the options the new host adds. */
typedef struct MrtParams {
    uint32_t window_width, window_height;
    uint32_t buffer_width, buffer_height;
    uint32_t samples_per_pixel;
    uint32_t tile_size;
    uint32_t num_threads;
    uint32_t max_bounces;
    uint32_t scene_select;
    uint32_t threading_mode;
    float max_luminance;
    uint32_t delay;
    /* additions (not in the reference) */
    uint32_t num_gpus;   /* -gpus  */
    uint64_t seed;       /* -seed  : PCG32 initstate of the per-(pixel,sample) streams */
    char out_path[512];  /* -out   : image file (.ppm tone-mapped, .pfm linear) */
    char asset_dir[512]; /* -assets: directory holding earthmap.ppm and obj/ */
} MrtParams;

void mrt_params_default(MrtParams *p);
/* Same options, defaults, range checks and warnings as ParseArgv (cmdline_parser.cpp:78-104).
 * Returns 1 if -help/--help/-? was given (help text printed), 0 otherwise. */
int mrt_params_parse(int argc, char **argv, MrtParams *p);

typedef struct MrtHostScene MrtHostScene;
/* The Cornell box and the final scene allocate a light list of two objects -- the ceiling light and a glass sphere -- but pass
 * count 1 (scene.cpp:326-329, 456-459), so only the rect is importance-sampled.  OR-ing this flag into `scene` builds the list
 * with both (the sphere then goes through sphere::pdf_value / pdf_generate, sphere.cpp:63-79).  Default = the reference's count. */
#define MRT_SCENE_ALL_LIGHTS 0x100u
/* Cornell box (scene 5) only: two triangle_scene_objects (triangle.cpp:5-175 -- a class no stock scene instantiates) appended to
 * the object list, one with a face normal, one with vertex normals; exists so that the class has a parity test. */
#define MRT_SCENE_EXTRA_TRIANGLES 0x200u
/* scene: the reference's enum scenes value (scene.h:6-17), optionally | MRT_SCENE_ALL_LIGHTS | MRT_SCENE_EXTRA_TRIANGLES;
 * aspect = width/height. */
int mrt_scene_create(uint32_t scene, float aspect, const char *asset_dir, MrtHostScene **out);
/* Flattened description (pointers stay valid until mrt_scene_free). */
const MrtSceneDesc *mrt_scene_desc(const MrtHostScene *s);
/* Canonical text dump of the scene graph (same format as the oracle's dump-scene). */
int mrt_scene_dump(const MrtHostScene *s, const char *path);
/* Flattened-scene file ("MRTSCN1"): every table of MrtSceneDesc, so that OBJ parsing, image decoding and the
 * BVH builds (obj_loader.cpp, triangle.h:77-168, scene_object.h:282-319) need not be repeated -- SURVEY.md 8(f)3.
 * A loaded scene has a description but no graph (mrt_scene_dump is not available for it). */
int mrt_scene_save(const MrtHostScene *s, const char *path);
int mrt_scene_load(const char *path, MrtHostScene **out);
void mrt_scene_free(MrtHostScene *s);

/* ----------------------------------------------------------------- device side
 * Replaces: the draw()/draw2() worker threads and work_queue (work_queue.cpp). */

typedef struct MrtDeviceInfo {
    int device;
    int sm_count;
    int clock_khz;
    int cc_major, cc_minor;
    uint64_t total_mem;
    char name[128];
} MrtDeviceInfo;

int mrt_gpu_init(int device, MrtDeviceInfo *info /* may be NULL */);

typedef struct MrtScene MrtScene;
/* Copies the flattened scene to the current device.  The description and everything it
 * points to may be freed afterwards. */
int mrt_gpu_scene_upload(const MrtSceneDesc *desc, MrtScene **out);
/* Scheduling knobs of the renderer (A/B measurements, tests).  Zero-initialised = the measured defaults; none of
 * them changes a result beyond float summation order (modes) -- see DESIGN.md section 4.  The library reads no
 * environment variables. */
typedef struct MrtTuning {
    uint32_t mode;         /* MRT_MODE_AUTO, or force MRT_MODE_PER_LANE / MRT_MODE_PER_WARP / MRT_MODE_BINNED */
    uint32_t bins;         /* mode B: 0 = classifier bins (default), 1 = one bin, 2 = classifier bins, 3 = + pending-weight bit */
    uint32_t min_blocks;   /* launch-bounds variant: resident blocks per SM, 5..8 (0 = by scene type) */
    uint32_t chunk_pixels; /* pixels per warp task (0 = automatic) */
    uint32_t variant_all;  /* 1 = the unspecialised kernel (all scene features compiled in) */
    uint32_t z_order;      /* 1 = hand out pixels along a Z-curve instead of row-major */
    uint32_t coop_trees;   /* BVH trees: 0 = default, 1 = per-lane traversal, 2 = warp-cooperative traversal */
    uint32_t coop_leaf_batch; /* cooperative traversal: leaves queued before a leaf step runs, 1..32 (0 = default) */
    uint32_t chunk_paths;  /* mode B: paths per big warp task (0 = by scene type); the guided tail of small tasks stays */
    uint32_t tail_tasks;   /* mode B: big tasks per resident warp handed out in small pieces at the end (0 = by scene type) */
    uint32_t blocks_per_sm; /* resident blocks per SM actually launched (0 = all that fit) */
    uint32_t reserved[5];
} MrtTuning;
#define MRT_MODE_AUTO 0u
#define MRT_MODE_PER_LANE 1u /* a lane owns a pixel and adds its samples in the reference's order (main.cpp:154-166) */
#define MRT_MODE_PER_WARP 2u /* a warp owns a chunk of pixels; lanes keep their paths */
#define MRT_MODE_BINNED 3u   /* a warp owns a chunk; paths are parked and regrouped between segments (default).  A pixel's samples are
                                summed in an order fixed by the launch's samples per pixel: <= 128 -> one after the other (the reference's
                                order), more -> lane-strided partial sums + a fixed tree; never dependent on the schedule */
int mrt_gpu_set_tuning(MrtScene *s, const MrtTuning *t /* NULL = defaults */);
/* Launch on this CUDA stream (a cudaStream_t passed as void*; NULL = default stream). */
int mrt_gpu_set_stream(MrtScene *s, void *cuda_stream);
/* Render into caller-owned device memory (width*height float4) instead of the library's own
 * accumulator -- e.g. a tensor that is then sum-reduced across GPUs.  NULL unbinds. */
int mrt_gpu_bind_accumulator(MrtScene *s, void *device_ptr, uint32_t width, uint32_t height);
/* Asynchronously renders samples [sample_begin, sample_end) of every pixel into the
 * accumulator: float4 = (sum of finite radiance samples, finite-sample count), row-major,
 * y up like G_linearBackBuffer.  Replaces spawning draw() (main.cpp:378-382). */
int mrt_gpu_render_async(MrtScene *s, const MrtRenderParams *p);
/* Progress in percent (work_queue::getPercentDone, main.cpp:395) and rays so far;
 * never blocks. */
int mrt_gpu_poll(MrtScene *s, float *pct_done, uint64_t *rays);
/* Blocks until the last render has finished (thread join, main.cpp:491-493). */
int mrt_gpu_wait(MrtScene *s);

typedef struct MrtRenderStats {
    uint64_t rays;       /* trace() calls = path segments (G_rayCounter, main.cpp:68) */
    uint64_t paths;      /* (pixel, sample) pairs */
    uint64_t nonfinite;  /* samples dropped by the finite check (main.cpp:163-165) */
    uint64_t warp_iterations; /* segment steps executed per warp, summed: rays / (32 * this) = share of lanes with a live path */
    float kernel_ms;     /* CUDA-event time of the render kernel on its stream */
    uint32_t grid, block, smem_bytes;
    uint32_t mode;       /* MRT_MODE_* of the kernel that ran */
    uint32_t coop_trees; /* 1 = BVH trees were traversed warp-cooperatively (coop_tree.cuh) */
    /* cooperative traversal: steps executed (summed over warps) and work items popped in them: items / (32 * steps) =
       lane fill of the node (box tests) and leaf (primitive tests) phases */
    uint64_t coop_node_steps, coop_node_items, coop_leaf_steps, coop_leaf_items;
    /* mode B: %globaltimer at entry and exit of every warp.  warp_time_sum_ns / (warps * warp_span_ns) = share of the launch the
       average warp was at work (the rest: warps that ran out of tickets wait for the slowest one) */
    uint64_t warp_time_sum_ns, warp_span_ns, first_exit_ns;   /* first_exit_ns: first warp exit, relative to the first warp entry */
    uint32_t warps, reserved;
    uint64_t stage_sum_ns;   /* mode B: warp time spent adding up the staged samples at the ends of the chunks (part of warp_time_sum_ns) */
} MrtRenderStats;
/* Statistics of the last finished render (blocks like mrt_gpu_wait). */
int mrt_gpu_stats(MrtScene *s, MrtRenderStats *out);

/* mean over finite samples + luminance clamp (main.cpp:168-173) on device buffers:
 * out[i].xyz = clamp(acc[i].xyz / acc[i].w), out[i].w = acc[i].w.  In place is allowed. */
int mrt_gpu_finalize_device(MrtScene *s, const void *acc_dev, void *out_dev, uint32_t width, uint32_t height,
                            float max_luminance);
/* Copies the accumulator (finalize = 0) or the finalised image (finalize = 1) to
 * rgba_host (width*height*4 floats); blocks.  Replaces reading G_linearBackBuffer. */
int mrt_gpu_readback(MrtScene *s, float *rgba_host, int finalize);
/* The reference's default worker draw2 (main.cpp:193-243, with work_queue_dynamic): sample-major passes and a RUNNING MEAN per
 * pixel -- a non-finite sample is replaced by the mean so far (0 for the first sample), and the luminance clamp is applied after
 * EVERY pass and feeds back into the mean (main.cpp:214-231).  mrt_gpu_running_mean_update applies one such pass: acc_dev holds one
 * sample per pixel (a one-sample launch: w = 1 if finite), mean_dev (float4 per pixel) is updated in place; `pass` = 0 for the first.
 * mrt_gpu_render_running_mean runs the whole thing for samples [sample_begin, sample_end) as one-sample launches into the scene's
 * own buffers and copies the final mean to rgba_host (may be NULL; w = number of passes).  The plain accumulate path (sum / count,
 * clamp once at the end) gives the same image unless a running mean crosses max_luminance on the way. */
int mrt_gpu_running_mean_update(MrtScene *s, const void *acc_dev, void *mean_dev, uint32_t width, uint32_t height, uint32_t pass,
                                float max_luminance);
int mrt_gpu_render_running_mean(MrtScene *s, const MrtRenderParams *p, float *rgba_host);
/* Adaptive logarithmic tone map + ARGB32 pack of the reference's preview loop
 * (main.cpp:416-444, vec3.h:327-333) from the finalised image; argb_host: width*height uint32. */
int mrt_gpu_tonemap(MrtScene *s, uint32_t *argb_host);
/* The same tone map on device buffers: img_dev = width*height float4 (finalised linear image), argb_dev = width*height
 * uint32; asynchronous on the scene's stream. */
int mrt_gpu_tonemap_device(MrtScene *s, const void *img_dev, void *argb_dev, uint32_t width, uint32_t height);
/* Samples-per-pixel sharding inside one process: the n scenes rendered the same frame (same size) on n GPUs, each its own
 * sample slice.  Sums the accumulators in scene order and applies mean + luminance clamp (main.cpp:168-173) in one kernel per
 * GPU over peer memory (NVLink / NVSwitch: each GPU reduces a stripe of the pixels and stores it into scenes[0]'s image), then
 * optionally tone-maps on scenes[0]'s GPU.  rgba_host (width*height*4 floats) and / or argb_host (width*height uint32) receive
 * the result; either may be NULL.  Blocks.  Scenes on the same device work too (no peer access involved). */
int mrt_gpu_reduce_finalize(MrtScene **scenes, int n, float max_luminance, float *rgba_host, uint32_t *argb_host);
/* Requests the running render to stop early (G_isRunning = false, main.cpp:274). */
int mrt_gpu_cancel(MrtScene *s);
void mrt_gpu_destroy(MrtScene *s);

#ifdef __cplusplus
}
#endif
#endif /* MRT_GPU_H */
