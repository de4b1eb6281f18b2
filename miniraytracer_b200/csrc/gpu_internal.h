// Internal state of the device half of the C ABI (render_kernel.cu).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "mrt_gpu.h"
#include "trace_core.h"

namespace mrt {
void set_error(const std::string &msg);   // host_api.cpp
constexpr int kBlock = 128;
}  // namespace mrt

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            mrt::set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                          \
            return MRT_E_CUDA;                                                                      \
        }                                                                                           \
    } while (0)

struct MrtScene {
    int device = 0;
    int sm_count = 0;
    void *scene_base = nullptr;   // the packed scene tables + control block (one cached device buffer)
    size_t scene_bytes = 0, own_acc_bytes = 0, order_bytes = 0, final_bytes = 0, argb_bytes = 0;
    mrt::SceneView view;
    uint32_t stack_words = 0, stack_words_coop = 0;
    MrtTuning tuning = {};        // mrt_gpu_set_tuning; all zero = measured defaults
    uint32_t has_trees = 0, n_node2 = 0;
    uint32_t features = 0;        // MRT_FEAT_* mask of the scene -> kernel variant (render_variants.h)
    cudaStream_t stream = nullptr;
    cudaStream_t poll_stream = nullptr;
    // accumulator
    float4 *own_acc = nullptr;
    size_t own_acc_pixels = 0;
    float4 *ext_acc = nullptr;
    uint32_t ext_w = 0, ext_h = 0;
    float4 *final_buf = nullptr;
    size_t final_pixels = 0;
    uint32_t *argb_buf = nullptr;
    // control block
    unsigned int *ticket = nullptr;
    unsigned long long *counters = nullptr;
    unsigned int *max_bits = nullptr;
    int *cancel_dev = nullptr;    // device flag polled by lane 0 when it takes a ticket (L2 hit)
    int *cancel_pinned = nullptr; // pinned staging word for the async write
    unsigned long long *poll_host = nullptr;   // pinned: [0] ticket [1] rays
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_up = nullptr;      // scene tables uploaded (recorded on poll_stream); the first render waits for it
    bool upload_pending = false;
    void *upload_pinned = nullptr;    // pinned staging of the upload (cached per process)
    size_t upload_pinned_bytes = 0;
    // last render
    bool rendered = false;
    MrtRenderParams last;
    uint32_t last_tasks = 0, last_grid = 0, last_block = mrt::kBlock, last_smem = 0, last_mode = 0;
    bool last_coop = false;
    bool final_is_running_mean = false;   // final_buf holds draw2's running mean (mrt_gpu_render_running_mean), not a finalised sum
    uint64_t stat_samples = 0;    // samples per pixel covered by the statistics (several launches with MRT_RENDER_CONTINUE)
    uint32_t last_w = 0, last_h = 0;   // size of the rendered window = of the accumulator
    float4 *last_acc = nullptr;
    // pixel work order (Z-curve), rebuilt when the frame size changes
    uint32_t *order_dev = nullptr;
    uint32_t order_w = 0, order_h = 0;
    // binned mode (render_pixel_binned): path pool + classifier boxes of the root list's composite children
    uint32_t *pool_dev = nullptr;
    size_t pool_words = 0;
    uint32_t *stage_dev = nullptr;   // finished samples of the chunks in flight (float4 per path)
    size_t stage_words = 0;
    uint32_t n_cls_boxes = 0;
    float cls_box[3][6];
};

