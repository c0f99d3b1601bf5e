// v5ela_fused_args.h — plain argument block between the C ABI (v5ela.cu) and the two builds of the fused kernel's launcher
// (v5ela_launch.cuh in namespaces v5 and v5m). No namespaced types: both translation units see the same struct.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "v5ela.h"

struct v5_fused_args {
    const uint8_t *rgb;
    int n, h, w;
    int64_t frame_stride, row_stride;
    v5ela_record *records;
    uint8_t *residual;              // optional
    uint32_t *tex_hist;             // optional
    int quality, seg_rows, target_items, max_ctas;
    unsigned int *ticket;           // zeroed by the caller on the same stream
    const void *lane_consts;        // 32 x mma::LaneConsts in device memory
    cudaStream_t stream;
    int check_only;                 // 1: validate the arguments and fill the outputs, launch nothing
    cudaEvent_t ev_start, ev_stop;  // optional: recorded right around the kernel (v5ela_profile_*)
    int inst;                       // out: V5ELA_INST_*
    long long total;                // out: work items
};

namespace v5 {
cudaError_t fused_prepare_smem();           // v5ela.cu
int fused_launch_smem(v5_fused_args &a);
}
namespace v5m {
cudaError_t fused_prepare_mma();            // v5ela_mma.cu
int fused_launch_mma(v5_fused_args &a);
bool lane_consts_host(void *dst128x32);     // fills 32 x 128 bytes (mma::LaneConsts); false = constants do not fit (cannot happen)
}
