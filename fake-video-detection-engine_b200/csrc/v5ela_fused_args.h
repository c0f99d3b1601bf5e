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
    const int *tune;                // optional decomposition knobs (fill_params)
    unsigned int *ticket;           // zeroed by the caller on the same stream
    const void *lane_consts;        // 32 x mma::LaneConsts in device memory
    cudaStream_t stream;
    int check_only;                 // 1: validate the arguments and fill the outputs, launch nothing
    cudaEvent_t ev_start, ev_stop;  // optional: recorded right around the kernel (v5ela_profile_*)
    int inst;                       // out: V5ELA_INST_*
    long long total;                // out: work items
};

#define V5_RAGGED_DESC_BYTES 56              // sizeof(FrameDesc), csrc/v5ela_device.cuh

struct v5_ragged_args {
    const v5ela_frame_desc *frames; // host array of n descriptors holding DEVICE pointers
    int n;
    v5ela_record *records;          // device, n records
    int quality, seg_rows, target_items, max_ctas;
    unsigned int *ticket;
    const void *lane_consts;
    void *table;                    // host staging for the kernel's frame table: table_bytes(n)
    const void *d_table;            // its device copy (uploaded by the caller between the two calls)
    cudaStream_t stream;
    cudaEvent_t ev_start, ev_stop;
    int check_only;
    long long total;                // out (check_only pass): work items
    static size_t table_bytes(int n) { return (size_t)V5_RAGGED_DESC_BYTES * (size_t)n; }
};

namespace v5 {
cudaError_t fused_prepare_smem();           // v5ela.cu
int fused_launch_smem(v5_fused_args &a);
int fused_launch_ragged_smem(v5_ragged_args &a);
}
namespace v5m {
cudaError_t fused_prepare_mma();            // v5ela_mma.cu
int fused_launch_mma(v5_fused_args &a);
int fused_launch_ragged_mma(v5_ragged_args &a);
bool lane_consts_host(void *dst128x32);     // fills 32 x 128 bytes (mma::LaneConsts); false = constants do not fit (cannot happen)
}
