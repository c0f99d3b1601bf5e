// v5ela_launch.cuh — the fused kernel's __global__ entry and its host-side launcher, written once and compiled twice: by v5ela.cu
// (namespace v5: the 4-threads-per-block shared-memory-transpose block stage) and by v5ela_mma.cu (namespace v5m, V5_MMA_BLOCKS=1: the
// tensor-core block stage). The C ABI (v5ela.cu) picks one per call through FusedArgs::..., see v5ela_set_block_stage in v5ela.h.
#pragma once
#include <cuda_runtime.h>

#include "v5ela.h"
#include "v5ela_fused_args.h"
#include "v5ela_host.h"
#include "v5ela_workitem.cuh"

namespace V5_NS {

static_assert(sizeof(KParams) <= 4096, "kernel parameters must fit the 4 KB parameter bank");
static_assert(sizeof(Smem) <= (227 * 1024) / MIN_CTAS - 1024, "MIN_CTAS CTAs per SM must fit in shared memory");

template <bool FAST, bool TEXHIST>
__global__ void __launch_bounds__(NT, MIN_CTAS) ela_fused_kernel(const __grid_constant__ KParams p, int total_work)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    ThreadAcc acc_store[1];
    acc_store[0].phase = 0;
    if (threadIdx.x == 0) {
        mbar_init(reinterpret_cast<uint64_t *>(&S.full_bar[0]), 1);
        mbar_init(reinterpret_cast<uint64_t *>(&S.full_bar[1]), 1);
        mbar_init(reinterpret_cast<uint64_t *>(&S.done_bar), NT);
        mbar_init_fence();
    }
    __syncthreads();
    // Work items are drawn from a global ticket counter: no tail of idle CTAs whatever the batch size / CTA count ratio.
    for (;;) {
        if (threadIdx.x == 0) S.next_work = atomicAdd(p.ticket, 1u);
        __syncthreads();
        const int work = (int)S.next_work;
        if (work >= total_work) break;
        process_work_item<FAST, TEXHIST>(S, p, work, acc_store);          // ends with a CTA barrier: next_work may be rewritten
    }
}

// once per device: opt in to the dynamic shared memory the kernel needs
inline cudaError_t fused_prepare()
{
    cudaError_t e = cudaFuncSetAttribute(ela_fused_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ela_fused_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ela_fused_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
    return e;
}

// Fills the kernel parameters, picks the instantiation and launches. Returns 0, -1 (bad arguments), -2 (batch too large) or a
// positive cudaError_t. a.inst / a.total report what was (or, with a.check_only, would be) launched.
inline int fused_launch(v5_fused_args &a)
{
    KParams p;
    if (fill_params(p, a.rgb, a.n, a.h, a.w, a.frame_stride, a.row_stride, a.records, a.residual, a.quality, a.seg_rows, a.target_items) != 0)
        return -1;
    p.tex_hist = a.tex_hist;
    p.ticket = a.ticket;
    p.lane_consts = static_cast<const mma::LaneConsts *>(a.lane_consts);
    const long long total = (long long)a.n * p.n_strips * p.n_segs;
    if (total > 0x7fffffffLL) return -2;
    a.total = total;
    const int grid = total < a.max_ctas ? (int)total : a.max_ctas;
    a.inst = p.tex_hist ? V5ELA_INST_TEXHIST : (fast_path_ok(p) ? V5ELA_INST_FAST : V5ELA_INST_GENERAL);
    if (a.check_only) return 0;
    if (a.ev_start) cudaEventRecord(a.ev_start, a.stream);
    if (a.inst == V5ELA_INST_TEXHIST) ela_fused_kernel<false, true><<<grid, NT, sizeof(Smem), a.stream>>>(p, (int)total);
    else if (a.inst == V5ELA_INST_FAST) ela_fused_kernel<true, false><<<grid, NT, sizeof(Smem), a.stream>>>(p, (int)total);
    else ela_fused_kernel<false, false><<<grid, NT, sizeof(Smem), a.stream>>>(p, (int)total);
    const cudaError_t e = cudaGetLastError();
    if (a.ev_stop) cudaEventRecord(a.ev_stop, a.stream);
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace V5_NS
