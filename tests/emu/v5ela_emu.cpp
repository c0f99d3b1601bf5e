// tests/emu/v5ela_emu.cpp — TEST INFRASTRUCTURE: runs the CUDA kernel's device code (csrc/v5ela_device.cuh,
// csrc/v5ela_workitem.cuh, compiled by g++ with the V5_* host shims) with the threads of each CTA emulated sequentially
// between barriers. It exists so the band/halo/edge indexing of the fused kernel can be checked against the oracle in
// the build container, which has no GPU. It is NOT a CPU fallback: nothing under fake-video-detection-engine_b200/ loads it.
//
// Build: g++ -O2 -shared -fPIC -I include -I fake-video-detection-engine_b200/csrc tests/emu/v5ela_emu.cpp -o tests/emu/libv5ela_emu.so
#include <cstdlib>
#include <cstring>
#include <vector>

#include "v5ela_workitem.cuh"
#include "v5ela_host.h"

extern "C" int v5emu_analyze(const uint8_t *rgb, int n, int h, int w, int64_t frame_stride, int64_t row_stride,
                             int quality, v5ela_record *records, uint8_t *residual, int seg_rows, uint32_t *tex_hist)
{
    v5::KParams p;
    if (v5::fill_params(p, rgb, n, h, w, frame_stride, row_stride, records, residual, quality, seg_rows) != 0) return -1;
    p.tex_hist = tex_hist;
    static v5::mma::LaneConsts lane_consts[32];
    if (!v5::mma::make_lane_consts(lane_consts)) return -2;
    p.lane_consts = lane_consts;
    memset(records, 0, sizeof(v5ela_record) * (size_t)n);
    if (tex_hist) memset(tex_hist, 0, sizeof(uint32_t) * 256 * (size_t)n);
    v5::Smem *S = (v5::Smem *)aligned_alloc(16, sizeof(v5::Smem));
    memset(S, 0xA5, sizeof(v5::Smem));                       // poison: uninitialised reads must not matter
    std::vector<v5::ThreadAcc> acc(v5::NT);
    const int total = n * p.n_strips * p.n_segs;
    for (int work = 0; work < total; work++) {
        if (p.tex_hist) v5::process_work_item<false, true>(*S, p, work, acc.data());         // the choice the C ABI makes
        else if (v5::fast_path_ok(p)) v5::process_work_item<true, false>(*S, p, work, acc.data());
        else v5::process_work_item<false, false>(*S, p, work, acc.data());
    }
    for (int i = 0; i < n; i++) v5::finalize_record(records[i]);
    free(S);
    return 0;
}

// 16-bit hand-offs of the block stage (int16 pairs / 8-bit limb pairs) that left their range, since the library was loaded
extern "C" long long v5emu_range_violations(void)
{
    return v5::range_violations();
}

extern "C" int v5emu_quant_selftest(void)
{
    // exhaustive: the reciprocal division equals libjpeg's rounding for every table entry and coefficient
    for (int t = 1; t <= 255; t++) {
        uint16_t tab[64];
        for (int i = 0; i < 64; i++) tab[i] = (uint16_t)t;
        v5::QuantTab q;
        v5::make_quant(tab, q);
        for (int c = -8192; c <= 8192; c++) {
            const int div = t << 3, a = c < 0 ? -c : c;
            int ref = (a + (div >> 1)) / div;
            if (c < 0) ref = -ref;
            const uint32_t x = (uint32_t)(c + (c >> 31) + q.bias[0]);
            const int got = (int)v5::umulhi32(x, q.recip[0]) * q.t[0] - q.unbias[0];
            if (got != ref * t) return t * 100000 + (c + 8192);
        }
    }
    return 0;
}
