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

static int g_last_split = -1, g_last_segs_b = -1, g_last_segs = -1;   // the decomposition of the most recent v5emu_analyze call
extern "C" int v5emu_last_split_frame(void) { return g_last_split; }
extern "C" int v5emu_last_segments(int tail) { return tail ? g_last_segs_b : g_last_segs; }
static int g_target_items = 0;       // v5emu_set_target_items: the host's small-batch decomposition (short segments, narrow strips)
extern "C" void v5emu_set_target_items(int t) { g_target_items = t; }

extern "C" int v5emu_analyze(const uint8_t *rgb, int n, int h, int w, int64_t frame_stride, int64_t row_stride,
                             int quality, v5ela_record *records, uint8_t *residual, int seg_rows, uint32_t *tex_hist)
{
    v5::KParams p;
    if (v5::fill_params(p, rgb, n, h, w, frame_stride, row_stride, records, residual, quality, seg_rows, g_target_items) != 0) return -1;
    p.tex_hist = tex_hist;
    g_last_split = p.split_frame; g_last_segs = p.n_segs; g_last_segs_b = p.n_segs_b;
    static v5::mma::LaneConsts lane_consts[32];
    if (!v5::mma::make_lane_consts(lane_consts)) return -2;
    p.lane_consts = lane_consts;
    memset(records, 0, sizeof(v5ela_record) * (size_t)n);
    if (tex_hist) memset(tex_hist, 0, sizeof(uint32_t) * 256 * (size_t)n);
    v5::Smem *S = (v5::Smem *)aligned_alloc(16, sizeof(v5::Smem));
    memset(S, 0xA5, sizeof(v5::Smem));                       // poison: uninitialised reads must not matter
    std::vector<v5::ThreadAcc> acc(v5::NT);
    const int total = (int)v5::total_work_items(p);
    for (int work = 0; work < total; work++) {
        if (p.tex_hist) v5::process_work_item<false, true>(*S, p, work, acc.data());         // the choice the C ABI makes
        else if (v5::fast_path_ok(p)) v5::process_work_item<true, false>(*S, p, work, acc.data());
        else v5::process_work_item<false, false>(*S, p, work, acc.data());
    }
    for (int i = 0; i < n; i++) v5::finalize_record(records[i]);
    free(S);
    return 0;
}

// The RAGGED instantiation: n frames of different sizes (tightly packed, one after the other in `rgb`), one work-item space.
extern "C" int v5emu_analyze_ragged(const uint8_t *rgb, const int *hw, int n, int quality, v5ela_record *records, uint8_t *residual,
                                    int seg_rows)
{
    std::vector<v5::FrameDesc> tab((size_t)n);
    size_t off = 0;
    uint32_t total = 0;
    for (int i = 0; i < n; i++) {
        const int h = hw[2 * i], w = hw[2 * i + 1];
        v5::FrameDesc &d = tab[(size_t)i];
        d.rgb = rgb + off;
        d.resid = residual ? residual + off : nullptr;
        d.row_stride = 3 * (int64_t)w;
        d.h = h; d.w = w; d.mw = (w + 15) / 16; d.mh = (h + 15) / 16;
        d.n_strips = (d.mw + v5::TW_MAX - 1) / v5::TW_MAX;
        const int sr = seg_rows > 0 ? seg_rows : 17;
        d.n_segs = (d.mh + sr - 1) / sr;
        d.work_base = total;
        total += (uint32_t)(d.n_strips * d.n_segs);
        d.flags = (((uintptr_t)d.rgb | (uintptr_t)d.row_stride) & 15) == 0 ? 1u : 0u;
        if (d.resid && (((uintptr_t)d.resid | (uintptr_t)(3 * w)) & 15) == 0) d.flags |= 2u;
        off += (size_t)h * w * 3;
    }
    v5::KParams p;
    memset(&p, 0, sizeof(p));
    p.records = records;
    p.n = n;
    p.frames = tab.data();
    static v5::mma::LaneConsts lane_consts[32];
    if (!v5::mma::make_lane_consts(lane_consts)) return -2;
    p.lane_consts = lane_consts;
    uint16_t ql[64], qc[64];
    v5::quant_tables(quality, ql, qc);
    v5::make_quant(ql, p.q[0]);
    v5::make_quant(qc, p.q[1]);
    memset(records, 0, sizeof(v5ela_record) * (size_t)n);
    v5::Smem *S = (v5::Smem *)aligned_alloc(16, sizeof(v5::Smem));
    memset(S, 0xA5, sizeof(v5::Smem));
    std::vector<v5::ThreadAcc> acc(v5::NT);
    for (uint32_t work = 0; work < total; work++) v5::process_work_item<false, false, true>(*S, p, (int)work, acc.data());
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
