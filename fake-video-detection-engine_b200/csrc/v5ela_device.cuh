// v5ela_device.cuh — device code of the fused V5 ELA + texture kernel (sm_100a).
//
// One CTA walks DOWN a vertical strip of a frame (<= TW_MAX MCUs = 480 px wide, plus one halo MCU column each side),
// one 16-pixel-tall MCU row ("band") per iteration, and for every band does, entirely in shared memory/registers:
//
//   load      : RGB band (16 lines x up to 512 px) global -> smem, edge-replicated            
//   convert   : RGB -> Y, Cb, Cr (16-bit fixed point), h2v2 chroma box downsample            (A.2, A.3)
//   blocks    : per 8x8 block, one thread: ISLOW fDCT -> quantise -> dequantise -> ISLOW IDCT (A.4, A.5, A.6)
//   residual  : h2v2 fancy upsample, YCbCr -> RGB, |orig - recon|, histogram; luma Laplacian  (A.7, A.8, A.9, §8a)
//
// The decoder's fancy upsampling needs one chroma sample beyond every 8x8 chroma block: horizontally that is what the
// halo MCU columns are for (their chroma is recomputed, (TW+2)/TW extra chroma work); vertically the strip walk keeps the
// previous band's decoded chroma in a two-band ring, and the residual stage of iteration r covers pixel rows
// 16r-1 .. 16r+14 — the last row of a band is finished one iteration later, when the chroma row below it exists.
// Each frame byte is read from HBM once (halo columns come from L2).
//
// The file is also compiled by g++ for tests/emu (a thread-emulating CPU harness used to debug indexing without a GPU);
// every CUDA-specific construct therefore goes through the small V5_* shims below. The product never runs that build.
#pragma once
#include <stdint.h>

#include "v5ela.h"

#ifdef __CUDACC__
#define V5_DEV __device__ __forceinline__
#define V5_HOSTDEV __host__ __device__ __forceinline__
#else
#define V5_DEV inline
#define V5_HOSTDEV inline
#endif

namespace v5 {

// ----------------------------------------------------------------------------------------------- geometry constants
constexpr int NT = 256;                     // threads per CTA
constexpr int TW_MAX = 30;                  // strip width, MCUs (16 px)
constexpr int BAND_MCUS = TW_MAX + 2;       // + halo MCU column each side
constexpr int BAND_PX = BAND_MCUS * 16;     // 512
constexpr int RGB_PITCH = BAND_PX * 3;      // 1536 bytes per band line
constexpr int Y_PITCH = BAND_PX;            // 512
constexpr int C_PITCH = BAND_PX / 2;        // 256
constexpr int CHROMA_TID0 = ((4 * TW_MAX + 31) / 32) * 32;   // first thread of the chroma block warps (128)

// Exact division constants for one quantisation table entry T (divisor d = 8T), see make_quant():
//   q_biased = umulhi(c + (c >> 31) + bias, recip);  dequantised = q_biased * t - unbias
struct QuantTab {
    uint32_t recip[64];
    int32_t bias[64];
    int32_t t[64];
    int32_t unbias[64];
};

struct KParams {
    const uint8_t *rgb;
    int64_t frame_stride;
    int64_t row_stride;
    v5ela_record *records;
    uint8_t *residual;          // optional
    int n, h, w;
    int mw, mh;                 // MCU columns / rows of the padded frame
    int n_strips, n_segs;       // work decomposition: strips x vertical segments per frame
    int vec_ok;                 // 1: every band line start is 16-byte aligned in global memory (128-bit loads)
    int resid_vec_ok;           // 1: residual rows are 16-byte aligned (3*W % 16 == 0 and base aligned): 128-bit stores
    QuantTab q[2];              // [0] luma, [1] chroma — lives in the kernel parameter constant bank
};

// Host side: fills the reciprocal constants (called by v5ela_set_quality).
inline void make_quant(const uint16_t tab[64], QuantTab &q)
{
    for (int i = 0; i < 64; i++) {
        const uint32_t t = tab[i], d = t << 3;
        const uint32_t b = (8192u + d - 1) / d;                     // b*d >= 8192 >= max |coef|
        q.recip[i] = (uint32_t)(0x100000000ull / d) + 1u;           // exact for x*d < 2^32, x < 2^15
        q.bias[i] = (int32_t)(d / 2 + b * d);
        q.t[i] = (int32_t)t;
        q.unbias[i] = (int32_t)(b * t);
    }
}

// ------------------------------------------------------------------------------------------------------- shims
#ifdef __CUDA_ARCH__
V5_DEV uint32_t umulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
V5_DEV int clamp255(int v) { return __vimin_s32_relu(v, 255); }
V5_DEV void smem_inc(uint32_t *p) { atomicAdd(p, 1u); }
#else
inline uint32_t umulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
inline void smem_inc(uint32_t *p) { *p += 1u; }
#endif

struct alignas(16) U4 { uint32_t x, y, z, w; };
struct alignas(8) U2 { uint32_t x, y; };

V5_DEV uint32_t byte_of(uint32_t w, int k) { return (w >> (8 * k)) & 0xffu; }
V5_DEV uint32_t pack4(int a, int b, int c, int d)
{
    return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}

// ------------------------------------------------------------------------------------------------- shared memory
struct alignas(16) Smem {
    uint8_t rgb[2][16][RGB_PITCH];      // band r in rgb[r & 1]
    uint8_t yorig[32][Y_PITCH];         // luma of the original; band r line l at [16*(r&1) + l]
    uint8_t ydec[32][Y_PITCH];          // luma after the JPEG round trip, same ring
    uint8_t cenc[2][8][C_PITCH];        // downsampled Cb/Cr of the current band (input of the block stage)
    uint8_t cdec[2][16][C_PITCH];       // decoded Cb/Cr; band r chroma line j at [8*(r&1) + j]
    uint32_t hist[3][256];
    unsigned long long tex_sumabs, tex_sumsq;
    uint32_t tex_maxabs;
    uint32_t pad_[3];
};

// Per work item geometry (uniform across the CTA).
struct Geo {
    const uint8_t *frame;       // first byte of this frame
    uint8_t *resid;             // first byte of this frame's residual map, or null
    int m0, m1;                 // strip MCU columns [m0, m1)
    int r0, r1;                 // segment MCU rows [r0, r1)
    int xb0;                    // pixel column of band smem column 0 (= 16*(m0-1), may be -16)
    int band_mcus;              // m1 - m0 + 2
};

struct ThreadAcc {              // per-thread accumulators that live across bands of one work item
    unsigned long long tex_sumsq;
    uint32_t tex_sumabs;
    uint32_t tex_maxabs;
};

V5_DEV int ring16(int r, int l) { return (16 * (r & 1) + l) & 31; }   // l in [-2, 15]
V5_DEV int ring8(int r, int j) { return (8 * (r & 1) + j) & 15; }     // j in [-1, 7]

// ------------------------------------------------------------------------------------------------ stage: load band
// Band r, lines 0..15 <- frame rows min(16r + l, H-1), pixel columns [xb0, xb0 + 16*band_mcus) clipped to [0, W);
// columns W .. Wm-1 replicate pixel W-1 (A.3: edges are replicated in full-resolution colour space).
V5_DEV void stage_load(int tid, Smem &S, const KParams &p, const Geo &g, int r)
{
    const int xs = g.xb0 < 0 ? 0 : g.xb0;
    int xe = g.xb0 + 16 * g.band_mcus;
    const int xpad_end = xe > 16 * p.mw ? 16 * p.mw : xe;       // last padded column (exclusive) inside this band
    if (xe > p.w) xe = p.w;
    const int nbytes = 3 * (xe - xs);
    const int dst0 = 3 * (xs - g.xb0);
    uint8_t(*dst)[RGB_PITCH] = S.rgb[r & 1];
    if (p.vec_ok) {
        const int nvec = nbytes >> 4;                           // 16-byte chunks per line
        for (int i = tid; i < 16 * nvec; i += NT) {
            const int l = i / nvec, v = i - l * nvec;
            int y = 16 * r + l;
            if (y > p.h - 1) y = p.h - 1;
            const U4 *src = reinterpret_cast<const U4 *>(g.frame + (int64_t)y * p.row_stride + 3 * xs);
            *reinterpret_cast<U4 *>(&dst[l][dst0 + 16 * v]) = src[v];
        }
        const int tail = nbytes & 15;
        for (int i = tid; i < 16 * tail; i += NT) {
            const int l = i / tail, b = (nvec << 4) + (i - l * tail);
            int y = 16 * r + l;
            if (y > p.h - 1) y = p.h - 1;
            dst[l][dst0 + b] = g.frame[(int64_t)y * p.row_stride + 3 * xs + b];
        }
    } else {
        for (int i = tid; i < 16 * nbytes; i += NT) {
            const int l = i / nbytes, b = i - l * nbytes;
            int y = 16 * r + l;
            if (y > p.h - 1) y = p.h - 1;
            dst[l][dst0 + b] = g.frame[(int64_t)y * p.row_stride + 3 * xs + b];
        }
    }
    const int npad = xpad_end - p.w;                            // > 0 only in the strip that holds the right edge
    if (npad > 0) {
        for (int i = tid; i < 16 * npad * 3; i += NT) {
            const int l = i / (npad * 3), b = i - l * npad * 3;
            int y = 16 * r + l;
            if (y > p.h - 1) y = p.h - 1;
            dst[l][3 * (p.w - g.xb0) + b] = g.frame[(int64_t)y * p.row_stride + 3 * (p.w - 1) + (b % 3)];
        }
    }
}

// ------------------------------------------------------------------------------------- stage: colour convert (A.2/A.3)
V5_DEV int rgb_to_y(int r, int g, int b) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
V5_DEV int rgb_to_cb(int r, int g, int b) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
V5_DEV int rgb_to_cr(int r, int g, int b) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

// 16 pixels of two band lines: optional luma (two packed lines) and the 8 downsampled Cb/Cr samples.
template <bool WANT_Y, bool WANT_C>
V5_DEV void convert16x2(const uint8_t *la, const uint8_t *lb, U4 &y0, U4 &y1, U2 &cbo, U2 &cro)
{
    uint32_t ya[4], yb[4], cbw[2], crw[2];
#pragma unroll
    for (int q4 = 0; q4 < 4; q4++) {                            // 4 pixels = 12 bytes = 3 words per line
        const uint32_t *wa = reinterpret_cast<const uint32_t *>(la) + 3 * q4;
        const uint32_t *wb = reinterpret_cast<const uint32_t *>(lb) + 3 * q4;
        const uint32_t a[3] = {wa[0], wa[1], wa[2]}, b[3] = {wb[0], wb[1], wb[2]};
        int yy[2][4], cb[2][4], cr[2][4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o = 3 * k;
            const int ra = byte_of(a[o >> 2], o & 3), ga = byte_of(a[(o + 1) >> 2], (o + 1) & 3),
                      ba = byte_of(a[(o + 2) >> 2], (o + 2) & 3);
            const int rb = byte_of(b[o >> 2], o & 3), gb = byte_of(b[(o + 1) >> 2], (o + 1) & 3),
                      bb = byte_of(b[(o + 2) >> 2], (o + 2) & 3);
            if (WANT_Y) {
                yy[0][k] = rgb_to_y(ra, ga, ba);
                yy[1][k] = rgb_to_y(rb, gb, bb);
            }
            if (WANT_C) {
                cb[0][k] = rgb_to_cb(ra, ga, ba);
                cb[1][k] = rgb_to_cb(rb, gb, bb);
                cr[0][k] = rgb_to_cr(ra, ga, ba);
                cr[1][k] = rgb_to_cr(rb, gb, bb);
            }
        }
        if (WANT_Y) {
            ya[q4] = pack4(yy[0][0], yy[0][1], yy[0][2], yy[0][3]);
            yb[q4] = pack4(yy[1][0], yy[1][1], yy[1][2], yy[1][3]);
        }
        if (WANT_C) {                                           // h2v2 box filter, bias 1,2,1,2 along x
            const int c0 = (cb[0][0] + cb[0][1] + cb[1][0] + cb[1][1] + 1) >> 2;
            const int c1 = (cb[0][2] + cb[0][3] + cb[1][2] + cb[1][3] + 2) >> 2;
            const int d0 = (cr[0][0] + cr[0][1] + cr[1][0] + cr[1][1] + 1) >> 2;
            const int d1 = (cr[0][2] + cr[0][3] + cr[1][2] + cr[1][3] + 2) >> 2;
            const uint32_t cbp = (uint32_t)c0 | ((uint32_t)c1 << 8), crp = (uint32_t)d0 | ((uint32_t)d1 << 8);
            if (q4 & 1) {
                cbw[q4 >> 1] |= cbp << 16;
                crw[q4 >> 1] |= crp << 16;
            } else {
                cbw[q4 >> 1] = cbp;
                crw[q4 >> 1] = crp;
            }
        }
    }
    if (WANT_Y) {
        y0 = U4{ya[0], ya[1], ya[2], ya[3]};
        y1 = U4{yb[0], yb[1], yb[2], yb[3]};
    }
    if (WANT_C) {
        cbo = U2{cbw[0], cbw[1]};
        cro = U2{crw[0], crw[1]};
    }
}

V5_DEV void stage_convert(int tid, Smem &S, const KParams &p, const Geo &g, int r)
{
    // Chroma line j of this band averages frame rows (2jc, min(2jc+1, H-1)) with jc = min(8r+j, He/2-1): below the
    // image the DOWNSAMPLED last row is replicated, which differs from the luma rule (replicate row H-1) when H is even.
    const int last_cline = ((p.h + 1) >> 1) - 1 - 8 * r;        // local index of the last real chroma line
    const int last_line = p.h - 1 - 16 * r;                     // local index of the last real pixel line
    const uint8_t(*src)[RGB_PITCH] = S.rgb[r & 1];
    for (int u = tid; u < 8 * BAND_MCUS; u += NT) {
        const int li = u / BAND_MCUS, ux = u - li * BAND_MCUS;
        const int mcu = g.m0 - 1 + ux;
        if (ux >= g.band_mcus || mcu < 0 || mcu >= p.mw) continue;
        U4 y0, y1;
        U2 cb, cr;
        convert16x2<true, true>(&src[2 * li][48 * ux], &src[2 * li + 1][48 * ux], y0, y1, cb, cr);
        const int jc = li < last_cline ? li : last_cline;
        int lb = 2 * jc + 1;
        if (lb > last_line) lb = last_line;
        if (2 * jc != 2 * li || lb != 2 * li + 1) {
            if (lb < 2 * jc) lb = 2 * jc;                       // (cannot happen: last_line >= 2*last_cline)
            U4 d0, d1;
            convert16x2<false, true>(&src[2 * jc][48 * ux], &src[lb][48 * ux], d0, d1, cb, cr);
        }
        *reinterpret_cast<U4 *>(&S.yorig[ring16(r, 2 * li)][16 * ux]) = y0;
        *reinterpret_cast<U4 *>(&S.yorig[ring16(r, 2 * li + 1)][16 * ux]) = y1;
        *reinterpret_cast<U2 *>(&S.cenc[0][li][8 * ux]) = cb;
        *reinterpret_cast<U2 *>(&S.cenc[1][li][8 * ux]) = cr;
    }
}

// --------------------------------------------------------------------------------- stage: 8x8 block round trip (A.4-A.6)
#define V5_C0_298 2446
#define V5_C0_390 3196
#define V5_C0_541 4433
#define V5_C0_765 6270
#define V5_C0_899 7373
#define V5_C1_175 9633
#define V5_C1_501 12299
#define V5_C1_847 15137
#define V5_C1_961 16069
#define V5_C2_053 16819
#define V5_C2_562 20995
#define V5_C3_072 25172

// Forward 8-point pass over v[0], v[S], ... v[7S]. PASS1: row pass on UNSIGNED samples (the -128 level shift only
// moves the DC term: -8*128 before the << 2), descale 11. Otherwise: column pass, descale 15, DC descale 2.
template <int S, bool PASS1>
V5_DEV void fdct8(int *v)
{
    const int n = PASS1 ? 11 : 15, rnd = 1 << (n - 1);
    const int t0 = v[0] + v[7 * S], t7 = v[0] - v[7 * S], t1 = v[S] + v[6 * S], t6 = v[S] - v[6 * S];
    const int t2 = v[2 * S] + v[5 * S], t5 = v[2 * S] - v[5 * S], t3 = v[3 * S] + v[4 * S], t4 = v[3 * S] - v[4 * S];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    if (PASS1) {
        v[0] = (t10 + t11 - 1024) * 4;
        v[4 * S] = (t10 - t11) * 4;
    } else {
        v[0] = (t10 + t11 + 2) >> 2;
        v[4 * S] = (t10 - t11 + 2) >> 2;
    }
    v[2 * S] = (t12 * V5_C0_541 + t13 * (V5_C0_541 + V5_C0_765) + rnd) >> n;
    v[6 * S] = (t13 * V5_C0_541 + t12 * (V5_C0_541 - V5_C1_847) + rnd) >> n;
    const int z1 = (t4 + t7) * -V5_C0_899, z2 = (t5 + t6) * -V5_C2_562;
    const int z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * V5_C1_175 + rnd;
    const int z3b = z3 * -V5_C1_961 + z5, z4b = z4 * -V5_C0_390 + z5;
    v[7 * S] = (t4 * V5_C0_298 + z1 + z3b) >> n;
    v[5 * S] = (t5 * V5_C2_053 + z2 + z4b) >> n;
    v[3 * S] = (t6 * V5_C3_072 + z2 + z3b) >> n;
    v[S] = (t7 * V5_C1_501 + z1 + z4b) >> n;
}

// Inverse 8-point pass. FINAL: row pass, descale 18, +128 and clamp to 0..255; otherwise column pass, descale 11.
template <int S, bool FINAL>
V5_DEV void idct8(int *v)
{
    const int n = FINAL ? 18 : 11;
    const int rnd = (1 << (n - 1)) + (FINAL ? (128 << 18) : 0);
    const int i2 = v[2 * S], i6 = v[6 * S];
    const int t2 = i2 * V5_C0_541 + i6 * (V5_C0_541 - V5_C1_847);
    const int t3 = i6 * V5_C0_541 + i2 * (V5_C0_541 + V5_C0_765);
    const int t0 = (v[0] + v[4 * S]) * 8192 + rnd, t1 = (v[0] - v[4 * S]) * 8192 + rnd;
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    int u0 = v[7 * S], u1 = v[5 * S], u2 = v[3 * S], u3 = v[S];
    const int z1 = (u0 + u3) * -V5_C0_899, z2 = (u1 + u2) * -V5_C2_562;
    const int z3 = u0 + u2, z4 = u1 + u3;
    const int z5 = (z3 + z4) * V5_C1_175;
    const int z3b = z3 * -V5_C1_961 + z5, z4b = z4 * -V5_C0_390 + z5;
    u0 = u0 * V5_C0_298 + z1 + z3b;
    u1 = u1 * V5_C2_053 + z2 + z4b;
    u2 = u2 * V5_C3_072 + z2 + z3b;
    u3 = u3 * V5_C1_501 + z1 + z4b;
    if (FINAL) {
        v[0] = clamp255((t10 + u3) >> n);
        v[7 * S] = clamp255((t10 - u3) >> n);
        v[S] = clamp255((t11 + u2) >> n);
        v[6 * S] = clamp255((t11 - u2) >> n);
        v[2 * S] = clamp255((t12 + u1) >> n);
        v[5 * S] = clamp255((t12 - u1) >> n);
        v[3 * S] = clamp255((t13 + u0) >> n);
        v[4 * S] = clamp255((t13 - u0) >> n);
    } else {
        v[0] = (t10 + u3) >> n;
        v[7 * S] = (t10 - u3) >> n;
        v[S] = (t11 + u2) >> n;
        v[6 * S] = (t11 - u2) >> n;
        v[2 * S] = (t12 + u1) >> n;
        v[5 * S] = (t12 - u1) >> n;
        v[3 * S] = (t13 + u0) >> n;
        v[4 * S] = (t13 - u0) >> n;
    }
}

// One 8x8 block held in registers: in[r] = 8 bytes of row r (unsigned samples); out likewise. All indices are compile
// time after unrolling, so the quantisation constants are read as constant-bank operands straight from the kernel params.
V5_DEV void block_roundtrip(const QuantTab &q, const uint8_t *in, int in_pitch, uint8_t *out, int out_pitch)
{
    int v[64];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const U2 w = *reinterpret_cast<const U2 *>(in + r * in_pitch);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            v[8 * r + k] = (int)byte_of(w.x, k);
            v[8 * r + 4 + k] = (int)byte_of(w.y, k);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; r++) fdct8<1, true>(v + 8 * r);
#pragma unroll
    for (int c = 0; c < 8; c++) fdct8<8, false>(v + c);
#pragma unroll
    for (int i = 0; i < 64; i++) {                              // A.5: round half away from zero of c / (8T), times T
        const int c = v[i];
        const uint32_t x = (uint32_t)(c + (c >> 31) + q.bias[i]);
        v[i] = (int)umulhi32(x, q.recip[i]) * q.t[i] - q.unbias[i];
    }
#pragma unroll
    for (int c = 0; c < 8; c++) idct8<8, false>(v + c);
#pragma unroll
    for (int r = 0; r < 8; r++) idct8<1, true>(v + 8 * r);
#pragma unroll
    for (int r = 0; r < 8; r++) {
        U2 w;
        w.x = pack4(v[8 * r], v[8 * r + 1], v[8 * r + 2], v[8 * r + 3]);
        w.y = pack4(v[8 * r + 4], v[8 * r + 5], v[8 * r + 6], v[8 * r + 7]);
        *reinterpret_cast<U2 *>(out + r * out_pitch) = w;
    }
}

// Threads [0, 4*TW) take the luma blocks of the strip proper (halo columns need no decoded luma); threads
// [CHROMA_TID0, CHROMA_TID0 + 2*(TW+2)) take the Cb and Cr blocks including the halo columns. `want_y` is false for
// the halo bands above/below a segment (only their chroma is needed).
V5_DEV void stage_blocks(int tid, Smem &S, const KParams &p, const Geo &g, int r, bool want_y)
{
    const int tw = g.m1 - g.m0;
    if (tid < CHROMA_TID0) {
        if (!want_y || tid >= 4 * tw) return;
        const int br = tid / (2 * tw), bc = tid - br * 2 * tw;
        // blocks entirely below / right of the image are libjpeg "dummy" data: never visible, skip them
        if (16 * r + 8 * br >= p.h || 16 * g.m0 + 8 * bc >= p.w) return;
        const int row = ring16(r, 8 * br), col = 16 + 8 * bc;
        block_roundtrip(p.q[0], &S.yorig[row][col], Y_PITCH, &S.ydec[row][col], Y_PITCH);
    } else {
        const int c = tid - CHROMA_TID0;
        if (c >= 2 * BAND_MCUS) return;
        const int comp = c / BAND_MCUS, cbk = c - comp * BAND_MCUS;
        const int mcu = g.m0 - 1 + cbk;
        if (cbk >= g.band_mcus || mcu < 0 || mcu >= p.mw) return;
        block_roundtrip(p.q[1], &S.cenc[comp][0][8 * cbk], C_PITCH, &S.cdec[comp][ring8(r, 0)][8 * cbk], C_PITCH);
    }
}

// ------------------------------------------------------------------- stage: upsample, reconstruct, residual, Laplacian
// Horizontal+vertical fancy upsample (A.7) of one chroma component for 16 output pixels of one line.
// lc / ln: decoded chroma lines (current row, neighbour row), pointing at this unit's first chroma column; columns -1 and
// 8 are the horizontal neighbours (halo). gcx0 = global chroma column of lc[0]; wc1 = Wc - 1.
V5_DEV void upsample16(const uint8_t *lc, const uint8_t *ln, int gcx0, int wc1, bool fancy, int out[16])
{
    int cs[10];                                                 // cs[j+1] = 3*c[r][j] + c[nb][j], j = -1..8
    const uint32_t c0 = *reinterpret_cast<const uint32_t *>(lc - 4), n0 = *reinterpret_cast<const uint32_t *>(ln - 4);
    const U2 c1 = *reinterpret_cast<const U2 *>(lc), n1 = *reinterpret_cast<const U2 *>(ln);
    const uint32_t c2 = *reinterpret_cast<const uint32_t *>(lc + 8), n2 = *reinterpret_cast<const uint32_t *>(ln + 8);
    if (!fancy) {                                               // Wc <= 2: libjpeg uses plain replication
#pragma unroll
        for (int j = 0; j < 4; j++) {
            out[2 * j] = out[2 * j + 1] = (int)byte_of(c1.x, j);
            out[8 + 2 * j] = out[9 + 2 * j] = (int)byte_of(c1.y, j);
        }
        return;
    }
    cs[0] = 3 * (int)byte_of(c0, 3) + (int)byte_of(n0, 3);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        cs[1 + j] = 3 * (int)byte_of(c1.x, j) + (int)byte_of(n1.x, j);
        cs[5 + j] = 3 * (int)byte_of(c1.y, j) + (int)byte_of(n1.y, j);
    }
    cs[9] = 3 * (int)byte_of(c2, 0) + (int)byte_of(n2, 0);
    if (gcx0 == 0) cs[0] = cs[1];                               // left image edge: neighbour clamps to column 0
    if (gcx0 + 8 > wc1) {                                       // right image edge inside / just after this unit
#pragma unroll
        for (int j = 1; j < 10; j++)
            if (gcx0 + j - 1 > wc1) cs[j] = cs[j - 1];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        out[2 * j] = (3 * cs[j + 1] + cs[j] + 8) >> 4;
        out[2 * j + 1] = (3 * cs[j + 1] + cs[j + 2] + 7) >> 4;
    }
}

template <bool FULL>
V5_DEV void residual_unit(Smem &S, const KParams &p, const Geo &g, ThreadAcc &acc, int r, int l, int ux)
{
    const int y = 16 * r + l;                                   // global pixel row, l in [-1, 14]
    const int gx0 = 16 * (g.m0 + ux);                           // global pixel column of this unit
    const int col = 16 + 16 * ux;                               // band smem column
    const int nvalid = FULL ? 16 : (p.w - gx0);                 // pixels of this unit inside the image

    // ---- chroma upsample (A.7)
    const int hc1 = ((p.h + 1) >> 1) - 1, wc1 = ((p.w + 1) >> 1) - 1;
    const int crow = y >> 1;
    int nrow = (y & 1) ? crow + 1 : crow - 1;
    nrow = nrow < 0 ? 0 : (nrow > hc1 ? hc1 : nrow);
    const int lcur = ring8(r, crow - 8 * r), lnb = ring8(r, nrow - 8 * r);
    const int ccol = col >> 1, gcx0 = gx0 >> 1;
    const bool fancy = wc1 > 1;
    int cb[16], cr[16];
    upsample16(&S.cdec[0][lcur][ccol], &S.cdec[0][lnb][ccol], gcx0, wc1, fancy, cb);
    upsample16(&S.cdec[1][lcur][ccol], &S.cdec[1][lnb][ccol], gcx0, wc1, fancy, cr);

    // ---- reconstruct (A.8), residual (A.9), histogram
    const U4 yd = *reinterpret_cast<const U4 *>(&S.ydec[ring16(r, l)][col]);
    const uint8_t *orig = l < 0 ? &S.rgb[(r - 1) & 1][15][3 * col] : &S.rgb[r & 1][l][3 * col];
    const uint32_t ydw[4] = {yd.x, yd.y, yd.z, yd.w};
    uint32_t ow[12];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const U4 t = reinterpret_cast<const U4 *>(orig)[i];
        ow[4 * i] = t.x; ow[4 * i + 1] = t.y; ow[4 * i + 2] = t.z; ow[4 * i + 3] = t.w;
    }
    uint32_t dw[12];
#pragma unroll
    for (int i = 0; i < 12; i++) dw[i] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int yy = (int)byte_of(ydw[k >> 2], k & 3);
        const int cbv = cb[k], crv = cr[k];
        const int rr = clamp255(yy + ((91881 * crv + (32768 - 91881 * 128)) >> 16));
        const int gg = clamp255(yy + ((-22554 * cbv - 46802 * crv + (32768 + (22554 + 46802) * 128)) >> 16));
        const int bb = clamp255(yy + ((116130 * cbv + (32768 - 116130 * 128)) >> 16));
        const int o = 3 * k;
        int dr = (int)byte_of(ow[o >> 2], o & 3) - rr;
        int dg = (int)byte_of(ow[(o + 1) >> 2], (o + 1) & 3) - gg;
        int db = (int)byte_of(ow[(o + 2) >> 2], (o + 2) & 3) - bb;
        dr = dr < 0 ? -dr : dr;
        dg = dg < 0 ? -dg : dg;
        db = db < 0 ? -db : db;
        if (FULL || k < nvalid) {
            smem_inc(&S.hist[0][dr]);
            smem_inc(&S.hist[1][dg]);
            smem_inc(&S.hist[2][db]);
        }
        dw[o >> 2] |= (uint32_t)dr << (8 * (o & 3));
        dw[(o + 1) >> 2] |= (uint32_t)dg << (8 * ((o + 1) & 3));
        dw[(o + 2) >> 2] |= (uint32_t)db << (8 * ((o + 2) & 3));
    }
    if (g.resid) {
        uint8_t *dst = g.resid + ((int64_t)y * p.w + gx0) * 3;
        if (FULL && p.resid_vec_ok) {
#pragma unroll
            for (int i = 0; i < 3; i++)
                reinterpret_cast<U4 *>(dst)[i] = U4{dw[4 * i], dw[4 * i + 1], dw[4 * i + 2], dw[4 * i + 3]};
        } else {
#pragma unroll
            for (int b = 0; b < 48; b++)
                if (b < 3 * nvalid) dst[b] = (uint8_t)byte_of(dw[b >> 2], b & 3);
        }
    }

    // ---- texture: Laplacian of the original luma, BORDER_REFLECT_101 (§8a)
    int lu = l - 1, ld = l + 1;
    if (y == 0) lu = p.h > 1 ? l + 1 : l;
    if (y == p.h - 1) ld = p.h > 1 ? l - 1 : l;
    const uint8_t *yc = &S.yorig[ring16(r, l)][col];
    const U4 cw = *reinterpret_cast<const U4 *>(yc);
    const U4 uw = *reinterpret_cast<const U4 *>(&S.yorig[ring16(r, lu)][col]);
    const U4 lw = *reinterpret_cast<const U4 *>(&S.yorig[ring16(r, ld)][col]);
    const uint32_t cww[4] = {cw.x, cw.y, cw.z, cw.w}, uww[4] = {uw.x, uw.y, uw.z, uw.w}, lww[4] = {lw.x, lw.y, lw.z, lw.w};
    int c[18];                                                  // c[k+1] = luma at column k, k = -1..16
    c[0] = yc[-1];
    c[17] = yc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) c[k + 1] = (int)byte_of(cww[k >> 2], k & 3);
    if (gx0 == 0) c[0] = p.w > 1 ? c[2] : c[1];
    if (!FULL || gx0 + 16 == p.w) {                             // the unit holds the right image edge
        const int ke = p.w - 1 - gx0;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (k == ke) c[k + 2] = p.w > 1 ? c[k] : c[k + 1];
    }
    uint32_t sabs = 0, ssq = 0, mx = acc.tex_maxabs;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int lap = c[k] + c[k + 2] + (int)byte_of(uww[k >> 2], k & 3) + (int)byte_of(lww[k >> 2], k & 3) - 4 * c[k + 1];
        lap = lap < 0 ? -lap : lap;
        if (FULL || k < nvalid) {
            sabs += (uint32_t)lap;
            ssq += (uint32_t)(lap * lap);
            mx = (uint32_t)lap > mx ? (uint32_t)lap : mx;
        }
    }
    acc.tex_sumabs += sabs;
    acc.tex_sumsq += ssq;
    acc.tex_maxabs = mx;
}

// Iteration r finishes pixel rows 16r-1 .. 16r+14 (clipped to the segment and the image).
V5_DEV void stage_residual(int tid, Smem &S, const KParams &p, const Geo &g, ThreadAcc &acc, int r)
{
    const int tw = g.m1 - g.m0;
    const int ylo = 16 * g.r0, yhi = 16 * g.r1 < p.h ? 16 * g.r1 : p.h;
    for (int u = tid; u < 16 * tw; u += NT) {
        const int wl = u / tw, ux = u - wl * tw;
        const int l = wl - 1, y = 16 * r + l;
        const int gx0 = 16 * (g.m0 + ux);
        if (y < ylo || y >= yhi || gx0 >= p.w) continue;
        if (gx0 + 16 <= p.w)
            residual_unit<true>(S, p, g, acc, r, l, ux);
        else
            residual_unit<false>(S, p, g, acc, r, l, ux);
    }
}

// ------------------------------------------------------------------------------------------------ work item set-up
V5_DEV void make_geo(const KParams &p, int work, Geo &g, int &frame)
{
    const int per_frame = p.n_strips * p.n_segs;
    frame = work / per_frame;
    const int rem = work - frame * per_frame;
    const int seg = rem / p.n_strips, strip = rem - seg * p.n_strips;
    g.m0 = (int)(((int64_t)strip * p.mw) / p.n_strips);
    g.m1 = (int)(((int64_t)(strip + 1) * p.mw) / p.n_strips);
    g.r0 = (int)(((int64_t)seg * p.mh) / p.n_segs);
    g.r1 = (int)(((int64_t)(seg + 1) * p.mh) / p.n_segs);
    g.xb0 = 16 * (g.m0 - 1);
    g.band_mcus = g.m1 - g.m0 + 2;
    g.frame = p.rgb + (int64_t)frame * p.frame_stride;
    g.resid = p.residual ? p.residual + (int64_t)frame * p.h * p.w * 3 : nullptr;
}

}  // namespace v5
