// effects.cu -- GaussianBlur / Glow / DropShadow (separable Q8.8 -> Q16.16 fixed point) and the
// pointwise colour effects ChromaKey / FillColor / HSLShift (sm_100a).  C ABI in
// include/movis_b200.h; arithmetic contract in SURVEY.md 8(a)-7..10 and oracle/mvo.c.
//
// Blur layout.  The reference pads the layer by k px per side (RGB edge-replicated, alpha 0) and
// blurs the padded image with cv2.GaussianBlur (reflect-101 border).  Because the pad is wider
// than the kernel radius, that equals blurring a *virtual* image whose RGB is the source clamped
// to its edge and whose alpha is 0 outside the source -- so no padded copy is ever materialised:
//   H pass:  mid[y][X] = sum_j w[j] * tap(X + j - r - k, y)        for the h source rows only
//            (uint16 x C per pixel: sum w = 256 keeps it inside 16 bits)
//   V pass:  out[Y][X] = (sum_i w[i] * mid[clamp(Y + i - r - k)][X] + 32768) >> 16
//            (alpha rows outside the source contribute 0), with the Glow composite fused in.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../include/movis_b200.h"
#include "host_util.h"
#include "pixel_math.cuh"
#include "chroma_key.cuh"

namespace mvb {

constexpr int kMaxKsize = 1023; // weights travel as kernel parameters

struct BlurParams {
    const uint8_t* src;
    int32_t w, h, sstride;
    int32_t k;          // ksize (odd); pad == k
    uint16_t* mid;      // h x (w + 2k) x C uint16
    uint8_t* dst;       // (h + 2k) x (w + 2k), RGBA8 (C == 4) or 1 byte per pixel (C == 1)
    int32_t dstride;
    int32_t mode;       // MVB_BLUR_PLAIN / MVB_BLUR_GLOW
    uint8_t lut[256];   // glow strength LUT
    uint16_t wt[kMaxKsize + 1];
};

// H pass, RGBA.  One thread per (X, source row).
__global__ void __launch_bounds__(256) k_blur_h_rgba(const __grid_constant__ BlurParams P)
{
    const int X = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int W2 = P.w + 2 * P.k;
    if (X >= W2 || y >= P.h) return;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(P.src + (int64_t)y * P.sstride);
    const int r = P.k >> 1;
    const int x0 = X - r - P.k; // source x of tap 0
    uint32_t ar = 0, ag = 0, ab = 0, aa = 0;
    for (int j = 0; j < P.k; j++) {
        const int sx = x0 + j;
        const uint32_t p = __ldg(row + min(max(sx, 0), P.w - 1));
        const uint32_t wj = P.wt[j];
        ar += wj * (p & 0xFFu);
        ag += wj * ((p >> 8) & 0xFFu);
        ab += wj * ((p >> 16) & 0xFFu);
        if ((unsigned)sx < (unsigned)P.w) aa += wj * (p >> 24);
    }
    uint2 o;
    o.x = ar | (ag << 16);
    o.y = ab | (aa << 16);
    reinterpret_cast<uint2*>(P.mid)[(int64_t)y * W2 + X] = o;
}

// V pass, RGBA (+ fused Glow composite).  One thread per output pixel.
__global__ void __launch_bounds__(256) k_blur_v_rgba(const __grid_constant__ BlurParams P)
{
    const int X = blockIdx.x * 32 + (threadIdx.x & 31);
    const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int W2 = P.w + 2 * P.k, H2 = P.h + 2 * P.k;
    if (X >= W2 || Y >= H2) return;
    const uint2* mid = reinterpret_cast<const uint2*>(P.mid);
    const int r = P.k >> 1;
    const int y0 = Y - r - P.k;
    uint32_t ar = 0, ag = 0, ab = 0, aa = 0;
    for (int i = 0; i < P.k; i++) {
        const int sy = y0 + i;
        const uint2 m = __ldg(mid + (int64_t)min(max(sy, 0), P.h - 1) * W2 + X);
        const uint32_t wi = P.wt[i];
        ar += wi * (m.x & 0xFFFFu);
        ag += wi * (m.x >> 16);
        ab += wi * (m.y & 0xFFFFu);
        if ((unsigned)sy < (unsigned)P.h) aa += wi * (m.y >> 16);
    }
    Rgba s{(ar + 32768u) >> 16, (ag + 32768u) >> 16, (ab + 32768u) >> 16, (aa + 32768u) >> 16};
    uint32_t o;
    if (P.mode == MVB_BLUR_GLOW) {
        // pad_image pixel (blur.py:74-76): rgb edge-replicated, alpha 0 outside the source
        const int sx = X - P.k, sy = Y - P.k;
        uint32_t d = __ldg(reinterpret_cast<const uint32_t*>(P.src + (int64_t)min(max(sy, 0), P.h - 1) * P.sstride) +
                           min(max(sx, 0), P.w - 1));
        if ((unsigned)sx >= (unsigned)P.w || (unsigned)sy >= (unsigned)P.h) d &= 0x00FFFFFFu;
        s.r = P.lut[s.r]; s.g = P.lut[s.g]; s.b = P.lut[s.b]; // blur.py:82
        o = over_numpy<MVB_BLEND_LINEAR_DODGE>(d, s, s.a, false); // blur.py:84-85
    } else {
        o = pack(s.r, s.g, s.b, s.a);
    }
    reinterpret_cast<uint32_t*>(P.dst + (int64_t)Y * P.dstride)[X] = o;
}

// ---- tiled RGBA blur (k <= kTiledMaxK, every weight <= 255) ---------------------------------------
// Same arithmetic as k_blur_h_rgba / k_blur_v_rgba; what changes is the data movement.
//   H: a CTA stages one row segment (1024 outputs + k + 3 taps, padding rule applied once) in shared
//      memory as four channel PLANES (4 x 4 byte transposes), so a word holds one channel of four adjacent
//      pixels; every thread produces 4 adjacent outputs, four taps of a channel per IDP.4A (u8 x u8,
//      32-bit accumulate; a Q8 weighted sum of bytes stays below 65 536), weights broadcast from shared memory.
//   V: every thread produces 4 vertically adjacent outputs of one column, so each intermediate
//      row is loaded once per 4 outputs (coalesced 8-byte loads); one IDP.2A per channel and PAIR of taps
//      (two 16-bit values x two 8-bit weights, 32-bit accumulate).
constexpr int kTiledMaxK = 255;
constexpr int kBlurHOut = 1024;   // outputs per CTA of the H pass (4 per thread)

__global__ void __launch_bounds__(256) k_blur_h_rgba_tiled(const __grid_constant__ BlurParams P)
{
    extern __shared__ uint32_t s_dyn[];
    const int k = P.k, r = k >> 1, W2 = P.w + 2 * k;
    const int G = kBlurHOut / 4 + (k + 2) / 4 + 1;         // groups of 4 staged pixels (taps of the last output included)
    uint32_t* s_pl = s_dyn;                                // plane c, group g at s_pl[c * G + g]: channel c of pixels 4 g .. 4 g + 3
    uint4* s_w = reinterpret_cast<uint4*>(s_dyn + 4 * G);  // s_w[q] = for m = 0 .. 3 the bytes w(4 q - m), .., w(4 q + 3 - m)
    const int y = blockIdx.y, X0 = blockIdx.x * kBlurHOut;
    const int t = threadIdx.x;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(P.src + (int64_t)y * P.sstride);
    for (int g = t; g < G; g += 256) {
        uint32_t p[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int sx = X0 + 4 * g + b - r - k;         // source x of staged pixel 4 g + b (blur.py:39-40 padding)
            p[b] = __ldg(row + min(max(sx, 0), P.w - 1));
            if ((unsigned)sx >= (unsigned)P.w) p[b] &= 0x00FFFFFFu;   // rgb edge-replicated, alpha zero-padded
        }
        // 4 x 4 byte transpose: pixels -> channel planes
        const uint32_t t0 = __byte_perm(p[0], p[1], 0x5140), t1 = __byte_perm(p[2], p[3], 0x5140);
        const uint32_t u0 = __byte_perm(p[0], p[1], 0x7362), u1 = __byte_perm(p[2], p[3], 0x7362);
        s_pl[g] = __byte_perm(t0, t1, 0x5410);
        s_pl[G + g] = __byte_perm(t0, t1, 0x7632);
        s_pl[2 * G + g] = __byte_perm(u0, u1, 0x5410);
        s_pl[3 * G + g] = __byte_perm(u0, u1, 0x7632);
    }
    const int Q = (k + 2) / 4 + 1;                         // groups that hold taps of a thread's 4 outputs
    for (int i = t; i < 4 * Q; i += 256) {
        const int q = i >> 2, m = i & 3;
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int tap = 4 * q + b - m;
            if (tap >= 0 && tap < k) w |= (uint32_t)P.wt[tap] << (8 * b);
        }
        reinterpret_cast<uint32_t*>(s_w)[i] = w;
    }
    __syncthreads();
    // Output X0 + 4 t + m takes staged pixel 4 (t + q) + b as tap 4 q + b - m: one IDP.4A per channel, group
    // and output (four taps each), the four weight alignments of a group in one broadcast 128-bit load.
    uint32_t acc[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0u;
    for (int q = 0; q < Q; ++q) {
        const uint4 w = s_w[q];
        const uint32_t wm[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t px = s_pl[c * G + t + q];
#pragma unroll
            for (int m = 0; m < 4; ++m) acc[m][c] = __dp4a(px, wm[m], acc[m][c]);
        }
    }
    uint2* out = reinterpret_cast<uint2*>(P.mid) + (int64_t)y * W2;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int X = X0 + 4 * t + m;
        if (X < W2) out[X] = make_uint2(acc[m][0] | (acc[m][1] << 16), acc[m][2] | (acc[m][3] << 16));
    }
}

constexpr int kBlurVOut = 8;   // vertically adjacent outputs per thread of the V pass

__global__ void __launch_bounds__(256) k_blur_v_rgba_tiled(const __grid_constant__ BlurParams P)
{
    __shared__ __align__(16) uint32_t s_w[kTiledMaxK + 2 * kBlurVOut + 4];   // s_w[t + kBlurVOut - 1] = w(t) | w(t + 1) << 8, w = 0 outside [0, k)
    const int k = P.k, r = k >> 1;
    const int W2 = P.w + 2 * k, H2 = P.h + 2 * k;
    for (int j = threadIdx.x; j < k + 2 * kBlurVOut + 4; j += 256) {
        const int t = j - (kBlurVOut - 1);
        const uint32_t w0 = (t >= 0 && t < k) ? P.wt[t] : 0u, w1 = (t + 1 >= 0 && t + 1 < k) ? P.wt[t + 1] : 0u;
        s_w[j] = w0 | (w1 << 8);
    }
    __syncthreads();
    const int X = blockIdx.x * 32 + (threadIdx.x & 31);
    const int Yb = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kBlurVOut;
    if (X >= W2 || Yb >= H2) return;
    const uint2* mid = reinterpret_cast<const uint2*>(P.mid) + X;
    uint32_t acc[kBlurVOut][4];
#pragma unroll
    for (int m = 0; m < kBlurVOut; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 32768u;   // rounding of (V + 32768) >> 16
    const int y0 = Yb - r - k;
    // Two intermediate rows per step: the same channel of rows j and j + 1 side by side (one PRMT) is the
    // 16-bit pair of one IDP.2A, whose byte pair is (w(j - m), w(j + 1 - m)) for output Yb + m: two taps
    // per instruction.  Row j is tap j - m of output m; rows past the last tap meet zero weights.
    auto taps = [&](const uint2 v0, const uint2 v1, int j) {
        const uint32_t rr = __byte_perm(v0.x, v1.x, 0x5410), gg = __byte_perm(v0.x, v1.x, 0x7632);
        const uint32_t bb = __byte_perm(v0.y, v1.y, 0x5410), aa = __byte_perm(v0.y, v1.y, 0x7632);
        // the weights of the 8 outputs are 8 consecutive words (j is even: four aligned 64-bit loads, broadcast)
        uint32_t wv[kBlurVOut];
#pragma unroll
        for (int q = 0; q < kBlurVOut / 2; ++q) {
            const uint2 t = *reinterpret_cast<const uint2*>(&s_w[j + 2 * q]);
            wv[2 * q] = t.x;
            wv[2 * q + 1] = t.y;
        }
#pragma unroll
        for (int m = 0; m < kBlurVOut; ++m) {
            const uint32_t w = wv[kBlurVOut - 1 - m];
            acc[m][0] = __dp2a_lo(rr, w, acc[m][0]);
            acc[m][1] = __dp2a_lo(gg, w, acc[m][1]);
            acc[m][2] = __dp2a_lo(bb, w, acc[m][2]);
            acc[m][3] = __dp2a_lo(aa, w, acc[m][3]);
        }
    };
    const int n_rows = (k + kBlurVOut) & ~1;                 // rows touched, two per step
    if (y0 >= 0 && y0 + n_rows <= P.h) {
        // every row exists (all but the top and bottom bands of the image): no clamping, no alpha padding.
        // (The row groups of a CTA re-read each other's rows, but through L1: staging the CTA's k + 63 rows in
        // shared memory first was measured 7 % slower.)
        const uint2* p = mid + (int64_t)y0 * W2;
        for (int j = 0; j < n_rows; j += 2, p += 2 * (int64_t)W2) taps(__ldg(p), __ldg(p + W2), j);
    } else {
        for (int j = 0; j < n_rows; j += 2) {
            const int sy0 = y0 + j, sy1 = sy0 + 1;
            uint2 v0 = __ldg(mid + (int64_t)min(max(sy0, 0), P.h - 1) * W2);
            uint2 v1 = __ldg(mid + (int64_t)min(max(sy1, 0), P.h - 1) * W2);
            if ((unsigned)sy0 >= (unsigned)P.h) v0.y &= 0xFFFFu;   // alpha zero-padded
            if ((unsigned)sy1 >= (unsigned)P.h) v1.y &= 0xFFFFu;
            taps(v0, v1, j);
        }
    }
#pragma unroll
    for (int m = 0; m < kBlurVOut; ++m) {
        const int Y = Yb + m;
        if (Y >= H2) break;
        Rgba s{acc[m][0] >> 16, acc[m][1] >> 16, acc[m][2] >> 16, acc[m][3] >> 16};
        uint32_t o;
        if (P.mode == MVB_BLUR_GLOW) {
            const int sx = X - P.k, sy = Y - P.k;          // pad_image pixel (blur.py:74-76)
            uint32_t d = __ldg(reinterpret_cast<const uint32_t*>(P.src + (int64_t)min(max(sy, 0), P.h - 1) * P.sstride) +
                               min(max(sx, 0), P.w - 1));
            if ((unsigned)sx >= (unsigned)P.w || (unsigned)sy >= (unsigned)P.h) d &= 0x00FFFFFFu;
            s.r = P.lut[s.r]; s.g = P.lut[s.g]; s.b = P.lut[s.b]; // blur.py:82 (staging the table in shared memory: measured 20 % slower)
            o = over_numpy<MVB_BLEND_LINEAR_DODGE>(d, s, s.a, false); // blur.py:84-85
        } else {
            o = pack(s.r, s.g, s.b, s.a);
        }
        reinterpret_cast<uint32_t*>(P.dst + (int64_t)Y * P.dstride)[X] = o;
    }
}

// Tiled H / V passes on the alpha channel only (DropShadow, style.py:60-62; zero padding): the same
// staging and 4-outputs-per-thread sliding window as the RGBA kernels, one channel.
__global__ void __launch_bounds__(256) k_blur_h_alpha_tiled(const __grid_constant__ BlurParams P)
{
    extern __shared__ uint32_t s_dyn[];
    const int k = P.k, r = k >> 1, W2 = P.w + 2 * k;
    const int n_in = kBlurHOut + ((k + 3 + 3) & ~3);
    const int Q = n_in >> 2;
    uint32_t* s_px = s_dyn;
    uint32_t* s_w = s_dyn + n_in;
    const int y = blockIdx.y, X0 = blockIdx.x * kBlurHOut;
    const int t = threadIdx.x;
    const uint8_t* row = P.src + (int64_t)y * P.sstride + 3;
    for (int i = t; i < n_in; i += 256) {
        const int sx = X0 + i - r - k;
        s_px[(i & 3) * Q + (i >> 2)] = (unsigned)sx < (unsigned)P.w ? row[(int64_t)sx * 4] : 0u;
    }
    for (int j = t; j < k + 12; j += 256) s_w[j] = (j >= 3 && j < k + 3) ? P.wt[j - 3] : 0u;
    __syncthreads();
    uint32_t a[4] = {0, 0, 0, 0};
    for (int j4 = 0; j4 * 4 < k + 3; ++j4) {
        const uint4 wa = *reinterpret_cast<const uint4*>(s_w + 4 * j4);
        const uint4 wb = *reinterpret_cast<const uint4*>(s_w + 4 * j4 + 4);
        const uint32_t w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const uint32_t p = s_px[jj * Q + t + j4];
#pragma unroll
            for (int m = 0; m < 4; ++m) a[m] += w[jj - m + 3] * p;
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int X = X0 + 4 * t + m;
        if (X < W2) P.mid[(int64_t)y * W2 + X] = (uint16_t)a[m];
    }
}

__global__ void __launch_bounds__(256) k_blur_v_alpha_tiled(const __grid_constant__ BlurParams P)
{
    __shared__ uint32_t s_w[kTiledMaxK + 9];
    const int k = P.k, r = k >> 1;
    const int W2 = P.w + 2 * k, H2 = P.h + 2 * k;
    for (int j = threadIdx.x; j < k + 8; j += 256) s_w[j] = (j >= 3 && j < k + 3) ? P.wt[j - 3] : 0u;
    __syncthreads();
    const int X = blockIdx.x * 32 + (threadIdx.x & 31);
    const int Yb = (blockIdx.y * 8 + (threadIdx.x >> 5)) * 4;
    if (X >= W2 || Yb >= H2) return;
    uint32_t acc[4] = {32768u, 32768u, 32768u, 32768u};
    const int y0 = Yb - r - k;
    for (int j = 0; j < k + 3; ++j) {
        const int sy = y0 + j;
        const uint32_t v = (unsigned)sy < (unsigned)P.h ? P.mid[(int64_t)sy * W2 + X] : 0u;
#pragma unroll
        for (int m = 0; m < 4; ++m) acc[m] += s_w[j - m + 3] * v;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
        if (Yb + m < H2) P.dst[(int64_t)(Yb + m) * P.dstride + X] = (uint8_t)(acc[m] >> 16);
}

// H / V passes on the alpha channel only (DropShadow, style.py:60-62): zero padding.
__global__ void __launch_bounds__(256) k_blur_h_alpha(const __grid_constant__ BlurParams P)
{
    const int X = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int W2 = P.w + 2 * P.k;
    if (X >= W2 || y >= P.h) return;
    const uint8_t* row = P.src + (int64_t)y * P.sstride + 3;
    const int r = P.k >> 1;
    const int x0 = X - r - P.k;
    uint32_t aa = 0;
    for (int j = 0; j < P.k; j++) {
        const int sx = x0 + j;
        if ((unsigned)sx < (unsigned)P.w) aa += (uint32_t)P.wt[j] * row[(int64_t)sx * 4];
    }
    P.mid[(int64_t)y * W2 + X] = (uint16_t)aa;
}

__global__ void __launch_bounds__(256) k_blur_v_alpha(const __grid_constant__ BlurParams P)
{
    const int X = blockIdx.x * 32 + (threadIdx.x & 31);
    const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int W2 = P.w + 2 * P.k, H2 = P.h + 2 * P.k;
    if (X >= W2 || Y >= H2) return;
    const int r = P.k >> 1;
    const int y0 = Y - r - P.k;
    uint32_t aa = 0;
    for (int i = 0; i < P.k; i++) {
        const int sy = y0 + i;
        if ((unsigned)sy < (unsigned)P.h) aa += (uint32_t)P.wt[i] * P.mid[(int64_t)sy * W2 + X];
    }
    P.dst[(int64_t)Y * P.dstride + X] = (uint8_t)((aa + 32768u) >> 16);
}

struct ShadowParams {
    const uint8_t* src;
    int32_t w, h, sstride;
    const uint8_t* a1;  // blurred alpha plane (Ws x Hs) or nullptr when k == 0
    int32_t Ws, Hs;
    uint8_t* dst;
    int32_t Wn, Hn, dstride;
    int32_t sx0, sy0;   // where the shadow is pasted (style.py:79-80)
    int32_t px, py;     // where the original is composited (style.py:83)
    uint32_t color;     // packed rgb
    uint8_t lut[256];   // uint8(opacity * v)
};

// DropShadow assembly (style.py:64-84): shadow pixel, then PIL-over of the original at the centre.
__global__ void __launch_bounds__(256) k_drop_shadow(const __grid_constant__ ShadowParams P)
{
    const int X = blockIdx.x * 32 + (threadIdx.x & 31);
    const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (X >= P.Wn || Y >= P.Hn) return;
    uint32_t d = 0;
    const int xs = X - P.sx0, ys = Y - P.sy0;
    if ((unsigned)xs < (unsigned)P.Ws && (unsigned)ys < (unsigned)P.Hs) {
        uint32_t a;
        if (P.a1) a = P.a1[(int64_t)ys * P.Ws + xs];
        else a = P.lut[P.src[(int64_t)ys * P.sstride + (int64_t)xs * 4 + 3]]; // style.py:64 (radius == 0)
        d = P.color | ((uint32_t)P.lut[a] << 24);                                // style.py:68
    }
    const int xo = X - P.px, yo = Y - P.py;
    if ((unsigned)xo < (unsigned)P.w && (unsigned)yo < (unsigned)P.h) {
        const uint32_t spx = __ldg(reinterpret_cast<const uint32_t*>(P.src + (int64_t)yo * P.sstride) + xo);
        const Rgba s = unpack(spx);
        d = over_pil(d, s, s.a);
    }
    reinterpret_cast<uint32_t*>(P.dst + (int64_t)Y * P.dstride)[X] = d;
}

// ---- colour conversions -----------------------------------------------------------------------

struct HsvTables { int32_t sdiv[256]; int32_t hdiv[256]; };
__device__ HsvTables g_hsv;   // global memory: every thread reads its own entry (a constant-bank read would replay per address)

// cv2.COLOR_RGB2HSV, 8-bit, H in [0, 180)
__device__ __forceinline__ void rgb_to_hsv(int r, int g, int b, const int32_t* sdiv, const int32_t* hdiv,
                                           int& h, int& s, int& v)
{
    v = max(max(r, g), b);
    const int vmin = min(min(r, g), b);
    const int diff = v - vmin;
    s = (diff * sdiv[v] + (1 << 11)) >> 12;
    int hh;
    if (v == r) hh = g - b;
    else if (v == g) hh = b - r + 2 * diff;
    else hh = r - g + 4 * diff;
    hh = (hh * hdiv[diff] + (1 << 11)) >> 12;
    if (hh < 0) hh += 180;
    h = min(max(hh, 0), 255);
}

// cv2.COLOR_HSV2RGB, 8-bit: float32 sector formula with one fused multiply-add per tab entry;
// `round_result` selects the rounding cv2 applies in the scalar tail of a row (see oracle/mvo.c).
__device__ __forceinline__ uint32_t hsv_to_rgb(int hi, int si, int vi, bool round_result)
{
    const float s = __fmul_rn((float)si, 1.0f / 255.0f), v = __fmul_rn((float)vi, 1.0f / 255.0f);
    float r, g, b;
    if (si == 0) {
        r = g = b = v;
    } else {
        float h = __fmul_rn((float)hi, 6.0f / 180.0f);
        int sector = (int)floorf(h);
        h = __fsub_rn(h, (float)sector);
        sector %= 6;
        const float t1 = __fmul_rn(v, __fsub_rn(1.0f, s));
        const float t2 = __fmul_rn(v, __fmaf_rn(-s, h, 1.0f));
        const float t3 = __fmul_rn(v, __fmaf_rn(-s, __fsub_rn(1.0f, h), 1.0f));
        switch (sector) { // sector_data of cv::HSV2RGB_native, (b, g, r) rows
        case 0: b = t1; g = t3; r = v; break;
        case 1: b = t1; g = v; r = t2; break;
        case 2: b = t3; g = v; r = t1; break;
        case 3: b = v; g = t2; r = t1; break;
        case 4: b = v; g = t1; r = t3; break;
        default: b = t2; g = t1; r = v; break;
        }
    }
    r = __fmul_rn(r, 255.0f); g = __fmul_rn(g, 255.0f); b = __fmul_rn(b, 255.0f);
    int ri, gi, bi;
    if (round_result) { ri = __float2int_rn(r); gi = __float2int_rn(g); bi = __float2int_rn(b); }
    else { ri = __float2int_rz(r); gi = __float2int_rz(g); bi = __float2int_rz(b); }
    return (uint32_t)min(max(ri, 0), 255) | ((uint32_t)min(max(gi, 0), 255) << 8) |
           ((uint32_t)min(max(bi, 0), 255) << 16);
}

struct PointParams {
    const uint8_t* src;
    uint8_t* dst;
    int32_t w, h, sstride, dstride;
    uint32_t rcp_mul, rcp_shift;   // i / vectors_per_row == umulhi(i, rcp_mul) >> rcp_shift (rcp_mul == 0: one vector per row)
    int32_t op;          // 0 chroma key, 1 fill colour, 2 HSL shift
    uint32_t color;      // fill colour (packed rgb)
    uint8_t lo[4], hi[4];
    uint8_t hlut[256], slut[256], vlut[256];
};

// One pixel of ChromaKey (op 0) / FillColor (op 1) / HSLShift (op 2).
__device__ __forceinline__ uint32_t pointwise_px(const PointParams& P, uint32_t p, int x, const int32_t* sdiv,
                                                 const int32_t* hdiv, const uint8_t* luts)
{
    if (P.op == 1) return P.color | (p & 0xFF000000u);
    int hh, ss, vv;
    rgb_to_hsv(p & 0xFF, (p >> 8) & 0xFF, (p >> 16) & 0xFF, sdiv, hdiv, hh, ss, vv);
    if (P.op == 0) {
        const bool in = hh >= P.lo[0] && hh <= P.hi[0] && ss >= P.lo[1] && ss <= P.hi[1] &&
                        vv >= P.lo[2] && vv <= P.hi[2];
        return (p & 0x00FFFFFFu) | (in ? 0u : 0xFF000000u);
    }
    const bool tail = x >= P.w - (P.w & 31);   // cv2 rounds instead of truncating in the scalar tail of a row
    return hsv_to_rgb(luts[hh], luts[256 + ss], luts[512 + vv], tail) | (p & 0xFF000000u);
}

// Persistent grid-stride kernel: the tables are staged in shared memory once per CTA, pixels move
// as 128-bit vectors (VEC = 4) when rows are 16-byte aligned.
template <int VEC>
__global__ void __launch_bounds__(256) k_pointwise(const __grid_constant__ PointParams P)
{
    __shared__ int32_t s_sdiv[256], s_hdiv[256];
    __shared__ uint8_t s_luts[768];
    if (P.op != 1) {
        s_sdiv[threadIdx.x] = g_hsv.sdiv[threadIdx.x];
        s_hdiv[threadIdx.x] = g_hsv.hdiv[threadIdx.x];
        s_luts[threadIdx.x] = P.hlut[threadIdx.x];
        s_luts[256 + threadIdx.x] = P.slut[threadIdx.x];
        s_luts[512 + threadIdx.x] = P.vlut[threadIdx.x];
        __syncthreads();
    }
    const uint32_t per_row = (uint32_t)(P.w + VEC - 1) / VEC;
    const uint32_t total = per_row * (uint32_t)P.h;                    // (the host keeps it below 2^31)
    for (uint32_t i = blockIdx.x * 256u + threadIdx.x; i < total; i += gridDim.x * 256u) {
        // row / column of vector i by a multiply-high with the host's reciprocal of the row length (no 64-bit division)
        const uint32_t yy = P.rcp_mul ? __umulhi(i, P.rcp_mul) >> P.rcp_shift : i;
        const int y = (int)yy, x = (int)(i - yy * per_row) * VEC;
        const uint8_t* sp = P.src + (int64_t)y * P.sstride + (int64_t)x * 4;
        uint8_t* dp = P.dst + (int64_t)y * P.dstride + (int64_t)x * 4;
        if (VEC == 4 && x + 4 <= P.w) {
            const uint4 v = *reinterpret_cast<const uint4*>(sp);
            uint4 o;
            o.x = pointwise_px(P, v.x, x, s_sdiv, s_hdiv, s_luts);
            o.y = pointwise_px(P, v.y, x + 1, s_sdiv, s_hdiv, s_luts);
            o.z = pointwise_px(P, v.z, x + 2, s_sdiv, s_hdiv, s_luts);
            o.w = pointwise_px(P, v.w, x + 3, s_sdiv, s_hdiv, s_luts);
            *reinterpret_cast<uint4*>(dp) = o;
        } else {
            for (int k = 0; k < VEC && x + k < P.w; ++k)
                reinterpret_cast<uint32_t*>(dp)[k] =
                    pointwise_px(P, reinterpret_cast<const uint32_t*>(sp)[k], x + k, s_sdiv, s_hdiv, s_luts);
        }
    }
}

// ChromaKey as its own streaming kernel: the range test of chroma_key.cuh on 128-bit vectors.  Persistent
// CTAs walk the image 512 vectors at a time (two per thread, 256 apart) and load the next pair before
// they key the current one, so every thread keeps four loads in flight; the vector index is split into
// (row, column) by a multiply-high with the host's reciprocal of the row length.
struct ChromaParams {
    const uint8_t* src;
    uint8_t* dst;
    int32_t w, h, sstride, dstride;
    uint32_t per_row, total;      // vectors per row, vectors in the image
    uint32_t rcp_mul, rcp_shift;  // i / per_row == umulhi(i, rcp_mul) >> rcp_shift for i < total
    const ChromaTables* tab;      // in global memory (chroma_tables_on_device): 3 KB, read once per CTA, coalesced
};

template <int VEC>
__global__ void __launch_bounds__(256) k_chroma_key(const __grid_constant__ ChromaParams P)
{
    __shared__ uint32_t s_sv[256];
    __shared__ uint2 s_hq[256];
    s_sv[threadIdx.x] = __ldg(&P.tab->sv[threadIdx.x]);
    s_hq[threadIdx.x] = __ldg(&P.tab->hq[threadIdx.x]);
    __syncthreads();
    typedef typename std::conditional<VEC == 4, uint4, uint32_t>::type Vec;
    auto locate = [&](uint32_t i, int64_t& so, int64_t& dof) {
        const uint32_t y = P.rcp_mul ? __umulhi(i, P.rcp_mul) >> P.rcp_shift : i, x = i - y * P.per_row;
        so = (int64_t)y * P.sstride + (int64_t)x * (4 * VEC);
        dof = (int64_t)y * P.dstride + (int64_t)x * (4 * VEC);
    };
    auto key = [&](uint32_t p) { return (p & 0x00FFFFFFu) | (chroma_in_range(p, s_sv, s_hq) ? 0u : 0xFF000000u); };
    const uint32_t step = gridDim.x * 512u;
    uint32_t i = blockIdx.x * 512u + threadIdx.x;
    Vec cur[2], nxt[2];
    int64_t dcur[2], dnxt[2];
    auto fetch = [&](uint32_t at, Vec (&v)[2], int64_t (&d)[2]) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            d[k] = -1;
            if (at + 256u * k < P.total) {
                int64_t so;
                locate(at + 256u * k, so, d[k]);
                v[k] = __ldcs(reinterpret_cast<const Vec*>(P.src + so));
            }
        }
    };
    if (i < P.total) fetch(i, cur, dcur);
    while (i < P.total) {
        const uint32_t in = i + step;
        dnxt[0] = dnxt[1] = -1;
        if (in < P.total && in > i) fetch(in, nxt, dnxt);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (dcur[k] < 0) continue;
            if constexpr (VEC == 4) {
                const uint4 o = make_uint4(key(cur[k].x), key(cur[k].y), key(cur[k].z), key(cur[k].w));
                *reinterpret_cast<uint4*>(P.dst + dcur[k]) = o;
            } else {
                *reinterpret_cast<uint32_t*>(P.dst + dcur[k]) = key(cur[k]);
            }
        }
        if (in <= i) break;   // (index overflow guard)
#pragma unroll
        for (int k = 0; k < 2; ++k) { cur[k] = nxt[k]; dcur[k] = dnxt[k]; }
        i = in;
    }
}

// ---- host side --------------------------------------------------------------------------------

static void host_hsv_tables(HsvTables& t)
{
    t.sdiv[0] = t.hdiv[0] = 0;
    for (int i = 1; i < 256; i++) {
        t.sdiv[i] = (int32_t)std::nearbyint((255 << 12) / (1. * i));
        t.hdiv[i] = (int32_t)std::nearbyint((180 << 12) / (6. * i));
    }
}

// cv2's division tables, uploaded once per device per process.  The copy comes from pageable memory on
// the legacy default stream and may return before the DMA has landed, while the kernels that read the
// tables run on non-blocking streams: the device is synchronised once, right after the copy, under a
// process-wide lock (no other thread can launch a reader before the flag is set).
static int upload_hsv_tables()
{
    static std::mutex mu;
    static uint64_t uploaded_mask[2] = {0, 0};   // devices 0..127
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 128) return set_error(MVB_ERR_NO_DEVICE, "no current CUDA device");
    std::lock_guard<std::mutex> lock(mu);
    if (uploaded_mask[dev >> 6] >> (dev & 63) & 1) return MVB_OK;
    static HsvTables t;
    static bool built = false;
    if (!built) { host_hsv_tables(t); built = true; }
    cudaError_t e = cudaMemcpyToSymbol(g_hsv, &t, sizeof t);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return set_error(MVB_ERR_CUDA, cudaGetErrorString(e));
    uploaded_mask[dev >> 6] |= 1ull << (dev & 63);
    return MVB_OK;
}

static int check_rgba(const void* p, int w, int h, int64_t stride, const char* what)
{
    if (!p || w <= 0 || h <= 0 || stride < (int64_t)w * 4 || (stride & 3) || stride > 0x7FFFFFFF ||
        (reinterpret_cast<uintptr_t>(p) & 3)) {
        char buf[160];
        std::snprintf(buf, sizeof buf, "%s: need non-null 4-byte-aligned RGBA8 image with stride >= 4*width", what);
        return set_error(MVB_ERR_INVALID_ARG, buf);
    }
    return 0;
}

static int check_ksize(int k, const uint16_t* weights)
{
    if (k < 3 || !(k & 1) || k > kMaxKsize)
        return set_error(MVB_ERR_UNSUPPORTED, "ksize must be odd and in [3, 1023]");
    if (!weights) return set_error(MVB_ERR_INVALID_ARG, "weights is null");
    return 0;
}

} // namespace mvb

using namespace mvb;

extern "C" int mvb_blur_ksize(double radius)
{
    double v = (4 * radius - 1) / 2;
    if (v < 1) v = 1;
    return 2 * (int)std::ceil(v) + 1;
}

extern "C" int mvb_gaussian_weights_q8(double sigma, int32_t k, uint16_t* w)
{
    if (!(sigma > 0) || k < 1 || !(k & 1) || !w) return set_error(MVB_ERR_INVALID_ARG, "need sigma > 0, odd ksize");
    // cv::getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED(fractionBits = 8) in IEEE double
    const int n2 = (k - 1) / 2;
    std::vector<double> val((size_t)n2 + 1);
    const double scale2X = -0.125 / (sigma * sigma);
    double sum = 0;
    for (int i = 0, x = 1 - k; i < n2; i++, x += 2) {
        const double t = std::exp((double)(x * x) * scale2X);
        val[i] = t;
        sum += t;
    }
    sum *= 2;
    sum += 1;
    const double mul1 = 1.0 / sum;
    double err = 0;
    int64_t isum = 0;
    for (int i = 0; i < n2; i++) {
        const double adj = (val[i] * mul1) * 256.0 + err;
        const int64_t v0 = (int64_t)std::nearbyint(adj);
        err = adj - (double)v0;
        w[i] = w[k - 1 - i] = (uint16_t)v0;
        isum += v0;
    }
    w[n2] = (uint16_t)(256 - 2 * isum);
    return MVB_OK;
}

extern "C" size_t mvb_effect_blur_scratch_bytes(int32_t w, int32_t h, int32_t k, int32_t channels)
{
    if (w <= 0 || h <= 0 || k < 0 || (channels != 1 && channels != 4)) return 0;
    size_t b = (size_t)h * (size_t)(w + 2 * k) * (size_t)channels * 2;
    return (b + 255) & ~(size_t)255;
}

extern "C" int mvb_effect_blur_rgba8(const void* src, int32_t w, int32_t h, int64_t sstride, void* dst,
                                     int64_t dstride, int32_t k, const uint16_t* weights, int32_t mode,
                                     const uint8_t* strength_lut, void* scratch, mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_rgba(src, w, h, sstride, "src")) return rc;
    if (int rc = check_ksize(k, weights)) return rc;
    if (int rc = check_rgba(dst, w + 2 * k, h + 2 * k, dstride, "dst")) return rc;
    if (!scratch || (reinterpret_cast<uintptr_t>(scratch) & 7)) return set_error(MVB_ERR_INVALID_ARG, "scratch must be 8-byte aligned");
    if (mode != MVB_BLUR_PLAIN && mode != MVB_BLUR_GLOW) return set_error(MVB_ERR_INVALID_ARG, "unknown blur mode");
    if (mode == MVB_BLUR_GLOW && !strength_lut) return set_error(MVB_ERR_INVALID_ARG, "glow needs strength_lut");
    BlurParams P;
    P.src = static_cast<const uint8_t*>(src);
    P.w = w; P.h = h; P.sstride = (int32_t)sstride; P.k = k;
    P.mid = static_cast<uint16_t*>(scratch);
    P.dst = static_cast<uint8_t*>(dst);
    P.dstride = (int32_t)dstride;
    P.mode = mode;
    if (strength_lut) std::memcpy(P.lut, strength_lut, 256); else std::memset(P.lut, 0, 256);
    std::memcpy(P.wt, weights, sizeof(uint16_t) * (size_t)k);
    const int W2 = w + 2 * k, H2 = h + 2 * k;
    cudaStream_t st = (cudaStream_t)stream;
    bool tiled = k <= kTiledMaxK;
    for (int j = 0; j < k; j++) tiled = tiled && weights[j] <= 255;   // IDP.2A takes 8-bit weights
    if (tiled) {
        const int groups = kBlurHOut / 4 + (k + 2) / 4 + 1, wq = (k + 2) / 4 + 1;
        k_blur_h_rgba_tiled<<<dim3((W2 + kBlurHOut - 1) / kBlurHOut, h), 256, (4 * groups + 4 * wq) * 4, st>>>(P);
        count_launches(1);
        k_blur_v_rgba_tiled<<<dim3((W2 + 31) / 32, (H2 + 8 * kBlurVOut - 1) / (8 * kBlurVOut)), 256, 0, st>>>(P);
        count_launches(1);
    } else {
        k_blur_h_rgba<<<dim3((W2 + 31) / 32, (h + 7) / 8), 256, 0, st>>>(P);
        count_launches(1);
        k_blur_v_rgba<<<dim3((W2 + 31) / 32, (H2 + 7) / 8), 256, 0, st>>>(P);
        count_launches(1);
    }
    return check_launch("mvb_effect_blur_rgba8");
}

extern "C" int mvb_effect_drop_shadow_rgba8(const void* src, int32_t w, int32_t h, int64_t sstride, void* dst,
                                            int64_t dstride, int32_t k, const uint16_t* weights,
                                            const uint8_t* opacity_lut, const uint8_t color[3], int32_t vx,
                                            int32_t vy, void* scratch, mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_rgba(src, w, h, sstride, "src")) return rc;
    if (k != 0) if (int rc = check_ksize(k, weights)) return rc;
    if (!opacity_lut || !color) return set_error(MVB_ERR_INVALID_ARG, "opacity_lut / color is null");
    const int Ws = w + 2 * k, Hs = h + 2 * k;
    const int dx = vx < 0 ? -vx : vx, dy = vy < 0 ? -vy : vy;
    const int Wn = Ws + 2 * dx, Hn = Hs + 2 * dy;
    if (int rc = check_rgba(dst, Wn, Hn, dstride, "dst")) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ShadowParams S;
    S.a1 = nullptr;
    if (k != 0) {
        if (!scratch || (reinterpret_cast<uintptr_t>(scratch) & 7)) return set_error(MVB_ERR_INVALID_ARG, "scratch must be 8-byte aligned");
        BlurParams P;
        P.src = static_cast<const uint8_t*>(src);
        P.w = w; P.h = h; P.sstride = (int32_t)sstride; P.k = k;
        P.mid = static_cast<uint16_t*>(scratch);
        uint8_t* plane = static_cast<uint8_t*>(scratch) + mvb_effect_blur_scratch_bytes(w, h, k, 1);
        P.dst = plane;
        P.dstride = Ws;
        P.mode = MVB_BLUR_PLAIN;
        std::memset(P.lut, 0, 256);
        std::memcpy(P.wt, weights, sizeof(uint16_t) * (size_t)k);
        if (k <= kTiledMaxK) {
            const int n_in = kBlurHOut + ((k + 3 + 3) & ~3);
            k_blur_h_alpha_tiled<<<dim3((Ws + kBlurHOut - 1) / kBlurHOut, h), 256, (n_in + k + 12) * 4, st>>>(P);
            count_launches(1);
            k_blur_v_alpha_tiled<<<dim3((Ws + 31) / 32, (Hs + 31) / 32), 256, 0, st>>>(P);
            count_launches(1);
        } else {
            k_blur_h_alpha<<<dim3((Ws + 31) / 32, (h + 7) / 8), 256, 0, st>>>(P);
            count_launches(1);
            k_blur_v_alpha<<<dim3((Ws + 31) / 32, (Hs + 7) / 8), 256, 0, st>>>(P);
            count_launches(1);
        }
        S.a1 = plane;
    }
    S.src = static_cast<const uint8_t*>(src);
    S.w = w; S.h = h; S.sstride = (int32_t)sstride;
    S.Ws = Ws; S.Hs = Hs;
    S.dst = static_cast<uint8_t*>(dst);
    S.Wn = Wn; S.Hn = Hn; S.dstride = (int32_t)dstride;
    S.sx0 = dx + vx; S.sy0 = dy + vy;
    S.px = (Wn - w) / 2; S.py = (Hn - h) / 2;
    S.color = (uint32_t)color[0] | ((uint32_t)color[1] << 8) | ((uint32_t)color[2] << 16);
    std::memcpy(S.lut, opacity_lut, 256);
    k_drop_shadow<<<dim3((Wn + 31) / 32, (Hn + 7) / 8), 256, 0, st>>>(S);
    count_launches(1);
    return check_launch("mvb_effect_drop_shadow_rgba8");
}

// floor(i / d) = umulhi(i, m) >> (l - 1) with l = ceil(log2 d), m = ceil(2^(31 + l) / d) <= 2^32 - 1 for d > 1:
// m d - 2^(31 + l) < d <= 2^l, so the error term i (m d - 2^(31 + l)) stays below 2^(31 + l) for i < 2^31.
// d == 1 is flagged by m == 0.
static void row_reciprocal(uint64_t d, uint32_t& mul, uint32_t& shift)
{
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    mul = d == 1 ? 0u : (uint32_t)(((1ull << (31 + l)) + d - 1) / d);
    shift = d == 1 ? 0u : l - 1;
}

static int launch_pointwise(PointParams& P, const void* src, int w, int h, int64_t sstride, void* dst,
                            int64_t dstride, mvb_stream_t stream, const char* what)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_rgba(src, w, h, sstride, "src")) return rc;
    if (int rc = check_rgba(dst, w, h, dstride, "dst")) return rc;
    if (P.op != 1) if (int rc = upload_hsv_tables()) return rc;
    P.src = static_cast<const uint8_t*>(src);
    P.dst = static_cast<uint8_t*>(dst);
    P.w = w; P.h = h; P.sstride = (int32_t)sstride; P.dstride = (int32_t)dstride;
    const bool vec = !((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) |
                        (uintptr_t)sstride | (uintptr_t)dstride) & 15);
    const long long per_row = (w + (vec ? 3 : 0)) / (vec ? 4 : 1), work = per_row * h;
    if (work >= (1ll << 31)) return set_error(MVB_ERR_UNSUPPORTED, "image too large");
    row_reciprocal((uint64_t)per_row, P.rcp_mul, P.rcp_shift);
    const int blocks = (int)std::min<long long>((work + 255) / 256, 148 * 8);
    if (vec) k_pointwise<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(P);
    else k_pointwise<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(P);
    count_launches(1);
    return check_launch(what);
}

// The bounds tables of a key, built once and kept in device memory for the life of the process (an effect
// instance keeps its key; a program uses a handful).  Passing the 3 KB as a kernel parameter made every CTA
// copy them out of the constant bank with a different address per lane, i.e. serially.  *out stays null
// when build_chroma_tables refuses the key.
static int chroma_tables_on_device(const uint8_t lo[3], const uint8_t hi[3], cudaStream_t stream, const ChromaTables** out)
{
    static std::mutex mu;
    static std::vector<std::pair<uint64_t, const ChromaTables*>> known;   // (device << 48 | lo << 24 | hi) -> tables
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t key = ((uint64_t)dev << 48) | ((uint64_t)lo[0] << 40) | ((uint64_t)lo[1] << 32) | ((uint64_t)lo[2] << 24) |
                         ((uint64_t)hi[0] << 16) | ((uint64_t)hi[1] << 8) | (uint64_t)hi[2];
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& e : known)
        if (e.first == key) { *out = e.second; return MVB_OK; }
    static ChromaTables host;
    const ChromaTables* devp = nullptr;
    if (build_chroma_tables(lo, hi, host)) {
        void* p = nullptr;
        if (cudaMalloc(&p, sizeof host) != cudaSuccess) { cudaGetLastError(); return set_error(MVB_ERR_CUDA, "no device memory for the chroma key tables"); }
        // visible to every stream before the pointer is published: copy on the caller's stream, then wait for it
        cudaError_t e = cudaMemcpyAsync(p, &host, sizeof host, cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) { cudaFree(p); return set_error(MVB_ERR_CUDA, cudaGetErrorString(e)); }
        devp = static_cast<const ChromaTables*>(p);
    }
    if (known.size() >= 256) {   // (unbounded key churn: start over; the old tables may still be in use, so they are kept)
        known.clear();
    }
    known.emplace_back(key, devp);
    *out = devp;
    return MVB_OK;
}

extern "C" int mvb_effect_chroma_key_rgba8(const void* src, int32_t w, int32_t h, int64_t sstride, void* dst,
                                           int64_t dstride, const uint8_t lo[3], const uint8_t hi[3],
                                           mvb_stream_t stream)
{
    if (!lo || !hi) return set_error(MVB_ERR_INVALID_ARG, "bounds are null");
    if (int rc = require_device()) return rc;
    const ChromaTables* tab = nullptr;
    if (int rc = chroma_tables_on_device(lo, hi, (cudaStream_t)stream, &tab)) return rc;
    if (!tab) {   // (this key's in-range set is not an arc: keep the full conversion)
        PointParams P;
        std::memset(&P, 0, sizeof P);
        P.op = 0;
        std::memcpy(P.lo, lo, 3);
        std::memcpy(P.hi, hi, 3);
        return launch_pointwise(P, src, w, h, sstride, dst, dstride, stream, "mvb_effect_chroma_key_rgba8");
    }
    ChromaParams C;
    C.tab = tab;
    if (int rc = check_rgba(src, w, h, sstride, "src")) return rc;
    if (int rc = check_rgba(dst, w, h, dstride, "dst")) return rc;
    C.src = static_cast<const uint8_t*>(src);
    C.dst = static_cast<uint8_t*>(dst);
    C.w = w; C.h = h; C.sstride = (int32_t)sstride; C.dstride = (int32_t)dstride;
    const bool vec = !((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) |
                        (uintptr_t)sstride | (uintptr_t)dstride) & 15) && !(w & 3);
    const uint64_t per_row = vec ? (uint64_t)w / 4 : (uint64_t)w, total = per_row * (uint64_t)h;
    if (total >= (1ull << 31)) return set_error(MVB_ERR_UNSUPPORTED, "image too large");
    C.per_row = (uint32_t)per_row;
    C.total = (uint32_t)total;
    row_reciprocal(per_row, C.rcp_mul, C.rcp_shift);
    int n_sm = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const unsigned grid = (unsigned)std::min<uint64_t>((total + 511) / 512, (uint64_t)n_sm * 6);
    if (vec) k_chroma_key<4><<<grid, 256, 0, (cudaStream_t)stream>>>(C);
    else k_chroma_key<1><<<grid, 256, 0, (cudaStream_t)stream>>>(C);
    count_launches(1);
    return check_launch("mvb_effect_chroma_key_rgba8");
}

extern "C" int mvb_effect_fill_color_rgba8(const void* src, int32_t w, int32_t h, int64_t sstride, void* dst,
                                           int64_t dstride, const uint8_t color[3], mvb_stream_t stream)
{
    if (!color) return set_error(MVB_ERR_INVALID_ARG, "color is null");
    PointParams P;
    std::memset(&P, 0, sizeof P);
    P.op = 1;
    P.color = (uint32_t)color[0] | ((uint32_t)color[1] << 8) | ((uint32_t)color[2] << 16);
    return launch_pointwise(P, src, w, h, sstride, dst, dstride, stream, "mvb_effect_fill_color_rgba8");
}

extern "C" int mvb_effect_hsl_shift_rgba8(const void* src, int32_t w, int32_t h, int64_t sstride, void* dst,
                                          int64_t dstride, const uint8_t h_lut[256], const uint8_t s_lut[256],
                                          const uint8_t v_lut[256], mvb_stream_t stream)
{
    if (!h_lut || !s_lut || !v_lut) return set_error(MVB_ERR_INVALID_ARG, "lut is null");
    PointParams P;
    std::memset(&P, 0, sizeof P);
    P.op = 2;
    std::memcpy(P.hlut, h_lut, 256);
    std::memcpy(P.slut, s_lut, 256);
    std::memcpy(P.vlut, v_lut, 256);
    return launch_pointwise(P, src, w, h, sstride, dst, dstride, stream, "mvb_effect_hsl_shift_rgba8");
}

// ---------------------------------------------------------------------------------------------
// Stripe source (movis/layer/texture.py:159-191): a procedural image, pure float64 numpy in the
// reference.  Every operation below is the IEEE double operation numpy performs, in numpy's order,
// with no FMA contraction (__dmul_rn / __dadd_rn / __ddiv_rn), so the pixels are bit-identical.
// ---------------------------------------------------------------------------------------------
struct StripeParams {
    uint8_t* dst;
    int64_t stride;
    int32_t w, h;
    int32_t solid;      // 1 / 2: ratio <= 0 / ratio >= 1, the whole image is colour 1 / colour 2
    int32_t pad_;
    double v0, v1;      // sin(theta), cos(theta)
    double cy, cx;      // height / 2, width / 2
    double width;       // total_width
    double phase;
    double edge0, span; // smoothstep: max(ratio - eps, 0), edge1 - edge0
    double c1[4], c2[4];
};

__global__ void __launch_bounds__(256)
k_stripe(const __grid_constant__ StripeParams P)
{
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= P.w || y >= P.h) return;
    uint32_t px = 0;
    if (P.solid) {
        const double* c = P.solid == 1 ? P.c1 : P.c2;
#pragma unroll
        for (int k = 0; k < 4; ++k) px |= ((uint32_t)(int)c[k] & 0xFFu) << (8 * k);   // astype(np.uint8)
    } else {
        const double iy = __dsub_rn((double)y, P.cy), ix = __dsub_rn((double)x, P.cx);    // np.mgrid - center
        double p = __dadd_rn(__dmul_rn(P.v0, iy), __dmul_rn(P.v1, ix));                   // (v * inds).sum(axis=0)
        p = __dadd_rn(__ddiv_rn(p, P.width), P.phase);                                    // / stripe_width + phase
        p = __dsub_rn(p, floor(p));                                                       // _fract
        double t = __ddiv_rn(__dsub_rn(p, P.edge0), P.span);                              // _smoothstep
        t = fmin(fmax(t, 0.0), 1.0);                                                      // np.clip
        const double sm = __dmul_rn(__dmul_rn(t, t), __dsub_rn(3.0, __dmul_rn(2.0, t)));
        const double om = __dsub_rn(1.0, sm);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double c = __dadd_rn(__dmul_rn(sm, P.c1[k]), __dmul_rn(om, P.c2[k]));
            px |= ((uint32_t)(int)c & 0xFFu) << (8 * k);                                  // astype(np.uint8): truncation
        }
    }
    reinterpret_cast<uint32_t*>(P.dst + (int64_t)y * P.stride)[x] = px;
}

extern "C" int mvb_source_stripe_rgba8(void* dst, int32_t w, int32_t h, int64_t dst_stride, const double params[16],
                                       int32_t solid, mvb_stream_t stream)
{
    if (!dst || !params || w <= 0 || h <= 0 || dst_stride < 4 * (int64_t)w || (dst_stride & 3) ||
        ((uintptr_t)dst & 3) || solid < 0 || solid > 2)
        return set_error(MVB_ERR_INVALID_ARG, "bad stripe image / parameters");
    if (int rc = require_device()) return rc;
    StripeParams P;
    std::memset(&P, 0, sizeof P);
    P.dst = static_cast<uint8_t*>(dst);
    P.stride = dst_stride;
    P.w = w;
    P.h = h;
    P.solid = solid;
    P.v0 = params[0]; P.v1 = params[1]; P.cy = params[2]; P.cx = params[3];
    P.width = params[4]; P.phase = params[5]; P.edge0 = params[6]; P.span = params[7];
    for (int k = 0; k < 4; k++) { P.c1[k] = params[8 + k]; P.c2[k] = params[12 + k]; }
    k_stripe<<<dim3((unsigned)((w + 63) / 64), (unsigned)((h + 3) / 4)), 256, 0, (cudaStream_t)stream>>>(P);
    count_launches(1);
    return check_launch("mvb_source_stripe_rgba8");
}

extern "C" int mvb_rgb_to_hsv_u8(const uint8_t rgb[3], uint8_t hsv[3])
{
    if (!rgb || !hsv) return set_error(MVB_ERR_INVALID_ARG, "null argument");
    static HsvTables t;
    static bool ready = false;
    if (!ready) { host_hsv_tables(t); ready = true; }
    const int r = rgb[0], g = rgb[1], b = rgb[2];
    const int v = r > g ? (r > b ? r : b) : (g > b ? g : b);
    const int vmin = r < g ? (r < b ? r : b) : (g < b ? g : b);
    const int diff = v - vmin;
    const int s = (diff * t.sdiv[v] + (1 << 11)) >> 12;
    int h;
    if (v == r) h = g - b;
    else if (v == g) h = b - r + 2 * diff;
    else h = r - g + 4 * diff;
    h = (h * t.hdiv[diff] + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    hsv[0] = (uint8_t)(h < 0 ? 0 : (h > 255 ? 255 : h));
    hsv[1] = (uint8_t)s;
    hsv[2] = (uint8_t)v;
    return MVB_OK;
}
