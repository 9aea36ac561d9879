// pixel_math.cuh -- exact integer pixel arithmetic shared by all kernels (sm_100a).
//
// Every function reproduces, bit for bit, what the reference computes through numpy / PIL / cv2;
// the contract is written out in SURVEY.md section 8(a) and restated on the CPU in oracle/mvo.c.
// Citations are path:line in the reference tree.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mvb {

// ---------------------------------------------------------------------------------------------
// small exact divisions
// ---------------------------------------------------------------------------------------------

// x / 255 for 0 <= x < 2^24: one IMAD.HI (ceil(2^32 / 255) = 0x01010102, error term 254 x < 2^32)
__device__ __forceinline__ uint32_t div255(uint32_t x) { return __umulhi(x, 0x01010102u); }

// x / 255 for any 32-bit x
__device__ __forceinline__ uint32_t div255_wide(uint32_t x) { return __umulhi(x, 0x80808081u) >> 7; }

__device__ __forceinline__ float rcp_approx(float d)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
}

// floor(n / d) for n < 2^24, 1 <= d < 2^24, given rd ~= 1/d (MUFU.RCP): one float multiply gives
// the quotient to within +-1, a single remainder test makes it exact.
__device__ __forceinline__ uint32_t div_exact(uint32_t n, uint32_t d, float rd)
{
    uint32_t q = __float2uint_rz(__uint2float_rz(n) * rd);
    int32_t r = (int32_t)(n - q * d);
    if (r < 0) { q -= 1; r += (int32_t)d; }
    if (r >= (int32_t)d) q += 1;
    return q;
}

// ---------------------------------------------------------------------------------------------
// blend targets T(b, f) -- movis/imgproc.py:11-133 under numpy-2 promotion (b = bg, f = fg)
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t t_overlay(uint32_t b, uint32_t f) // imgproc.py:19-20
{
    return b < 128 ? div255(2 * b * f) : 255 - div255(2 * (255 - b) * (255 - f));
}

// min(255, floor(num / den)) with num <= 65025, 1 <= den <= 511
__device__ __forceinline__ uint32_t t_ratio255(uint32_t num, uint32_t den)
{
    float rd = rcp_approx(__uint2float_rz(den));
    uint32_t q = div_exact(num, den, rd);
    return min(q, 255u);
}

__device__ __forceinline__ uint32_t t_soft_light(uint32_t b, uint32_t f) // imgproc.py:58-72
{
    if (f < 128) {
        uint32_t t = (255u - 2u * f) * b * (255u - b); // <= 255^3, no wrap
        return (b - t / 65025u) & 0xFFu;
    }
    int32_t g;
    if (b < 64) {
        uint32_t t = 16u * b - 780300u * b + 260100u; // wraps mod 2^32: part of the contract
        t *= b;
        g = (int32_t)(t / 65025u);
    } else {
        g = (int32_t)__float2uint_rz(sqrtf(__uint2float_rz(255u * b)));
    }
    int32_t n = (int32_t)(2u * f - 255u) * (g - (int32_t)b); // |n| < 2^25
    int32_t q = n / 255;
    q -= (n - q * 255) < 0; // floor division
    return (uint32_t)((int32_t)b + q) & 0xFFu;
}

template <int MODE>
__device__ __forceinline__ uint32_t blend_target(uint32_t b, uint32_t f)
{
    if constexpr (MODE == 0) return f;                                   // normal
    else if constexpr (MODE == 1) return div255(b * f);                  // multiply
    else if constexpr (MODE == 2) return 255 - div255((255 - b) * (255 - f)); // screen
    else if constexpr (MODE == 3) return t_overlay(b, f);                // overlay
    else if constexpr (MODE == 4) return min(b, f);                      // darken
    else if constexpr (MODE == 5) return max(b, f);                      // lighten
    else if constexpr (MODE == 6) return t_ratio255(255 * b, 256 - f);   // color_dodge
    else if constexpr (MODE == 7) return 255 - t_ratio255(255 * (255 - b), f + 1); // color_burn
    else if constexpr (MODE == 8) return min(255u, b + f);               // linear_dodge
    else if constexpr (MODE == 9) return (uint32_t)max(0, (int)(b + f) - 255); // linear_burn
    else if constexpr (MODE == 10) return t_overlay(f, b);               // hard_light
    else if constexpr (MODE == 11) return t_soft_light(b, f);            // soft_light
    else if constexpr (MODE == 12)                                       // vivid_light (sic)
        return f > 127 ? 0u : 255 - t_ratio255(255 * (255 - b), 2 * f + 1);
    else if constexpr (MODE == 13) {                                     // linear_light
        int v = (int)b + 2 * (int)f - 255;
        return (uint32_t)(f > 127 ? min(255, v) : max(0, v));
    } else if constexpr (MODE == 14) {                                   // pin_light
        int f2 = 2 * (int)f;
        return (uint32_t)(f > 127 ? max((int)b, f2 - 255) : min((int)b, f2));
    } else if constexpr (MODE == 15) return (uint32_t)abs((int)b - (int)f); // difference
    else if constexpr (MODE == 16)                                       // exclusion
        return (uint32_t)min(255, max(0, (int)(b + f) - (int)div255_wide(2 * b * f))); // 2bf can reach 130050
    else return (uint32_t)max(0, (int)b - (int)f);                       // subtract
}

// Table-assisted blend targets for the fused kernel.  The four division / sqrt heavy modes read
//   rcp24[d] = ceil(2^24 / d), d in 1..256:  floor(n / d) == umulhi(n << 8, rcp24[d]) for n <= 65025
//              (error term n * (rcp24[d] * d - 2^24) < 2^24, checked in tests/test_oracle_golden.py)
//   softg[b] = g_w3c(b) of soft_light (imgproc.py:64-68), incl. its uint32 wrap-around for b < 64
// from shared memory; every other mode is the closed form of blend_target<MODE>.
struct BlendTables {
    const uint32_t* rcp24; // 257 entries
    const uint32_t* softg; // 256 entries
};

__device__ __forceinline__ uint32_t ratio255_tab(uint32_t num, uint32_t den, const uint32_t* rcp24)
{
    return min(__umulhi(num << 8, rcp24[den]), 255u);
}

template <int MODE>
__device__ __forceinline__ uint32_t blend_target_tab(uint32_t b, uint32_t f, const BlendTables& t)
{
    if constexpr (MODE == 6) return ratio255_tab(255 * b, 256 - f, t.rcp24);
    else if constexpr (MODE == 7) return 255 - ratio255_tab(255 * (255 - b), f + 1, t.rcp24);
    else if constexpr (MODE == 12) {
        const uint32_t q = 255 - ratio255_tab(255 * (255 - b), (2 * f + 1) & 255u, t.rcp24);
        return f > 127 ? 0u : q;
    } else if constexpr (MODE == 11) {
        const uint32_t td = (255u - 2u * f) * b * (255u - b); // meaningful for f < 128 only
        const uint32_t dark = (b - td / 65025u) & 0xFFu;
        const int32_t n = (int32_t)(2u * f - 255u) * ((int32_t)t.softg[b] - (int32_t)b);
        int32_t q = n / 255;
        q -= (n - q * 255) < 0; // floor division
        const uint32_t light = (uint32_t)((int32_t)b + q) & 0xFFu;
        return f < 128 ? dark : light;
    } else return blend_target<MODE>(b, f);
}

// ---------------------------------------------------------------------------------------------
// "over" operators on packed RGBA (byte 0 = R ... byte 3 = A)
// ---------------------------------------------------------------------------------------------

struct Rgba { uint32_t r, g, b, a; };

__device__ __forceinline__ Rgba unpack(uint32_t p)
{
    return Rgba{p & 0xFFu, (p >> 8) & 0xFFu, (p >> 16) & 0xFFu, p >> 24};
}
__device__ __forceinline__ uint32_t pack(uint32_t r, uint32_t g, uint32_t b, uint32_t a)
{
    return r | (g << 8) | (b << 16) | (a << 24);
}

// PIL ImagingAlphaComposite (reached from imgproc.py:210-213); sa already opacity-scaled.
__device__ __forceinline__ uint32_t over_pil(uint32_t dpx, const Rgba& s, uint32_t sa)
{
    if (sa == 0) return dpx;
    Rgba d = unpack(dpx);
    if (d.a == 255) { // coef1 = 128*sa, coef2 = 128*(255-sa), out alpha 255: no division
        uint32_t ia = 255 - sa;
        uint32_t tr = 128 * (s.r * sa + d.r * ia) + 0x4000;
        uint32_t tg = 128 * (s.g * sa + d.g * ia) + 0x4000;
        uint32_t tb = 128 * (s.b * sa + d.b * ia) + 0x4000;
        return pack((((tr >> 8) + tr) >> 8) >> 7, (((tg >> 8) + tg) >> 8) >> 7,
                    (((tb >> 8) + tb) >> 8) >> 7, 255);
    }
    uint32_t blend = d.a * (255 - sa);
    uint32_t outa255 = sa * 255 + blend;
    uint32_t coef1 = sa * 255 * 255 * 128 / outa255;
    uint32_t coef2 = 255 * 128 - coef1;
    uint32_t tr = s.r * coef1 + d.r * coef2 + 0x4000;
    uint32_t tg = s.g * coef1 + d.g * coef2 + 0x4000;
    uint32_t tb = s.b * coef1 + d.b * coef2 + 0x4000;
    uint32_t u = outa255 + 0x80;
    return pack((((tr >> 8) + tr) >> 8) >> 7, (((tg >> 8) + tg) >> 8) >> 7,
                (((tb >> 8) + tb) >> 8) >> 7, ((u >> 8) + u) >> 8);
}

// numpy _overlay, matte NONE or ALPHA (imgproc.py:156-170); fa already opacity-scaled.
template <int MODE>
__device__ __forceinline__ uint32_t over_numpy(uint32_t dpx, const Rgba& s, uint32_t fa, bool keep_alpha)
{
    Rgba d = unpack(dpx);
    uint32_t tr = blend_target<MODE>(d.r, s.r);
    uint32_t tg = blend_target<MODE>(d.g, s.g);
    uint32_t tb = blend_target<MODE>(d.b, s.b);
    if (d.a == 255) { // oa = 65025: (255 fa T + 255 (255-fa) b) // 65025 == (fa T + (255-fa) b) // 255
        uint32_t ia = 255 - fa;
        return pack(div255(fa * tr + ia * d.r), div255(fa * tg + ia * d.g), div255(fa * tb + ia * d.b), 255);
    }
    uint32_t c1 = 255 * fa, c2 = d.a * (255 - fa);
    uint32_t oa = c1 + c2;
    if (oa == 0) return keep_alpha ? (dpx & 0xFF000000u) : 0u; // rgb zeroed, alpha 0 (== d.a)
    float rd = rcp_approx(__uint2float_rz(oa));
    uint32_t r = div_exact(c1 * tr + c2 * d.r, oa, rd);
    uint32_t g = div_exact(c1 * tg + c2 * d.g, oa, rd);
    uint32_t b = div_exact(c1 * tb + c2 * d.b, oa, rd);
    return pack(r, g, b, keep_alpha ? d.a : oa / 255);
}

__device__ __forceinline__ uint32_t over_numpy_dyn(int mode, uint32_t dpx, const Rgba& s, uint32_t fa,
                                                   bool keep_alpha)
{
    switch (mode) {
#define MVB_CASE(M) case M: return over_numpy<M>(dpx, s, fa, keep_alpha);
        MVB_CASE(0) MVB_CASE(1) MVB_CASE(2) MVB_CASE(3) MVB_CASE(4) MVB_CASE(5) MVB_CASE(6) MVB_CASE(7)
        MVB_CASE(8) MVB_CASE(9) MVB_CASE(10) MVB_CASE(11) MVB_CASE(12) MVB_CASE(13) MVB_CASE(14)
        MVB_CASE(15) MVB_CASE(16)
#undef MVB_CASE
    default: return over_numpy<17>(dpx, s, fa, keep_alpha);
    }
}

// cv2.cvtColor(RGB2GRAY) (imgproc.py:148)
__device__ __forceinline__ uint32_t rgb_to_gray(const Rgba& p)
{
    return (p.r * 9798u + p.g * 19235u + p.b * 3735u + 16384u) >> 15;
}

// ---------------------------------------------------------------------------------------------
// cv2.warpAffine's fixed-point bilinear sample (composition.py:811-813)
// ---------------------------------------------------------------------------------------------

// out = (sum w_ij p_ij + 512) >> 10 per channel, w = {(32-fx)(32-fy), fx(32-fy), (32-fx)fy, fx fy}.
// Horizontal lerp on two channels per register (each lane <= 255*32 < 2^13), then one dp2a per
// channel for the vertical lerp (16-bit lane pair x 8-bit weight pair).
__device__ __forceinline__ Rgba bilinear_q5(uint32_t p00, uint32_t p01, uint32_t p10, uint32_t p11,
                                            uint32_t fx, uint32_t fy)
{
    const uint32_t wx1 = fx, wx0 = 32 - fx;
    const uint32_t wy = (32 - fy) | (fy << 8);
    uint32_t t_rb = __byte_perm(p00, 0, 0x4240) * wx0 + __byte_perm(p01, 0, 0x4240) * wx1;
    uint32_t t_ga = __byte_perm(p00, 0, 0x4341) * wx0 + __byte_perm(p01, 0, 0x4341) * wx1;
    uint32_t b_rb = __byte_perm(p10, 0, 0x4240) * wx0 + __byte_perm(p11, 0, 0x4240) * wx1;
    uint32_t b_ga = __byte_perm(p10, 0, 0x4341) * wx0 + __byte_perm(p11, 0, 0x4341) * wx1;
    Rgba o;
    o.r = __dp2a_lo(__byte_perm(t_rb, b_rb, 0x5410), wy, 512u) >> 10;
    o.b = __dp2a_lo(__byte_perm(t_rb, b_rb, 0x7632), wy, 512u) >> 10;
    o.g = __dp2a_lo(__byte_perm(t_ga, b_ga, 0x5410), wy, 512u) >> 10;
    o.a = __dp2a_lo(__byte_perm(t_ga, b_ga, 0x7632), wy, 512u) >> 10;
    return o;
}

__device__ __forceinline__ uint32_t ldg_u32(const uint8_t* p)
{
    return __ldg(reinterpret_cast<const unsigned int*>(p));
}

// The four taps of one bilinear sample plus its 5-bit fractions, fetched ahead of their use.
struct Taps {
    uint32_t p00, p01, p10, p11;
    uint32_t fx, fy;
    bool any; // false: all four taps are outside the image (sample is transparent black)
};

// Fetch the taps at fixed-point (X, Y) with 5 fractional bits; taps outside the image read 0
// (BORDER_CONSTANT).  Byte offsets are 32-bit: a source image is < 4 GiB (checked on the host).
__device__ __forceinline__ Taps fetch_taps(const uint8_t* __restrict__ src, int sw, int sh, int stride, int X, int Y)
{
    Taps t;
    const int sx = X >> 5, sy = Y >> 5;
    t.fx = X & 31;
    t.fy = Y & 31;
    t.any = true;
    if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
        const uint8_t* base = src + ((uint32_t)sy * (uint32_t)stride + (uint32_t)sx * 4u);
        t.p00 = ldg_u32(base);
        t.p01 = ldg_u32(base + 4);
        t.p10 = ldg_u32(base + stride);
        t.p11 = ldg_u32(base + stride + 4);
    } else {
        const bool x0 = (unsigned)sx < (unsigned)sw, x1 = (unsigned)(sx + 1) < (unsigned)sw;
        const bool y0 = (unsigned)sy < (unsigned)sh, y1 = (unsigned)(sy + 1) < (unsigned)sh;
        t.any = (x0 | x1) & (y0 | y1);
        const uint8_t* base = src + (int64_t)sy * stride + (int64_t)sx * 4;
        t.p00 = (x0 & y0) ? ldg_u32(base) : 0u;
        t.p01 = (x1 & y0) ? ldg_u32(base + 4) : 0u;
        t.p10 = (x0 & y1) ? ldg_u32(base + stride) : 0u;
        t.p11 = (x1 & y1) ? ldg_u32(base + stride + 4) : 0u;
    }
    return t;
}

// Sample `src` at fixed-point (X, Y); returns false when all four taps are outside the image.
__device__ __forceinline__ bool sample_q5(const uint8_t* __restrict__ src, int sw, int sh, int stride,
                                          int X, int Y, Rgba& out)
{
    const Taps t = fetch_taps(src, sw, sh, stride, X, Y);
    if (!t.any) return false;
    out = bilinear_q5(t.p00, t.p01, t.p10, t.p11, t.fx, t.fy); // fx == fy == 0 reproduces p00 exactly
    return true;
}

// cv::saturate_cast<int>(double): round-half-even; out of range -> INT_MIN (x86 cvtsd2si).
__device__ __forceinline__ int cv_round(double v)
{
    return (fabs(v) < 2147483647.5) ? __double2int_rn(v) : (int)0x80000000;
}

// The four integer tables of cv::warpAffine, from the inverted matrix m (no FMA contraction).
__device__ __forceinline__ int warp_adelta(const double* m, int u) { return cv_round(__dmul_rn(__dmul_rn(m[0], (double)u), 1024.0)); }
__device__ __forceinline__ int warp_bdelta(const double* m, int u) { return cv_round(__dmul_rn(__dmul_rn(m[3], (double)u), 1024.0)); }
__device__ __forceinline__ int warp_x0(const double* m, int v) { return cv_round(__dmul_rn(__dadd_rn(__dmul_rn(m[1], (double)v), m[2]), 1024.0)) + 16; }
__device__ __forceinline__ int warp_y0(const double* m, int v) { return cv_round(__dmul_rn(__dadd_rn(__dmul_rn(m[4], (double)v), m[5]), 1024.0)) + 16; }

} // namespace mvb
