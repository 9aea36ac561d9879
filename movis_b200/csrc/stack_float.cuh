// stack_float.cuh -- the fused warp + blend kernel for OPAQUE frames (bg_color alpha 255), sm_100a.
//
// Same contract as k_composite_stack (composite.cu): per frame it replaces the reference's
// cv2.warpAffine -> alpha_composite pair of every layer (movis/layer/composition.py:811-819) and
// the frame walk around it (composition.py:363-368), bit for bit.  An opaque background keeps the
// frame's alpha at 255 through every layer (both "over" operators map dst alpha 255 to 255), so
// both reduce to   out = floor((fa*T + (255-fa)*b [+127 for PIL]) / 255)   per colour channel
// (imgproc.py:156-165 with oa = 65025; PIL's ImagingAlphaComposite with outa255 = 65025).
//
// Why float: on B200 the kernel is bound by instruction issue, not HBM (tools/experiments/
// pipe_rates*.cu: LOP3/SHF/PRMT/IMAD/IDP issue at 0.5 per clock per scheduler, IMAD.HI at 0.21,
// FFMA at 1.0 and the packed FFMA2 does two lanes per issue slot).  All quantities here are
// integers below 2^24, so fp32 arithmetic is exact, and a floor division becomes ONE fused
// multiply-add rounded towards -inf against the constant 1.5*2^23 ("magic"): the quotient lands
// in the low mantissa bits.  Pixels, samples and quotients are carried as `magic + value`
// ("biased"); differences of biased numbers are plain differences.
//   * bilinear:  cv2's 2 x 2 weights as 16-bit pairs, IDP.2A (16 x 8 bit dot product) on taps paired by
//                channel, accumulator preset to the bits of 2^23 + 32*512, so a channel's exact weighted
//                sum comes out as a float with no conversion; (S+512)>>10 is one FFMA2.RM for TWO rows;
//   * blend + composite + floor(./255): four FFMA2/FADD2 per channel for TWO rows (NORMAL).
// The tile's pixels stay in registers for the whole layer walk (thread = 1 column x 4 rows).
#pragma once

#include <type_traits>

#include "pixel_math.cuh"

namespace mvb {

typedef unsigned long long u64;

constexpr float kMagic = 12582912.0f;          // 1.5 * 2^23: ulp 1, so RN/RM of (magic + x) rounds x to an integer
constexpr uint32_t kMagicBits = 0x4B400000u;   // float bits of kMagic; bits of (kMagic + k) = kMagicBits + k
constexpr uint32_t kTwo23Bits = 0x4B000000u;   // float bits of 2^23
constexpr float kInv65025 = 1.5378702300949953e-05f;   // 0x37810183: the float above 1/65025 (soft_light)
constexpr float kInv255 = 1.0f / 255.0f;       // 0x3B808081 = 0.00392156886 > 1/255: floor(x * kInv255) == x / 255
                                               // for integers 0 <= x <= 130050 (tests/test_oracle_golden.py)
constexpr float kNeg255Magic = -3208642560.0f; // -255 * kMagic = -765 * 2^22, exactly representable

__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ u64 bc(float v) { return pk(v, v); }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 fma2_rm(u64 a, u64 b, u64 c)
{
    u64 d;
    asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b)
{
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ float lds_f32(uint32_t shared_addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(shared_addr));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Blend + composite of one colour channel of two rows.  Inputs are biased (dB = magic + b,
// sB = magic + f); `diff` below is T(b, f) - b with T the blend target of movis/imgproc.py:11-133
// in the closed forms of SURVEY.md 8(a-5).  Two algebraic identities keep the product modes short
// (N integer => floor(N + x) = N + floor(x)):
//   screen   255 - floor((255-b)(255-f)/255)   = b + f - floor(b f / 255)
//   overlay  255 - floor(2(255-b)(255-f)/255)  = 2b + 2f - 255 - floor(2 b f / 255)     (b >= 128)
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ u64 add2_rm(u64 a, u64 b)
{
    u64 d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// f32x2 has no min / max: two scalar FMNMX on the halves of the register pair
__device__ __forceinline__ u64 min2(u64 a, u64 b)
{
    float a0, a1, b0, b1;
    upk(a, a0, a1);
    upk(b, b0, b1);
    return pk(fminf(a0, b0), fminf(a1, b1));
}
__device__ __forceinline__ u64 max2(u64 a, u64 b)
{
    float a0, a1, b0, b1;
    upk(a, a0, a1);
    upk(b, b0, b1);
    return pk(fmaxf(a0, b0), fmaxf(a1, b1));
}
// per lane: key < thr ? a : b
__device__ __forceinline__ u64 select_lt(u64 key, float thr, u64 a, u64 b)
{
    float k0, k1, a0, a1, b0, b1;
    upk(key, k0, k1);
    upk(a, a0, a1);
    upk(b, b0, b1);
    return pk(k0 < thr ? a0 : b0, k1 < thr ? a1 : b1);
}
__device__ __forceinline__ u64 rcp2(u64 a)
{
    float a0, a1;
    upk(a, a0, a1);
    return pk(rcp_approx(a0), rcp_approx(a1));
}

// magic + min(255, floor(255 n / den)), n <= 255 plain, 1 <= den <= 256 plain.  The quotient of
// (255 n + 0.5) by den is at least 0.5/256 away from every integer, MUFU.RCP and the multiply are
// within 2^-22 relative, and only quotients below 256 matter (the rest clamp).
__device__ __forceinline__ u64 ratio255_biased(u64 n, u64 den)
{
    const u64 t = mul2(fma2(n, bc(255.0f), bc(0.5f)), rcp2(den));
    return min2(add2_rm(t, bc(kMagic)), bc(kMagic + 255.0f));
}

// T(b, f) - b for one colour channel of two rows, biased inputs.
template <int MODE>
__device__ __forceinline__ u64 blend_diff2(u64 dB, u64 sB, const float2* softg)
{
    const u64 kM = bc(kMagic), kM255 = bc(kMagic + 255.0f);
    u64 diff;
    if constexpr (MODE == 0 || MODE == 18) {                                      // normal / PIL over
        diff = sub2(sB, dB);
    } else if constexpr (MODE == 1 || MODE == 2 || MODE == 16) {                  // multiply / screen / exclusion
        const u64 p = mul2(sub2(dB, kM), sub2(sB, kM));
        const u64 qB = fma2_rm(p, bc(MODE == 16 ? 2.0f * kInv255 : kInv255), kM);
        diff = MODE == 1 ? sub2(qB, dB) : sub2(sB, qB);                           // q - b  |  f - q
    } else if constexpr (MODE == 3 || MODE == 10) {                               // overlay / hard_light
        const u64 b = sub2(dB, kM), f = sub2(sB, kM);
        const u64 qB = fma2_rm(mul2(b, f), bc(2.0f * kInv255), kM);               // floor(2 b f / 255)
        const u64 lo = sub2(qB, dB);                                              // q - b
        const u64 hi = sub2(fma2(f, bc(2.0f), sub2(dB, bc(255.0f))), qB);         // (b + 2f - 255) - q
        diff = select_lt(MODE == 3 ? b : f, 128.0f, lo, hi);
    } else if constexpr (MODE == 4) {                                             // darken: min(f - b, 0)
        diff = min2(sub2(sB, dB), bc(0.0f));
    } else if constexpr (MODE == 5) {                                             // lighten
        diff = max2(sub2(sB, dB), bc(0.0f));
    } else if constexpr (MODE == 6) {                                             // color_dodge: 255 b / (256 - f)
        diff = sub2(ratio255_biased(sub2(dB, kM), sub2(bc(kMagic + 256.0f), sB)), dB);
    } else if constexpr (MODE == 7) {                                             // color_burn: 255 - 255 (255-b) / (f+1)
        const u64 nb = sub2(kM255, dB);
        diff = sub2(add2(nb, kM), ratio255_biased(nb, sub2(sB, bc(kMagic - 1.0f))));  // 255 - q - b
    } else if constexpr (MODE == 8) {                                             // linear_dodge: min(255 - b, f)
        diff = min2(sub2(kM255, dB), sub2(sB, kM));
    } else if constexpr (MODE == 9) {                                             // linear_burn: max(-b, f - 255)
        diff = max2(sub2(kM, dB), sub2(sB, kM255));
    } else if constexpr (MODE == 11) {                                            // soft_light (imgproc.py:58-72)
        const u64 b = sub2(dB, kM), f = sub2(sB, kM);
        // f < 128: T - b = -floor((255 - 2f) b (255 - b) / 65025); the numerator stays below 2^22 and
        // fl_up(1/65025) makes the single fma.rm exact for every (b, f) (tests/test_oracle_golden.py)
        const u64 kd = fma2(f, bc(-2.0f), bc(255.0f));
        const u64 qd = fma2_rm(mul2(mul2(b, sub2(kM255, dB)), kd), bc(kInv65025), kM);
        const u64 dark = sub2(kM, qd);
        // f >= 128: T = (b + (2f - 255) D[b] // 255) mod 256 with D = g_w3c(b) - b in uint32, as numpy computes
        // it.  Only T mod 256 survives astype(uint8), and floor(k (D + 65280 m) / 255) = floor(k D / 255)
        // + 256 k m, so D is reduced mod 65280 and split as 255 A + R: the sum b + k A + floor(k R / 255)
        // stays exact in fp32.  softg[b] = (A, R).
        float d0, d1;
        upk(dB, d0, d1);
        const float2 ar0 = softg[__float_as_uint(d0) & 0xFFu], ar1 = softg[__float_as_uint(d1) & 0xFFu];
        const u64 k = fma2(f, bc(2.0f), bc(-255.0f));
        const u64 q2 = fma2_rm(mul2(k, pk(ar0.y, ar1.y)), bc(kInv255), kM);
        float t0, t1;
        upk(add2(fma2(k, pk(ar0.x, ar1.x), q2), b), t0, t1);                      // magic + b + k A + floor(k R / 255)
        const u64 light = sub2(pk(__uint_as_float((__float_as_uint(t0) & 0xFFu) | kMagicBits),
                                  __uint_as_float((__float_as_uint(t1) & 0xFFu) | kMagicBits)), dB);
        diff = select_lt(f, 128.0f, dark, light);
    } else if constexpr (MODE == 12) {                                            // vivid_light (sic)
        const u64 f = sub2(sB, kM), nb = sub2(kM255, dB);
        const u64 d = sub2(add2(nb, kM), ratio255_biased(nb, fma2(f, bc(2.0f), bc(1.0f))));
        diff = select_lt(f, 127.5f, d, sub2(kM, dB));                             // f > 127: T = 0
    } else if constexpr (MODE == 13) {                                            // linear_light
        const u64 negb = sub2(kM, dB);                                            // clamp(2f - 255, -b, 255 - b)
        diff = min2(max2(fma2(sub2(sB, kM), bc(2.0f), bc(-255.0f)), negb), add2(negb, bc(255.0f)));
    } else if constexpr (MODE == 14) {                                            // pin_light
        const u64 hB = fma2(sub2(sB, kM), bc(2.0f), kM);                          // magic + 2f
        diff = sub2(min2(max2(dB, sub2(hB, bc(255.0f))), hB), dB);                // min(max(b, 2f - 255), 2f) - b
    } else if constexpr (MODE == 15) {                                            // difference: |b - f| - b
        float t0, t1, b0, b1;
        upk(sub2(dB, sB), t0, t1);
        upk(sub2(dB, kM), b0, b1);
        diff = pk(fabsf(t0) - b0, fabsf(t1) - b1);
    } else {                                                                      // subtract: max(-b, -f)
        static_assert(MODE == 17, "unknown blend mode");
        diff = max2(sub2(kM, dB), sub2(kM, sB));
    }
    return diff;
}

// Opaque frame (alpha 255 before and after): blend + composite + requantise one colour channel of
// two rows.  x = fa*T + (255 - fa)*b = fa*(T - b) + 255*b is an integer in [0, 65025].
template <int MODE>
__device__ __forceinline__ u64 blend_chan(u64 dB, u64 sB, u64 fa, const float2* softg)
{
    const u64 kM = bc(kMagic);
    const u64 x = fma2(fa, blend_diff2<MODE>(dB, sB, softg), fma2(dB, bc(255.0f), bc(kNeg255Magic)));
    if constexpr (MODE == 18) return fma2(x, bc(kInv255), kM);                    // PIL: floor((x+127)/255) = rint(x/255)
    else return fma2_rm(x, bc(kInv255), kM);                                      // numpy: floor(x/255)
}

constexpr int kFWarps = 8;   // 256 threads: 8 warps x 4 rows = one 32 x 32 tile
constexpr int kFRows = 4;
constexpr int kFTileH = kFWarps * kFRows;
constexpr int kFSlots = 8;   // layers whose per-tile tables are resident in shared memory at a time

struct RowPairState { u64 c[4]; };   // r, g, b, a of rows (2p, 2p+1), biased (a is unused on opaque frames)

// Everything the layer loop reads about one resident layer, contiguous so that one base address
// serves all of it (offsets become LDS immediates).
struct __align__(16) SlotMem {
    int4 info;               // see the table-building code
    int2 col[kTileW];        // (adelta, bdelta) per tile column
    int2 row[kFTileH];       // (X0, Y0) per tile row (relative to the box origin for TMA layers)
    float lut[256];          // opacity-scaled alpha, as float
};

// ---- frames whose alpha is not known to be 255 ("general") ----------------------------------------
// numpy _overlay (imgproc.py:156-170), matte NONE, two rows: with c1 = 255 fa, c2 = ba (255 - fa),
// oa = c1 + c2:  rgb = (c1 T + c2 b) // max(oa, 1),  alpha = oa // 255.  The numerator
// c1 (T - b) + oa b is below 2^24, so it is exact in fp32; the quotient by the variable oa is
// q0 = rint(n * rcp(oa)) corrected by the sign of the exact remainder n - q0 oa (|n/oa - q0| < 1).
template <int MODE>
__device__ __forceinline__ void over_numpy_rows(RowPairState& d, const u64 (&s)[4], u64 fa, const float2* softg)
{
    const u64 kM = bc(kMagic);
    const u64 ba = sub2(d.c[3], kM);
    const u64 c1 = mul2(fa, bc(255.0f));
    const u64 oa = fma2(ba, sub2(bc(255.0f), fa), c1);
    const u64 oam = max2(oa, bc(1.0f));                       // oa == 0: numerator 0, quotient 0 (imgproc.py:164)
    const u64 rd = rcp2(oam);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const u64 diff = blend_diff2<MODE>(d.c[c], s[c], softg);
        const u64 n = fma2(c1, diff, mul2(oa, sub2(d.c[c], kM)));
        const u64 t = fma2(n, rd, kM);                        // magic + rint(n / oa)
        float r0, r1, t0, t1;
        upk(fma2(sub2(kM, t), oam, n), r0, r1);               // n - q0 oa, exact, in (-oa, oa)
        upk(t, t0, t1);
        d.c[c] = pk(__uint_as_float(__float_as_uint(t0) + (uint32_t)(__float_as_int(r0) >> 31)),   // q0 - (r < 0)
                    __uint_as_float(__float_as_uint(t1) + (uint32_t)(__float_as_int(r1) >> 31)));
    }
    d.c[3] = fma2_rm(oa, bc(kInv255), kM);
}

// PIL ImagingAlphaComposite (reached from imgproc.py:210-213), two rows:
//   outa255 = 255 sa + da (255 - sa);  coef1 = sa*255*255*128 // outa255;  coef2 = 255*128 - coef1
//   t = s coef1 + d coef2 + 0x4000;  out = (((t >> 8) + t) >> 8) >> 7 = (t + (t >> 8)) >> 15
//   alpha: u = outa255 + 0x80;  out = ((u >> 8) + u) >> 8
// sa == 0 gives coef1 = 0 and reproduces d exactly, which is PIL's "leave the pixel alone" rule.
// coef1 needs a 31-bit numerator: integer division, once per pixel.
__device__ __forceinline__ void over_pil_rows(RowPairState& d, const u64 (&s)[4], u64 fa)
{
    const u64 kM = bc(kMagic);
    float f0, f1, a0, a1;
    upk(fa, f0, f1);
    upk(d.c[3], a0, a1);
    const uint32_t sa0 = __float_as_uint(f0 + kMagic) & 0xFFu, sa1 = __float_as_uint(f1 + kMagic) & 0xFFu;
    const uint32_t da0 = __float_as_uint(a0) & 0xFFu, da1 = __float_as_uint(a1) & 0xFFu;
    const uint32_t oa0 = sa0 * 255u + da0 * (255u - sa0), oa1 = sa1 * 255u + da1 * (255u - sa1);
    const uint32_t k0 = (sa0 * 8323200u) / max(oa0, 1u), k1 = (sa1 * 8323200u) / max(oa1, 1u);
    const u64 c1 = pk((float)k0, (float)k1), c2 = sub2(bc(32640.0f), c1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const u64 t = fma2(sub2(s[c], kM), c1, fma2(sub2(d.c[c], kM), c2, bc(16384.0f)));   // < 2^24
        const u64 v = add2(t, sub2(fma2_rm(t, bc(1.0f / 256.0f), kM), kM));                 // t + (t >> 8)
        d.c[c] = fma2_rm(v, bc(1.0f / 32768.0f), kM);
    }
    const u64 u = pk((float)(oa0 + 128u), (float)(oa1 + 128u));
    const u64 v = add2(u, sub2(fma2_rm(u, bc(1.0f / 256.0f), kM), kM));
    d.c[3] = fma2_rm(v, bc(1.0f / 256.0f), kM);
}

// Blend the samples of four rows into the thread's pixels.  `keep`: bit q set = row q of this thread
// lies inside the layer's bbox (only consulted on general frames, where a transparent sample is
// not a no-op for the numpy arithmetic).
template <int MODE, bool OPAQUE>
__device__ __forceinline__ void blend_rows(RowPairState (&d)[2], const u64 (&s)[2][4], const u64 (&fa)[2],
                                           const float2* softg, unsigned keep)
{
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        if constexpr (OPAQUE) {
#pragma unroll
            for (int c = 0; c < 3; ++c) d[p].c[c] = blend_chan<MODE>(d[p].c[c], s[p][c], fa[p], softg);
        } else {
            RowPairState nd = d[p];
            if constexpr (MODE == 18) over_pil_rows(nd, s[p], fa[p]);
            else over_numpy_rows<MODE>(nd, s[p], fa[p], softg);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float o0, o1, n0, n1;
                upk(d[p].c[c], o0, o1);
                upk(nd.c[c], n0, n1);
                d[p].c[c] = pk((keep >> (2 * p) & 1) ? n0 : o0, (keep >> (2 * p + 1) & 1) ? n1 : o1);
            }
        }
    }
}

// ---- TMA / mbarrier plumbing (PTX; sm_90+ async proxy) -------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One arrival for the whole (converged) warp: elect.sync waits for all 32 lanes, so every lane has issued
// its reads of the buffer being released before the elected lane signals (two instructions instead of
// the __syncwarp + lane test + branch sequence).
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar)
{
    asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; @P mbarrier.arrive.shared::cta.b64 _, [%0]; }"
                 ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Blocks until the phase with the given parity has completed.  A wait that never completes means
// a broken descriptor: trap instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (spins > (1u << 22)) __trap();
    }
}
// One non-blocking probe of the phase with the given parity.
__device__ __forceinline__ int mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return (int)done;
}
// One box of the tiled tensor map -> shared memory; completion is signalled on `bar` (bytes).
__device__ __forceinline__ void tma_load_box(uint32_t dst, const void* map, int x, int y, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar) : "memory");
}
// a * b + c on the FMA pipe (IMAD), for address arithmetic the compiler would put on the busier ALU pipe
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t shared_addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr));
    return v;
}

constexpr int kMaxStages = 8;   // shared-memory stages of TMA boxes in flight (2, 4 or 8 are used)
enum { kTileTma = 3, kTileTmaPart = 7 };   // next to kTileEdge / kTileEmpty / kTileInterior: taps come from a TMA-staged box

// cv2's bilinear weights of a sample at (sumx, sumy) (1/1024 px; 5 fractional bits are used), times 32:
// wt = 32 (32 - fx)(32 - fy) | 32 fx (32 - fy) << 16,  wb = 32 (32 - fx) fy | 32 fx fy << 16  (imgwarp.cpp:
// the INTER_LINEAR table of initInterTab2D, whose entries are these exact products).  The factor 32 is
// what masking fx in place (sumx & 0x3E0 = 32 fx) leaves instead of shifting it down; every half stays
// below 2^16 and the weighted sums below 2^23, and the final fma divides it out.
__device__ __forceinline__ void tap_weights(int sumx, int sumy, uint32_t& wt, uint32_t& wb)
{
    const uint32_t fy = (uint32_t)(sumy >> 5) & 31u;
    const uint32_t a = imad((uint32_t)sumx & 0x3E0u, 0xFFFFu, 1024u);   // 32 (32 - fx) | 32 fx << 16
    wb = a * fy;
    wt = (a << 5) - wb;
}

// What identifies a thread's pixels inside the tile being walked.
struct PixelCtx {
    int tx0, ty0;   // tile origin in the frame
    int x;          // the thread's column
    int rbeg;       // first of its kFRows rows, relative to ty0
    int lane;
    bool x_ok;      // x < frame width
};

// One layer of the walk for one thread: fetch the taps of its four rows (from the TMA-staged box,
// or through L1 for layers that cannot be staged), cv2's fixed-point bilinear, opacity LUT, blend
// into the pixels `d`.  `before_tma_wait(info)` runs ahead of the wait on the stage's barrier.
template <bool OPAQUE, class Slot, class BeforeTmaWait>
__device__ __forceinline__ void walk_layer(const StackParams& P, const Slot& S, const int* layer_index, const PixelCtx& C,
                                           RowPairState (&d)[2], const float2* softg, BeforeTmaWait&& before_tma_wait)
{
    const int tx0 = C.tx0, ty0 = C.ty0, x = C.x, rbeg = C.rbeg, lane = C.lane;
    const bool x_ok = C.x_ok;
    (void)tx0;
    const int4 info = S.info;
    const int cls = (info.x >> 8) & 7;
    const int2 col = S.col[lane];
    const int2* __restrict__ rowtab = &S.row[rbeg];

    // ---- taps of the four rows, all requested before the first is used ----------------------
    uint32_t t[kFRows][4];
    uint32_t wt[kFRows], wb[kFRows];   // cv2's 2 x 2 weights, two 16-bit values each: (w00, w01), (w10, w11)
    unsigned ok = 0xFu;   // bit q: row q of this thread lies inside the layer's bbox and the frame
    auto staged = [&](auto part_tag) {
        before_tma_wait(info);
        const uint32_t pitch = (uint32_t)info.x >> 16, stage = (uint32_t)info.w;
        // Tiles the bbox covers only partly run their own copy of the loop (cls == kTileTmaPart): the box
        // spans the covered part alone, so samples outside it read a dummy address (softg) with weight 0, which
        // makes them transparent (alpha 0 -> lut[0] = 0).  Fully covered tiles carry none of that.
        {
            constexpr bool PART = decltype(part_tag)::value;
            if (PART) {
                const DevLayer& L = P.layers[*layer_index & 255];
                const bool col_in = x_ok && (unsigned)(x - L.bx) < (unsigned)L.bw;
                ok = 0u;
#pragma unroll
                for (int q = 0; q < kFRows; ++q) {
                    const int yy = ty0 + rbeg + q;
                    if (col_in && (unsigned)(yy - L.by) < (unsigned)L.bh && yy < P.H) ok |= 1u << q;
                }
            }
            // addresses and weights first: they do not need the texels, so they overlap the box's arrival
            uint32_t a[kFRows], down[kFRows];
#pragma unroll
            for (int q = 0; q < kFRows; ++q) {
                const int2 r = rowtab[q];
                const int sumx = r.x + col.x, sumy = r.y + col.y;      // 1/1024 px, relative to the box origin
                a[q] = imad((uint32_t)(sumy >> 10), pitch, imad((uint32_t)(sumx >> 10), 4u, stage));
                tap_weights(sumx, sumy, wt[q], wb[q]);
                down[q] = pitch;
                if (PART && !(ok >> q & 1)) { a[q] = smem_addr(softg); down[q] = 0u; wt[q] = wb[q] = 0u; }
            }
            mbar_wait((uint32_t)info.y, (info.x >> 11) & 1);
#pragma unroll
            for (int q = 0; q < kFRows; ++q) {
                t[q][0] = lds_u32(a[q]);
                t[q][1] = lds_u32(a[q] + 4);
                t[q][2] = lds_u32(a[q] + down[q]);
                t[q][3] = lds_u32(a[q] + down[q] + 4);
            }
        }
        mbar_arrive_warp((uint32_t)info.y + 8 * kMaxStages);   // this warp is done with the stage
    };
    // (an opaque copy of the info word keeps the compiler from folding this test into the switch over the
    // other classes, which it orders with the common case last)
    uint32_t hot;
    asm("mov.b32 %0, %1;" : "=r"(hot) : "r"(info.x));
    if ((hot & 0x700u) == (uint32_t)kTileTma << 8) {
        staged(std::false_type{});
    } else if (cls == kTileTmaPart) {
        staged(std::true_type{});
    } else if (cls == kTileInterior) {
        const DevLayer& L = P.layers[*layer_index & 255];
        const uint8_t* __restrict__ src = L.src;
        const int ss = L.src_stride;
#pragma unroll
        for (int k = 0; k < kFRows; ++k) {
            const int2 r = rowtab[k];
            const int sumx = r.x + col.x, sumy = r.y + col.y;      // 1/1024 px (cv2: X0 + adelta)
            const uint8_t* p = src + ((uint32_t)(sumy >> 10) * (uint32_t)ss + (uint32_t)(sumx >> 10) * 4u);
            t[k][0] = ldg_u32(p);
            t[k][1] = ldg_u32(p + 4);
            t[k][2] = ldg_u32(p + ss);
            t[k][3] = ldg_u32(p + ss + 4);
            tap_weights(sumx, sumy, wt[k], wb[k]);
        }
    } else {
        const DevLayer& L = P.layers[*layer_index & 255];
        const uint8_t* __restrict__ src = L.src;
        const int ss = L.src_stride;
        const int sw = L.src_w, sh = L.src_h;
        const bool col_in = x_ok && (unsigned)(x - L.bx) < (unsigned)L.bw;
#pragma unroll
        for (int k = 0; k < kFRows; ++k) {
            const int2 r = rowtab[k];
            const int sumx = r.x + col.x, sumy = r.y + col.y;
            const int yy = ty0 + rbeg + k;
            t[k][0] = t[k][1] = t[k][2] = t[k][3] = 0u;
            if (col_in && (unsigned)(yy - L.by) < (unsigned)L.bh && yy < P.H) {
                const Taps tp = fetch_taps(src, sw, sh, ss, sumx >> 5, sumy >> 5);
                t[k][0] = tp.p00; t[k][1] = tp.p01; t[k][2] = tp.p10; t[k][3] = tp.p11;
            } else {
                ok &= ~(1u << k);
            }
            tap_weights(sumx, sumy, wt[k], wb[k]);
        }
    }

    // ---- cv2's fixed-point bilinear: out = (sum w p + 512) >> 10 -------------------------------
    // IDP.2A multiplies the two 16-bit weights of a tap row with two bytes: with the taps of a row paired
    // by channel, four of them give one channel's exact sum; started from the bits of 2^23 + 32 * 512 the
    // accumulator IS the float 2^23 + 32 (512 + sum), and one fma.rm takes floor(. / 1024) of a row pair.
    u64 s[2][4], fa[2];
    // &lut[bits - kMagicBits] for the biased alpha `bits`: one LEA (the shift drops the exponent bits)
    const uint32_t lut_base = (uint32_t)info.z;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        float v[2][4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = 2 * p + q;
            const uint32_t t_rg = __byte_perm(t[k][0], t[k][1], 0x5140), t_ba = __byte_perm(t[k][0], t[k][1], 0x7362);
            const uint32_t b_rg = __byte_perm(t[k][2], t[k][3], 0x5140), b_ba = __byte_perm(t[k][2], t[k][3], 0x7362);
            v[q][0] = __uint_as_float(__dp2a_lo(wb[k], b_rg, __dp2a_lo(wt[k], t_rg, kTwo23Bits + 32u * 512u)));
            v[q][1] = __uint_as_float(__dp2a_hi(wb[k], b_rg, __dp2a_hi(wt[k], t_rg, kTwo23Bits + 32u * 512u)));
            v[q][2] = __uint_as_float(__dp2a_lo(wb[k], b_ba, __dp2a_lo(wt[k], t_ba, kTwo23Bits + 32u * 512u)));
            v[q][3] = __uint_as_float(__dp2a_hi(wb[k], b_ba, __dp2a_hi(wt[k], t_ba, kTwo23Bits + 32u * 512u)));
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)   // (2^23 + 32 (512 + sum)) / 32768 = 256 + (sum + 512) / 1024
            s[p][c] = fma2_rm(pk(v[0][c], v[1][c]), bc(1.0f / 32768.0f), bc(kMagic - 256.0f));
        float a0, a1;
        upk(s[p][3], a0, a1);
        const uint32_t i0 = __float_as_uint(a0), i1 = __float_as_uint(a1);
        fa[p] = pk(lds_f32(lut_base + (i0 << 2)), lds_f32(lut_base + (i1 << 2)));
    }

    // ---- blend, in the code specialised for this layer's mode ------------------------------------
    switch (info.x & 255) {
#define MVB_ROWS(M) case M: blend_rows<M, OPAQUE>(d, s, fa, softg, ok); break;
        MVB_ROWS(0) MVB_ROWS(1) MVB_ROWS(2) MVB_ROWS(3) MVB_ROWS(4) MVB_ROWS(5) MVB_ROWS(6) MVB_ROWS(7)
        MVB_ROWS(8) MVB_ROWS(9) MVB_ROWS(10) MVB_ROWS(11) MVB_ROWS(12) MVB_ROWS(13) MVB_ROWS(14)
        MVB_ROWS(15) MVB_ROWS(16) MVB_ROWS(17) MVB_ROWS(18)
#undef MVB_ROWS
    default: break;
    }
}

// Position of the n-th (0-based) set bit of `mask` (n below its popcount, n small).
__device__ __forceinline__ int nth_set_bit(unsigned mask, int n)
{
    for (int q = 0; q < n; ++q) mask &= mask - 1;
    return __ffs((int)mask) - 1;
}

// One CTA per 32 x 32 output tile, 256 threads, thread = 1 column x 4 rows, pixels in registers.
//
// Source texels reach the SM through TMA: for every layer that touches the tile, one elected
// thread requests the box of the source image that covers the tile's footprint (DevLayer::box_w
// x box_h px at the tile's own origin, zero-filled outside the image = cv2's BORDER_CONSTANT)
// into one of two shared-memory stages, one layer ahead of the arithmetic; the 4 taps of a
// sample are then 4 LDS at tile-local offsets with no bounds logic.  Layers whose source is not
// 16-byte aligned / too strongly minified for a stage gather through L1 instead (LDG).
__global__ void __launch_bounds__(kFWarps * 32, 4)
k_stack_opaque(const __grid_constant__ StackParamsTma PT)
{
    const StackParams& P = PT.s;
    extern __shared__ __align__(1024) uint8_t s_stage[];   // 2 stages of PT.stage_bytes
    __shared__ SlotMem s_slot[kFSlots];         // per resident layer: info, bbox, warp tables, alpha LUT
    __shared__ float2 s_softg[256];
    __shared__ __align__(8) unsigned long long s_bar[2 * kMaxStages];   // full[], empty[]
    __shared__ int s_class[kMaxLayers];         // per layer: -1 (misses the tile) or kTile* | (partial coverage ? 256 : 0)
    __shared__ int2 s_org[kMaxLayers];          // per layer: TMA box origin in source px
    __shared__ int s_layer[kFSlots];            // per resident layer: layer index | index among the tile's TMA layers << 8

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kFTileH;
    const uint32_t bar_full = smem_addr(&s_bar[0]), bar_empty = smem_addr(&s_bar[kMaxStages]);
    const uint32_t stage0 = smem_addr(s_stage);
    const int ns_log2 = PT.stages_log2, ns_mask = (1 << ns_log2) - 1;   // 2 or 4 stages

    // ---- cull + classify the layers against this tile ---------------------------------------------
    // Warp 0 classifies one layer per lane; the others meanwhile set up their pixels.
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, kFWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    s_softg[tid] = g_blend_tables.softg[tid];   // unconditional: a predicated LDG makes ptxas re-derive the
                                                // global-memory descriptor (R2UR) in front of every load
    if (warp == 0) {
        int cls = -1;   // -1: the layer does not touch this tile
        int2 org = make_int2(0, 0);
        if (lane < P.n_layers) {
            const DevLayer& L = P.layers[lane];
            const int bx = L.bx, by = L.by, bw = L.bw, bh = L.bh, src_w = L.src_w, src_h = L.src_h;
            const int box_w = L.box_w, box_h = L.box_h;
            const float m0 = L.mf[0], m1 = L.mf[1], m2 = L.mf[2], m3 = L.mf[3], m4 = L.mf[4], m5 = L.mf[5];
            // tile ∩ bbox ∩ frame, in bbox coordinates (inclusive corners)
            const int u0 = max(tx0, bx) - bx, u1 = min(min(tx0 + kTileW, bx + bw), P.W) - 1 - bx;
            const int v0 = max(ty0, by) - by, v1 = min(min(ty0 + kFTileH, by + bh), P.H) - 1 - by;
            if (bw > 0 && bh > 0 && u0 <= u1 && v0 <= v1) {
                // source-space extent of that rectangle (affine: centre +- sum |coef| * half size); 0.25 px
                // of slack covers the 1/32 px coordinate quantisation and its +16/1024 rounding offset
                // (fp32: positions below 32768 px are good to 0.01 px, far inside that slack)
                const float uc = 0.5f * (float)(u0 + u1), vc = 0.5f * (float)(v0 + v1);
                const float hu = 0.5f * (float)(u1 - u0), hv = 0.5f * (float)(v1 - v0);
                const float xc = m0 * uc + m1 * vc + m2, ex = fabsf(m0) * hu + fabsf(m1) * hv;
                const float yc = m3 * uc + m4 * vc + m5, ey = fabsf(m3) * hu + fabsf(m4) * hv;
                const float xmin = xc - ex, xmax = xc + ex, ymin = yc - ey, ymax = yc + ey;
                const bool full = u1 - u0 == kTileW - 1 && v1 - v0 == kFTileH - 1;
                // (no tap inside the source: a no-op on an opaque frame -> stays -1)
                if (!(xmax < -1.25f || xmin > src_w + 0.25f || ymax < -1.25f || ymin > src_h + 0.25f)) {
                    cls = kTileEdge;
                    // taps span [floor(xmin) - 1, floor(xmax) + 2] at most (same slack as above); TMA wants
                    // the box to start on a 16-byte boundary of the row: x origin a multiple of 4 px
                    const int fx0 = ((int)floorf(xmin) - 1) & ~3, fy0 = (int)floorf(ymin) - 1;
                    if (box_w > 0 && (int)floorf(xmax) + 3 - fx0 <= box_w && (int)floorf(ymax) + 3 - fy0 <= box_h) {
                        cls = full ? kTileTma : kTileTmaPart;
                        org = make_int2(fx0, fy0);
                    } else if (full && xmin > 0.25f && xmax < src_w - 1.25f && ymin > 0.25f && ymax < src_h - 1.25f) {
                        cls = kTileInterior;   // whole tile covered and every tap inside the source
                    }
                }
            }
        }
        s_class[lane] = cls;
        s_org[lane] = org;
    }

    // ---- the thread's pixels: column `lane`, rows [4 warp, 4 warp + 4), in registers --------------
    const int x = tx0 + lane;
    const int rbeg = warp * kFRows;
    const bool x_ok = x < P.W;
    const PixelCtx px{tx0, ty0, x, rbeg, lane, x_ok};
    RowPairState d[2];
    if (P.keep_frame) {   // continuing a stack of more than kMaxLayers layers: alpha is 255 already
        uint32_t v[kFRows];
#pragma unroll
        for (int k = 0; k < kFRows; ++k)   // clamped, not predicated (see s_softg above); out-of-frame threads never store
            v[k] = reinterpret_cast<const uint32_t*>(P.frame + (int64_t)min(ty0 + rbeg + k, P.H - 1) * P.stride)[min(x, P.W - 1)];
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                d[p].c[c] = pk(__uint_as_float(__byte_perm(v[2 * p], kMagicBits, 0x7650 + c)),
                               __uint_as_float(__byte_perm(v[2 * p + 1], kMagicBits, 0x7650 + c)));
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) d[0].c[c] = d[1].c[c] = bc(kMagic + (float)((P.bg >> (8 * c)) & 0xFFu));
    }
    __syncthreads();
    // every warp derives the paint-ordered list of active layers from the class table by itself
    const int my_cls = s_class[lane];
    const unsigned m_act = __ballot_sync(0xFFFFFFFFu, my_cls >= 0);
    const unsigned m_tma = __ballot_sync(0xFFFFFFFFu, my_cls >= 0 && (my_cls & 3) == kTileTma);
    const int n_active = __popc(m_act);
    const int n_tma = __popc(m_tma);

    // Request the box of the k-th TMA layer of this tile into stage k mod n_stages (thread 0 only).
    auto request = [&](int k) {
        const int li = nth_set_bit(m_tma, k);
        const DevLayer& L = P.layers[li];
        const int2 org = s_org[li];
        const uint32_t bar = bar_full + 8 * (k & ns_mask);
        mbar_arrive_expect_tx(bar, (uint32_t)(L.box_w * L.box_h * 4));
        tma_load_box(stage0 + (k & ns_mask) * PT.stage_bytes, &PT.maps[li], org.x, org.y, bar);
    };
    if (tid == 0)
        for (int k = 0; k < min(n_tma, ns_mask + 1); ++k) request(k);

    unsigned m_rest = m_act;   // active layers not yet walked, in paint order
    for (int base = 0; base < n_active; base += kFSlots) {
        const int n_here = min(kFSlots, n_active - base);
        if (base > 0) __syncthreads();
        for (int i = tid; i < n_here * (kTileW + kFTileH); i += kFWarps * 32) {
            const int slot = i / (kTileW + kFTileH), j = i % (kTileW + kFTileH);
            const int li = nth_set_bit(m_rest, slot);
            const DevLayer& L = P.layers[li];
            if (j < kTileW) {
                const int u = tx0 + j - L.bx;
                s_slot[slot].col[j] = make_int2(warp_adelta(L.m, u), warp_bdelta(L.m, u));
            } else {
                const int v = ty0 + (j - kTileW) - L.by;
                const int2 org = s_org[li];   // (0, 0) unless the layer is staged by TMA
                s_slot[slot].row[j - kTileW] = make_int2(warp_x0(L.m, v) - org.x * 1024, warp_y0(L.m, v) - org.y * 1024);
            }
        }
        if (tid < n_here) {   // what the layer loop needs, one 16-byte read per layer
            const int li = nth_set_bit(m_rest, tid);
            const DevLayer& L = P.layers[li];
            const int cls = s_class[li];
            const int k = (cls & 3) == kTileTma ? __popc(m_tma & ((1u << li) - 1u)) : 0;
            // x: mode (8 bits) | class (3) | phase parity (bit 11) |
            //    refill: thread 0 re-arms the stage of TMA layer k-1 before this layer (bit 12) | row pitch << 16
            // y: address of the stage's "full" barrier ("empty" is 8 * kMaxStages bytes on)   z: &lut[-kMagicBits]   w: stage address
            const int refill = k >= 1 && k + ns_mask < n_tma;
            s_slot[tid].info = make_int4((L.pil ? kPilOver : L.mode) | ((cls & 7) << 8) |
                                        (((k >> ns_log2) & 1) << 11) | (refill << 12) | ((L.box_w * 4) << 16),
                                    (int)(bar_full + 8 * (k & ns_mask)),
                                    (int)(smem_addr(s_slot[tid].lut) - (kMagicBits << 2)),
                                    (int)(stage0 + (k & ns_mask) * PT.stage_bytes));
            s_layer[tid] = li | (k << 8);
        }
        for (int slot = 0; slot < n_here; ++slot, m_rest &= m_rest - 1) {   // 256 threads: one LUT entry each
            const DevLayer& L = P.layers[__ffs((int)m_rest) - 1];
            uint32_t v = (uint32_t)tid;
            if (L.opacity < 1.0)
                v = L.pil ? (uint32_t)(tid * L.pil_a) / 255u                        // imgproc.py:206-208
                          : __double2uint_rz(__dmul_rn((double)tid, L.opacity));    // imgproc.py:159
            s_slot[slot].lut[tid] = (float)v;
        }
        __syncthreads();

#pragma unroll 1
        for (int slot = 0; slot < n_here; ++slot) {
            walk_layer<true>(P, s_slot[slot], &s_layer[slot], px, d, s_softg, [&](const int4& info) {
                if (tid == 0 && (info.x & (1 << 12))) {   // the stage of layer k-1 is free once every warp has read it
                    const int k = s_layer[slot] >> 8;
                    mbar_wait(bar_empty + 8 * ((k - 1) & ns_mask), ((k - 1) >> ns_log2) & 1);
                    request(k + ns_mask);
                }
            });
        }
    }

    // ---- epilogue: one coalesced 128-byte store per warp row ---------------------------------------
    if (x_ok) {
#pragma unroll
        for (int k = 0; k < kFRows; ++k) {
            if (ty0 + rbeg + k >= P.H) break;
            float r0, r1, g0, g1, b0, b1;
            upk(d[k >> 1].c[0], r0, r1);
            upk(d[k >> 1].c[1], g0, g1);
            upk(d[k >> 1].c[2], b0, b1);
            const uint32_t rb = __float_as_uint((k & 1) ? r1 : r0), gb = __float_as_uint((k & 1) ? g1 : g0);
            const uint32_t bb = __float_as_uint((k & 1) ? b1 : b0);
            const uint32_t px = __byte_perm(__byte_perm(rb, gb, 0x0040), bb, 0x0410) | 0xFF000000u;
            reinterpret_cast<uint32_t*>(P.frame + (int64_t)(ty0 + rbeg + k) * P.stride)[x] = px;
        }
    }
}

} // namespace mvb
