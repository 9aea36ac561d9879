// stack_float.cuh -- arithmetic of the fused warp + blend kernel (k_stack_frames, stack_frames.cuh), sm_100a:
// exact float forms of cv2's bilinear, of both "over" operators and of the 18 blend modes, the TMA /
// mbarrier plumbing, and walk_layer, one layer of the walk for one thread.
//
// Contract: per frame the kernel replaces the reference's cv2.warpAffine -> alpha_composite pair of every
// layer (movis/layer/composition.py:811-819) and the frame walk around it (composition.py:363-368), bit
// for bit.  An opaque background keeps the frame's alpha at 255 through every layer (both "over"
// operators map dst alpha 255 to 255), so on OPAQUE frames both reduce to
//     out = floor((fa*T + (255-fa)*b [+127 for PIL]) / 255)   per colour channel
// (imgproc.py:156-165 with oa = 65025; PIL's ImagingAlphaComposite with outa255 = 65025); translucent
// frames carry alpha as a fourth channel and use the general forms (over_numpy_rows / over_pil_rows).
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
// two rows.  The reference's x = fa*T + (255 - fa)*b = 255*b + y with y = fa*(T - b) an integer in
// [-65025, 65025], so floor(x / 255) = b + floor(y / 255) (numpy) and PIL's rint(x / 255) =
// b + floor((y + 127.5) / 255).  (y + h) / 255 with h = 0.5 resp. 127.5 is never closer than 0.5 / 255
// to an integer, so multiplying by fl(1/255) instead (relative error 6e-8, absolute below 2e-5) keeps
// the floor for either sign of y, and the fma rounded towards -inf adds it to the biased b exactly.
template <int MODE>
__device__ __forceinline__ u64 blend_chan(u64 dB, u64 sB, u64 fa, const float2* softg)
{
    const u64 yh = fma2(fa, blend_diff2<MODE>(dB, sB, softg), bc(MODE == 18 ? 127.5f : 0.5f));
    return fma2_rm(yh, bc(kInv255), dB);
}

struct RowPairState { u64 c[4]; };   // r, g, b, a of rows (2p, 2p+1), biased (a is unused on opaque frames)

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

// Blend the samples of 2 * PAIRS rows into the thread's pixels.  `keep`: bit q set = row q of this
// thread lies inside the layer's bbox (only consulted on general frames, where a transparent sample
// is not a no-op for the numpy arithmetic).
template <int MODE, bool OPAQUE, int PAIRS>
__device__ __forceinline__ void blend_rows(RowPairState (&d)[PAIRS], const u64 (&s)[PAIRS][4], const u64 (&fa)[PAIRS],
                                           const float2* softg, unsigned keep)
{
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) {
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
// Blocks until the phase with the given parity has completed.  The common case is one try_wait and a
// branch.  mbarrier.try_wait suspends the thread in the barrier unit for a hardware-defined time, so a
// longer wait is a run of probes: eight straight ones (most waits in this kernel are for data that is
// about to land), then a counted loop.  It is bounded by WALL TIME (%globaltimer, looked at every 2^18
// probes), not by a spin count: a slow but live run -- under a debugger, a profiler replay or time
// slicing -- is left alone, and only a wait that has made no progress for 20 s (a broken descriptor,
// i.e. a bug) traps instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
#define MVB_PROBE "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra MVB_DONE;\n"
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .u32 n;\n"
        ".reg .u64 t0, t1;\n"
        MVB_PROBE MVB_PROBE MVB_PROBE MVB_PROBE MVB_PROBE MVB_PROBE MVB_PROBE MVB_PROBE MVB_PROBE
        "mov.u64 t0, %%globaltimer;\n"
        "MVB_AGAIN:\n"
        "mov.u32 n, 0;\n"
        "MVB_WAIT:\n"
        MVB_PROBE MVB_PROBE MVB_PROBE MVB_PROBE
        "add.u32 n, n, 1;\n"
        "setp.lt.u32 q, n, 65536;\n"
        "@q bra MVB_WAIT;\n"
        "mov.u64 t1, %%globaltimer;\n"
        "sub.u64 t1, t1, t0;\n"
        "setp.gt.u64 q, t1, 20000000000;\n"
        "@q trap;\n"
        "bra MVB_AGAIN;\n"
        "MVB_DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
#undef MVB_PROBE
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
// 1-D bulk copy global -> shared memory, completion signalled on `bar` in bytes (16-byte granular).
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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
// How a layer reaches a tile (decided per tile by k_stack_prep, refined by the producer warp):
//   kTileEdge      taps gathered through L1 with full bounds / coverage logic (always valid)
//   kTileInterior  tile fully inside the bbox and every tap inside the source: unchecked L1 gather
//   kTileTma       taps come from a TMA-staged box, tile fully inside the bbox (the common case)
//   kTileTmaPart   TMA-staged box, tile partly outside the bbox
//   kTileTmaLut / kTileTmaPartLut   the same for a layer whose opacity needs its table (kLayerLut)
// Odd codes are the ones that use a TMA stage.
enum { kTileEdge = 0, kTileTmaLut = 1, kTileInterior = 2, kTileTma = 3, kTileTmaPartLut = 5, kTileTmaPart = 7 };
// With kTileTma only: the layer is shifted by whole pixels (kLayerIntShift), so a sample IS a texel: one tap
// per pixel, no weights, no bilinear sums (cv2: fx = fy = 0, (1024 p + 512) >> 10 = p).
constexpr int kInfoOneTap = 1 << 12;

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

// Frame geometry shared by every frame of a launch.
struct FrameGeom { int W, H, stride; };

// What identifies a thread's pixels inside the tile being walked.
struct PixelCtx {
    int tx0, ty0;   // tile origin in the frame
    int x;          // the thread's column
    int rbeg;       // first of its rows, relative to ty0
    int lane;
    bool x_ok;      // x < frame width
};

// One layer of the walk for one thread: fetch the taps of its ROWS = 4 * QUADS rows (from the TMA-staged
// box, or through L1 for layers that cannot be staged), cv2's fixed-point bilinear, the opacity scale,
// blend into the pixels `d`.
//
// Opacity: the reference scales the sampled alpha a by the layer's opacity as an integer table
//   numpy path  fa = trunc(float64(a) * opacity)          (imgproc.py:159)
//   PIL path    fa = a * int(round(255 opacity)) // 255   (imgproc.py:206-208)
// k_stack_prep finds, per layer, a constant c = m / 2^18 with floor(a c) == table[a] for all 256 a
// (it exists unless float64 rounding made the table non-monotone in a / fa, which happens for
// opacities a few ulps below a fraction of small integers).  Then, with the alpha biased as
// everything else here (magic + a):   fma.rm(magic + a, c, magic - magic c) = magic + floor(a c),
// exactly: magic c = 48 m and the addend are representable, the product is exact inside the fma.
// Layers without such a constant (opacity 0.7 is one) carry their 256-entry table in global memory and
// travel through tile classes of their own (kTile*Lut), which look the value up; layers with a
// constant pay nothing for that.
template <bool OPAQUE, int QUADS, bool ONE_TAP, class Slot>
__device__ __forceinline__ void walk_layer(const FrameGeom& G, const DevLayer* __restrict__ layers, const Slot& S,
                                           const int* layer_index, const PixelCtx& C,
                                           RowPairState (&d)[2 * QUADS], const float2* softg)
{
    constexpr int ROWS = 4 * QUADS, PAIRS = 2 * QUADS;
    const int ty0 = C.ty0, x = C.x, rbeg = C.rbeg, lane = C.lane;
    const bool x_ok = C.x_ok;
    const int4 info = S.info;
    const int cls = (info.x >> 8) & 7;
    const int2 col = S.col[lane];
    const int2* __restrict__ rowtab = &S.row[rbeg];

    // ---- taps of the rows, all requested before the first is used ----------------------------
    uint32_t t[ROWS][4];
    uint32_t wt[ROWS], wb[ROWS];   // cv2's 2 x 2 weights, two 16-bit values each: (w00, w01), (w10, w11)
    unsigned ok = (1u << ROWS) - 1u;   // bit q: row q of this thread lies inside the layer's bbox and the frame
    float oc = __int_as_float(info.z); // the layer's opacity constant c (see above)
    // A layer whose opacity table is not floor(a c) for any c: take the bilinear alpha here, look the scaled
    // value v up and give all four taps the alpha v.  The common code below then samples exactly v (the
    // weights sum to 1024) and, with c = 1, leaves it as it is.  Off the common path: only layers flagged
    // kLayerLut come here, through their own tile classes.
    auto apply_table = [&](const float* __restrict__ lut) {
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
            const uint32_t t_ba = __byte_perm(t[k][0], t[k][1], 0x7362), b_ba = __byte_perm(t[k][2], t[k][3], 0x7362);
            const uint32_t a = __dp2a_hi(wb[k], b_ba, __dp2a_hi(wt[k], t_ba, 32u * 512u)) >> 15;
            const uint32_t v = (uint32_t)__ldg(lut + a) << 24;
#pragma unroll
            for (int j = 0; j < 4; ++j) t[k][j] = (t[k][j] & 0x00FFFFFFu) | v;
        }
        oc = 1.0f;
    };
    auto staged = [&](auto part_tag, auto lut_tag) {
        const uint32_t pitch = (uint32_t)info.x >> 16, stage = (uint32_t)info.w;
        // Tiles the bbox covers only partly run their own copy of the loop (cls == kTileTmaPart): the box
        // spans the covered part alone, so samples outside it read a dummy address (softg) with weight 0, which
        // makes them transparent (alpha 0 -> fa = 0).  Fully covered tiles carry none of that.
        constexpr bool PART = decltype(part_tag)::value;
        if (PART) {
            const DevLayer& L = layers[*layer_index];
            const bool col_in = x_ok && (unsigned)(x - L.bx) < (unsigned)L.bw;
            ok = 0u;
#pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                const int yy = ty0 + rbeg + q;
                if (col_in && (unsigned)(yy - L.by) < (unsigned)L.bh && yy < G.H) ok |= 1u << q;
            }
        }
        // addresses and weights first: they do not need the texels, so they overlap the box's arrival
        uint32_t a[ROWS], down[ROWS];
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            const int2 r = rowtab[q];
            const int sumx = r.x + col.x, sumy = r.y + col.y;      // 1/1024 px, relative to the box origin
            a[q] = imad((uint32_t)(sumy >> 10), pitch, imad((uint32_t)(sumx >> 10), 4u, stage));
            tap_weights(sumx, sumy, wt[q], wb[q]);
            down[q] = pitch;
            if (PART && !(ok >> q & 1)) { a[q] = smem_addr(softg); down[q] = 0u; wt[q] = wb[q] = 0u; }
        }
        mbar_wait((uint32_t)info.y, (info.x >> 11) & 1);
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            t[q][0] = lds_u32(a[q]);
            t[q][1] = lds_u32(a[q] + 4);
            t[q][2] = lds_u32(a[q] + down[q]);
            t[q][3] = lds_u32(a[q] + down[q] + 4);
        }
        mbar_arrive_warp((uint32_t)info.y + 8 * kMaxStages);   // this warp is done with the stage
        if (decltype(lut_tag)::value) apply_table(layers[*layer_index].lut);
    };
    // (an opaque copy of the info word keeps the compiler from folding this test into the switch over the
    // other classes, which it orders with the common case last)
    uint32_t hot;
    asm("mov.b32 %0, %1;" : "=r"(hot) : "r"(info.x));
    if (ONE_TAP) {
        // whole-pixel shift (kInfoOneTap): texel (x + m2, y + m5) of the staged box, rows one pitch apart
        const uint32_t pitch = (uint32_t)info.x >> 16;
        const int2 r0 = rowtab[0];
        const uint32_t a0 = imad((uint32_t)((r0.y + col.y) >> 10), pitch, imad((uint32_t)((r0.x + col.x) >> 10), 4u, (uint32_t)info.w));
        mbar_wait((uint32_t)info.y, (info.x >> 11) & 1);
#pragma unroll
        for (int q = 0; q < ROWS; ++q) t[q][0] = lds_u32(a0 + (uint32_t)q * pitch);
        mbar_arrive_warp((uint32_t)info.y + 8 * kMaxStages);
    } else if ((hot & 0x700u) == (uint32_t)kTileTma << 8) {
        staged(std::false_type{}, std::false_type{});
    } else if (cls == kTileTmaPart) {
        staged(std::true_type{}, std::false_type{});
    } else if (cls == kTileTmaLut) {
        staged(std::false_type{}, std::true_type{});
    } else if (cls == kTileTmaPartLut) {
        staged(std::true_type{}, std::true_type{});
    } else if (cls == kTileInterior) {
        const DevLayer& L = layers[*layer_index];
        const uint8_t* __restrict__ src = L.src;
        const int ss = L.src_stride;
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
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
        const DevLayer& L = layers[*layer_index];
        const uint8_t* __restrict__ src = L.src;
        const int ss = L.src_stride;
        const int sw = L.src_w, sh = L.src_h;
        const bool col_in = x_ok && (unsigned)(x - L.bx) < (unsigned)L.bw;
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
            const int2 r = rowtab[k];
            const int sumx = r.x + col.x, sumy = r.y + col.y;
            const int yy = ty0 + rbeg + k;
            t[k][0] = t[k][1] = t[k][2] = t[k][3] = 0u;
            if (col_in && (unsigned)(yy - L.by) < (unsigned)L.bh && yy < G.H) {
                const Taps tp = fetch_taps(src, sw, sh, ss, sumx >> 5, sumy >> 5);
                t[k][0] = tp.p00; t[k][1] = tp.p01; t[k][2] = tp.p10; t[k][3] = tp.p11;
            } else {
                ok &= ~(1u << k);
            }
            tap_weights(sumx, sumy, wt[k], wb[k]);
        }
        if (L.flags & kLayerLut) apply_table(L.lut);
    }

    // ---- cv2's fixed-point bilinear: out = (sum w p + 512) >> 10 -------------------------------
    // IDP.2A multiplies the two 16-bit weights of a tap row with two bytes: with the taps of a row paired
    // by channel, four of them give one channel's exact sum; started from the bits of 2^23 + 32 * 512 the
    // accumulator IS the float 2^23 + 32 (512 + sum), and one fma.rm takes floor(. / 1024) of a row pair.
    u64 s[PAIRS][4], fa[PAIRS];
    const u64 oc2 = bc(oc), ok2 = bc(fmaf(oc, -kMagic, kMagic));
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) {
        if (ONE_TAP) {   // the texel is the sample: bytes to biased floats
#pragma unroll
            for (int c = 0; c < 4; ++c)
                s[p][c] = pk(__uint_as_float(__byte_perm(t[2 * p][0], kMagicBits, 0x7650 + c)),
                             __uint_as_float(__byte_perm(t[2 * p + 1][0], kMagicBits, 0x7650 + c)));
        } else {
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
        }
        // opacity-scaled alpha, plain (not biased): floor(a c) for the two rows in one fma.rm + one subtract
        fa[p] = sub2(fma2_rm(s[p][3], oc2, ok2), bc(kMagic));
    }

    // ---- blend, in the code specialised for this layer's mode ------------------------------------
    if (ONE_TAP) {   // only NORMAL layers take this instantiation (numpy's or PIL's arithmetic): the kernel's
                     // code must stay within reach of the instruction cache, a second copy of all 19 bodies does not
        if ((info.x & 255) == 18) blend_rows<18, OPAQUE, PAIRS>(d, s, fa, softg, ok);
        else blend_rows<0, OPAQUE, PAIRS>(d, s, fa, softg, ok);
        return;
    }
#ifdef MVB_DIAG_MODE0      // (diagnostic build: every layer blends as NORMAL)
    switch (0) {
#else
    switch (info.x & 255) {
#endif
#define MVB_ROWS(M) case M: blend_rows<M, OPAQUE, PAIRS>(d, s, fa, softg, ok); break;
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

} // namespace mvb
