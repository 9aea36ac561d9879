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
//   * vertical lerp:  IDP.4A on byte-interleaved taps, accumulator preset to the bits of 2^23,
//                     so the dot product comes out as a float (2^23 + v) with no conversion;
//   * horizontal lerp + (S+512)>>10: three FFMA2 and one FADD2 per channel for TWO rows;
//   * blend + composite + floor(./255): four FFMA2/FADD2 per channel for TWO rows (NORMAL).
// The tile's pixels stay in registers for the whole layer walk (thread = 1 column x 4 rows).
#pragma once

#include "pixel_math.cuh"

namespace mvb {

typedef unsigned long long u64;

constexpr float kMagic = 12582912.0f;          // 1.5 * 2^23: ulp 1, so RN/RM of (magic + x) rounds x to an integer
constexpr uint32_t kMagicBits = 0x4B400000u;   // float bits of kMagic; bits of (kMagic + k) = kMagicBits + k
constexpr uint32_t kTwo23Bits = 0x4B000000u;   // float bits of 2^23
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
// T(b, f) - b for one lane, biased inputs (dB = magic + b, sB = magic + f): the blend targets of
// movis/imgproc.py:11-133 in the closed forms of SURVEY.md 8(a-5), minus the background.
// ---------------------------------------------------------------------------------------------

// min(255, floor(num255 / den)) + magic, num255 = 255 * n (n <= 255), 1 <= den <= 256.  The
// quotient of (num + 0.5) by den is at least 0.5/256 away from every integer, MUFU.RCP and the
// multiply are within 2^-22 relative, and only quotients below 256 matter (the rest clamp).
__device__ __forceinline__ float ratio255_biased(float n, float den)
{
    const float t = fmaf(n, 255.0f, 0.5f) * rcp_approx(den);
    return fminf(__fadd_rd(t, kMagic), kMagic + 255.0f);
}

template <int MODE>
__device__ __forceinline__ float blend_diff(float dB, float sB, const uint32_t* softg)
{
    if constexpr (MODE == 4) return fminf(sB, dB) - dB;                           // darken
    else if constexpr (MODE == 5) return fmaxf(sB, dB) - dB;                      // lighten
    else if constexpr (MODE == 3 || MODE == 10) {                                 // overlay / hard_light
        const float b = dB - kMagic, f = sB - kMagic;
        const float nb = 255.0f - b, nf = 255.0f - f;
        const bool lo = (MODE == 3 ? b : f) < 128.0f;
        const float bb = lo ? b : nb, ff = lo ? f : nf;
        const float q = __fmaf_rd(bb * ff, 2.0f * kInv255, kMagic) - kMagic;      // floor(2 bb ff / 255)
        const float t = q - bb;                                                   // lo: T - b; else -(T - b)
        return lo ? t : -t;
    } else if constexpr (MODE == 6) {                                             // color_dodge
        return ratio255_biased(dB - kMagic, (kMagic + 256.0f) - sB) - dB;
    } else if constexpr (MODE == 7) {                                             // color_burn
        const float nb = (kMagic + 255.0f) - dB;
        return (nb + kMagic) - ratio255_biased(nb, sB - (kMagic - 1.0f));         // 255 - q - b
    } else if constexpr (MODE == 8) {                                             // linear_dodge
        return fminf((kMagic + 255.0f) - dB, sB - kMagic);                        // min(255 - b, f)
    } else if constexpr (MODE == 9) {                                             // linear_burn
        return fmaxf(kMagic - dB, sB - (kMagic + 255.0f));                        // max(-b, f - 255)
    } else if constexpr (MODE == 11) {                                            // soft_light (integer form)
        const uint32_t bi = __float_as_uint(dB) & 0xFFu, fi = __float_as_uint(sB) & 0xFFu;
        const BlendTables tabs{nullptr, softg};
        return __uint_as_float(blend_target_tab<11>(bi, fi, tabs) | kMagicBits) - dB;
    } else if constexpr (MODE == 12) {                                            // vivid_light (sic)
        const float f = sB - kMagic;
        const float nb = (kMagic + 255.0f) - dB;
        const float d = (nb + kMagic) - ratio255_biased(nb, fmaf(f, 2.0f, 1.0f));
        return f > 127.0f ? kMagic - dB : d;                                      // f > 127: T = 0
    } else if constexpr (MODE == 13) {                                            // linear_light
        const float g = fmaf(sB - kMagic, 2.0f, -255.0f);                         // clamp(b + 2f - 255, 0, 255) - b
        return fminf(fmaxf(g, kMagic - dB), (kMagic + 255.0f) - dB);
    } else if constexpr (MODE == 14) {                                            // pin_light
        const float hB = fmaf(sB - kMagic, 2.0f, kMagic);                         // magic + 2f
        return fminf(fmaxf(dB, hB - 255.0f), hB) - dB;                            // min(max(b, 2f - 255), 2f) - b
    } else if constexpr (MODE == 15) {                                            // difference
        return fabsf(dB - sB) - (dB - kMagic);
    } else {                                                                      // subtract
        static_assert(MODE == 17, "scalar blend_diff covers the min/max/select modes only");
        return fmaxf(kMagic - dB, kMagic - sB);                                   // max(0, b - f) - b
    }
}

// One colour channel of two rows: blend + composite + requantise.  Returns magic + new value.
template <int MODE>
__device__ __forceinline__ u64 blend_chan(u64 dB, u64 sB, u64 fa, const uint32_t* softg)
{
    u64 diff;
    if constexpr (MODE == 0 || MODE == 18) {                                      // normal / PIL over
        diff = sub2(sB, dB);
    } else if constexpr (MODE == 1) {                                             // multiply
        const u64 p = mul2(add2(dB, bc(-kMagic)), add2(sB, bc(-kMagic)));
        diff = sub2(fma2_rm(p, bc(kInv255), bc(kMagic)), dB);
    } else if constexpr (MODE == 2) {                                             // screen
        const u64 nb = sub2(bc(kMagic + 255.0f), dB), nf = sub2(bc(kMagic + 255.0f), sB);
        diff = sub2(add2(nb, bc(kMagic)), fma2_rm(mul2(nb, nf), bc(kInv255), bc(kMagic))); // 255 - q - b
    } else if constexpr (MODE == 16) {                                            // exclusion: f - floor(2bf/255)
        const u64 p = mul2(add2(dB, bc(-kMagic)), add2(sB, bc(-kMagic)));
        diff = sub2(sB, fma2_rm(p, bc(2.0f * kInv255), bc(kMagic)));
    } else {
        float d0, d1, s0, s1;
        upk(dB, d0, d1);
        upk(sB, s0, s1);
        diff = pk(blend_diff<MODE>(d0, s0, softg), blend_diff<MODE>(d1, s1, softg));
    }
    // x = fa*T + (255 - fa)*b = fa*(T - b) + 255*b, an integer in [0, 65025]
    const u64 x = fma2(fa, diff, fma2(dB, bc(255.0f), bc(kNeg255Magic)));
    if constexpr (MODE == 18) return fma2(x, bc(kInv255), bc(kMagic));            // PIL: floor((x+127)/255) = rint(x/255)
    else return fma2_rm(x, bc(kInv255), bc(kMagic));                              // numpy: floor(x/255)
}

constexpr int kFWarps = 8;   // 256 threads: 8 warps x 4 rows = one 32 x 32 tile
constexpr int kFRows = 4;
constexpr int kFTileH = kFWarps * kFRows;
constexpr int kFSlots = 8;   // layers whose per-tile tables are resident in shared memory at a time

struct RowPairState { u64 c[3]; };   // r, g, b of rows (2p, 2p+1), biased

template <int MODE>
__device__ __forceinline__ void blend_rows(RowPairState (&d)[2], const u64 (&s)[2][4], const u64 (&fa)[2],
                                           const uint32_t* softg)
{
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int c = 0; c < 3; ++c) d[p].c[c] = blend_chan<MODE>(d[p].c[c], s[p][c], fa[p], softg);
}

__global__ void __launch_bounds__(kFWarps * 32, 4)
k_stack_opaque(const __grid_constant__ StackParams P)
{
    __shared__ int2 s_col[kFSlots][kTileW];     // (adelta, bdelta) per tile column
    __shared__ int2 s_row[kFSlots][kFTileH];    // (X0, Y0) per tile row
    __shared__ float s_lut[kFSlots][256];       // opacity-scaled alpha, as float
    __shared__ uint32_t s_softg[256];
    __shared__ int s_active[kMaxLayers];
    __shared__ int s_class[kMaxLayers];
    __shared__ int s_nactive;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * kTileW, ty0 = blockIdx.y * kFTileH;

    // ---- cull + classify the layers against this tile (as in k_composite_stack) -----------------
    if (warp == 0) {
        bool hit = false;
        int cls = kTileEdge;
        if (lane < P.n_layers) {
            const DevLayer& L = P.layers[lane];
            const int u0 = max(tx0, L.bx) - L.bx, u1 = min(min(tx0 + kTileW, L.bx + L.bw), P.W) - 1 - L.bx;
            const int v0 = max(ty0, L.by) - L.by, v1 = min(min(ty0 + kFTileH, L.by + L.bh), P.H) - 1 - L.by;
            hit = L.bw > 0 && L.bh > 0 && u0 <= u1 && v0 <= v1;
            if (hit) {
                double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double u = (k & 1) ? u1 : u0, v = (k & 2) ? v1 : v0;
                    const double xs = L.m[0] * u + L.m[1] * v + L.m[2];
                    const double ys = L.m[3] * u + L.m[4] * v + L.m[5];
                    xmin = fmin(xmin, xs); xmax = fmax(xmax, xs);
                    ymin = fmin(ymin, ys); ymax = fmax(ymax, ys);
                }
                const bool full = u1 - u0 == kTileW - 1 && v1 - v0 == kFTileH - 1;
                if (xmax < -1.25 || xmin > L.src_w + 0.25 || ymax < -1.25 || ymin > L.src_h + 0.25)
                    hit = false;           // no tap inside the source: a no-op on an opaque frame
                else if (full && xmin > 0.25 && xmax < L.src_w - 1.25 && ymin > 0.25 && ymax < L.src_h - 1.25)
                    cls = kTileInterior;   // whole tile covered and every tap inside the source
            }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, hit);
        if (hit) {
            const int slot = __popc(m & ((1u << lane) - 1u));
            s_active[slot] = lane;
            s_class[slot] = cls;
        }
        if (lane == 0) s_nactive = __popc(m);
    }
    s_softg[tid] = g_blend_tables.softg[tid];   // unconditional: a predicated LDG makes ptxas re-derive the
                                                // global-memory descriptor (R2UR) in front of every load

    // ---- the thread's pixels: column `lane`, rows [4 warp, 4 warp + 4), in registers --------------
    const int x = tx0 + lane;
    const int rbeg = warp * kFRows;
    const bool x_ok = x < P.W;
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
    const int n_active = s_nactive;

    for (int base = 0; base < n_active; base += kFSlots) {
        const int n_here = min(kFSlots, n_active - base);
        if (base > 0) __syncthreads();
        for (int i = tid; i < n_here * (kTileW + kFTileH); i += kFWarps * 32) {
            const int slot = i / (kTileW + kFTileH), j = i % (kTileW + kFTileH);
            const DevLayer& L = P.layers[s_active[base + slot]];
            if (j < kTileW) {
                const int u = tx0 + j - L.bx;
                s_col[slot][j] = make_int2(warp_adelta(L.m, u), warp_bdelta(L.m, u));
            } else {
                const int v = ty0 + (j - kTileW) - L.by;
                s_row[slot][j - kTileW] = make_int2(warp_x0(L.m, v), warp_y0(L.m, v));
            }
        }
        for (int slot = 0; slot < n_here; ++slot) {   // 256 threads: one LUT entry each
            const DevLayer& L = P.layers[s_active[base + slot]];
            uint32_t v = (uint32_t)tid;
            if (L.opacity < 1.0)
                v = L.pil ? (uint32_t)(tid * L.pil_a) / 255u                        // imgproc.py:206-208
                          : __double2uint_rz(__dmul_rn((double)tid, L.opacity));    // imgproc.py:159
            s_lut[slot][tid] = (float)v;
        }
        __syncthreads();

#pragma unroll 1
        for (int slot = 0; slot < n_here; ++slot) {
            const DevLayer& L = P.layers[s_active[base + slot]];
            const int cls = s_class[base + slot];
            const uint8_t* __restrict__ src = L.src;
            const int ss = L.src_stride;
            const int2 col = s_col[slot][lane];
            const int2* __restrict__ rowtab = &s_row[slot][rbeg];

            // ---- taps of the four rows, all requested before the first is used ----------------------
            uint32_t t[kFRows][4];
            float fxs[kFRows];
            uint32_t wy[kFRows];
            if (cls == kTileInterior) {
#pragma unroll
                for (int k = 0; k < kFRows; ++k) {
                    const int2 r = rowtab[k];
                    const int sumx = r.x + col.x, sumy = r.y + col.y;      // 1/1024 px (cv2: X0 + adelta)
                    const uint8_t* p = src + ((uint32_t)(sumy >> 10) * (uint32_t)ss + (uint32_t)(sumx >> 10) * 4u);
                    t[k][0] = ldg_u32(p);
                    t[k][1] = ldg_u32(p + 4);
                    t[k][2] = ldg_u32(p + ss);
                    t[k][3] = ldg_u32(p + ss + 4);
                    fxs[k] = (float)(sumx & 0x3E0);                        // 32 * fx
                    wy[k] = (uint32_t)((sumy >> 5) & 31) * 255u + 32u;     // bytes (32 - fy, fy, 0, 0)
                }
            } else {
                const int sw = L.src_w, sh = L.src_h;
                const bool col_in = x_ok && (unsigned)(x - L.bx) < (unsigned)L.bw;
#pragma unroll
                for (int k = 0; k < kFRows; ++k) {
                    const int2 r = rowtab[k];
                    const int sumx = r.x + col.x, sumy = r.y + col.y;
                    const int yy = ty0 + rbeg + k;
                    t[k][0] = t[k][1] = t[k][2] = t[k][3] = 0u;            // outside the bbox: fa = lut[0] = 0, a no-op
                    if (col_in && (unsigned)(yy - L.by) < (unsigned)L.bh && yy < P.H) {
                        const Taps tp = fetch_taps(src, sw, sh, ss, sumx >> 5, sumy >> 5);
                        t[k][0] = tp.p00; t[k][1] = tp.p01; t[k][2] = tp.p10; t[k][3] = tp.p11;
                    }
                    fxs[k] = (float)(sumx & 0x3E0);
                    wy[k] = (uint32_t)((sumy >> 5) & 31) * 255u + 32u;
                }
            }

            // ---- cv2's fixed-point bilinear: out = (sum w p + 512) >> 10 -------------------------------
            u64 s[2][4], fa[2];
            // &lut[bits - kMagicBits] for the biased alpha `bits`: one LEA (the shift drops the exponent bits)
            const uint32_t lut_base = (uint32_t)__cvta_generic_to_shared(s_lut[slot]) - (kMagicBits << 2);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                float vl[2][4], vr[2][4];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int k = 2 * p + q;
                    const uint32_t wa = wy[k], wb = wy[k] << 16;
                    const uint32_t l_rg = __byte_perm(t[k][0], t[k][2], 0x5140), l_ba = __byte_perm(t[k][0], t[k][2], 0x7362);
                    const uint32_t r_rg = __byte_perm(t[k][1], t[k][3], 0x5140), r_ba = __byte_perm(t[k][1], t[k][3], 0x7362);
                    vl[q][0] = __uint_as_float(__dp4a(l_rg, wa, kTwo23Bits));   // 2^23 + p00 (32 - fy) + p10 fy
                    vl[q][1] = __uint_as_float(__dp4a(l_rg, wb, kTwo23Bits));
                    vl[q][2] = __uint_as_float(__dp4a(l_ba, wa, kTwo23Bits));
                    vl[q][3] = __uint_as_float(__dp4a(l_ba, wb, kTwo23Bits));
                    vr[q][0] = __uint_as_float(__dp4a(r_rg, wa, kTwo23Bits));
                    vr[q][1] = __uint_as_float(__dp4a(r_rg, wb, kTwo23Bits));
                    vr[q][2] = __uint_as_float(__dp4a(r_ba, wa, kTwo23Bits));
                    vr[q][3] = __uint_as_float(__dp4a(r_ba, wb, kTwo23Bits));
                }
                const u64 fx2 = pk(fxs[2 * p], fxs[2 * p + 1]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // 32 (32 vl + fx (vr - vl) + 512) = 1024 vl + (32 fx)(vr - vl) + 16384  < 2^24
                    const u64 vl2 = pk(vl[0][c], vl[1][c]);
                    const u64 dd = sub2(pk(vr[0][c], vr[1][c]), vl2);
                    const u64 tt = fma2(vl2, bc(1024.0f), bc(16384.0f - 8589934592.0f));
                    s[p][c] = fma2_rm(fma2(fx2, dd, tt), bc(1.0f / 32768.0f), bc(kMagic));
                }
                float a0, a1;
                upk(s[p][3], a0, a1);
                fa[p] = pk(lds_f32(lut_base + (__float_as_uint(a0) << 2)), lds_f32(lut_base + (__float_as_uint(a1) << 2)));
            }

            // ---- blend, in the code specialised for this layer's mode ------------------------------------
            switch (L.pil ? kPilOver : L.mode) {
#define MVB_ROWS(M) case M: blend_rows<M>(d, s, fa, s_softg); break;
                MVB_ROWS(0) MVB_ROWS(1) MVB_ROWS(2) MVB_ROWS(3) MVB_ROWS(4) MVB_ROWS(5) MVB_ROWS(6) MVB_ROWS(7)
                MVB_ROWS(8) MVB_ROWS(9) MVB_ROWS(10) MVB_ROWS(11) MVB_ROWS(12) MVB_ROWS(13) MVB_ROWS(14)
                MVB_ROWS(15) MVB_ROWS(16) MVB_ROWS(17) MVB_ROWS(18)
#undef MVB_ROWS
            default: break;
            }
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
