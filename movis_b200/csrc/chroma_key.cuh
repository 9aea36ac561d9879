// chroma_key.cuh -- ChromaKey (movis/contrib/segmentation.py:58-64) as an integer range test, sm_100a.
//
// The reference converts the image with cv2.cvtColor(COLOR_RGB2HSV), tests cv2.inRange(hsv, lower,
// upper) and writes alpha = 255 outside the range, 0 inside.  cv2's 8-bit conversion (color_hsv.simd.hpp,
// RGB2HSV_b, hrange 180) is integer arithmetic:
//     v = max(r, g, b);  diff = v - min(r, g, b)
//     s = (diff * sdiv[v] + 2^11) >> 12                 sdiv[v]    = rint(255 * 4096 / v)
//     n = v == r ? g - b : v == g ? b - r + 2 diff : r - g + 4 diff
//     h = (n * hdiv[diff] + 2^11) >> 12;  h += 180 if h < 0        hdiv[diff] = rint(180 * 4096 / (6 diff))
// Nothing but the three range tests survives, so the divisions are folded into bounds on the integers
// they are applied to, once per key on the host:
//   * s is non-decreasing in diff for a fixed v: "lo_s <= s <= hi_s and lo_v <= v <= hi_v" is
//     "diff in [d0(v), d1(v)]"                                     -> sv[v]
//   * for a fixed diff, n runs over [-diff, 5 diff] and h is non-decreasing along the CYCLE
//     0, 1, .., 5 diff, -diff, .., -1, (0): it climbs from 0 to 150, the wrap lifts the negative
//     numerators to [150, 180), and the last few of them round to 0 again.  So "lo_h <= h <= hi_h" is an
//     arc of that cycle: "(n - n0(diff)) mod (6 diff + 1) <= span(diff)"      -> hq[diff]
// The per-texel test is then max/min, the hue numerator, two table reads and three packed range tests
// (1 KB + 2 KB of tables; build_chroma_tables checks the arc structure instead of assuming it).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mvb {

// Table entries are "packed range tests": for a range [a, b] of a quantity x (all below 2^15) the word
// E = b * 65536 - a gives  Z = x * 0xFFFF0001 + E = (x - a) + (b - x) * 65536,  whose bits 15 and 31 are
// both clear exactly when a <= x <= b (a negative low half borrows from the high half, but then bit 15 is
// set already): one IMAD and one LOP3 per test instead of unpacking two bounds.  An empty range is
// a = 0x7FFF, b = 0.
struct ChromaTables {
    uint32_t sv[256];   // by v:    diff in [d0, d1]
    uint2 hq[256];      // by diff: n + 256 in [a1, b1] or in [a2, b2] (the arc of the cycle, cut where it wraps)
};

constexpr uint32_t kChromaEmpty = 0u * 65536u - 0x7FFFu;
__host__ __device__ inline uint32_t chroma_range(int a, int b) { return (uint32_t)b * 65536u - (uint32_t)a; }

// Host: false when the in-range set is not of the form above (cannot happen for cv2's tables; the
// caller then keeps the full conversion).
inline bool build_chroma_tables(const uint8_t lo[3], const uint8_t hi[3], ChromaTables& T)
{
    int32_t sdiv[256], hdiv[256];
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; i++) {
        sdiv[i] = (int32_t)__builtin_nearbyint((255 << 12) / (1. * i));
        hdiv[i] = (int32_t)__builtin_nearbyint((180 << 12) / (6. * i));
    }
    for (int v = 0; v < 256; v++) {
        int d0 = -1, d1 = -1;
        bool closed = false;
        if (v >= lo[2] && v <= hi[2]) {
            for (int diff = 0; diff <= v; diff++) {
                int s = (diff * sdiv[v] + (1 << 11)) >> 12;
                s = s < 0 ? 0 : s > 255 ? 255 : s;               // saturate_cast<uchar>
                const bool in = s >= lo[1] && s <= hi[1];
                if (in) {
                    if (closed) return false;
                    if (d0 < 0) d0 = diff;
                    d1 = diff;
                } else if (d0 >= 0) {
                    closed = true;
                }
            }
        }
        T.sv[v] = d0 < 0 ? kChromaEmpty : chroma_range(d0, d1);
    }
    for (int diff = 0; diff < 256; diff++) {
        const int M = 6 * diff + 1;                              // numerators -diff .. 5 diff, cyclically
        bool in[6 * 255 + 1];
        int count = 0;
        for (int j = 0; j < M; j++) {
            const int n = j - diff;
            int h = (n * hdiv[diff] + (1 << 11)) >> 12;          // arithmetic shift, as cv2's int code
            if (h < 0) h += 180;
            h = h < 0 ? 0 : h > 255 ? 255 : h;                   // saturate_cast<uchar>
            in[j] = h >= lo[0] && h <= hi[0];
            count += in[j];
        }
        T.hq[diff] = make_uint2(kChromaEmpty, kChromaEmpty);
        if (count == 0) continue;
        int starts = 0, first = 0;
        for (int j = 0; j < M; j++)
            if (in[j] && !in[(j + M - 1) % M]) { ++starts; first = j; }
        if (count == M) { starts = 1; first = 0; }
        if (starts != 1) return false;
        // the arc [first, first + count) of the cycle of length M, in j = n + diff; as n + 256: j + 256 - diff
        const int off = 256 - diff, last = first + count - 1;
        if (last < M) {
            T.hq[diff].x = chroma_range(first + off, last + off);
        } else {
            T.hq[diff].x = chroma_range(first + off, M - 1 + off);
            T.hq[diff].y = chroma_range(0 + off, last - M + off);
        }
    }
    return true;
}

#ifdef __CUDACC__
// One texel: true when cv2.inRange(cvtColor(rgb -> hsv), lower, upper) holds (the pixel is keyed out).
// Bytes are picked and differenced by IDP.4A (the FMA pipe is idle here, the ALU pipe is the busy one).
__device__ __forceinline__ int dp4a_us(uint32_t a, int32_t b_signed_bytes, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_signed_bytes), "r"(c));
    return d;
}
__device__ __forceinline__ bool chroma_in_range(uint32_t p, const uint32_t* __restrict__ sv, const uint2* __restrict__ hq)
{
    const int r = (int)__dp4a(p, 0x00000001u, 0u), g = (int)__dp4a(p, 0x00000100u, 0u), b = (int)__dp4a(p, 0x00010000u, 0u);
    const int v = max(max(r, g), b), diff = v - min(min(r, g), b);
    // hue numerators + 256: g - b, b - r + 2 diff, r - g + 4 diff (signed byte weights 0x00FF0100 = (0, +1, -1, 0) ...)
    const int n_r = dp4a_us(p, 0x00FF0100, 256), n_g = dp4a_us(p, 0x000100FF, 2 * diff + 256), n_b = dp4a_us(p, 0x0000FF01, 4 * diff + 256);
    int n;
    asm("{\n.reg .pred pr, pg;\nsetp.eq.s32 pr, %1, %2;\nsetp.eq.s32 pg, %1, %3;\n"
        "selp.s32 %0, %5, %6, pg;\nselp.s32 %0, %4, %0, pr;\n}"
        : "=r"(n) : "r"(v), "r"(r), "r"(g), "r"(n_r), "r"(n_g), "r"(n_b));
    const uint32_t e = sv[v];
    const uint2 q = hq[diff];
    const uint32_t zs = (uint32_t)diff * 0xFFFF0001u + e;
    const uint32_t z1 = (uint32_t)n * 0xFFFF0001u + q.x, z2 = (uint32_t)n * 0xFFFF0001u + q.y;
    return !(zs & 0x80008000u) && (!(z1 & 0x80008000u) || !(z2 & 0x80008000u));
}
#endif

} // namespace mvb
