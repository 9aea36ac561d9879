/*
 * mvo.c -- CPU oracle for the movis per-frame compositing path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.  The product path
 * (movis_b200/) never calls into it and has no CPU fallback.
 *
 * It restates, in plain C, the integer arithmetic that the reference (rezoo/movis 0.7.1)
 * delegates to third-party wheels whose sources are not under /root/reference:
 *   opencv-python (requirements.txt:7, >=4.8.0; checked against 4.13.0)  warpAffine, GaussianBlur,
 *                                                  cvtColor(RGB2HSV/HSV2RGB/RGB2GRAY), inRange
 *   Pillow        (requirements.txt:3, >=8.2.0; checked against 12.2.0)  Image.alpha_composite
 *   numpy         (requirements.txt:1; checked against 2.3.5)            ufunc chains of imgproc.py
 * Parity is pinned by tests/golden/ (outputs of the reference itself, generated in the build
 * container by oracle/gen_golden.py) -- see tests/test_oracle_golden.py.
 *
 * Every function cites the reference call site (path:line relative to /root/reference) whose
 * result it reproduces.  All images are uint8 RGBA, HWC, with an explicit row stride in bytes.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off; no -ffast-math: the fp64 set-up
 * arithmetic must round exactly like OpenCV's / numpy's).
 */
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MVO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* helpers                                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* cv::saturate_cast<int>(double) == cvRound == round-half-even (SSE2 cvtsd2si); values outside
 * the int range give INT_MIN ("integer indefinite"). */
static inline int cv_round(double v)
{
    if (!(v > -2147483648.5 && v < 2147483647.5)) return INT_MIN;
    return (int)lrint(v);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int mini(int a, int b) { return a < b ? a : b; }
static inline int maxi(int a, int b) { return a > b ? a : b; }

MVO_API int mvo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

MVO_API void mvo_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* cv2.warpAffine(fg, M, dsize, INTER_LINEAR, BORDER_CONSTANT)  -- layer/composition.py:811-813 */
/* ------------------------------------------------------------------------------------------ */

/* The inversion cv::warpAffine performs on the forward 2x3 matrix (no WARP_INVERSE_MAP). */
MVO_API void mvo_invert_affine(const double M[6], double m[6])
{
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    m[0] = A11;
    m[1] = M[1] * (-D);
    m[3] = M[3] * (-D);
    m[4] = A22;
    double b1 = -m[0] * M[2] - m[1] * M[5];
    double b2 = -m[3] * M[2] - m[4] * M[5];
    m[2] = b1;
    m[5] = b2;
}

static inline int sat_short(int v) { return clampi(v, -32768, 32767); }

/* One bilinear sample at fixed-point coordinates X, Y (5 fractional bits); taps outside the
 * source are 0 (BORDER_CONSTANT, borderValue 0). */
static inline void sample_bilinear(const uint8_t* src, int sw, int sh, ptrdiff_t ss, int X, int Y,
                                   uint8_t out[4])
{
    int sx = sat_short(X >> 5), sy = sat_short(Y >> 5);
    int fx = X & 31, fy = Y & 31;
    int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
    static const uint8_t zero[4] = {0, 0, 0, 0};
    const uint8_t *p00 = zero, *p01 = zero, *p10 = zero, *p11 = zero;
    int x0ok = (unsigned)sx < (unsigned)sw, x1ok = (unsigned)(sx + 1) < (unsigned)sw;
    int y0ok = (unsigned)sy < (unsigned)sh, y1ok = (unsigned)(sy + 1) < (unsigned)sh;
    if (y0ok) {
        const uint8_t* row = src + (ptrdiff_t)sy * ss;
        if (x0ok) p00 = row + 4 * sx;
        if (x1ok) p01 = row + 4 * (sx + 1);
    }
    if (y1ok) {
        const uint8_t* row = src + (ptrdiff_t)(sy + 1) * ss;
        if (x0ok) p10 = row + 4 * sx;
        if (x1ok) p11 = row + 4 * (sx + 1);
    }
    for (int c = 0; c < 4; c++)
        out[c] = (uint8_t)((p00[c] * w00 + p01[c] * w01 + p10[c] * w10 + p11[c] * w11 + 512) >> 10);
}

/* dst (dw x dh) = warpAffine(src, Mfwd).  Mfwd maps source pixels to destination pixels. */
MVO_API void mvo_warp_affine_rgba8(const uint8_t* src, int sw, int sh, ptrdiff_t ss, uint8_t* dst,
                                   int dw, int dh, ptrdiff_t ds, const double Mfwd[6])
{
    double m[6];
    mvo_invert_affine(Mfwd, m);
    int* adelta = (int*)malloc(sizeof(int) * 2 * (size_t)(dw > 0 ? dw : 1));
    int* bdelta = adelta + dw;
    for (int x = 0; x < dw; x++) {
        adelta[x] = cv_round(m[0] * x * 1024);
        bdelta[x] = cv_round(m[3] * x * 1024);
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; y++) {
        int X0 = cv_round((m[1] * y + m[2]) * 1024) + 16;
        int Y0 = cv_round((m[4] * y + m[5]) * 1024) + 16;
        uint8_t* drow = dst + (ptrdiff_t)y * ds;
        for (int x = 0; x < dw; x++) {
            int X = (int)((unsigned)X0 + (unsigned)adelta[x]) >> 5;
            int Y = (int)((unsigned)Y0 + (unsigned)bdelta[x]) >> 5;
            sample_bilinear(src, sw, sh, ss, X, Y, drow + 4 * x);
        }
    }
    free(adelta);
}

/* ------------------------------------------------------------------------------------------ */
/* blend targets -- imgproc.py:11-133, evaluated under numpy-2 (NEP 50) promotion on uint32     */
/* ------------------------------------------------------------------------------------------ */

enum {
    MVO_NORMAL = 0, MVO_MULTIPLY, MVO_SCREEN, MVO_OVERLAY, MVO_DARKEN, MVO_LIGHTEN,
    MVO_COLOR_DODGE, MVO_COLOR_BURN, MVO_LINEAR_DODGE, MVO_LINEAR_BURN, MVO_HARD_LIGHT,
    MVO_SOFT_LIGHT, MVO_VIVID_LIGHT, MVO_LINEAR_LIGHT, MVO_PIN_LIGHT, MVO_DIFFERENCE,
    MVO_EXCLUSION, MVO_SUBTRACT, MVO_N_MODES
};

static inline uint32_t overlay_u32(uint32_t b, uint32_t f) /* imgproc.py:19-20 */
{
    return b < 128 ? 2 * b * f / 255 : 255 - 2 * (255 - b) * (255 - f) / 255;
}

static inline int64_t floordiv_i64(int64_t a, int64_t b)
{
    int64_t q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

static uint32_t soft_light_u8(uint32_t b, uint32_t f) /* imgproc.py:58-72 */
{
    if (f < 128) {
        uint32_t t = (uint32_t)(255u - 2u * f) * b; /* all mod 2^32 like numpy uint32 */
        t = t * (255u - b);
        return (uint32_t)(b - t / 65025u) & 0xFFu;
    }
    int64_t g;
    if (b < 64) {
        uint32_t t = 16u * b - 780300u * b + 260100u; /* wraps: part of the contract */
        t = t * b;
        g = (int64_t)(t / 65025u);
    } else {
        g = (int64_t)sqrt((double)(255u * b)); /* np.sqrt(uint32)->float64, astype(int32) truncs */
    }
    int64_t v = (int64_t)b + floordiv_i64((int64_t)(2u * f - 255u) * (g - (int64_t)b), 255);
    return (uint32_t)((uint64_t)v & 0xFFu);
}

static uint32_t blend_target(int mode, uint32_t b, uint32_t f)
{
    switch (mode) {
    case MVO_NORMAL: return f;
    case MVO_MULTIPLY: return b * f / 255;
    case MVO_SCREEN: return 255 - (255 - b) * (255 - f) / 255;
    case MVO_OVERLAY: return overlay_u32(b, f);
    case MVO_DARKEN: return b < f ? b : f;
    case MVO_LIGHTEN: return b > f ? b : f;
    case MVO_COLOR_DODGE: { /* imgproc.py:36-38: float64 true-divide, clip, trunc */
        double x = 255.0 * (double)b / (double)(255 - (int)f + 1);
        return (uint32_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
    }
    case MVO_COLOR_BURN: {
        double x = 255.0 * (double)(255 - (int)b) / (double)((int)f + 1);
        return 255 - (uint32_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
    }
    case MVO_LINEAR_DODGE: return b + f > 255 ? 255 : b + f;
    case MVO_LINEAR_BURN: return b + f < 255 ? 0 : b + f - 255;
    case MVO_HARD_LIGHT: return overlay_u32(f, b);
    case MVO_SOFT_LIGHT: return soft_light_u8(b, f);
    case MVO_VIVID_LIGHT: { /* imgproc.py:75-80: the dodge arm divides by a negative number */
        int f2 = 2 * (int)f;
        if (f > 127) {
            double x = 255.0 * (double)b / (double)(255 - 2 * (f2 - 127) + 1);
            return (uint32_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); /* always 0 */
        }
        double x = 255.0 * (double)(255 - (int)b) / (double)(f2 + 1);
        return 255 - (uint32_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
    }
    case MVO_LINEAR_LIGHT: {
        int v = (int)b + 2 * (int)f - 255;
        return (uint32_t)(f > 127 ? mini(255, v) : maxi(0, v));
    }
    case MVO_PIN_LIGHT: {
        int f2 = 2 * (int)f;
        return (uint32_t)(f > 127 ? maxi((int)b, f2 - 255) : mini((int)b, f2));
    }
    case MVO_DIFFERENCE: return b > f ? b - f : f - b;
    case MVO_EXCLUSION: return (uint32_t)clampi((int)b + (int)f - (int)(2 * b * f / 255), 0, 255);
    case MVO_SUBTRACT: return b > f ? b - f : 0;
    default: return f;
    }
}

/* 256x256 table T[b][f] for one mode (golden vectors: SURVEY.md 8c). */
MVO_API void mvo_blend_lut(int mode, uint8_t* lut)
{
    for (uint32_t b = 0; b < 256; b++)
        for (uint32_t f = 0; f < 256; f++) lut[b * 256 + f] = (uint8_t)blend_target(mode, b, f);
}

/* ------------------------------------------------------------------------------------------ */
/* "over" operators                                                                           */
/* ------------------------------------------------------------------------------------------ */

/* PIL ImagingAlphaComposite per pixel -- reached from imgproc.py:210-213 */
static inline void pil_over(const uint8_t s[4], uint32_t sa, uint8_t d[4])
{
    if (sa == 0) return;
    uint32_t blend = d[3] * (255 - sa);
    uint32_t outa255 = sa * 255 + blend;
    uint32_t coef1 = sa * 255 * 255 * 128 / outa255;
    uint32_t coef2 = 255 * 128 - coef1;
    for (int c = 0; c < 3; c++) {
        uint32_t t = s[c] * coef1 + d[c] * coef2 + (0x80 << 7);
        d[c] = (uint8_t)((((t >> 8) + t) >> 8) >> 7);
    }
    uint32_t u = outa255 + 0x80;
    d[3] = (uint8_t)(((u >> 8) + u) >> 8);
}

/* numpy _overlay per pixel, matte NONE / ALPHA -- imgproc.py:156-170 */
static inline void numpy_over(int mode, const uint8_t s[4], uint32_t fa, uint8_t d[4], int keep_alpha)
{
    uint32_t ba = d[3];
    uint32_t c1 = 255 * fa, c2 = ba * (255 - fa);
    uint32_t oa = c1 + c2;
    uint32_t den = oa + (oa == 0);
    for (int c = 0; c < 3; c++) {
        uint32_t T = blend_target(mode, d[c], s[c]);
        d[c] = (uint8_t)((c1 * T + c2 * d[c]) / den);
    }
    if (!keep_alpha) d[3] = (uint8_t)(oa / 255);
}

/* cv2.cvtColor(RGB2GRAY) -- imgproc.py:148 */
static inline uint32_t rgb2gray(const uint8_t p[4])
{
    return (uint32_t)(p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + 16384) >> 15;
}

/* alpha LUTs: how `opacity` rescales fg alpha on either path. */
static void alpha_lut_pil(double opacity, uint8_t lut[256]) /* imgproc.py:204-209 */
{
    if (!(opacity < 1.0)) {
        for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i;
        return;
    }
    int a = (int)lrint(opacity * 255); /* int(np.round(.)) : half-even */
    for (int i = 0; i < 256; i++) lut[i] = (uint8_t)(i * a / 255);
}

static void alpha_lut_numpy(double opacity, uint8_t lut[256]) /* imgproc.py:149-150,159 */
{
    if (!(opacity < 1.0)) {
        for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i;
        return;
    }
    for (int i = 0; i < 256; i++) lut[i] = (uint8_t)(uint32_t)((double)i * opacity);
}

/*
 * movis.imgproc.alpha_composite -- imgproc.py:216-273.  bg is modified in place.
 *   use_pil != 0 : the NORMAL + MatteMode.NONE enum fast path (imgproc.py:265-268)
 *   matte: 0 NONE, 1 ALPHA, 2 LUMINANCE
 */
MVO_API void mvo_alpha_composite(uint8_t* bg, int bw, int bh, ptrdiff_t bs, const uint8_t* fg, int fw,
                                 int fh, ptrdiff_t fs, int px, int py, double opacity, int mode,
                                 int matte, int use_pil)
{
    int x1 = maxi(0, px), y1 = maxi(0, py);
    int x2 = -mini(0, px), y2 = -mini(0, py);
    int w = mini(px + fw, bw) - x1, h = mini(py + fh, bh) - y1;
    if (w <= 0 || h <= 0) return;
    uint8_t lut[256];
    if (use_pil) alpha_lut_pil(opacity, lut); else alpha_lut_numpy(opacity, lut);

    if (!use_pil && matte == 2) { /* imgproc.py:147-154 */
#pragma omp parallel for schedule(static)
        for (int y = 0; y < bh; y++) {
            uint8_t* brow = bg + (ptrdiff_t)y * bs;
            int in_y = y >= y1 && y < y1 + h;
            for (int x = 0; x < bw; x++) {
                uint8_t* d = brow + 4 * x;
                if (in_y && x >= x1 && x < x1 + w) {
                    const uint8_t* s = fg + (ptrdiff_t)(y - y1 + y2) * fs + 4 * (x - x1 + x2);
                    uint32_t mask = rgb2gray(d);
                    uint32_t fa = lut[s[3]];
                    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
                    d[3] = (uint8_t)(mask * fa / 255);
                } else {
                    d[3] = 0;
                }
            }
        }
        return;
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        uint8_t* brow = bg + (ptrdiff_t)(y1 + y) * bs + 4 * x1;
        const uint8_t* frow = fg + (ptrdiff_t)(y2 + y) * fs + 4 * x2;
        for (int x = 0; x < w; x++) {
            const uint8_t* s = frow + 4 * x;
            uint8_t* d = brow + 4 * x;
            uint32_t fa = lut[s[3]];
            if (use_pil) pil_over(s, fa, d);
            else numpy_over(mode, s, fa, d, matte == 1);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* One Composition.__call__ frame walk -- layer/composition.py:363-368 + :791-820             */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const uint8_t* src;   /* post-effect layer image */
    int32_t src_w, src_h;
    int64_t src_stride;
    double M[6];          /* affine_matrix_fixed (composition.py:906): layer px -> bbox px */
    int32_t bbox_x, bbox_y, bbox_w, bbox_h; /* (offset_x, offset_y), (W, H) composition.py:896-900 */
    int32_t blend_mode;   /* BlendingMode value 0..17 */
    int32_t use_pil;      /* 1 when mode is the NORMAL enum (composition always uses matte NONE) */
    double opacity;
} mvo_layer;

/* frame := bg_color; for each layer: warp into a bbox-sized scratch image exactly like the
 * reference does, then alpha_composite it at (bbox_x, bbox_y). */
MVO_API int mvo_composite_stack(uint8_t* frame, int W, int H, ptrdiff_t fs, const uint8_t bg_color[4],
                                int n_layers, const mvo_layer* layers)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        uint8_t* row = frame + (ptrdiff_t)y * fs;
        for (int x = 0; x < W; x++) memcpy(row + 4 * x, bg_color, 4);
    }
    for (int i = 0; i < n_layers; i++) {
        const mvo_layer* L = &layers[i];
        if (L->bbox_w <= 0 || L->bbox_h <= 0) continue;
        size_t stride = (size_t)L->bbox_w * 4;
        uint8_t* tmp = (uint8_t*)malloc(stride * (size_t)L->bbox_h);
        if (!tmp) return -1;
        mvo_warp_affine_rgba8(L->src, L->src_w, L->src_h, (ptrdiff_t)L->src_stride, tmp, L->bbox_w,
                              L->bbox_h, (ptrdiff_t)stride, L->M);
        mvo_alpha_composite(frame, W, H, fs, tmp, L->bbox_w, L->bbox_h, (ptrdiff_t)stride, L->bbox_x,
                            L->bbox_y, L->opacity, L->blend_mode, 0, L->use_pil);
        free(tmp);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* cv2.GaussianBlur on uint8 -- effect/blur.py:41-43,77-79; effect/style.py:60-62              */
/* ------------------------------------------------------------------------------------------ */

/* ksize rule shared by the three effects: k = 2*ceil(max(1,(4r-1)/2))+1  (blur.py:37) */
MVO_API int mvo_blur_ksize(double radius)
{
    double v = (4 * radius - 1) / 2;
    if (v < 1) v = 1;
    return 2 * (int)ceil(v) + 1;
}

/* OpenCV's bit-exact Q0.8 Gaussian kernel (getGaussianKernelBitExact +
 * getGaussianKernelFixedPoint_ED, fractionBits = 8), restated with IEEE double + libm exp.
 * k must be odd.  w receives k integers that sum to 256. */
MVO_API void mvo_gaussian_weights_q8(double sigma, int k, int32_t* w)
{
    int n2 = (k - 1) / 2;
    double* val = (double*)malloc(sizeof(double) * (size_t)(n2 + 1));
    double scale2X = -0.125 / (sigma * sigma);
    double sum = 0;
    for (int i = 0, x = 1 - k; i < n2; i++, x += 2) {
        double t = exp((double)(x * x) * scale2X);
        val[i] = t;
        sum += t;
    }
    sum *= 2;
    sum += 1;
    double mul1 = 1.0 / sum;
    double err = 0;
    int64_t isum = 0;
    for (int i = 0; i < n2; i++) {
        double adj = (val[i] * mul1) * 256.0 + err;
        int64_t v0 = (int64_t)lrint(adj);
        err = adj - (double)v0;
        w[i] = w[k - 1 - i] = (int32_t)v0;
        isum += v0;
    }
    w[n2] = (int32_t)(256 - 2 * isum);
    free(val);
}

static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* dst = GaussianBlur(src, (k,k), sigma) with BORDER_REFLECT_101; C = 1 or 4 interleaved channels.
 * Horizontal pass to Q8.8 (<= 16 bit), vertical pass to Q16.16, out = (V + 32768) >> 16. */
MVO_API void mvo_gaussian_blur_u8(const uint8_t* src, int w, int h, ptrdiff_t ss, int C, int k,
                                  const int32_t* wt, uint8_t* dst, ptrdiff_t ds)
{
    int r = k / 2;
    uint16_t* tmp = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h * (size_t)C);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        const uint8_t* srow = src + (ptrdiff_t)y * ss;
        uint16_t* trow = tmp + (size_t)y * (size_t)w * (size_t)C;
        for (int x = 0; x < w; x++)
            for (int c = 0; c < C; c++) {
                uint32_t acc = 0;
                for (int j = 0; j < k; j++) acc += (uint32_t)wt[j] * srow[reflect101(x + j - r, w) * C + c];
                trow[x * C + c] = (uint16_t)acc;
            }
    }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++) {
        uint8_t* drow = dst + (ptrdiff_t)y * ds;
        for (int x = 0; x < w * C; x++) {
            uint32_t acc = 0;
            for (int j = 0; j < k; j++)
                acc += (uint32_t)wt[j] * tmp[(size_t)reflect101(y + j - r, h) * (size_t)w * (size_t)C + (size_t)x];
            drow[x] = (uint8_t)((acc + 32768) >> 16);
        }
    }
    free(tmp);
}

/* np.pad of blur.py:39-40 / :74-75: k px per side, RGB mode='edge', alpha zeros. */
MVO_API void mvo_pad_edge_rgb_zero_alpha(const uint8_t* src, int w, int h, ptrdiff_t ss, int k,
                                         uint8_t* dst, ptrdiff_t ds)
{
    int W = w + 2 * k, H = h + 2 * k;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++) {
        int sy = clampi(y - k, 0, h - 1);
        int in_y = y >= k && y < k + h;
        for (int x = 0; x < W; x++) {
            int sx = clampi(x - k, 0, w - 1);
            const uint8_t* s = src + (ptrdiff_t)sy * ss + 4 * sx;
            uint8_t* d = dst + (ptrdiff_t)y * ds + 4 * x;
            d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
            d[3] = (in_y && x >= k && x < k + w) ? s[3] : 0;
        }
    }
}

/* GaussianBlur.__call__ -- effect/blur.py:29-43.  dst is (h+2k) x (w+2k), k = mvo_blur_ksize. */
MVO_API void mvo_effect_gaussian_blur(const uint8_t* src, int w, int h, ptrdiff_t ss, double radius,
                                      uint8_t* dst, ptrdiff_t ds)
{
    int k = mvo_blur_ksize(radius);
    int W = w + 2 * k, H = h + 2 * k;
    uint8_t* pad = (uint8_t*)malloc((size_t)W * (size_t)H * 4);
    int32_t* wt = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
    mvo_pad_edge_rgb_zero_alpha(src, w, h, ss, k, pad, (ptrdiff_t)W * 4);
    mvo_gaussian_weights_q8(radius, k, wt);
    mvo_gaussian_blur_u8(pad, W, H, (ptrdiff_t)W * 4, 4, k, wt, dst, ds);
    free(wt);
    free(pad);
}

/* Glow.__call__ -- effect/blur.py:66-85.  strength_lut[v] = trunc(clip(strength*float32(v),0,255))
 * is evaluated by the caller with numpy so the float32 rounding is numpy's own. */
MVO_API void mvo_effect_glow(const uint8_t* src, int w, int h, ptrdiff_t ss, double radius,
                             const uint8_t strength_lut[256], uint8_t* dst, ptrdiff_t ds)
{
    int k = mvo_blur_ksize(radius);
    int W = w + 2 * k, H = h + 2 * k;
    uint8_t* blurred = (uint8_t*)malloc((size_t)W * (size_t)H * 4);
    int32_t* wt = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
    mvo_pad_edge_rgb_zero_alpha(src, w, h, ss, k, dst, ds); /* dst starts as pad_image */
    mvo_gaussian_weights_q8(radius, k, wt);
    mvo_gaussian_blur_u8(dst, W, H, ds, 4, k, wt, blurred, (ptrdiff_t)W * 4);
    for (size_t i = 0; i < (size_t)W * (size_t)H; i++) {
        blurred[4 * i + 0] = strength_lut[blurred[4 * i + 0]];
        blurred[4 * i + 1] = strength_lut[blurred[4 * i + 1]];
        blurred[4 * i + 2] = strength_lut[blurred[4 * i + 2]];
    }
    mvo_alpha_composite(dst, W, H, ds, blurred, W, H, (ptrdiff_t)W * 4, 0, 0, 1.0, MVO_LINEAR_DODGE, 0, 0);
    free(wt);
    free(blurred);
}

/* DropShadow.__call__ -- effect/style.py:49-84.
 *   opacity_lut[v] = uint8(opacity * v) (float64, numpy-evaluated by the caller)
 *   color: round(color(t)) as 3 bytes; (vx, vy) = round(offset * (cos, sin)) ints.
 * dst is (h + 2k + 2|vy|) x (w + 2k + 2|vx|) with k = 0 when radius == 0. */
MVO_API void mvo_effect_drop_shadow(const uint8_t* src, int w, int h, ptrdiff_t ss, double radius,
                                    const uint8_t opacity_lut[256], const uint8_t color[3], int vx,
                                    int vy, uint8_t* dst, ptrdiff_t ds)
{
    int k = radius > 0 ? mvo_blur_ksize(radius) : 0;
    int Ws = w + 2 * k, Hs = h + 2 * k;
    uint8_t* a0 = (uint8_t*)calloc((size_t)Ws * (size_t)Hs, 1);
    uint8_t* a1 = (uint8_t*)malloc((size_t)Ws * (size_t)Hs);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) a0[(size_t)(y + k) * Ws + (x + k)] = src[(ptrdiff_t)y * ss + 4 * x + 3];
    if (radius > 0) {
        int32_t* wt = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
        mvo_gaussian_weights_q8(radius, k, wt);
        mvo_gaussian_blur_u8(a0, Ws, Hs, Ws, 1, k, wt, a1, Ws);
        free(wt);
    } else {
        for (size_t i = 0; i < (size_t)Ws * Hs; i++) a1[i] = opacity_lut[a0[i]]; /* style.py:64 */
    }
    int dx = abs(vx), dy = abs(vy);
    int Wn = Ws + 2 * dx, Hn = Hs + 2 * dy;
    for (int y = 0; y < Hn; y++) memset(dst + (ptrdiff_t)y * ds, 0, (size_t)Wn * 4);
    int sx0 = dx + vx, sy0 = dy + vy;
    for (int y = 0; y < Hs; y++)
        for (int x = 0; x < Ws; x++) {
            uint8_t* d = dst + (ptrdiff_t)(sy0 + y) * ds + 4 * (sx0 + x);
            d[0] = color[0]; d[1] = color[1]; d[2] = color[2];
            d[3] = opacity_lut[a1[(size_t)y * Ws + x]]; /* style.py:68 */
        }
    int px = (Wn - w) / 2, py = (Hn - h) / 2;
    mvo_alpha_composite(dst, Wn, Hn, ds, src, w, h, ss, px, py, 1.0, MVO_NORMAL, 0, 1);
    free(a0);
    free(a1);
}

/* ------------------------------------------------------------------------------------------ */
/* colour conversions -- effect/color.py:59,67; contrib/segmentation.py:47,59-61               */
/* ------------------------------------------------------------------------------------------ */

static int g_sdiv[256], g_hdiv[256], g_hsv_ready = 0;

static void hsv_tables(void)
{
    if (g_hsv_ready) return;
    g_sdiv[0] = g_hdiv[0] = 0;
    for (int i = 1; i < 256; i++) {
        g_sdiv[i] = cv_round((255 << 12) / (1. * i));
        g_hdiv[i] = cv_round((180 << 12) / (6. * i));
    }
    g_hsv_ready = 1;
}

MVO_API void mvo_hsv_tables(int32_t sdiv[256], int32_t hdiv[256])
{
    hsv_tables();
    for (int i = 0; i < 256; i++) { sdiv[i] = g_sdiv[i]; hdiv[i] = g_hdiv[i]; }
}

/* cv2.COLOR_RGB2HSV on uint8 (H in [0,180)). */
static inline void rgb2hsv_px(int r, int g, int b, uint8_t hsv[3])
{
    int v = maxi(maxi(r, g), b), vmin = mini(mini(r, g), b);
    int diff = v - vmin;
    int s = (diff * g_sdiv[v] + (1 << 11)) >> 12;
    int h;
    if (v == r) h = g - b;
    else if (v == g) h = b - r + 2 * diff;
    else h = r - g + 4 * diff;
    h = (h * g_hdiv[diff] + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    hsv[0] = (uint8_t)clampi(h, 0, 255);
    hsv[1] = (uint8_t)s;
    hsv[2] = (uint8_t)v;
}

/* n pixels, `cn` bytes per input pixel (3 or 4), 3 bytes per output pixel. */
MVO_API void mvo_rgb2hsv(const uint8_t* rgb, size_t n, int cn, uint8_t* hsv)
{
    hsv_tables();
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; i++)
        rgb2hsv_px(rgb[i * cn], rgb[i * cn + 1], rgb[i * cn + 2], hsv + 3 * i);
}

/* cv2.COLOR_HSV2RGB on uint8.  OpenCV converts through float32; the exact operation sequence was
 * identified by exhaustive search against cv2 4.13.0 (AVX2 dispatch) over all 180*256*256 inputs
 * (0 mismatches): s,v scaled by the float 1/255, h by the float 6/180, tab2/tab3 use one fused
 * multiply-add each.  The vector body of each row (the first 32*floor(w/32) pixels) TRUNCATES
 * after scaling by 255, whereas the scalar tail of the row (the last w%32 pixels) ROUNDS
 * half-even (saturate_cast) -- a position-dependent quirk of cv2 that is part of the contract. */
#define MVO_CV_SIMD_BLOCK 32

static inline void hsv2rgb_px(int hi, int si, int vi, int round_result, uint8_t rgb[3])
{
    static const int sector_data[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
    float s = (float)si * (1.0f / 255.0f), v = (float)vi * (1.0f / 255.0f);
    float b, g, r;
    if (si == 0) {
        b = g = r = v;
    } else {
        float h = (float)hi * (6.0f / 180.0f);
        int sector = (int)floorf(h);
        h -= (float)sector;
        sector %= 6;
        if (sector < 0) sector += 6;
        float tab[4];
        tab[0] = v;
        tab[1] = v * (1.0f - s);
        tab[2] = v * fmaf(-s, h, 1.0f);
        tab[3] = v * fmaf(-s, 1.0f - h, 1.0f);
        b = tab[sector_data[sector][0]];
        g = tab[sector_data[sector][1]];
        r = tab[sector_data[sector][2]];
    }
    r *= 255.0f; g *= 255.0f; b *= 255.0f;
    int ri, gi, bi;
    if (round_result) { ri = (int)lrintf(r); gi = (int)lrintf(g); bi = (int)lrintf(b); }
    else { ri = (int)r; gi = (int)g; bi = (int)b; }
    rgb[0] = (uint8_t)clampi(ri, 0, 255);
    rgb[1] = (uint8_t)clampi(gi, 0, 255);
    rgb[2] = (uint8_t)clampi(bi, 0, 255);
}

/* hsv: h rows of w pixels, 3 bytes each, densely packed (what np.stack gives at color.py:67). */
MVO_API void mvo_hsv2rgb(const uint8_t* hsv, int w, int h, uint8_t* rgb)
{
    int body = w - w % MVO_CV_SIMD_BLOCK;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            size_t i = ((size_t)y * (size_t)w + (size_t)x) * 3;
            hsv2rgb_px(hsv[i], hsv[i + 1], hsv[i + 2], x >= body, rgb + i);
        }
}

/* ChromaKey.__call__ -- contrib/segmentation.py:58-64: alpha := inRange(hsv, lo, hi) ? 0 : 255 */
MVO_API void mvo_effect_chroma_key(const uint8_t* src, int w, int h, ptrdiff_t ss, const uint8_t lo[3],
                                   const uint8_t hi[3], uint8_t* dst, ptrdiff_t ds)
{
    hsv_tables();
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* s = src + (ptrdiff_t)y * ss + 4 * x;
            uint8_t* d = dst + (ptrdiff_t)y * ds + 4 * x;
            uint8_t hsv[3];
            rgb2hsv_px(s[0], s[1], s[2], hsv);
            int in = hsv[0] >= lo[0] && hsv[0] <= hi[0] && hsv[1] >= lo[1] && hsv[1] <= hi[1] &&
                     hsv[2] >= lo[2] && hsv[2] <= hi[2];
            d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
            d[3] = in ? 0 : 255;
        }
}

/* FillColor.__call__ -- effect/color.py:25-31 */
MVO_API void mvo_effect_fill_color(const uint8_t* src, int w, int h, ptrdiff_t ss, const uint8_t color[3],
                                   uint8_t* dst, ptrdiff_t ds)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint8_t* d = dst + (ptrdiff_t)y * ds + 4 * x;
            d[0] = color[0]; d[1] = color[1]; d[2] = color[2];
            d[3] = src[(ptrdiff_t)y * ss + 4 * x + 3];
        }
}

/* HSLShift.__call__ -- effect/color.py:56-69.  The three float32 remaps (color.py:64-66) are
 * 256-entry tables evaluated by the caller with numpy (h table only uses entries 0..179). */
MVO_API void mvo_effect_hsl_shift(const uint8_t* src, int w, int h, ptrdiff_t ss, const uint8_t hlut[256],
                                  const uint8_t slut[256], const uint8_t vlut[256], uint8_t* dst,
                                  ptrdiff_t ds)
{
    hsv_tables();
    int body = w - w % MVO_CV_SIMD_BLOCK;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* s = src + (ptrdiff_t)y * ss + 4 * x;
            uint8_t* d = dst + (ptrdiff_t)y * ds + 4 * x;
            uint8_t hsv[3];
            rgb2hsv_px(s[0], s[1], s[2], hsv);
            hsv2rgb_px(hlut[hsv[0]], slut[hsv[1]], vlut[hsv[2]], x >= body, d);
            d[3] = s[3];
        }
}

/* Stripe.__call__ -- layer/texture.py:159-191.  float64 throughout, one rounding per numpy operation
 * (this file is compiled with -ffp-contract=off): p = (v0 (y - h/2) + v1 (x - w/2)) / total_width +
 * phase; p -= floor(p); t = clip((p - edge0) / (edge1 - edge0), 0, 1); s = t t (3 - 2 t);
 * colour = s c1 + (1 - s) c2, truncated to uint8.  solid = 1 / 2: the early returns of :170-175. */
MVO_API void mvo_stripe(uint8_t* dst, int w, int h, ptrdiff_t ds, const double v[2], double total_width,
                        double phase, double edge0, double edge1, const double c1[4], const double c2[4],
                        int solid)
{
    const double cy = (double)h / 2, cx = (double)w / 2;
    const double span = edge1 - edge0;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            uint8_t* d = dst + (ptrdiff_t)y * ds + 4 * x;
            if (solid) {
                for (int k = 0; k < 4; k++) d[k] = (uint8_t)(solid == 1 ? c1[k] : c2[k]);
                continue;
            }
            const double a = v[0] * ((double)y - cy);
            const double b = v[1] * ((double)x - cx);
            double p = (a + b) / total_width + phase;
            p = p - floor(p);
            double t = (p - edge0) / span;
            t = t < 0.0 ? 0.0 : t;
            t = t > 1.0 ? 1.0 : t;
            const double tt = t * t;
            const double sm = tt * (3.0 - 2.0 * t);
            const double om = 1.0 - sm;
            for (int k = 0; k < 4; k++) {
                const double u = sm * c1[k];
                const double q = om * c2[k];
                d[k] = (uint8_t)(u + q);
            }
        }
}
