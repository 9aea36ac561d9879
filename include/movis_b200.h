/*
 * movis_b200.h -- C ABI of the B200-native (sm_100a) frame compositor for movis.
 *
 * The reference (rezoo/movis 0.7.1) has no FFI of its own: its per-frame render path
 * (Composition.__call__, movis/layer/composition.py:345-370) calls three third-party native
 * wheels from Python.  Each entry point below replaces one of those call sites; the citation
 * next to it (path:line, relative to the reference tree) names the interface it stands in for.
 * INTEGRATION.md shows the ctypes stub a movis maintainer would add at each site.
 *
 * Conventions
 *   - Plain C types only.  Every image is uint8 RGBA, HWC, non-premultiplied, in DEVICE memory,
 *     described by (pointer, width, height, row stride in bytes).  Row strides must be multiples
 *     of 4 and pointers 4-byte aligned.  Source images must be smaller than 32768 px per side
 *     (cv2's own limit for warpAffine).  Layer sources whose pointer AND stride are multiples of 16
 *     are staged through TMA by mvb_composite_stack; others are gathered through L1 (same results).
 *   - No ownership transfer and no hidden allocation: all buffers are caller-allocated.
 *   - Every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising.  Small host-side parameter blocks (layer descriptors, LUTs, weights) are
 *     consumed before the call returns.
 *   - Return value: MVB_OK (0) or a negative error code; mvb_last_error() gives a thread-local
 *     message.  There is no CPU fallback: without a usable sm_100 device every compute call
 *     fails with MVB_ERR_NO_DEVICE / MVB_ERR_CUDA.
 */
#ifndef MOVIS_B200_H
#define MOVIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVB_ABI_VERSION 2

#if defined(__GNUC__)
#define MVB_API __attribute__((visibility("default")))
#else
#define MVB_API
#endif

typedef void* mvb_stream_t; /* cudaStream_t */

enum {
    MVB_OK = 0,
    MVB_ERR_INVALID_ARG = -1,
    MVB_ERR_CUDA = -2,
    MVB_ERR_NO_DEVICE = -3,
    MVB_ERR_UNSUPPORTED = -4
};

/* movis.enum.BlendingMode values (movis/enum.py:173-192). */
enum {
    MVB_BLEND_NORMAL = 0, MVB_BLEND_MULTIPLY = 1, MVB_BLEND_SCREEN = 2, MVB_BLEND_OVERLAY = 3,
    MVB_BLEND_DARKEN = 4, MVB_BLEND_LIGHTEN = 5, MVB_BLEND_COLOR_DODGE = 6, MVB_BLEND_COLOR_BURN = 7,
    MVB_BLEND_LINEAR_DODGE = 8, MVB_BLEND_LINEAR_BURN = 9, MVB_BLEND_HARD_LIGHT = 10,
    MVB_BLEND_SOFT_LIGHT = 11, MVB_BLEND_VIVID_LIGHT = 12, MVB_BLEND_LINEAR_LIGHT = 13,
    MVB_BLEND_PIN_LIGHT = 14, MVB_BLEND_DIFFERENCE = 15, MVB_BLEND_EXCLUSION = 16,
    MVB_BLEND_SUBTRACT = 17, MVB_BLEND_COUNT = 18
};

/* movis.enum.MatteMode values (movis/enum.py:225-229). */
enum { MVB_MATTE_NONE = 0, MVB_MATTE_ALPHA = 1, MVB_MATTE_LUMINANCE = 2 };

/* mvb_layer.flags / mvb_alpha_composite flags */
enum {
    /* Use PIL's "over" arithmetic (movis/imgproc.py:197-213).  The reference takes this path iff
     * blending_mode is the NORMAL *enum* and matte_mode is NONE (imgproc.py:265-268); everything
     * else uses the numpy arithmetic of _overlay (imgproc.py:136-170). */
    MVB_LAYER_PIL_OVER = 1,
    /* mvb_layer only, optional: the caller knows that every pixel of `src` has alpha 255 (one reduction
     * when the image is uploaded).  Pure work elimination: a NORMAL layer with opacity 1 and such a source
     * replaces whatever lies below it wherever all of its taps fall inside the source (both "over"
     * operators return the sample exactly there, imgproc.py:156-165 / 197-213), so the kernel does not
     * walk the layers underneath in those tiles.  Never changes a pixel. */
    MVB_LAYER_OPAQUE_SOURCE = 2
};

/* mvb_composite_stack flags */
enum {
    /* Start from the pixels already in `frame` instead of filling it with bg_color (used to
     * continue a stack that has more layers than one launch takes). */
    MVB_STACK_KEEP_FRAME = 1
};

MVB_API const char* mvb_last_error(void);
MVB_API int mvb_abi_version(void);

/* 0 when the current CUDA device can run this library (compute capability 10.x), else an error. */
MVB_API int mvb_device_check(void);

/* Number of kernels this library has launched in this process so far (all entry points, all threads).
 * Diagnostic: lets a benchmark report how many of the library's kernels ran inside a timed region
 * without guessing at the launch structure. */
MVB_API uint64_t mvb_kernel_launches(void);

/* Upper bound on n_layers for a single mvb_composite_stack launch (longer stacks are chunked
 * internally; this is informational). */
MVB_API int mvb_max_layers_per_launch(void);

/*
 * One visible layer of one frame, after the host evaluated its keyframes.
 * Replaces the argument set of the pair of calls at movis/layer/composition.py:811-819:
 *     cv2.warpAffine(fg_image, affine_matrix_fixed, dsize=(W, H), INTER_LINEAR, BORDER_CONSTANT)
 *     alpha_composite(bg_image, warped, position=(offset_x, offset_y), opacity, blending_mode)
 */
typedef struct mvb_layer {
    const void* src;      /* device pointer: post-effect layer image (fg_image) */
    int32_t src_w, src_h;
    int64_t src_stride;   /* bytes */
    double M[6];          /* affine_matrix_fixed, row-major 2x3: layer px -> bbox px (composition.py:906) */
    int32_t bbox_x, bbox_y; /* (offset_x, offset_y) in frame pixels (composition.py:900) */
    int32_t bbox_w, bbox_h; /* (W, H) = dsize (composition.py:896-897); <= 0 skips the layer */
    int32_t blend_mode;   /* MVB_BLEND_* */
    int32_t flags;        /* MVB_LAYER_* */
    double opacity;       /* TransformValue.opacity in [0, 1] */
    const void* alpha_map; /* optional (NULL: none): device pointer to the source's alpha map written by
                            * mvb_alpha_map_rgba8 -- ceil(src_h / 32) rows of ceil(src_w / 32) bytes, the largest
                            * alpha of each 32 x 32 block.  Pure work elimination: tiles whose taps all fall into
                            * blocks of alpha 0 skip the layer where that cannot change a pixel (PIL over, or an
                            * opaque frame); keyed-out green screens, sprites with wide transparent margins. */
} mvb_layer;

/*
 * The frame walk of Composition.__call__ (composition.py:363-368): fill `frame` with bg_color,
 * then warp + blend every layer in paint order, fused into one pass over the frame (each output
 * tile walks the layer stack in registers; the frame is written once).
 * `layers` is a HOST array.  Results are bit-identical to running the reference's
 * warpAffine -> alpha_composite pair per layer.  Frames and layer sources are limited to 32767 px
 * per side (MVB_ERR_UNSUPPORTED beyond).
 * The library keeps a small device scratch buffer per (device, stream) for cv2's coordinate tables
 * (8 bytes x (width + height) per layer); calls that share a stream must not be made concurrently
 * from several host threads.
 */
MVB_API int mvb_composite_stack(void* frame, int32_t frame_w, int32_t frame_h, int64_t frame_stride,
                        const uint8_t bg_color[4], int32_t n_layers, const mvb_layer* layers,
                        int32_t flags, mvb_stream_t stream);

/*
 * A run of frames in one call (the loop of Composition._write_video, composition.py:409-412):
 * frame f uses layers[layer_offsets[f] .. layer_offsets[f+1]) and is written to frames[f].
 * All arrays are HOST arrays; frames[] holds device pointers.
 */
MVB_API int mvb_composite_frames(int32_t n_frames, void* const* frames, int32_t frame_w, int32_t frame_h,
                         int64_t frame_stride, const uint8_t bg_color[4],
                         const int32_t* layer_offsets, const mvb_layer* layers, int32_t flags,
                         mvb_stream_t stream);

/* cv2.warpAffine(src, M, dsize=(dst_w, dst_h), INTER_LINEAR, BORDER_CONSTANT, 0) for RGBA8
 * (composition.py:811-813) as a stand-alone operator.  M is the FORWARD 2x3 matrix. */
MVB_API int mvb_warp_affine_rgba8(const void* src, int32_t src_w, int32_t src_h, int64_t src_stride,
                          void* dst, int32_t dst_w, int32_t dst_h, int64_t dst_stride,
                          const double M[6], mvb_stream_t stream);

/* The alpha map of an RGBA8 image for mvb_layer.alpha_map: map[(y / 32) * ceil(w / 32) + x / 32] = the
 * largest alpha among the pixels of that 32 x 32 block.  No counterpart in the reference (its warpAffine /
 * alpha_composite pair always touches every pixel of the layer's bbox, composition.py:811-819). */
MVB_API int mvb_alpha_map_rgba8(const void* src, int32_t w, int32_t h, int64_t stride, void* map,
                        mvb_stream_t stream);

/* movis.imgproc.alpha_composite(bg, fg, position, opacity, blending_mode, matte_mode)
 * (imgproc.py:216-273).  bg is updated in place.  flags: MVB_LAYER_PIL_OVER selects the PIL
 * arithmetic (only meaningful with MVB_BLEND_NORMAL + MVB_MATTE_NONE). */
MVB_API int mvb_alpha_composite(void* bg, int32_t bg_w, int32_t bg_h, int64_t bg_stride, const void* fg,
                        int32_t fg_w, int32_t fg_h, int64_t fg_stride, int32_t pos_x, int32_t pos_y,
                        double opacity, int32_t blend_mode, int32_t matte_mode, int32_t flags,
                        mvb_stream_t stream);

/* k = 2*ceil(max(1,(4r-1)/2))+1, the kernel size GaussianBlur / Glow / DropShadow derive from
 * `radius` (movis/effect/blur.py:37). */
MVB_API int mvb_blur_ksize(double radius);

/* OpenCV's bit-exact Q0.8 Gaussian kernel for (ksize, sigma): k uint16 weights summing to 256.
 * Host-side helper; replaces the kernel generation inside cv2.GaussianBlur (blur.py:41-43). */
MVB_API int mvb_gaussian_weights_q8(double sigma, int32_t ksize, uint16_t* weights);

/* Blur modes for mvb_effect_blur_rgba8 */
enum {
    MVB_BLUR_PLAIN = 0, /* GaussianBlur.__call__ (movis/effect/blur.py:29-43) */
    MVB_BLUR_GLOW = 1   /* Glow.__call__ (movis/effect/blur.py:66-85): LINEAR_DODGE of the
                           strength-scaled blur over the padded original */
};

/*
 * GaussianBlur / Glow on an RGBA image: pads by `ksize` px per side (RGB edge-replicated, alpha
 * zero; blur.py:39-40), applies cv2.GaussianBlur((ksize,ksize), sigma) in Q8.8/Q16.16 fixed
 * point, and (GLOW) composites.  dst is (src_h + 2*ksize) x (src_w + 2*ksize).
 *   weights      host, ksize entries from mvb_gaussian_weights_q8
 *   strength_lut host, 256 entries = uint8(clip(strength * float32(v), 0, 255)) (blur.py:82);
 *                ignored for MVB_BLUR_PLAIN (may be NULL)
 *   scratch      device, >= mvb_effect_blur_scratch_bytes(src_w, src_h, ksize, 4)
 */
MVB_API size_t mvb_effect_blur_scratch_bytes(int32_t src_w, int32_t src_h, int32_t ksize, int32_t channels);
MVB_API int mvb_effect_blur_rgba8(const void* src, int32_t src_w, int32_t src_h, int64_t src_stride, void* dst,
                          int64_t dst_stride, int32_t ksize, const uint16_t* weights, int32_t mode,
                          const uint8_t* strength_lut, void* scratch, mvb_stream_t stream);

/*
 * DropShadow.__call__ (movis/effect/style.py:49-84).  dst is
 * (src_h + 2k + 2|vy|) x (src_w + 2k + 2|vx|) with k = ksize (0 when radius == 0).
 *   weights     host, ksize entries (ignored when ksize == 0)
 *   opacity_lut host, 256 entries = uint8(opacity * v) in float64 (style.py:64,68)
 *   color       host, round(color(t)) as 3 bytes (style.py:65-67)
 *   (vx, vy)    round(offset * (cos, sin)) (style.py:72-73)
 *   scratch     device, >= mvb_effect_blur_scratch_bytes(src_w, src_h, ksize, 1) + (src_w+2k)*(src_h+2k)
 */
MVB_API int mvb_effect_drop_shadow_rgba8(const void* src, int32_t src_w, int32_t src_h, int64_t src_stride,
                                 void* dst, int64_t dst_stride, int32_t ksize, const uint16_t* weights,
                                 const uint8_t* opacity_lut, const uint8_t color[3], int32_t vx,
                                 int32_t vy, void* scratch, mvb_stream_t stream);

/* ChromaKey.__call__ (movis/contrib/segmentation.py:58-64): alpha := inRange(RGB2HSV(rgb), lo, hi)
 * ? 0 : 255; rgb copied.  src may equal dst. */
MVB_API int mvb_effect_chroma_key_rgba8(const void* src, int32_t w, int32_t h, int64_t src_stride, void* dst,
                                int64_t dst_stride, const uint8_t lower_hsv[3],
                                const uint8_t upper_hsv[3], mvb_stream_t stream);

/* FillColor.__call__ (movis/effect/color.py:25-31): rgb := color, alpha kept. */
MVB_API int mvb_effect_fill_color_rgba8(const void* src, int32_t w, int32_t h, int64_t src_stride, void* dst,
                                int64_t dst_stride, const uint8_t color[3], mvb_stream_t stream);

/* Stripe.__call__ (movis/layer/texture.py:159-191): the procedural stripe image, evaluated with the
 * reference's float64 operations in the reference's order (bit-identical pixels).  `params` is a HOST
 * array of 16 doubles: sin(theta), cos(theta), height / 2, width / 2, total_width, phase,
 * edge0 = max(ratio - 0.01, 0), edge1 - edge0 with edge1 = min(ratio + 0.01, 1), then the two RGBA
 * colours as the reference rounds them (np.round(color), np.round(255 * opacity)).
 * solid = 1 / 2: the reference's early returns for ratio <= 0 / ratio >= 1 (all colour 1 / colour 2). */
MVB_API int mvb_source_stripe_rgba8(void* dst, int32_t w, int32_t h, int64_t dst_stride, const double params[16],
                            int32_t solid, mvb_stream_t stream);

/* HSLShift.__call__ (movis/effect/color.py:56-69).  The three float32 remaps of color.py:64-66
 * are passed as host 256-entry tables (evaluated by the caller with numpy, so their rounding is
 * numpy's own); RGB->HSV and HSV->RGB follow cv2's 8-bit arithmetic, including the different
 * rounding cv2 applies to the last (w mod 32) pixels of each row. */
MVB_API int mvb_effect_hsl_shift_rgba8(const void* src, int32_t w, int32_t h, int64_t src_stride, void* dst,
                               int64_t dst_stride, const uint8_t h_lut[256], const uint8_t s_lut[256],
                               const uint8_t v_lut[256], mvb_stream_t stream);

/* cv2.cvtColor(COLOR_RGB2HSV) of one colour on the host (ChromaKey.__init__,
 * contrib/segmentation.py:46-47).  No device needed. */
MVB_API int mvb_rgb_to_hsv_u8(const uint8_t rgb[3], uint8_t hsv[3]);

#ifdef __cplusplus
}
#endif
#endif /* MOVIS_B200_H */
