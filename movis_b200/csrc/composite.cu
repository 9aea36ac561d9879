// composite.cu -- C-ABI entry points and host side of the fused multi-layer warp + blend kernel,
// plus the stand-alone warp / alpha_composite operators (sm_100a).  C ABI in include/movis_b200.h.
//
// mvb_composite_frames replaces the frame loop of Composition._write_video (movis/layer/
// composition.py:405-413): per frame the reference's per-layer pair cv2.warpAffine -> alpha_composite
// (composition.py:811-819) and the frame walk around it (composition.py:363-368).  The device code
// lives in stack_float.cuh (arithmetic, walk_layer) and stack_frames.cuh (k_stack_prep and the
// persistent warp-specialised k_stack_frames); this file turns mvb_layer records into device records,
// picks the TMA box of every layer, keeps the cache of TMA descriptors, stages a run of frames and
// launches: one H2D copy + two kernels per run of up to kMaxRunFrames frames.  Frames and sources
// are uint8 RGBA HWC in HBM; every frame is written exactly once.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include <cuda.h>   // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved through the runtime

#include "../../include/movis_b200.h"
#include "host_util.h"
#include "pixel_math.cuh"

namespace mvb {

constexpr int kTileW = 32;
constexpr int kMaxLayers = 32;   // layers of one frame per launch (one per producer lane); longer stacks continue from the stored frame
constexpr int kMaxMaps = 128;    // distinct TMA descriptors per launch: they travel as kernel parameters (16 KB)

enum { kLayerPil = 1,        // PIL's "over" arithmetic (MVB_LAYER_PIL_OVER)
       kLayerLut = 2,        // set by k_stack_prep: the opacity table is not floor(a c) for any c; DevLayer::lut holds it
       kLayerIntShift = 4,   // the inverse map is a translation by whole pixels: every sample is one texel
       kLayerOccluder = 8 }; // NORMAL, opacity 1, source alpha 255 throughout: hides what is below wherever it covers

constexpr int kPilOver = 18; // dispatch code of PIL's "over" next to the 18 numpy blend modes

struct DevLayer {
    // --- first 80 bytes: what the producer warp keeps in shared memory per frame (kCullInts) ---
    int32_t bx, by, bw, bh;
    int32_t src_w, src_h;
    int32_t box_w, box_h; // TMA box (source px) that covers the footprint of any tile; 0: no TMA
    float mf[6];          // m rounded to float, for the cull only
    int32_t mode;         // dispatch code: MVB_BLEND_* or kPilOver
    int32_t flags;        // kLayer*
    float opac_c;         // k_stack_prep: opacity constant c (walk_layer, stack_float.cuh)
    int32_t pad0_;
    int32_t map_index;    // FramesParams::maps[] entry of (src, box)
    int32_t pad1_;
    // --- the rest ---
    const uint8_t* src;
    float* lut;           // 256 floats of this layer in the run's scratch (written only with kLayerLut)
    const uint8_t* amap;  // mvb_layer::alpha_map or null
    int32_t src_stride;
    int32_t pil_a;        // int(round(opacity * 255)) for the PIL path (imgproc.py:207)
    double m[6];          // inverse map: bbox px -> layer px (OpenCV's own inversion, done on host)
    double opacity;
};
static_assert(sizeof(DevLayer) == 168 && offsetof(DevLayer, src) == 80 && offsetof(DevLayer, opac_c) == 64 &&
              offsetof(DevLayer, map_index) == 72, "descriptor layout");

// soft_light's table (see blend_diff2<11>, stack_float.cuh): 256 pairs (A, R), a compile-time constant.
__device__ const float g_softg[512] = {
#include "softg_table.inc"
};

} // namespace mvb

#include "stack_float.cuh"    // arithmetic + walk_layer
#include "stack_frames.cuh"   // k_stack_prep, k_stack_frames

namespace mvb {

// cv2.warpAffine stand-alone: one thread per destination pixel.
__global__ void __launch_bounds__(256)
k_warp_affine(const uint8_t* __restrict__ src, int sw, int sh, int sstride, uint8_t* __restrict__ dst,
              int dw, int dh, int dstride, const double m0, const double m1, const double m2,
              const double m3, const double m4, const double m5)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const double m[6] = {m0, m1, m2, m3, m4, m5};
    const int X = (warp_x0(m, y) + warp_adelta(m, x)) >> 5;
    const int Y = (warp_y0(m, y) + warp_bdelta(m, x)) >> 5;
    Rgba s;
    uint32_t o = 0;
    if (sample_q5(src, sw, sh, sstride, X, Y, s)) o = pack(s.r, s.g, s.b, s.a);
    reinterpret_cast<uint32_t*>(dst + (int64_t)y * dstride)[x] = o;
}

struct OverParams {
    uint8_t* bg;
    int32_t bw, bh, bstride;
    const uint8_t* fg;
    int32_t fstride;
    int32_t x1, y1, x2, y2, w, h; // intersection: bg origin, fg origin, size (imgproc.py:184-187)
    int32_t mode, matte, pil, pil_a;
    double opacity;
};

// movis.imgproc.alpha_composite: one thread per bg pixel of the region (whole bg for LUMINANCE).
__global__ void __launch_bounds__(256)
k_alpha_composite(const __grid_constant__ OverParams P)
{
    __shared__ uint8_t s_lut[256];
    {
        const int a = threadIdx.x;
        uint32_t v = (uint32_t)a;
        if (P.opacity < 1.0)
            v = P.pil ? (uint32_t)(a * P.pil_a) / 255u : __double2uint_rz(__dmul_rn((double)a, P.opacity));
        s_lut[a] = (uint8_t)v;
    }
    __syncthreads();
    const bool lum = (P.matte == MVB_MATTE_LUMINANCE) && !P.pil;
    const int gx = blockIdx.x * 32 + (threadIdx.x & 31);
    const int gy = blockIdx.y * 8 + (threadIdx.x >> 5);
    // LUMINANCE touches the whole background (imgproc.py:152); other modes only the region.
    const int x = lum ? gx : P.x1 + gx;
    const int y = lum ? gy : P.y1 + gy;
    if (lum ? (x >= P.bw || y >= P.bh) : (gx >= P.w || gy >= P.h)) return;
    uint32_t* dp = reinterpret_cast<uint32_t*>(P.bg + (int64_t)y * P.bstride) + x;
    const bool inside = x >= P.x1 && x < P.x1 + P.w && y >= P.y1 && y < P.y1 + P.h;
    if (!inside) { // only reachable for LUMINANCE
        *dp &= 0x00FFFFFFu;
        return;
    }
    const uint32_t spx = reinterpret_cast<const uint32_t*>(
        P.fg + (int64_t)(y - P.y1 + P.y2) * P.fstride)[x - P.x1 + P.x2];
    const Rgba s = unpack(spx);
    const uint32_t fa = s_lut[s.a];
    uint32_t d = *dp;
    if (lum) { // imgproc.py:147-154
        const uint32_t mask = rgb_to_gray(unpack(d));
        d = (spx & 0x00FFFFFFu) | (((mask * fa) / 255u) << 24);
    } else if (P.pil) {
        d = over_pil(d, s, fa);
    } else {
        d = over_numpy_dyn(P.mode, d, s, fa, P.matte == MVB_MATTE_ALPHA);
    }
    *dp = d;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

static int fill_dev_layer(const mvb_layer& in, DevLayer& out)
{
    std::memset(&out, 0, sizeof(out));
    if (in.bbox_w <= 0 || in.bbox_h <= 0) return 0; // skipped by the kernel's cull (bw <= 0)
    if (!in.src || in.src_w <= 0 || in.src_h <= 0 || in.src_w >= 32768 || in.src_h >= 32768)
        return set_error(MVB_ERR_INVALID_ARG, "layer source must be non-null and 1..32767 px per side");
    if (in.src_stride < (int64_t)in.src_w * 4 || (in.src_stride & 3) || in.src_stride > 0x7FFFFFFF ||
        (reinterpret_cast<uintptr_t>(in.src) & 3))
        return set_error(MVB_ERR_INVALID_ARG, "layer source stride/pointer must be 4-byte aligned and >= 4*width");
    if ((int64_t)in.src_h * in.src_stride > 0xFFFFFFFFll)
        return set_error(MVB_ERR_UNSUPPORTED, "layer source must span less than 4 GiB");
    if (in.blend_mode < 0 || in.blend_mode >= MVB_BLEND_COUNT)
        return set_error(MVB_ERR_INVALID_ARG, "unknown blend mode");
    if (!(in.opacity >= 0.0 && in.opacity <= 1.0))
        return set_error(MVB_ERR_INVALID_ARG, "opacity must be in [0, 1]");
    out.src = static_cast<const uint8_t*>(in.src);
    out.amap = static_cast<const uint8_t*>(in.alpha_map);
    out.src_w = in.src_w;
    out.src_h = in.src_h;
    out.src_stride = (int32_t)in.src_stride;
    const bool pil = (in.flags & MVB_LAYER_PIL_OVER) != 0;
    out.mode = pil ? kPilOver : in.blend_mode;
    out.flags = pil ? kLayerPil : 0;
    invert_affine_cv(in.M, out.m);
    out.bx = in.bbox_x;
    out.by = in.bbox_y;
    out.bw = in.bbox_w;
    out.bh = in.bbox_h;
    out.opacity = in.opacity;
    out.pil_a = (int32_t)std::nearbyint(in.opacity * 255.0);
    for (int i = 0; i < 6; i++) out.mf[i] = (float)out.m[i];
    // cv2's tables for such a map are adelta = 1024 x, bdelta = 0, X0 = 1024 m2 + 16, Y0 = 1024 (y + m5) + 16,
    // all exact: the fractions are 0 and the sample is the texel (x + m2, y + m5) itself
    const double* m = out.m;
    static const bool one_tap_off = std::getenv("MVB_DISABLE_ONE_TAP") != nullptr;   // diagnostics
    if (!one_tap_off && m[0] == 1.0 && m[4] == 1.0 && m[1] == 0.0 && m[3] == 0.0 && m[2] == std::nearbyint(m[2]) &&
        m[5] == std::nearbyint(m[5]) && std::fabs(m[2]) < 1e6 && std::fabs(m[5]) < 1e6)
        out.flags |= kLayerIntShift;
    if ((in.flags & MVB_LAYER_OPAQUE_SOURCE) && in.opacity >= 1.0 && (pil || in.blend_mode == MVB_BLEND_NORMAL))
        out.flags |= kLayerOccluder;
    return 0;
}

static int check_image(const void* p, int w, int h, int64_t stride, const char* what)
{
    if (!p || w <= 0 || h <= 0 || stride < (int64_t)w * 4 || (stride & 3) || stride > 0x7FFFFFFF ||
        (reinterpret_cast<uintptr_t>(p) & 3)) {
        char buf[160];
        std::snprintf(buf, sizeof buf, "%s: need non-null 4-byte-aligned RGBA8 image with stride >= 4*width", what);
        return set_error(MVB_ERR_INVALID_ARG, buf);
    }
    return 0;
}

// ---- TMA descriptors for layer sources ----------------------------------------------------------
constexpr int kMaxStageBytes = 16384;   // per shared-memory stage: boxes up to 64 x 64 px

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// The box a 32 x tile_h output tile needs from the source: extent of the tile under the inverse map
// (|m0| * 31 + |m1| * (tile_h - 1), likewise in y) plus the second tap and the 1/32 px quantisation
// (5 px), plus up to 3 px because the box must start on a 16-byte boundary of the source row; width
// a multiple of 8 px and height a multiple of 4 so animated scales share a few descriptors.
static void choose_box(DevLayer& L, int tile_h)
{
    L.box_w = L.box_h = 0;
    static const bool disabled = std::getenv("MVB_DISABLE_TMA") != nullptr;   // diagnostics: gather through L1 instead
    if (disabled) return;
    if ((reinterpret_cast<uintptr_t>(L.src) & 15) || (L.src_stride & 15)) return;   // TMA alignment rules
    const double ex = 31.0 * std::fabs(L.m[0]) + (tile_h - 1.0) * std::fabs(L.m[1]);
    const double ey = 31.0 * std::fabs(L.m[3]) + (tile_h - 1.0) * std::fabs(L.m[4]);
    if (!(ex < 250.0 && ey < 250.0)) return;
    const int bw = (((int)std::ceil(ex) + 8) + 7) & ~7, bh = (((int)std::ceil(ey) + 5) + 3) & ~3;
    if (bw > 256 || bh > 256 || bw * bh * 4 > kMaxStageBytes) return;
    L.box_w = bw;
    L.box_h = bh;
}

struct MapKey {
    const void* src; int32_t w, h, stride, bw, bh;
    bool operator==(const MapKey& o) const
    { return src == o.src && w == o.w && h == o.h && stride == o.stride && bw == o.bw && bh == o.bh; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const
    {
        uint64_t x = reinterpret_cast<uintptr_t>(k.src) * 0x9E3779B97F4A7C15ull;
        x ^= ((uint64_t)(uint32_t)k.w << 32 | (uint32_t)k.h) * 0xC2B2AE3D27D4EB4Full;
        x ^= ((uint64_t)(uint32_t)k.stride << 20 ^ (uint64_t)k.bw << 10 ^ (uint64_t)k.bh) * 0x165667B19E3779F9ull;
        return (size_t)(x ^ (x >> 29));
    }
};

// A descriptor is a pure function of its key, so entries never go stale; the cache is only bounded.
static bool tensor_map_for(const MapKey& key, CUtensorMap* out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)key.w, (cuuint64_t)key.h};
    const cuuint64_t strides[1] = {(cuuint64_t)key.stride};
    const cuuint32_t box[2] = {(cuuint32_t)key.bw, (cuuint32_t)key.bh};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(key.src), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return true;
}

// ---- per (device, stream) state -------------------------------------------------------------------
// A run's records, opacity tables and coordinate tables live in ONE device buffer per stream, grown on
// demand and kept: work on a stream is ordered, so the next run's upload cannot overtake this run's
// kernels.  The host side of the upload is a ring of pinned staging buffers, each guarded by an event
// recorded after its copy, so the caller can plan several runs ahead of the device.  Contexts are
// never evicted (a program uses a handful of streams); they are released at process exit.
constexpr int kRing = 8;
struct StreamCtx {
    uint8_t* dev = nullptr;
    size_t dev_cap = 0;
    struct HostSlot { uint8_t* p = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; bool pending = false; } ring[kRing];
    int next = 0;
};

static StreamCtx* stream_ctx(cudaStream_t stream)
{
    static std::mutex mu;
    static std::unordered_map<uint64_t, std::unique_ptr<StreamCtx>> all;
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t key = ((uint64_t)(uintptr_t)stream << 6) ^ (uint64_t)(dev & 63);
    std::lock_guard<std::mutex> lock(mu);
    auto& slot = all[key];
    if (!slot) slot.reset(new StreamCtx());
    return slot.get();
}

static int env_int(const char* name, int fallback)
{
    const char* e = std::getenv(name);
    return e && *e ? std::atoi(e) : fallback;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// One run: frames[0..nF) with layers[beg[f] .. beg[f] + cnt[f]) each (cnt <= kMaxLayers), one
// H2D copy + k_stack_prep + k_stack_frames.  Returns the number of frames launched through
// *launched (fewer than nF when the run needs more than kMaxMaps descriptors; at least one).
static int launch_run(int nF, void* const* frames, const int32_t* beg, const int32_t* cnt, const mvb_layer* layers,
                      int W, int H, int64_t stride, const uint8_t bg[4], bool keep, bool opaque, cudaStream_t stream,
                      int* launched)
{
    static const int rows = (MVB_UNITS <= 4 && env_int("MVB_ROWS", 4) == 8) ? 8 : 4;   // rows per thread: tile = 32 x (8 * rows)
    const int tile_h = kWsConsumers * rows;
    const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + tile_h - 1) / tile_h;
    const int tab_w = tiles_x * kTileW, tab_h = tiles_y * tile_h;
    int nL = 0;
    for (int f = 0; f < nF; f++) nL += cnt[f];

    int max_cnt = 0;
    for (int f = 0; f < nF; f++) max_cnt = std::max(max_cnt, (int)cnt[f]);
    const int tile_stride = 1 + max_cnt;   // int2 entries per (frame, tile) list: count + one per layer

    // device layout: [counter | FrameRec x nF | DevLayer x nL | 256 floats x nL | tables x nL | tile lists]; the first three are uploaded
    const size_t off_frames = 64;   // [0, 64): the run's work counter, zeroed by the upload
    const size_t off_layers = align_up(off_frames + (size_t)nF * sizeof(FrameRec), 64);
    const size_t off_luts = align_up(off_layers + (size_t)nL * sizeof(DevLayer), 256);
    const size_t off_tab = off_luts + (size_t)nL * 256 * sizeof(float);
    const size_t off_tiles = align_up(off_tab + (size_t)std::max(nL, 1) * (size_t)(tab_w + tab_h) * sizeof(int2), 256);
    const size_t total = off_tiles + (size_t)nF * (size_t)(tiles_x * tiles_y) * (size_t)tile_stride * sizeof(int2);
    StreamCtx* C = stream_ctx(stream);
    if (C->dev_cap < total) {
        if (C->dev) cudaFree(C->dev);   // synchronises: nothing in flight reads the old buffer afterwards
        C->dev = nullptr;
        C->dev_cap = 0;
        const size_t cap = std::max<size_t>(total + total / 4, (size_t)1 << 20);
        if (cudaMalloc(&C->dev, cap) != cudaSuccess) {
            cudaGetLastError();
            return set_error(MVB_ERR_CUDA, "no device memory for the run's records and warp tables");
        }
        C->dev_cap = cap;
    }
    StreamCtx::HostSlot& S = C->ring[C->next];
    C->next = (C->next + 1) % kRing;
    if (!S.done && cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming) != cudaSuccess)
        return set_error(MVB_ERR_CUDA, "cudaEventCreate failed");
    if (S.pending) { cudaEventSynchronize(S.done); S.pending = false; }   // its last upload has been consumed
    if (S.cap < off_luts) {
        if (S.p) cudaFreeHost(S.p);
        S.p = nullptr;
        S.cap = 0;
        const size_t cap = std::max<size_t>(off_luts + off_luts / 4, (size_t)64 << 10);
        if (cudaHostAlloc(&S.p, cap, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return set_error(MVB_ERR_CUDA, "no pinned host memory for the run's records");
        }
        S.cap = cap;
    }
    std::memset(S.p, 0, off_frames);
    FrameRec* h_frames = reinterpret_cast<FrameRec*>(S.p + off_frames);
    DevLayer* h_layers = reinterpret_cast<DevLayer*>(S.p + off_layers);
    DevLayer* d_layers = reinterpret_cast<DevLayer*>(C->dev + off_layers);
    float* d_luts = reinterpret_cast<float*>(C->dev + off_luts);
    int2* d_tab = reinterpret_cast<int2*>(C->dev + off_tab);
    int2* d_tiles = reinterpret_cast<int2*>(C->dev + off_tiles);

    static thread_local FramesParams P;   // 16 KB: kept off the stack, reused between launches
    static thread_local std::unordered_map<MapKey, int, MapKeyHash> run_maps;
    run_maps.clear();
    int n_maps = 0, stage = 0, li = 0, done = 0;
    for (int f = 0; f < nF; f++) {
        const int maps_before = n_maps, li_before = li, stage_before = stage;
        FrameRec& F = h_frames[f];
        std::memset(&F, 0, sizeof F);
        F.frame = static_cast<uint8_t*>(frames[f]);
        F.layers = d_layers + li;
        F.tab = d_tab + (size_t)li * (size_t)(tab_w + tab_h);
        F.n_layers = cnt[f];
        F.ux0 = W; F.uy0 = H; F.ux1 = 0; F.uy1 = 0;
        bool overflow = false;
        for (int i = 0; i < cnt[f] && !overflow; i++, li++) {
            DevLayer& L = h_layers[li];
            if (int rc = fill_dev_layer(layers[beg[f] + i], L)) return rc;
            if (L.bw <= 0 || L.bh <= 0) continue;
            L.lut = d_luts + (size_t)li * 256;
            F.ux0 = std::min(F.ux0, std::max(L.bx, 0));
            F.uy0 = std::min(F.uy0, std::max(L.by, 0));
            F.ux1 = std::max(F.ux1, (int32_t)std::min<int64_t>((int64_t)L.bx + L.bw, W));
            F.uy1 = std::max(F.uy1, (int32_t)std::min<int64_t>((int64_t)L.by + L.bh, H));
            choose_box(L, tile_h);
            if (!L.box_w) continue;
            const MapKey key{L.src, L.src_w, L.src_h, L.src_stride, L.box_w, L.box_h};
            auto it = run_maps.find(key);
            if (it != run_maps.end()) {
                L.map_index = it->second;
            } else if (n_maps == kMaxMaps) {
                overflow = true;   // this frame starts the next run
                break;
            } else if (tensor_map_for(key, &P.maps[n_maps])) {
                run_maps.emplace(key, n_maps);
                L.map_index = n_maps++;
            } else {
                L.box_w = L.box_h = 0;   // no descriptor: gather through L1
            }
            if (L.box_w) stage = std::max(stage, L.box_w * L.box_h * 4);
        }
        if (overflow) {
            for (auto it = run_maps.begin(); it != run_maps.end();)
                it = it->second >= maps_before ? run_maps.erase(it) : std::next(it);
            n_maps = maps_before;
            li = li_before;
            stage = stage_before;
            break;
        }
        done = f + 1;
    }
    if (done == 0) return set_error(MVB_ERR_UNSUPPORTED, "a single frame needs more TMA descriptors than a launch takes");
    nF = done;
    nL = li;
    *launched = nF;

    const size_t upload = off_layers + (size_t)nL * sizeof(DevLayer);
    if (cudaMemcpyAsync(C->dev, S.p, upload, cudaMemcpyHostToDevice, stream) != cudaSuccess)
        return check_launch("mvb_composite_frames (upload)");
    cudaEventRecord(S.done, stream);
    S.pending = true;

    {
        PrepParams T;
        T.layers = d_layers;
        T.tab = d_tab;
        T.frames = reinterpret_cast<const FrameRec*>(C->dev + off_frames);
        T.tiles = d_tiles;
        T.n_layers = nL;
        T.n_frames = nF;
        T.tab_w = tab_w;
        T.tab_h = tab_h;
        T.tab_blocks = (tab_w + tab_h + 255) / 256 + 1;
        T.W = W;
        T.H = H;
        T.tile_h = tile_h;
        T.tiles_x = tiles_x;
        T.n_tiles = tiles_x * tiles_y;
        T.tile_stride = tile_stride;
        T.opaque = opaque ? 1 : 0;
        T.stride = (int32_t)stride;
        std::memcpy(&T.bg, bg, 4);
        T.keep_frame = keep ? 1 : 0;
        const long long blocks = (long long)nL * T.tab_blocks + (long long)nF * ((T.n_tiles + 7) / 8);
        if (blocks > 0x7FFFFFFFll) return set_error(MVB_ERR_UNSUPPORTED, "run too large");
        k_stack_prep<<<(unsigned)blocks, 256, 0, stream>>>(T);
        count_launches(1);
    }

    P.frames = reinterpret_cast<const FrameRec*>(C->dev + off_frames);
    P.next_item = reinterpret_cast<float*>(C->dev);
    P.n_frames = nF;
    P.W = W;
    P.H = H;
    P.stride = (int32_t)stride;
    std::memcpy(&P.bg, bg, 4);
    P.keep_frame = keep ? 1 : 0;
    P.tab_w = tab_w;
    P.tab_h = tab_h;
    P.tiles_x = tiles_x;
    P.n_tiles = tiles_x * tiles_y;
    P.tile_stride = tile_stride;
    P.tiles = d_tiles;
    P.force_units = 0;
    P.stage_bytes = (stage + 127) & ~127;

    // kernel variant + its launch geometry (cached per device)
    typedef void (*Kernel)(const FramesParams);
    const int variant = (opaque ? 1 : 0) | (rows == 8 ? 2 : 0);
#if MVB_UNITS <= 4
    static const Kernel kernels[4] = {k_stack_frames<false, 4>, k_stack_frames<true, 4>,
                                      k_stack_frames<false, 8>, k_stack_frames<true, 8>};
#else
    static const Kernel kernels[4] = {k_stack_frames<false, 4>, k_stack_frames<true, 4>,
                                      k_stack_frames<false, 4>, k_stack_frames<true, 4>};
#endif
    struct Geometry { int dev = -1, n_sm = 0, static_smem = 0, dyn_max = 0, occ_smem = -1, occ = 0; };
    static thread_local Geometry geo[4];
    Geometry& Gm = geo[variant];
    int dev = -1;
    cudaGetDevice(&dev);
    if (Gm.dev != dev) {
        Gm = Geometry();
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kernels[variant]) != cudaSuccess) return check_launch("mvb_composite_frames");
        Gm.static_smem = (int)fa.sharedSizeBytes;
        Gm.dyn_max = 6 * kMaxStageBytes;
        // static + dynamic shared memory can exceed 48 KB: opt in once
        cudaFuncSetAttribute(kernels[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, Gm.dyn_max);
        cudaDeviceGetAttribute(&Gm.n_sm, cudaDevAttrMultiProcessorCount, dev);
        Gm.dev = dev;
    }
    // as many stages as leave room for the CTAs per SM the register budget allows (3 with 4 rows, 2 with 8);
    // (228 KB per SM, 1 KB reserved per CTA)
    const int per_cta = (rows == 8 ? 113 : 75) * 1024 - Gm.static_smem - 1024;
    static const int force_stages = env_int("MVB_STAGES", 0);   // diagnostics: 2..8
    int stages = P.stage_bytes > 0 ? per_cta / P.stage_bytes : kMaxStages;
    if (force_stages) stages = force_stages;
    stages = std::max(2, std::min(stages, kMaxStages));
    while (stages > 2 && stages * P.stage_bytes > Gm.dyn_max) --stages;
    P.stages = stages;
    const int smem = P.stage_bytes * P.stages;
    if (Gm.occ_smem != smem) {
        int v = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernels[variant], kWsThreads, (size_t)smem);
        Gm.occ_smem = smem;
        Gm.occ = v < 1 ? 1 : v;
    }
    // persistent CTAs: exactly as many as are resident at once (a CTA that has to wait for a slot would
    // start its share of the items when the others are done)
    const long long n_items = (long long)nF * P.n_tiles;
    if (n_items + (long long)Gm.occ * Gm.n_sm >= (1ll << 24)) return set_error(MVB_ERR_UNSUPPORTED, "run too large");
    const int ctas = (int)std::min<long long>(n_items, (long long)Gm.occ * Gm.n_sm);
    // Programmatic dependent launch: the CTAs may start their prologue while k_stack_prep drains; the
    // producer warps wait for its results (griddepcontrol.wait) before they read the first record.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ctas);
    cfg.blockDim = dim3(kWsThreads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernels[variant], P);
    count_launches(1);
    return check_launch("mvb_composite_frames");
}

// Frames [0, n_frames) with layers[offsets[f] .. offsets[f + 1]) each, split into runs.
static int launch_frames(int n_frames, void* const* frames, int W, int H, int64_t stride, const uint8_t bg[4],
                         const int32_t* offsets, const mvb_layer* layers, int flags, cudaStream_t stream)
{
    // (tile origins travel as 16-bit halves of one word, k_stack_prep's list header)
    if (W > 32767 || H > 32767) return set_error(MVB_ERR_UNSUPPORTED, "frames are limited to 32767 px per side");
    // An opaque bg_color keeps the frame's alpha at 255 through every layer (both "over" operators
    // map dst alpha 255 to 255), so the division-free kernel instantiation applies.
    const bool keep = (flags & MVB_STACK_KEEP_FRAME) != 0;
    const bool opaque = bg[3] == 255 && !keep;
    static const int max_run = std::max(1, env_int("MVB_RUN_FRAMES", 32));
    static const size_t max_tab = (size_t)std::max(1, env_int("MVB_RUN_TABLE_MB", 32)) << 20;
    const int tiles_x = (W + kTileW - 1) / kTileW;
    const size_t tab_per_layer = ((size_t)tiles_x * kTileW + (size_t)H + 64) * sizeof(int2);
    static thread_local std::vector<int32_t> beg, cnt;
    int f = 0;
    while (f < n_frames) {
        const int n = offsets[f + 1] - offsets[f];
        if (n > kMaxLayers) {
            // a stack longer than one launch takes: passes of kMaxLayers layers, each continuing from the stored frame
            for (int done = 0; done < n; done += kMaxLayers) {
                const int32_t b = offsets[f] + done, c = std::min(kMaxLayers, n - done);
                int launched = 0;
                if (int rc = launch_run(1, frames + f, &b, &c, layers, W, H, stride, bg, keep || done > 0, opaque, stream, &launched))
                    return rc;
            }
            f += 1;
            continue;
        }
        beg.clear();
        cnt.clear();
        size_t nL = 0;
        int f1 = f;
        // (the run's work counter counts (frame, tile) items exactly up to 2^24: launch_run checks it)
        const long long tiles_per_frame = (long long)tiles_x * ((H + 31) / 32);
        const int fit = (int)std::max<long long>(1, ((1ll << 24) - 4096) / std::max<long long>(tiles_per_frame, 1));
        while (f1 < n_frames && f1 - f < std::min(max_run, fit)) {
            const int c = offsets[f1 + 1] - offsets[f1];
            if (c > kMaxLayers || (f1 > f && (nL + c) * tab_per_layer > max_tab)) break;
            beg.push_back(offsets[f1]);
            cnt.push_back(c);
            nL += c;
            ++f1;
        }
        int launched = 0;
        if (int rc = launch_run(f1 - f, frames + f, beg.data(), cnt.data(), layers, W, H, stride, bg, keep, opaque, stream, &launched))
            return rc;
        f += launched;
    }
    return MVB_OK;
}

} // namespace mvb

using namespace mvb;

// One CTA per 32 x 32 block of the image: the largest alpha of the block (mvb_layer::alpha_map).
__global__ void __launch_bounds__(256)
k_alpha_map(const uint8_t* __restrict__ src, int w, int h, int64_t stride, uint8_t* __restrict__ map, int gw)
{
    __shared__ uint32_t s_max[8];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int x = bx + (threadIdx.x & 31);
    uint32_t m = 0;
    if (x < w) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int y = by + (threadIdx.x >> 5) * 4 + r;
            if (y < h) m = max(m, reinterpret_cast<const uint32_t*>(src + (int64_t)y * stride)[x] >> 24);
        }
    }
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) m = max(m, s_max[i]);
        map[(size_t)blockIdx.y * (size_t)gw + blockIdx.x] = (uint8_t)m;
    }
}

extern "C" int mvb_alpha_map_rgba8(const void* src, int32_t w, int32_t h, int64_t stride, void* map, mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_image(src, w, h, stride, "src")) return rc;
    if (!map) return set_error(MVB_ERR_INVALID_ARG, "map is null");
    const int gw = (w + 31) / 32, gh = (h + 31) / 32;
    k_alpha_map<<<dim3((unsigned)gw, (unsigned)gh), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint8_t*>(src), w, h, stride, static_cast<uint8_t*>(map), gw);
    count_launches(1);
    return check_launch("mvb_alpha_map_rgba8");
}

extern "C" int mvb_max_layers_per_launch(void) { return kMaxLayers; }

extern "C" int mvb_composite_stack(void* frame, int32_t W, int32_t H, int64_t stride, const uint8_t bg_color[4],
                                   int32_t n_layers, const mvb_layer* layers, int32_t flags, mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_image(frame, W, H, stride, "frame")) return rc;
    if (!bg_color || n_layers < 0 || (n_layers > 0 && !layers))
        return set_error(MVB_ERR_INVALID_ARG, "bad bg_color / layers");
    const int32_t offsets[2] = {0, n_layers};
    void* frames[1] = {frame};
    return launch_frames(1, frames, W, H, stride, bg_color, offsets, layers, flags, (cudaStream_t)stream);
}

extern "C" int mvb_composite_frames(int32_t n_frames, void* const* frames, int32_t W, int32_t H, int64_t stride,
                                    const uint8_t bg_color[4], const int32_t* layer_offsets,
                                    const mvb_layer* layers, int32_t flags, mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (n_frames < 0 || (n_frames > 0 && (!frames || !layer_offsets)) || !bg_color)
        return set_error(MVB_ERR_INVALID_ARG, "bad frames / layer_offsets / bg_color");
    for (int f = 0; f < n_frames; f++) {
        if (int rc = check_image(frames[f], W, H, stride, "frame")) return rc;
        const int n = layer_offsets[f + 1] - layer_offsets[f];
        if (n < 0 || (n > 0 && !layers)) return set_error(MVB_ERR_INVALID_ARG, "layer_offsets must be non-decreasing");
    }
    return launch_frames(n_frames, frames, W, H, stride, bg_color, layer_offsets, layers, flags, (cudaStream_t)stream);
}

extern "C" int mvb_warp_affine_rgba8(const void* src, int32_t sw, int32_t sh, int64_t sstride, void* dst,
                                     int32_t dw, int32_t dh, int64_t dstride, const double M[6],
                                     mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_image(src, sw, sh, sstride, "src")) return rc;
    if (int rc = check_image(dst, dw, dh, dstride, "dst")) return rc;
    if (!M || sw >= 32768 || sh >= 32768 || (int64_t)sh * sstride > 0xFFFFFFFFll)
        return set_error(MVB_ERR_INVALID_ARG, "bad matrix or source too large");
    double m[6];
    invert_affine_cv(M, m);
    const dim3 grid((dw + 31) / 32, (dh + 7) / 8);
    k_warp_affine<<<grid, 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint8_t*>(src), sw, sh, (int)sstride, static_cast<uint8_t*>(dst), dw, dh,
        (int)dstride, m[0], m[1], m[2], m[3], m[4], m[5]);
    count_launches(1);
    return check_launch("mvb_warp_affine_rgba8");
}

extern "C" int mvb_alpha_composite(void* bg, int32_t bw, int32_t bh, int64_t bstride, const void* fg, int32_t fw,
                                   int32_t fh, int64_t fstride, int32_t px, int32_t py, double opacity,
                                   int32_t blend_mode, int32_t matte_mode, int32_t flags, mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_image(bg, bw, bh, bstride, "bg")) return rc;
    if (int rc = check_image(fg, fw, fh, fstride, "fg")) return rc;
    if (blend_mode < 0 || blend_mode >= MVB_BLEND_COUNT || matte_mode < 0 || matte_mode > 2)
        return set_error(MVB_ERR_INVALID_ARG, "unknown blend or matte mode");
    if (!(opacity >= 0.0 && opacity <= 1.0)) return set_error(MVB_ERR_INVALID_ARG, "opacity must be in [0, 1]");
    const bool pil = (flags & MVB_LAYER_PIL_OVER) != 0;
    if (pil && (blend_mode != MVB_BLEND_NORMAL || matte_mode != MVB_MATTE_NONE))
        return set_error(MVB_ERR_INVALID_ARG, "MVB_LAYER_PIL_OVER requires NORMAL blending and no matte");
    OverParams P;
    P.bg = static_cast<uint8_t*>(bg);
    P.bw = bw; P.bh = bh; P.bstride = (int32_t)bstride;
    P.fg = static_cast<const uint8_t*>(fg);
    P.fstride = (int32_t)fstride;
    P.x1 = px > 0 ? px : 0; P.y1 = py > 0 ? py : 0;
    P.x2 = px < 0 ? -px : 0; P.y2 = py < 0 ? -py : 0;
    const int64_t w = (int64_t)(px + (int64_t)fw < bw ? px + (int64_t)fw : bw) - P.x1;
    const int64_t h = (int64_t)(py + (int64_t)fh < bh ? py + (int64_t)fh : bh) - P.y1;
    if (w <= 0 || h <= 0) return MVB_OK; // imgproc.py:189-190 (and PIL's clipped paste)
    P.w = (int32_t)w; P.h = (int32_t)h;
    P.mode = blend_mode; P.matte = matte_mode; P.pil = pil ? 1 : 0;
    P.pil_a = (int32_t)std::nearbyint(opacity * 255.0);
    P.opacity = opacity;
    const bool lum = matte_mode == MVB_MATTE_LUMINANCE;
    const dim3 grid(((lum ? bw : P.w) + 31) / 32, ((lum ? bh : P.h) + 7) / 8);
    k_alpha_composite<<<grid, 256, 0, (cudaStream_t)stream>>>(P);
    count_launches(1);
    return check_launch("mvb_alpha_composite");
}
