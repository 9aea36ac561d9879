// composite.cu -- C-ABI entry points and host side of the fused multi-layer warp + blend kernel,
// plus the stand-alone warp / alpha_composite operators (sm_100a).  C ABI in include/movis_b200.h.
//
// mvb_composite_stack replaces, per frame, the reference's per-layer pair cv2.warpAffine ->
// alpha_composite (movis/layer/composition.py:811-819) and the frame walk around it
// (composition.py:363-368).  The device code lives in stack_float.cuh (arithmetic, walk_layer, the
// one-CTA-per-tile form) and stack_ws.cuh (the persistent warp-specialised kernel that is launched);
// this file turns mvb_layer records into device records, picks the TMA box of every layer, keeps the
// cache of TMA descriptors and launches.  Frame and sources are uint8 RGBA HWC in HBM; the frame is
// written exactly once per launch.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
constexpr int kTileH = 32;
constexpr int kMaxLayers = 32; // per launch: device records + TMA descriptors travel as kernel parameters (8.6 KB)

struct DevLayer {
    // --- first 64 bytes: what the per-tile cull / classification reads (copied to shared memory) ---
    int32_t bx, by, bw, bh;
    int32_t src_w, src_h;
    int32_t box_w, box_h; // TMA box (source px) that covers the footprint of any 32 x 32 tile; 0: no TMA
    float mf[6];        // m rounded to float, for the cull only
    int32_t mode;       // MVB_BLEND_*
    int32_t pil;        // 1: PIL over
    // --- the rest ---
    const uint8_t* src;
    int32_t src_stride;
    int32_t pil_a;      // int(round(opacity * 255)) for the PIL path (imgproc.py:207)
    double m[6];        // inverse map: bbox px -> layer px (OpenCV's own inversion, done on host)
    double opacity;
};
static_assert(sizeof(DevLayer) == 136 && offsetof(DevLayer, src) == 64, "descriptor layout");

struct StackParams {
    uint8_t* frame;
    int32_t W, H;
    int32_t stride;
    uint32_t bg;
    int32_t n_layers;
    int32_t keep_frame;
    int32_t needs_tables; // some layer uses color_dodge / color_burn / soft_light / vivid_light
    int32_t pad_;
    DevLayer layers[kMaxLayers];
};

// Parameters of the opaque-frame kernel: the stack plus one TMA descriptor per layer (a 2-D tiled
// map over the layer's source image, box = DevLayer::box_w x box_h pixels, out-of-bounds = 0).
struct StackParamsTma {
    StackParams s;
    int32_t stage_bytes;   // bytes of one shared-memory stage (the largest box of this launch)
    int32_t stages_log2;   // 1 or 2: two or four stages in flight
    int32_t ux0, uy0, ux1, uy1;   // union of the layers' bboxes, clipped to the frame: tiles outside it are background only
    // cv2's integer tables of every layer for the whole frame (k_stack_tables): per layer tab_w entries
    // (adelta, bdelta) by frame column, then tab_h entries (X0, Y0) by frame row; both padded to tiles
    const int2* tab;
    int32_t tab_w, tab_h;
    int32_t pad_[6];
    alignas(64) CUtensorMap maps[kMaxLayers];
};

// Parameters of k_stack_tables: what cv::warpAffine's table set-up needs of each layer.
struct TableLayer { double m[6]; int32_t bx, by; };
struct TableParams {
    int2* tab;
    int32_t tab_w, tab_h, n_layers, pad_;
    TableLayer l[kMaxLayers];
};

constexpr int kPilOver = 18; // dispatch code of PIL's "over" next to the 18 numpy blend modes

// Host-initialised tables for the division / sqrt heavy blend modes (see BlendTables).
struct BlendTableData { uint32_t rcp24[260]; float2 softg[256]; };   // softg: see blend_diff2<11> (stack_float.cuh)
__device__ BlendTableData g_blend_tables;

enum { kTileEdge = 0, kTileEmpty = 1, kTileInterior = 2 };

} // namespace mvb

#include "stack_float.cuh"   // k_stack_opaque: the float-arithmetic instantiation for opaque frames
#include "stack_ws.cuh"      // k_stack_ws: its persistent, warp-specialised form (default)

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
    if (in.bbox_w <= 0 || in.bbox_h <= 0) {
        std::memset(&out, 0, sizeof(out));
        return 0; // skipped by the kernel's cull (bw <= 0)
    }
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
    out.src_w = in.src_w;
    out.src_h = in.src_h;
    out.src_stride = (int32_t)in.src_stride;
    out.mode = in.blend_mode;
    invert_affine_cv(in.M, out.m);
    out.bx = in.bbox_x;
    out.by = in.bbox_y;
    out.bw = in.bbox_w;
    out.bh = in.bbox_h;
    out.opacity = in.opacity;
    out.pil = (in.flags & MVB_LAYER_PIL_OVER) ? 1 : 0;
    out.pil_a = (int32_t)std::nearbyint(in.opacity * 255.0);
    out.box_w = out.box_h = 0;
    for (int i = 0; i < 6; i++) out.mf[i] = (float)out.m[i];
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

static int upload_blend_tables()
{
    static thread_local int uploaded_dev = -1;
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev == uploaded_dev) return MVB_OK;
    static BlendTableData t;
    for (uint32_t d = 1; d <= 256; d++) t.rcp24[d] = (uint32_t)(((1ull << 24) + d - 1) / d);
    t.rcp24[0] = 0;
    for (int i = 257; i < 260; i++) t.rcp24[i] = 0;
    for (uint32_t b = 0; b < 256; b++) { // g_w3c of imgproc.py:64-68 under uint32 wrap-around
        uint32_t g;
        if (b < 64) {
            uint32_t v = 16u * b - 780300u * b + 260100u;
            v *= b;
            g = v / 65025u;
        } else {
            g = (uint32_t)std::sqrt((double)(255u * b));
        }
        // the kernel needs D = g - b (uint32) only modulo 255 * 256, split as 255 A + R
        const uint32_t D = (g - b) % 65280u;
        t.softg[b] = make_float2((float)(D / 255u), (float)(D % 255u));
    }
    cudaError_t e = cudaMemcpyToSymbol(g_blend_tables, &t, sizeof t);
    if (e != cudaSuccess) return set_error(MVB_ERR_CUDA, cudaGetErrorString(e));
    uploaded_dev = dev;
    return MVB_OK;
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

// The box a 32 x 32 output tile needs from the source: extent of the tile under the inverse map
// (|m0| + |m1|) * 31 plus the second tap and the 1/32 px quantisation (5 px), plus up to 3 px
// because the box must start on a 16-byte boundary of the source row; width a multiple of 8 px
// and height a multiple of 4 so animated scales share a few descriptors.
static void choose_box(DevLayer& L)
{
    L.box_w = L.box_h = 0;
    static const bool disabled = std::getenv("MVB_DISABLE_TMA") != nullptr;   // diagnostics: gather through L1 instead
    if (disabled) return;
    if ((reinterpret_cast<uintptr_t>(L.src) & 15) || (L.src_stride & 15)) return;   // TMA alignment rules
    const double ex = 31.0 * (std::fabs(L.m[0]) + std::fabs(L.m[1]));
    const double ey = 31.0 * (std::fabs(L.m[3]) + std::fabs(L.m[4]));
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
static bool tensor_map_for(const DevLayer& L, CUtensorMap* out)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{L.src, L.src_w, L.src_h, L.src_stride, L.box_w, L.box_h};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)L.src_w, (cuuint64_t)L.src_h};
    const cuuint64_t strides[1] = {(cuuint64_t)L.src_stride};
    const cuuint32_t box[2] = {(cuuint32_t)L.box_w, (cuuint32_t)L.box_h};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(L.src), dims, strides, box, estr,
           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return true;
}

// Device scratch for the tables of one launch, one buffer per (device, stream), grown on demand and kept.
// (cudaMallocAsync was tried: its pool trims at every synchronisation and re-allocating stalls the stream.)
static void* table_scratch(cudaStream_t stream, size_t bytes)
{
    struct Buf { void* p; size_t cap; };
    static std::mutex mu;
    static std::unordered_map<uint64_t, Buf> bufs;
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t key = ((uint64_t)(uintptr_t)stream << 6) | (uint64_t)(dev & 63);
    std::lock_guard<std::mutex> lock(mu);
    if (bufs.size() > 64 && !bufs.count(key)) {   // streams come and go: drop this device's idle buffers
        for (auto it = bufs.begin(); it != bufs.end();) {
            if ((int)(it->first & 63) == dev) { cudaFree(it->second.p); it = bufs.erase(it); }
            else ++it;
        }
    }
    Buf& b = bufs[key];
    if (b.cap < bytes) {
        if (b.p) cudaFree(b.p);   // synchronises: nothing in flight reads the old buffer afterwards
        b.p = nullptr;
        b.cap = 0;
        const size_t cap = std::max<size_t>(bytes, (size_t)1 << 20);
        if (cudaMalloc(&b.p, cap) != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return nullptr; }
        b.cap = cap;
    }
    return b.p;
}

static int launch_stack(void* frame, int W, int H, int64_t stride, const uint8_t bg[4], int n_layers,
                        const mvb_layer* layers, int flags, cudaStream_t stream)
{
    static thread_local StackParamsTma PT;   // 7.6 KB: kept off the stack, reused between launches
    StackParams& P = PT.s;
    P.frame = static_cast<uint8_t*>(frame);
    P.W = W;
    P.H = H;
    P.stride = (int32_t)stride;
    std::memcpy(&P.bg, bg, 4);
    P.pad_ = 0;
    if (int rc = upload_blend_tables()) return rc;
    // An opaque bg_color keeps the frame's alpha at 255 through every layer (both "over" operators
    // map dst alpha 255 to 255), so the division-free kernel instantiation applies.
    const bool opaque = bg[3] == 255 && !(flags & MVB_STACK_KEEP_FRAME);
    const dim3 grid((W + kTileW - 1) / kTileW, (H + kTileH - 1) / kTileH);
    int done = 0;
    bool first = true;
    do {
        const int n = n_layers - done < kMaxLayers ? n_layers - done : kMaxLayers;
        for (int i = 0; i < n; i++) {
            int rc = fill_dev_layer(layers[done + i], P.layers[i]);
            if (rc) return rc;
        }
        P.n_layers = n;
        P.keep_frame = (!first || (flags & MVB_STACK_KEEP_FRAME)) ? 1 : 0;
        P.needs_tables = 0;
        for (int i = 0; i < n; i++) {
            const int m = P.layers[i].mode;
            if (!P.layers[i].pil && (m == 6 || m == 7 || m == 11 || m == 12)) P.needs_tables = 1;
        }
        static const std::string which = [] {   // diagnostics: MVB_STACK_KERNEL=tile selects the one-CTA-per-tile form
            const char* e = std::getenv("MVB_STACK_KERNEL");
            return std::string(e ? e : "");
        }();
        {
            int stage = 0;
            for (int i = 0; i < n; i++) {
                DevLayer& L = P.layers[i];
                choose_box(L);
                if (L.box_w && !tensor_map_for(L, &PT.maps[i])) L.box_w = L.box_h = 0;
                if (L.box_w && L.box_w * L.box_h * 4 > stage) stage = L.box_w * L.box_h * 4;
            }
            static thread_local int smem_dev = -1, n_sm = 0;   // static + dynamic shared memory can exceed 48 KB: opt in once
            int dev = -1;
            cudaGetDevice(&dev);
            if (dev != smem_dev) {
                cudaFuncSetAttribute(k_stack_opaque, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kMaxStageBytes);
                cudaFuncSetAttribute(k_stack_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     5 * kMaxStageBytes + kMaxLayers * 1024);
                cudaFuncSetAttribute(k_stack_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     5 * kMaxStageBytes + kMaxLayers * 1024);
                cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
                smem_dev = dev;
            }
            PT.ux0 = W; PT.uy0 = H; PT.ux1 = 0; PT.uy1 = 0;
            for (int i = 0; i < n; i++) {
                const DevLayer& L = P.layers[i];
                if (L.bw <= 0 || L.bh <= 0) continue;
                PT.ux0 = std::min(PT.ux0, std::max(L.bx, 0));
                PT.uy0 = std::min(PT.uy0, std::max(L.by, 0));
                PT.ux1 = std::max(PT.ux1, (int32_t)std::min<int64_t>((int64_t)L.bx + L.bw, W));
                PT.uy1 = std::max(PT.uy1, (int32_t)std::min<int64_t>((int64_t)L.by + L.bh, H));
            }
            PT.stage_bytes = (stage + 127) & ~127;
            PT.stages_log2 = 4 * PT.stage_bytes <= 36 * 1024 ? 2 : 1;   // keep 3-4 CTAs per SM resident
            static const int force_stages = [] {   // diagnostics: MVB_STAGES_LOG2=1..3
                const char* e = std::getenv("MVB_STAGES_LOG2");
                return e ? std::atoi(e) : 0;
            }();
            if (force_stages >= 1 && force_stages <= 3 && ((size_t)PT.stage_bytes << force_stages) <= 5 * (size_t)kMaxStageBytes)
                PT.stages_log2 = force_stages;
            if (which == "tile" && opaque) {
                if (PT.stages_log2 > 2) PT.stages_log2 = 2;
                k_stack_opaque<<<grid, kFWarps * 32, PT.stage_bytes << PT.stages_log2, stream>>>(PT);
            } else {
                const int n_tiles = (int)(grid.x * grid.y);
                const size_t smem = ((size_t)PT.stage_bytes << PT.stages_log2) + (size_t)n * 1024;
                // persistent CTAs: exactly as many as are resident at once (a CTA that has to wait for a
                // slot would start its share of the tiles when the others are done)
                static thread_local size_t occ_smem[2] = {~(size_t)0, ~(size_t)0};
                static thread_local int occ_dev[2] = {-1, -1}, occ_val[2] = {0, 0};
                if (occ_smem[opaque] != smem || occ_dev[opaque] != dev) {
                    int v = 0;
                    if (opaque) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_stack_ws<true>, kWsThreads, smem);
                    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k_stack_ws<false>, kWsThreads, smem);
                    occ_smem[opaque] = smem;
                    occ_dev[opaque] = dev;
                    occ_val[opaque] = v < 1 ? 1 : v;
                }
                const int per_sm = occ_val[opaque];
                const int ctas = n_tiles < per_sm * n_sm ? n_tiles : per_sm * n_sm;
                // cv2's tables for the whole frame, once per launch (scratch owned by the stream: launches on a
                // stream are ordered, so the next frame's tables cannot overtake this frame's walk)
                static thread_local TableParams TP;
                TP.tab_w = PT.tab_w = (int32_t)grid.x * kTileW;
                TP.tab_h = PT.tab_h = (int32_t)grid.y * kFTileH;
                TP.n_layers = n;
                for (int i = 0; i < n; i++) {
                    std::memcpy(TP.l[i].m, P.layers[i].m, sizeof(TP.l[i].m));
                    TP.l[i].bx = P.layers[i].bx;
                    TP.l[i].by = P.layers[i].by;
                }
                const size_t tab_bytes = (size_t)(n > 0 ? n : 1) * (size_t)(TP.tab_w + TP.tab_h) * sizeof(int2);
                void* tab = table_scratch(stream, tab_bytes);
                if (!tab) return set_error(MVB_ERR_CUDA, "no device memory for the warp tables");
                TP.tab = static_cast<int2*>(tab);
                PT.tab = TP.tab;
                if (n > 0) k_stack_tables<<<dim3((unsigned)((TP.tab_w + TP.tab_h + 255) / 256), (unsigned)n), 256, 0, stream>>>(TP);
                // Programmatic dependent launch: the stack kernel's CTAs may start their prologue (barriers,
                // opacity LUTs) while the tables kernel drains; its producer warps wait for the tables
                // (griddepcontrol.wait) before they copy the first one.
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)ctas);
                cfg.blockDim = dim3(kWsThreads);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = stream;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = n > 0 ? 1 : 0;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                if (opaque) cudaLaunchKernelEx(&cfg, k_stack_ws<true>, PT);
                else cudaLaunchKernelEx(&cfg, k_stack_ws<false>, PT);
            }
        }
        done += n;
        first = false;
    } while (done < n_layers);
    return check_launch("mvb_composite_stack");
}

} // namespace mvb

using namespace mvb;

// Two side streams (+ fork / join events) per device for mvb_composite_frames.
constexpr int kLanes = 2;
struct FrameLanes { cudaStream_t stream[kLanes]; cudaEvent_t fork, join[kLanes]; };

static FrameLanes* frame_lanes()
{
    static thread_local std::vector<std::pair<int, FrameLanes>> per_device;
    int dev = -1;
    cudaGetDevice(&dev);
    for (auto& e : per_device)
        if (e.first == dev) return &e.second;
    FrameLanes L;
    for (int k = 0; k < kLanes; k++) {
        if (cudaStreamCreateWithFlags(&L.stream[k], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&L.join[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    per_device.emplace_back(dev, L);
    return &per_device.back().second;
}

extern "C" int mvb_max_layers_per_launch(void) { return kMaxLayers; }

extern "C" int mvb_composite_stack(void* frame, int32_t W, int32_t H, int64_t stride, const uint8_t bg_color[4],
                                   int32_t n_layers, const mvb_layer* layers, int32_t flags, mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (int rc = check_image(frame, W, H, stride, "frame")) return rc;
    if (!bg_color || n_layers < 0 || (n_layers > 0 && !layers))
        return set_error(MVB_ERR_INVALID_ARG, "bad bg_color / layers");
    return launch_stack(frame, W, H, stride, bg_color, n_layers, layers, flags, (cudaStream_t)stream);
}

extern "C" int mvb_composite_frames(int32_t n_frames, void* const* frames, int32_t W, int32_t H, int64_t stride,
                                    const uint8_t bg_color[4], const int32_t* layer_offsets,
                                    const mvb_layer* layers, int32_t flags, mvb_stream_t stream)
{
    if (int rc = require_device()) return rc;
    if (n_frames < 0 || (n_frames > 0 && (!frames || !layer_offsets)) || !bg_color)
        return set_error(MVB_ERR_INVALID_ARG, "bad frames / layer_offsets / bg_color");
    // Frames are independent, and the persistent kernel of one frame drains unevenly (the last CTAs
    // of a launch leave SMs idle, and the next launch on the same stream waits for all of them): the
    // frames of a run alternate between two side streams forked from `stream` and joined back into
    // it, so the next frame's CTAs start on SMs as they come free.
    FrameLanes* lanes = n_frames >= 2 ? frame_lanes() : nullptr;
    cudaStream_t user = (cudaStream_t)stream;
    if (lanes) {
        cudaEventRecord(lanes->fork, user);
        for (int k = 0; k < kLanes; k++) cudaStreamWaitEvent(lanes->stream[k], lanes->fork, 0);
    }
    int rc = MVB_OK;
    for (int f = 0; f < n_frames && rc == MVB_OK; f++) {
        rc = check_image(frames[f], W, H, stride, "frame");
        if (rc) break;
        const int n = layer_offsets[f + 1] - layer_offsets[f];
        if (n < 0 || (n > 0 && !layers)) {
            rc = set_error(MVB_ERR_INVALID_ARG, "layer_offsets must be non-decreasing");
            break;
        }
        rc = launch_stack(frames[f], W, H, stride, bg_color, n, layers + layer_offsets[f], flags,
                          lanes ? lanes->stream[f % kLanes] : user);
    }
    if (lanes)
        for (int k = 0; k < kLanes; k++) {   // join, also after an error: `stream` must see whatever was enqueued
            cudaEventRecord(lanes->join[k], lanes->stream[k]);
            cudaStreamWaitEvent(user, lanes->join[k], 0);
        }
    return rc;
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
    return check_launch("mvb_alpha_composite");
}
