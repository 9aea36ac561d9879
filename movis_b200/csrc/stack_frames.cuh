// stack_frames.cuh -- the fused warp + blend kernel: ONE persistent, warp-specialised launch over a
// whole run of frames (sm_100a), plus the small preparation kernel that precedes it.
//
// Contract: the frame walk of Composition.__call__ (movis/layer/composition.py:363-368) with
// cv2.warpAffine -> alpha_composite per layer (composition.py:811-819), bit for bit, for every frame
// of the loop in Composition._write_video (composition.py:405-413).  Frames are independent, so a
// launch takes a work list of (frame, tile) items; every CTA is persistent (grid = the CTAs that are
// resident at once, items taken round-robin) and runs straight through frame boundaries: there is no
// per-frame launch, no per-frame prologue and no tail until the last frame of the run.
//
//   k_stack_prep   one launch per run, everything that is embarrassingly parallel: per layer cv2's
//                  integer coordinate tables for the whole frame and the layer's opacity constant (see
//                  walk_layer in stack_float.cuh); per (frame, tile) the cull + classification of the
//                  frame's layers against the tile, as a compact list in paint order
//   k_stack_frames 8 CONSUMER warps own one 32 x (8 * ROWS) tile at a time, pixels in registers for the
//                  whole layer walk; a ninth, PRODUCER warp runs ahead of them:
//     producer   reads the next tile's list (prefetched one tile ahead); for up to 8 layers ("unit") writes
//                the info words and bulk-copies the tile's slice of the tables into one of four
//                shared-memory unit buffers, whose mbarrier publishes the unit when the copies have
//                landed; requests the layers' source boxes through TMA into a ring of stages as they
//                come free.  It is on the consumers' critical path (they wait for units and boxes), so
//                it does as little as possible, and when nothing can proceed it waits in
//                mbarrier.try_wait on the event it needs next (a free stage if a box is waiting for
//                one, else a free unit buffer);
//     consumers  wait for the unit, walk its layers (taps from the TMA stages, float arithmetic),
//                release the unit, store the tile after its last unit.
//
// There is no __syncthreads() after start-up: all hand-offs are mbarrier phases; waits are bounded by
// wall time (20 s) and trap instead of hanging the device.
#pragma once

#include "stack_float.cuh"

namespace mvb {

#ifndef MVB_CONSUMERS
#define MVB_CONSUMERS 8
#endif
constexpr int kWsConsumers = MVB_CONSUMERS;       // consumer warps x ROWS rows = one tile
constexpr int kWsThreads = (kWsConsumers + 1) * 32;
constexpr int kWsSlots = 8;                       // layers per unit
#ifndef MVB_UNITS
#define MVB_UNITS 4
#endif
constexpr int kWsUnits = MVB_UNITS;               // unit buffers: how far the producer may run ahead (power of two)
constexpr int kCullInts = 20;                     // leading ints of a DevLayer the producer keeps in shared memory
constexpr int kCullPitch = 21;                    // conflict-free by lane

// Per layer of a unit.
template <int ROWS>
struct __align__(16) UnitSlot {
    // x: mode (8 bits) | class (3) | phase parity of the stage (bit 11) | kInfoOneTap (bit 12) | row pitch of the box << 16
    // y: address of the stage's "full" barrier ("empty" is 8 * kMaxStages bytes on)
    // z: the layer's opacity constant c (float bits)
    // w: address the stage would have at source pixel (0, 0): stage - origin
    int4 info;
    int2 col[kTileW];                 // (adelta, bdelta) per tile column
    int2 row[kWsConsumers * ROWS];    // (X0, Y0) per tile row
};

template <int ROWS>
struct __align__(16) UnitMem {
    int4 hdr;                  // x: tile x0, y: tile y0, z: layers in this unit | first << 8 | last << 9 | stop << 10
    uint8_t* frame;            // the frame this tile belongs to
    const DevLayer* layers;    // that frame's layer records (for the layers that gather through L1)
    int layer[kWsSlots];       // layer index of each slot
    UnitSlot<ROWS> slot[kWsSlots];
};

// One frame of a run.
struct FrameRec {
    uint8_t* frame;
    const DevLayer* layers;
    const int2* tab;           // per layer tab_w entries (adelta, bdelta) by frame column, then tab_h entries (X0, Y0) by row
    int32_t n_layers;
    int32_t ux0, uy0, ux1, uy1;   // union of the layers' bboxes, clipped to the frame: tiles outside it are background only
    int32_t pad_[5];
};
static_assert(sizeof(FrameRec) == 64, "frame record layout");

// ---------------------------------------------------------------------------------------------------
// k_stack_prep: 1-D grid of 256-thread blocks with three roles.
//   blocks [0, n_layers * tab_blocks): layer = b / tab_blocks, part = b % tab_blocks
//     part < tab_blocks - 1   cv::warpAffine's integer tables (imgwarp.cpp: adelta / bdelta by
//         destination column, X0 / Y0 by destination row; fp64 without FMA contraction, pixel_math.cuh)
//         of the layer for the whole frame.  They depend on the layer and the frame column / row
//         only; the producer warp of k_stack_frames copies a tile's slice with cp.async.bulk.
//     part == tab_blocks - 1  the layer's opacity constant: thread a evaluates the reference's table
//         entry k_a; c = m / 2^18 with m = max_a ceil(k_a 2^18 / a) is the smallest candidate, and it is
//         the answer iff m a < (k_a + 1) 2^18 for every a.  Otherwise the table itself is stored.
//   the remaining blocks: one WARP per (frame, tile), lane = layer: cull the layer against the tile and
//     decide how its taps reach it (kTile*, with the TMA box origin); the layers that touch the tile
//     are written as a compact list in paint order:  rec[0] = (count | frame << 8, tile x0 | y0 << 16),
//     rec[1 + s] = (layer | class << 8, origin x (16 bits) | origin y << 16).  A tile no layer touches is finished right here (filled with
//     the background colour, or left alone when the run continues stored frames): k_stack_frames
//     skips it.
// ---------------------------------------------------------------------------------------------------
struct PrepParams {
    DevLayer* layers;
    int2* tab;
    const FrameRec* frames;
    int2* tiles;               // n_frames * n_tiles records of tile_stride int2 each
    int32_t n_layers, n_frames;
    int32_t tab_w, tab_h, tab_blocks;
    int32_t W, H, tile_h, tiles_x, n_tiles, tile_stride;
    int32_t opaque;            // the frames' alpha is 255 throughout
    int32_t stride;            // frame row stride, bytes
    uint32_t bg;
    int32_t keep_frame;
};

__global__ void __launch_bounds__(256)
k_stack_prep(const PrepParams T)
{
    asm volatile("griddepcontrol.launch_dependents;");   // k_stack_frames may start its prologue now
    const int n_layer_blocks = T.n_layers * T.tab_blocks;
    if ((int)blockIdx.x >= n_layer_blocks) {
        // ---- cull + classify: warp = tile, lane = layer ---------------------------------------------
        const int groups = (T.n_tiles + 7) >> 3;
        const int b = blockIdx.x - n_layer_blocks;
        const int f = b / groups, t = (b - f * groups) * 8 + (threadIdx.x >> 5);
        const int lane = threadIdx.x & 31;
        if (t >= T.n_tiles) return;
        const FrameRec& F = T.frames[f];
        const int tx0 = (t % T.tiles_x) * kTileW, ty0 = (t / T.tiles_x) * T.tile_h;
        int cls = -1;
        bool hides = false;   // the layer replaces everything below it in this tile
        int2 org = make_int2(0, 0);
        // (tiles outside the union of the layers' bboxes take the short way: background only)
        const bool covered = tx0 < F.ux1 && tx0 + kTileW > F.ux0 && ty0 < F.uy1 && ty0 + T.tile_h > F.uy0;
        if (covered && lane < F.n_layers) {
            const DevLayer& L = F.layers[lane];
            const int bx = L.bx, by = L.by, bw = L.bw, bh = L.bh, src_w = L.src_w, src_h = L.src_h;
            const int box_w = L.box_w, box_h = L.box_h;
            const float m0 = L.mf[0], m1 = L.mf[1], m2 = L.mf[2], m3 = L.mf[3], m4 = L.mf[4], m5 = L.mf[5];
            // tile ∩ bbox ∩ frame, in bbox coordinates (inclusive corners)
            const int u0 = max(tx0, bx) - bx, u1 = min(min(tx0 + kTileW, bx + bw), T.W) - 1 - bx;
            const int v0 = max(ty0, by) - by, v1 = min(min(ty0 + T.tile_h, by + bh), T.H) - 1 - by;
            if (bw > 0 && bh > 0 && u0 <= u1 && v0 <= v1) {
                // source-space extent of that rectangle (affine: centre +- sum |coef| * half size); 0.25 px
                // of slack covers the 1/32 px coordinate quantisation and its +16/1024 rounding offset
                // (fp32: positions below 32768 px are good to 0.01 px, far inside that slack)
                const float uc = 0.5f * (float)(u0 + u1), vc = 0.5f * (float)(v0 + v1);
                const float hu = 0.5f * (float)(u1 - u0), hv = 0.5f * (float)(v1 - v0);
                const float xc = m0 * uc + m1 * vc + m2, ex = fabsf(m0) * hu + fabsf(m1) * hv;
                const float yc = m3 * uc + m4 * vc + m5, ey = fabsf(m3) * hu + fabsf(m4) * hv;
                const float xmin = xc - ex, xmax = xc + ex, ymin = yc - ey, ymax = yc + ey;
                const bool full = u1 - u0 == kTileW - 1 && v1 - v0 == T.tile_h - 1;
                // every tap of every pixel of the tile inside the source
                const bool inside = xmin > 0.25f && xmax < src_w - 1.25f && ymin > 0.25f && ymax < src_h - 1.25f;
                // (a tile the frame cuts short is as good as full for this purpose: the pixels that exist are covered)
                hides = (L.flags & kLayerOccluder) && inside && u0 == max(tx0, 0) - bx && v0 == max(ty0, 0) - by &&
                        u1 == min(tx0 + kTileW, T.W) - 1 - bx && v1 == min(ty0 + T.tile_h, T.H) - 1 - by;
                // No tap inside the source: every sample is transparent.  A no-op for PIL's over and on an
                // opaque frame; numpy's over still zeroes rgb where the frame's alpha is 0 (imgproc.py:164).
                bool nothing = xmax < -1.25f || xmin > src_w + 0.25f || ymax < -1.25f || ymin > src_h + 0.25f;
                // ... or every tap falls into 32 x 32 blocks of the source whose alpha is 0 throughout (the
                // layer's alpha map, if it has one): the same "every sample is transparent"
                if (!nothing && L.amap) {
                    const int cx0 = max((int)floorf(xmin) - 1, 0) >> 5, cx1 = min((int)floorf(xmax) + 2, src_w - 1) >> 5;
                    const int cy0 = max((int)floorf(ymin) - 1, 0) >> 5, cy1 = min((int)floorf(ymax) + 2, src_h - 1) >> 5;
                    if (cx0 <= cx1 && cy0 <= cy1 && (cx1 - cx0 + 1) * (cy1 - cy0 + 1) <= 16) {
                        const int gw = (src_w + 31) >> 5;
                        uint32_t seen = 0;
                        for (int cy = cy0; cy <= cy1; ++cy)
                            for (int cx = cx0; cx <= cx1; ++cx) seen |= __ldg(L.amap + cy * gw + cx);
                        nothing = seen == 0;
                    }
                }
                if (!(nothing && (T.opaque || (L.flags & kLayerPil)))) {
                    cls = kTileEdge;
                    // taps span [floor(xmin) - 1, floor(xmax) + 2] at most (same slack as above); TMA wants
                    // the box to start on a 16-byte boundary of the row: x origin a multiple of 4 px
                    const int fx0 = ((int)floorf(xmin) - 1) & ~3, fy0 = (int)floorf(ymin) - 1;
                    if (box_w > 0 && (int)floorf(xmax) + 3 - fx0 <= box_w && (int)floorf(ymax) + 3 - fy0 <= box_h) {
                        cls = full ? kTileTma : kTileTmaPart;
                        org = make_int2(fx0, fy0);
                    } else if (full && inside) {
                        cls = kTileInterior;   // whole tile covered and every tap inside the source
                    }
                }
            }
        }
        // An occluder (NORMAL, opacity 1, source alpha 255 everywhere) whose taps all lie inside its source
        // samples alpha 255, and both "over" operators then return the sample whatever is underneath
        // (numpy: (65025 T + 0) // 65025, alpha 65025 // 255; PIL: coef1 = 32640, coef2 = 0): the layers
        // below the topmost one are dropped from the tile's list.  Bit-exact by construction.
        unsigned active = __ballot_sync(0xFFFFFFFFu, cls >= 0);
        const unsigned hiding = __ballot_sync(0xFFFFFFFFu, cls >= 0 && hides);
        if (hiding) {
            const int top = 31 - __clz((int)hiding);
            active &= ~((1u << top) - 1u);
            if (lane < top) cls = -1;
        }
        int2* rec = T.tiles + ((size_t)f * (size_t)T.n_tiles + (size_t)t) * (size_t)T.tile_stride;
        if (lane == 0) rec[0] = make_int2(__popc(active) | (f << 8), tx0 | (ty0 << 16));   // count | frame << 8, tile origin
        if (active == 0u && !T.keep_frame && tx0 + lane < T.W) {
            uint8_t* p = F.frame + (size_t)ty0 * (size_t)T.stride + (size_t)(tx0 + lane) * 4;
            const int rows = min(T.tile_h, T.H - ty0);
            for (int r = 0; r < rows; ++r, p += T.stride) *reinterpret_cast<uint32_t*>(p) = T.bg;
        }
        if (cls >= 0)
            rec[1 + __popc(active & ((1u << lane) - 1u))] = make_int2(lane | (cls << 8), (org.x & 0xFFFF) | (org.y << 16));
        return;
    }
    const int part = blockIdx.x % T.tab_blocks, layer = blockIdx.x / T.tab_blocks;
    DevLayer& L = T.layers[layer];
    if (part == T.tab_blocks - 1) {
        __shared__ uint32_t s_need[8];
        const uint32_t a = threadIdx.x;
        uint32_t k = a;
        if (L.opacity < 1.0)
            k = (L.flags & kLayerPil) ? (a * (uint32_t)L.pil_a) / 255u                  // imgproc.py:206-208
                                      : __double2uint_rz(__dmul_rn((double)a, L.opacity));   // imgproc.py:159
        uint32_t need = a ? ((k << 18) + a - 1u) / a : 0u;
        need = __reduce_max_sync(0xFFFFFFFFu, need);
        if ((threadIdx.x & 31) == 0) s_need[threadIdx.x >> 5] = need;
        __syncthreads();
        uint32_t m = s_need[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) m = max(m, s_need[i]);
        const bool fits = m <= (1u << 18) &&
                          (a ? (unsigned long long)m * a < ((unsigned long long)(k + 1u) << 18) : k == 0u);
        const int all = __syncthreads_and(fits);
        if (a == 0) {
            L.opac_c = all ? (float)m * (1.0f / 262144.0f) : 0.0f;
            if (!all) L.flags |= kLayerLut;
        }
        if (!all) L.lut[a] = (float)k;
        return;
    }
    const int i = part * 256 + threadIdx.x;
    int2* out = T.tab + (size_t)layer * (size_t)(T.tab_w + T.tab_h);
    if (i < T.tab_w) {
        out[i] = make_int2(warp_adelta(L.m, i - L.bx), warp_bdelta(L.m, i - L.bx));
    } else if (i < T.tab_w + T.tab_h) {
        const int v = i - T.tab_w - L.by;
        out[i] = make_int2(warp_x0(L.m, v), warp_y0(L.m, v));
    }
}

struct FramesParams {
    const FrameRec* frames;
    int32_t n_frames;
    int32_t W, H, stride;      // every frame of a run has the same geometry
    uint32_t bg;
    int32_t keep_frame;        // start from the pixels in the frames (continuing a stack of more than kMaxLayers layers)
    int32_t tab_w, tab_h;
    int32_t stage_bytes;       // bytes of one shared-memory stage (the largest box of this run)
    int32_t stages;            // 2..kMaxStages stages in flight
    int32_t tiles_x, n_tiles;  // per frame
    int32_t tile_stride;       // int2 entries per tile record
    int32_t force_units;       // diagnostics: walk background-only tiles too
    int32_t pad_;
    float* next_item;          // work counter of the run (starts at 0; items below gridDim.x are taken statically)
    const int2* tiles;         // k_stack_prep's per (frame, tile) lists, frame-major
    alignas(64) CUtensorMap maps[kMaxMaps];   // distinct (source, box) descriptors of the run; DevLayer::map_index
};

// OPAQUE: the frames' alpha is 255 throughout (opaque bg_color); otherwise alpha is carried as a
// fourth channel and both "over" operators use their general forms (over_numpy_rows / over_pil_rows).
// ROWS: rows per thread (4 or 8); the tile is 32 x (8 * ROWS).
template <bool OPAQUE, int ROWS>
#ifndef MVB_MINBLOCKS
#define MVB_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(kWsThreads, ROWS == 4 ? MVB_MINBLOCKS : 2)
k_stack_frames(const __grid_constant__ FramesParams P)
{
    constexpr int QUADS = ROWS / 4, PAIRS = ROWS / 2, TILE_H = kWsConsumers * ROWS;
    extern __shared__ __align__(1024) uint8_t s_stage[];   // 2, 4 or 8 stages of P.stage_bytes
    __shared__ UnitMem<ROWS> s_unit[kWsUnits];
    __shared__ float2 s_softg[256];
    __shared__ int32_t s_cull[kMaxLayers][kCullPitch];     // the current frame's records, producer only
    __shared__ int4 s_req[kWsUnits * kWsSlots];                           // queued TMA requests: map index, bytes, box origin x, y
    __shared__ __align__(8) unsigned long long s_bar[2 * kMaxStages + 2 * kWsUnits];  // full[], empty[], unit_full[], unit_empty[]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_full = smem_addr(&s_bar[0]), bar_empty = smem_addr(&s_bar[kMaxStages]);
    const uint32_t unit_full = smem_addr(&s_bar[2 * kMaxStages]), unit_empty = smem_addr(&s_bar[2 * kMaxStages + kWsUnits]);
    const uint32_t stage0 = smem_addr(s_stage);
    const int ns = P.stages;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, kWsConsumers);
        }
#pragma unroll
        for (int i = 0; i < kWsUnits; ++i) {
            mbar_init(unit_full + 8 * i, 1);
            mbar_init(unit_empty + 8 * i, kWsConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 256) s_softg[tid] = make_float2(g_softg[2 * tid], g_softg[2 * tid + 1]);   // (kWsThreads >= 256)
    __syncthreads();

    if (warp == kWsConsumers) {
        // =========================== producer warp ===============================================
        int u = 0;   // units published so far
        int g = 0;   // TMA boxes assigned so far (position in the stage ring)
        // Boxes assigned but not yet requested, because their stage was still being read, wait in s_req
        // (request number q in entry q mod 8 kWsUnits: at most kWsUnits units x 8 layers are ever
        // outstanding).  Lane st < stages owns stage st: it requests the boxes q = st (mod stages) in order, as soon as the producer passes by and the
        // stage has come free -- the lanes work side by side, so one pass costs one barrier probe + one
        // TMA issue however many stages came free.
        int my_head = lane;   // lanes < stages: next request number for the lane's stage (lane + my_k * stages)
        int my_k = 0;         // how many boxes the lane has requested: the stage's phase count
        int g_stage = 0, g_round = 0;   // stage and use count of request number g
        asm volatile("griddepcontrol.wait;" ::: "memory");   // k_stack_prep's results are complete and visible
        auto service = [&]() {
            if (lane < ns) {
                while (my_head < g) {
                    if (my_k > 0 && !mbar_test(bar_empty + 8 * lane, (my_k - 1) & 1)) break;
                    const int4 r = s_req[my_head & (kWsUnits * kWsSlots - 1)];
                    mbar_arrive_expect_tx(bar_full + 8 * lane, (uint32_t)r.y);
                    tma_load_box(stage0 + lane * P.stage_bytes, &P.maps[r.x], r.z, r.w, bar_full + 8 * lane);
                    my_head += ns;
                    ++my_k;
                }
            }
            __syncwarp();
        };
        // Waits for the stage of the oldest request that is still queued; false: nothing is queued.
        auto wait_for_oldest = [&]() {
            const bool pend = lane < ns && my_head < g;
            const int head = __reduce_min_sync(0xFFFFFFFFu, pend ? my_head : 0x7FFFFFFF);
            if (head == 0x7FFFFFFF) return false;
            const int owner = __ffs((int)__ballot_sync(0xFFFFFFFFu, pend && my_head == head)) - 1;
            const int k = __shfl_sync(0xFFFFFFFFu, my_k, owner);   // > 0: a first use never waits
            mbar_wait(bar_empty + 8 * owner, (k - 1) & 1);
            __syncwarp();
            return true;
        };
        // Items are handed out through a counter in the run's scratch (zeroed by the upload): a static
        // round-robin pins a CTA to the same few tile columns of every frame (its stride and the tiles
        // per row share factors), and the cost of a column -- how many layers cross it -- differs, so the
        // imbalance would never average out over the run.  The counter runs TWO items ahead: the atomic
        // issued while tile k is set up is only looked at when tile k + 1 is, so its round trip (about a
        // microsecond, a third of a two-layer tile) is never waited for.
        const int n_items = P.n_frames * P.n_tiles;   // (the host keeps it, plus the grid size, below 2^24)
        int w = blockIdx.x;                // next (frame, tile) item of this CTA
        // (The counter is a FLOAT: an integer atom.add in a one-lane branch is compiled to vote + atomic +
        // shuffle, and the shuffle waits for the result on the spot.  Counts stay below 2^24, so they are exact.)
        float raw = 0.0f;                  // lane 0: the counter value behind the item after that
        auto take = [&]() {
            if (lane == 0) asm volatile("atom.global.add.f32 %0, [%1], 0f3F800000;" : "=f"(raw) : "l"(P.next_item) : "memory");
        };
        auto taken = [&]() {
            const unsigned int v = (unsigned int)__shfl_sync(0xFFFFFFFFu, raw, 0) + gridDim.x;
            return v < (unsigned int)n_items ? (int)v : n_items;
        };
        take();
        // its list, fetched one tile ahead of its use: header (count, frame, origin) and this lane's entry
        int2 nx_hdr = make_int2(0, 0), nx_e = make_int2(0, 0);
        auto fetch = [&](int item) {
            if (item < n_items) {
                const int2* rec = P.tiles + (size_t)item * (size_t)P.tile_stride;
                nx_hdr = __ldg(rec);
                if (lane + 1 < P.tile_stride) nx_e = __ldg(rec + 1 + lane);
            }
        };
        fetch(w);
        int cur_f = -1;                    // frame whose records are in s_cull
        uint8_t* fr_frame = nullptr;
        const DevLayer* fr_layers = nullptr;
        const int2* fr_tab = nullptr;
        bool have_tile = false, stop_sent = false, first = true;
        int n_act = 0, base = 0;           // layers of the planned tile, and how many of them are in units already
        int tx0 = 0, ty0 = 0;
        int my_li = 0, my_c = 0;           // lane s < n_act: s-th layer (paint order) touching the tile, its class
        int2 my_org = make_int2(0, 0);     // and its TMA box origin

        for (;;) {
            service();
            if (!have_tile && w < n_items) {
                // ---- the next tile --------------------------------------------------------------------
                const int f = nx_hdr.x >> 8;
                if (f != cur_f) {   // entering a new frame: its record and the leading ints of its layers
                    const FrameRec& F = P.frames[f];
                    fr_frame = F.frame; fr_layers = F.layers; fr_tab = F.tab;
                    const int fr_n = F.n_layers;
                    for (int i = lane; i < fr_n * kCullInts; i += 32) {
                        const int li = i / kCullInts, j = i - li * kCullInts;
                        s_cull[li][j] = reinterpret_cast<const int32_t*>(&fr_layers[li])[j];
                    }
                    cur_f = f;
                    __syncwarp();
                }
                tx0 = nx_hdr.y & 0xFFFF;
                ty0 = (int)((unsigned)nx_hdr.y >> 16);
                n_act = nx_hdr.x & 255;
                my_li = nx_e.x & 255;
                my_c = (nx_e.x >> 8) & 7;
                my_org = make_int2((int)(short)(nx_e.y & 0xFFFF), nx_e.y >> 16);
                if (lane >= n_act) { my_li = 0; my_c = kTileEdge; }
                // a layer whose opacity needs its table travels through the classes that consult it
                if (s_cull[my_li][15] & kLayerLut)
                    my_c = my_c == kTileTma ? kTileTmaLut : my_c == kTileTmaPart ? kTileTmaPartLut : kTileEdge;
                base = 0;
                first = true;
                have_tile = n_act > 0 || P.force_units;   // a tile no layer touches was finished by k_stack_prep
                w = taken();         // the list of the item after this one is on its way while this one is built,
                fetch(w);            // and so is the number of the one after that
                if (w < n_items) take();
                if (!have_tile) continue;
            }
            if (!have_tile && stop_sent) {
                if (!wait_for_oldest()) break;   // everything requested: done (else the consumers still wait for boxes)
                continue;
            }
            // ---- the next unit (or the stop mark) needs a free unit buffer ---------------------------------
            const int ub = u & (kWsUnits - 1);
            const uint32_t ubar_empty = unit_empty + 8 * ub, ubar_par = ((u / kWsUnits) - 1) & 1;
            int free_ = 1;
            if (u >= kWsUnits) {
                if (lane == 0) free_ = mbar_test(ubar_empty, ubar_par);
                free_ = __shfl_sync(0xFFFFFFFFu, free_, 0);
            }
            if (!free_) {
                // Nothing can proceed without an event: wait for the one needed first.  A queued box gates the
                // consumers sooner than a unit that is four units ahead of them.
                if (!wait_for_oldest()) {
                    mbar_wait(ubar_empty, ubar_par);
                    __syncwarp();
                }
                continue;
            }
            UnitMem<ROWS>& U = s_unit[ub];
            const uint32_t ubar = unit_full + 8 * ub;
            if (!have_tile) {   // all items planned: the stop mark
                if (lane == 0) {
                    U.hdr = make_int4(0, 0, 1 << 10, 0);
                    mbar_arrive(ubar);
                }
                ++u;
                stop_sent = true;
                continue;
            }
            // ---- one unit: the next (up to) 8 layers of the tile's list; a tile without layers still gets
            //      one (background only).  Lane base + s owns slot s. ---------------------------------------
            const int n_here = min(kWsSlots, n_act - base);
            const bool mine = lane >= base && lane < base + n_here;
            const unsigned want = __ballot_sync(0xFFFFFFFFu, mine && (my_c & 1));
            const int my_off = __popc(want & ((1u << lane) - 1u));
            const int my_g = g + my_off;   // request number of this lane's box, its stage and that stage's use count
            int my_st = g_stage + my_off, my_rd = g_round;
            while (my_st >= ns) { my_st -= ns; ++my_rd; }
            const int32_t* C = s_cull[my_li];
            if (mine && (my_c & 1))
                s_req[my_g & (kWsUnits * kWsSlots - 1)] = make_int4(C[18], C[6] * C[7] * 4, my_org.x, my_org.y);
            g += __popc(want);
            g_stage += __popc(want);
            while (g_stage >= ns) { g_stage -= ns; ++g_round; }
            __syncwarp();
            service();
            // the info word, and two bulk copies of the layer's tables restricted to the tile (columns
            // tx0.., rows ty0..), which complete on the unit's barrier
            const bool last = base + n_here >= n_act;
            if (lane == 0) {
                U.hdr = make_int4(tx0, ty0, n_here | (first ? 256 : 0) | (last ? 512 : 0), 0);
                U.frame = fr_frame;
                U.layers = fr_layers;
            }
            if (mine) {
                const int pitch = C[6] * 4;   // box_w px
                UnitSlot<ROWS>& S = U.slot[lane - base];
                const int one_tap = (my_c == kTileTma && (C[15] & kLayerIntShift) && (C[14] == 0 || C[14] == kPilOver)) ? kInfoOneTap : 0;
                S.info = make_int4(C[14] | ((my_c & 7) << 8) | ((my_rd & 1) << 11) | one_tap | (pitch << 16),
                                   (int)(bar_full + 8 * my_st),
                                   C[16],
                                   (int)(stage0 + my_st * P.stage_bytes - my_org.y * pitch - my_org.x * 4));
                U.layer[lane - base] = my_li;
                const int2* tab = fr_tab + (size_t)my_li * (size_t)(P.tab_w + P.tab_h);
                bulk_load(smem_addr(S.col), tab + tx0, sizeof(int2) * kTileW, ubar);
                bulk_load(smem_addr(S.row), tab + P.tab_w + ty0, sizeof(int2) * TILE_H, ubar);
            }
            __syncwarp();
            if (lane == 0) {   // publish (release: the warp's writes above); the phase ends when the copies have landed
                if (n_here) mbar_arrive_expect_tx(ubar, (uint32_t)n_here * (uint32_t)(sizeof(int2) * (kTileW + TILE_H)));
                else mbar_arrive(ubar);
            }
            ++u;
            first = false;
            base += n_here;
            if (last) have_tile = false;
        }
    } else {
        // =========================== consumer warps ==============================================
        const FrameGeom G{P.W, P.H, P.stride};
        RowPairState d[PAIRS];
        for (int u = 0;; ++u) {
            const UnitMem<ROWS>& U = s_unit[u & (kWsUnits - 1)];
            mbar_wait(unit_full + 8 * (u & (kWsUnits - 1)), (u / kWsUnits) & 1);
            const int4 hdr = U.hdr;
            if (hdr.z & (1 << 10)) break;
            uint8_t* const frame = U.frame;
            const DevLayer* const layers = U.layers;
            const int tx0 = hdr.x, ty0 = hdr.y, n_here = hdr.z & 255;
            const int x = tx0 + lane, rbeg = warp * ROWS;
            const bool x_ok = x < P.W;
            const PixelCtx px{tx0, ty0, x, rbeg, lane, x_ok};
            if (hdr.z & 256) {   // first unit of the tile: background
                if (P.keep_frame) {   // continuing a stack of more than kMaxLayers layers
                    uint32_t v[ROWS];
#pragma unroll
                    for (int k = 0; k < ROWS; ++k)   // clamped, not predicated; out-of-frame threads never store
                        v[k] = reinterpret_cast<const uint32_t*>(frame + (int64_t)min(ty0 + rbeg + k, P.H - 1) * P.stride)[min(x, P.W - 1)];
#pragma unroll
                    for (int p = 0; p < PAIRS; ++p)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            d[p].c[c] = pk(__uint_as_float(__byte_perm(v[2 * p], kMagicBits, 0x7650 + c)),
                                           __uint_as_float(__byte_perm(v[2 * p + 1], kMagicBits, 0x7650 + c)));
                } else {
#pragma unroll
                    for (int p = 0; p < PAIRS; ++p)
#pragma unroll
                        for (int c = 0; c < 4; ++c) d[p].c[c] = bc(kMagic + (float)((P.bg >> (8 * c)) & 0xFFu));
                }
            }
#pragma unroll 1
            for (int slot = 0; slot < n_here; ++slot) {
                // (whole-pixel shifts run their own instantiation, blend code included, so that the general
                // walk is compiled exactly as if they did not exist)
                if (__builtin_expect(!(U.slot[slot].info.x & kInfoOneTap), 1))
                    walk_layer<OPAQUE, QUADS, false>(G, layers, U.slot[slot], &U.layer[slot], px, d, s_softg);
                else
                    walk_layer<OPAQUE, QUADS, true>(G, layers, U.slot[slot], &U.layer[slot], px, d, s_softg);
            }
            mbar_arrive_warp(unit_empty + 8 * (u & (kWsUnits - 1)));   // this warp no longer reads the unit
            if ((hdr.z & 512) && x_ok) {   // last unit of the tile: one coalesced 128-byte store per warp row
#pragma unroll
                for (int k = 0; k < ROWS; ++k) {
                    if (ty0 + rbeg + k >= P.H) break;
                    float r0, r1, g0, g1, b0, b1;
                    upk(d[k >> 1].c[0], r0, r1);
                    upk(d[k >> 1].c[1], g0, g1);
                    upk(d[k >> 1].c[2], b0, b1);
                    const uint32_t rb = __float_as_uint((k & 1) ? r1 : r0), gb = __float_as_uint((k & 1) ? g1 : g0);
                    const uint32_t bb = __float_as_uint((k & 1) ? b1 : b0);
                    uint32_t pxl = __byte_perm(__byte_perm(rb, gb, 0x0040), bb, 0x0410);
                    if constexpr (OPAQUE) {
                        pxl |= 0xFF000000u;
                    } else {
                        float al0, al1;
                        upk(d[k >> 1].c[3], al0, al1);
                        pxl = __byte_perm(pxl, __float_as_uint((k & 1) ? al1 : al0), 0x4210);
                    }
                    reinterpret_cast<uint32_t*>(frame + (int64_t)(ty0 + rbeg + k) * P.stride)[x] = pxl;
                }
            }
        }
    }
}

} // namespace mvb
