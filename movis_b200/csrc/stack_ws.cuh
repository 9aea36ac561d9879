// stack_ws.cuh -- persistent, warp-specialised form of the opaque-frame stack kernel (sm_100a).
//
// Same arithmetic as k_stack_opaque (stack_float.cuh: walk_layer) and the same contract: the frame
// walk of Composition.__call__ (movis/layer/composition.py:363-368) with cv2.warpAffine ->
// alpha_composite per layer (composition.py:811-819), bit for bit.  What changes is who does the
// per-tile bookkeeping.  In k_stack_opaque all eight warps of a tile stop at two barriers while
// warp 0 culls the layer list and the OpenCV tables are built; here every CTA is persistent
// (grid = a few CTAs per SM, tiles taken round-robin) and has a ninth, PRODUCER warp that runs one
// or two tiles ahead of the eight CONSUMER warps:
//
//   producer   cull + classify the layers against the next tile; for up to 8 layers ("unit") write the
//              info words and bulk-copy the tile's slice of cv2's integer tables (built for the whole
//              frame by k_stack_tables) into one of four shared-memory unit buffers, whose mbarrier
//              publishes the unit when the copies have landed; request the layers' source boxes
//              through TMA into a ring of stages as they become free;
//   consumers  wait for the unit, walk its layers (taps from the TMA stages, float arithmetic,
//              pixels in registers), release the unit, store the tile after its last unit.
//
// There is no __syncthreads() after start-up: all hand-offs are mbarrier phases, and the waits
// are bounded (a wait that cannot complete traps instead of hanging the device).
#pragma once

#include "stack_float.cuh"

namespace mvb {

constexpr int kWsConsumers = kFWarps;             // 8 warps x 4 rows = one 32 x 32 tile
constexpr int kWsThreads = (kWsConsumers + 1) * 32;
constexpr int kWsSlots = 8;                       // layers per unit
constexpr int kWsUnits = 4;                       // unit buffers: how far the producer may run ahead (power of two)

// Per layer of a unit: as SlotMem, without the alpha LUT (one per layer per CTA here, see s_lut).
struct __align__(16) UnitSlot {
    int4 info;
    int2 col[kTileW];
    int2 row[kFTileH];
};

struct __align__(16) UnitMem {
    int4 hdr;                 // x: tile x0, y: tile y0, z: layers in this unit | first << 8 | last << 9 | stop << 10
    int layer[kWsSlots];      // layer index of each slot (for the layers that gather through L1)
    UnitSlot slot[kWsSlots];
};

// cv::warpAffine's integer tables (imgwarp.cpp: adelta / bdelta by destination column, X0 / Y0 by
// destination row; fp64 without FMA contraction, see pixel_math.cuh) of every layer of a launch for the
// whole frame.  They depend on the layer and the frame column / row only, so they are built once per
// launch instead of once per tile; the producer warp of k_stack_ws copies a tile's 2 x 256 bytes per
// layer into shared memory with cp.async.bulk.
__global__ void __launch_bounds__(256)
k_stack_tables(const __grid_constant__ TableParams T)
{
    asm volatile("griddepcontrol.launch_dependents;");   // k_stack_ws may start its prologue now
    const int i = blockIdx.x * 256 + threadIdx.x;
    const TableLayer& L = T.l[blockIdx.y];
    int2* out = T.tab + (size_t)blockIdx.y * (size_t)(T.tab_w + T.tab_h);
    if (i < T.tab_w) {
        out[i] = make_int2(warp_adelta(L.m, i - L.bx), warp_bdelta(L.m, i - L.bx));
    } else if (i < T.tab_w + T.tab_h) {
        const int v = i - T.tab_w - L.by;
        out[i] = make_int2(warp_x0(L.m, v), warp_y0(L.m, v));
    }
}

// 1-D bulk copy global -> shared memory, completion signalled on `bar` in bytes (16-byte granular).
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// OPAQUE: the frame's alpha is 255 throughout (opaque bg_color); otherwise alpha is carried as a
// fourth channel and both "over" operators use their general forms (over_numpy_rows / over_pil_rows).
template <bool OPAQUE>
__global__ void __launch_bounds__(kWsThreads, 3)
k_stack_ws(const __grid_constant__ StackParamsTma PT)
{
    const StackParams& P = PT.s;
    extern __shared__ __align__(1024) uint8_t s_stage[];   // 2 or 4 stages of PT.stage_bytes, then the alpha LUTs
    __shared__ UnitMem s_unit[kWsUnits];
    __shared__ float2 s_softg[256];
    __shared__ int32_t s_cull[kMaxLayers][17];             // first 64 bytes of every DevLayer (17: conflict-free by lane)
    __shared__ __align__(8) unsigned long long s_bar[2 * kMaxStages + 2 * kWsUnits];  // full[], empty[], unit_full[], unit_empty[]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t bar_full = smem_addr(&s_bar[0]), bar_empty = smem_addr(&s_bar[kMaxStages]);
    const uint32_t unit_full = smem_addr(&s_bar[2 * kMaxStages]), unit_empty = smem_addr(&s_bar[2 * kMaxStages + kWsUnits]);
    const uint32_t stage0 = smem_addr(s_stage);
    const int ns_log2 = PT.stages_log2, ns_mask = (1 << ns_log2) - 1;
    // opacity-scaled alpha as float, 256 entries per layer: the same for every tile, so built once per CTA
    float* const s_lut = reinterpret_cast<float*>(s_stage + ((size_t)PT.stage_bytes << ns_log2));
    const int tiles_x = (P.W + kTileW - 1) / kTileW, tiles_y = (P.H + kFTileH - 1) / kFTileH;
    const int n_tiles = tiles_x * tiles_y;

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
    if (tid < 256) s_softg[tid] = g_blend_tables.softg[tid];
    // a constant-bank read with a different address in every lane is replayed per address: the producer's
    // one-layer-per-lane cull reads its records from shared memory instead
    for (int i = tid; i < P.n_layers * 16; i += kWsThreads)
        s_cull[i >> 4][i & 15] = reinterpret_cast<const int32_t*>(&P.layers[i >> 4])[i & 15];
    for (int i = tid; i < P.n_layers * 256; i += kWsThreads) {
        const DevLayer& L = P.layers[i >> 8];
        const int a = i & 255;
        uint32_t v = (uint32_t)a;
        if (L.opacity < 1.0)
            v = L.pil ? (uint32_t)(a * L.pil_a) / 255u                          // imgproc.py:206-208
                      : __double2uint_rz(__dmul_rn((double)a, L.opacity));      // imgproc.py:159
        s_lut[i] = (float)v;
    }
    __syncthreads();

    if (warp == kWsConsumers) {
        // =========================== producer warp ===============================================
        int u = 0;   // units published so far
        int g = 0;   // TMA boxes assigned so far (position in the stage ring)
        // Boxes assigned but not yet requested, because their stage was still being read: request number
        // q waits in lane q & 31 (at most 4 units x 8 layers are ever outstanding).  They are requested
        // in order, whenever the producer passes by and the stage has come free -- never blocking the
        // building of the next unit, which is what the consumers run out of first.
        int head = 0;   // next request number to issue
        int q_li = 0, q_ox = 0, q_oy = 0;
        asm volatile("griddepcontrol.wait;" ::: "memory");   // the tables of k_stack_tables are complete and visible
        auto service = [&]() {
            while (head < g) {
                const int st = head & ns_mask;
                int go = 1;
                if (lane == 0 && head > ns_mask) go = mbar_test(bar_empty + 8 * st, ((head >> ns_log2) - 1) & 1);
                go = __shfl_sync(0xFFFFFFFFu, go, 0);
                if (!go) break;
                const int li = __shfl_sync(0xFFFFFFFFu, q_li, head & 31);
                const int ox = __shfl_sync(0xFFFFFFFFu, q_ox, head & 31), oy = __shfl_sync(0xFFFFFFFFu, q_oy, head & 31);
                if (lane == 0) {
                    const DevLayer& L = P.layers[li];
                    mbar_arrive_expect_tx(bar_full + 8 * st, (uint32_t)(L.box_w * L.box_h * 4));
                    tma_load_box(stage0 + st * PT.stage_bytes, &PT.maps[li], ox, oy, bar_full + 8 * st);
                }
                __syncwarp();
                ++head;
            }
        };
        // Wait for a phase while keeping the request queue moving (the consumers may need those boxes
        // to get to the arrival this wait is for).
        auto wait_servicing = [&](uint32_t bar, uint32_t parity) {
            for (uint32_t spins = 0;; ++spins) {
                int done = 0;
                if (lane == 0) done = mbar_test(bar, parity);
                if (__shfl_sync(0xFFFFFFFFu, done, 0)) return;
                service();
                if (spins > (1u << 24)) __trap();
            }
        };
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int tx0 = (t % tiles_x) * kTileW, ty0 = (t / tiles_x) * kFTileH;
            // ---- cull + classify: one layer per lane (as in k_stack_opaque) -------------------------
            int cls = -1;
            int2 org = make_int2(0, 0);
            // (tiles outside the union of the layers' bboxes take the short way: background only)
            const bool covered = tx0 < PT.ux1 && tx0 + kTileW > PT.ux0 && ty0 < PT.uy1 && ty0 + kFTileH > PT.uy0;
            if (covered && lane < P.n_layers) {
                const int32_t* C = s_cull[lane];
                const int bx = C[0], by = C[1], bw = C[2], bh = C[3], src_w = C[4], src_h = C[5], box_w = C[6], box_h = C[7];
                const float m0 = __int_as_float(C[8]), m1 = __int_as_float(C[9]), m2 = __int_as_float(C[10]);
                const float m3 = __int_as_float(C[11]), m4 = __int_as_float(C[12]), m5 = __int_as_float(C[13]);
                const int u0 = max(tx0, bx) - bx, u1 = min(min(tx0 + kTileW, bx + bw), P.W) - 1 - bx;
                const int v0 = max(ty0, by) - by, v1 = min(min(ty0 + kFTileH, by + bh), P.H) - 1 - by;
                if (bw > 0 && bh > 0 && u0 <= u1 && v0 <= v1) {
                    const float uc = 0.5f * (float)(u0 + u1), vc = 0.5f * (float)(v0 + v1);
                    const float hu = 0.5f * (float)(u1 - u0), hv = 0.5f * (float)(v1 - v0);
                    const float xc = m0 * uc + m1 * vc + m2, ex = fabsf(m0) * hu + fabsf(m1) * hv;
                    const float yc = m3 * uc + m4 * vc + m5, ey = fabsf(m3) * hu + fabsf(m4) * hv;
                    const float xmin = xc - ex, xmax = xc + ex, ymin = yc - ey, ymax = yc + ey;
                    const bool full = u1 - u0 == kTileW - 1 && v1 - v0 == kFTileH - 1;
                    // No tap inside the source: every sample is transparent.  A no-op for PIL's over and on an
                    // opaque frame; numpy's over still zeroes rgb where the frame's alpha is 0 (imgproc.py:164).
                    const bool nothing = xmax < -1.25f || xmin > src_w + 0.25f || ymax < -1.25f || ymin > src_h + 0.25f;
                    if (!(nothing && (OPAQUE || C[15] != 0))) {
                        cls = kTileEdge;
                        const int fx0 = ((int)floorf(xmin) - 1) & ~3, fy0 = (int)floorf(ymin) - 1;
                        if (box_w > 0 && (int)floorf(xmax) + 3 - fx0 <= box_w && (int)floorf(ymax) + 3 - fy0 <= box_h) {
                            cls = full ? kTileTma : kTileTmaPart;
                            org = make_int2(fx0, fy0);
                        } else if (full && xmin > 0.25f && xmax < src_w - 1.25f && ymin > 0.25f && ymax < src_h - 1.25f) {
                            cls = kTileInterior;
                        }
                    }
                }
            }
            unsigned rest = __ballot_sync(0xFFFFFFFFu, cls >= 0);   // active layers, paint order = bit order
            bool first = true;
            do {   // one unit per 8 active layers; a tile without layers still gets one (background only)
                const int n_here = min(kWsSlots, __popc(rest));
                // lane s < n_here owns slot s of the unit: its layer, class, box origin and ring position
                int my_li = 0;
                if (lane < n_here) my_li = nth_set_bit(rest, lane);
                const int my_c = __shfl_sync(0xFFFFFFFFu, cls, my_li);
                const int2 my_org = make_int2(__shfl_sync(0xFFFFFFFFu, org.x, my_li), __shfl_sync(0xFFFFFFFFu, org.y, my_li));
                const unsigned want = __ballot_sync(0xFFFFFFFFu, lane < n_here && (my_c & 3) == kTileTma);
                const int my_g = g + __popc(want & ((1u << lane) - 1u));
                for (int i = 0; i < n_here; ++i) rest &= rest - 1;

                UnitMem& U = s_unit[u & (kWsUnits - 1)];
                if (u >= kWsUnits)   // consumers are done with this buffer
                    wait_servicing(unit_empty + 8 * (u & (kWsUnits - 1)), ((u / kWsUnits) - 1) & 1);
                // queue the unit's TMA requests (slot order = ring order)
                for (unsigned w = want; w; w &= w - 1) {
                    const int sl = __ffs((int)w) - 1;
                    const int gq = __shfl_sync(0xFFFFFFFFu, my_g, sl), li = __shfl_sync(0xFFFFFFFFu, my_li, sl);
                    const int ox = __shfl_sync(0xFFFFFFFFu, my_org.x, sl), oy = __shfl_sync(0xFFFFFFFFu, my_org.y, sl);
                    if (lane == (gq & 31)) { q_li = li; q_ox = ox; q_oy = oy; }
                }
                g += __popc(want);
                service();
                // lane s < n_here fills slot s: the info word, and two bulk copies of the layer's tables
                // restricted to the tile (columns tx0.., rows ty0..), which complete on the unit's barrier
                const uint32_t ubar = unit_full + 8 * (u & (kWsUnits - 1));
                if (lane == 0) U.hdr = make_int4(tx0, ty0, n_here | (first ? 256 : 0) | (rest == 0 ? 512 : 0), 0);
                if (lane < n_here) {
                    const int32_t* C = s_cull[my_li];
                    const int pitch = C[6] * 4;   // box_w px
                    UnitSlot& S = U.slot[lane];
                    // x: mode (8 bits) | class (3) | phase parity (bit 11) | row pitch << 16
                    // y: address of the stage's "full" barrier ("empty" is 8 * kMaxStages bytes on)   z: &lut[-kMagicBits]
                    // w: address the stage would have at source pixel (0, 0): stage - origin
                    S.info = make_int4((C[15] ? kPilOver : C[14]) | ((my_c & 7) << 8) |
                                           (((my_g >> ns_log2) & 1) << 11) | (pitch << 16),
                                       (int)(bar_full + 8 * (my_g & ns_mask)),
                                       (int)(smem_addr(s_lut + my_li * 256) - (kMagicBits << 2)),
                                       (int)(stage0 + (my_g & ns_mask) * PT.stage_bytes - my_org.y * pitch - my_org.x * 4));
                    U.layer[lane] = my_li;
                    const int2* tab = PT.tab + (size_t)my_li * (size_t)(PT.tab_w + PT.tab_h);
                    bulk_load(smem_addr(S.col), tab + tx0, sizeof(int2) * kTileW, ubar);
                    bulk_load(smem_addr(S.row), tab + PT.tab_w + ty0, sizeof(int2) * kFTileH, ubar);
                }
                __syncwarp();
                if (lane == 0) {   // publish (release: the warp's writes above); the phase ends when the copies have landed
                    if (n_here) mbar_arrive_expect_tx(ubar, (uint32_t)n_here * (uint32_t)(sizeof(int2) * (kTileW + kFTileH)));
                    else mbar_arrive(ubar);
                }
                service();
                ++u;
                first = false;
            } while (rest);
        }
        // ---- stop unit --------------------------------------------------------------------------------
        if (u >= kWsUnits) wait_servicing(unit_empty + 8 * (u & (kWsUnits - 1)), ((u / kWsUnits) - 1) & 1);
        if (lane == 0) {
            s_unit[u & (kWsUnits - 1)].hdr = make_int4(0, 0, 1 << 10, 0);
            mbar_arrive(unit_full + 8 * (u & (kWsUnits - 1)));
        }
        for (uint32_t spins = 0; head < g; ++spins) {   // the boxes the consumers are still waiting for
            service();
            if (spins > (1u << 24)) __trap();
        }
    } else {
        // =========================== consumer warps ==============================================
        RowPairState d[2];
        for (int u = 0;; ++u) {
            const UnitMem& U = s_unit[u & (kWsUnits - 1)];
            mbar_wait(unit_full + 8 * (u & (kWsUnits - 1)), (u / kWsUnits) & 1);
            const int4 hdr = U.hdr;
            if (hdr.z & (1 << 10)) break;
            const int tx0 = hdr.x, ty0 = hdr.y, n_here = hdr.z & 255;
            const int x = tx0 + lane, rbeg = warp * kFRows;
            const bool x_ok = x < P.W;
            const PixelCtx px{tx0, ty0, x, rbeg, lane, x_ok};
            if (hdr.z & 256) {   // first unit of the tile: background
                if (P.keep_frame) {   // continuing a stack of more than kMaxLayers layers: alpha is 255 already
                    uint32_t v[kFRows];
#pragma unroll
                    for (int k = 0; k < kFRows; ++k)   // clamped, not predicated; out-of-frame threads never store
                        v[k] = reinterpret_cast<const uint32_t*>(P.frame + (int64_t)min(ty0 + rbeg + k, P.H - 1) * P.stride)[min(x, P.W - 1)];
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            d[p].c[c] = pk(__uint_as_float(__byte_perm(v[2 * p], kMagicBits, 0x7650 + c)),
                                           __uint_as_float(__byte_perm(v[2 * p + 1], kMagicBits, 0x7650 + c)));
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) d[0].c[c] = d[1].c[c] = bc(kMagic + (float)((P.bg >> (8 * c)) & 0xFFu));
                }
            }
#pragma unroll 1
            for (int slot = 0; slot < n_here; ++slot)
                walk_layer<OPAQUE>(P, U.slot[slot], &U.layer[slot], px, d, s_softg, [](const int4&) {});
            mbar_arrive_warp(unit_empty + 8 * (u & (kWsUnits - 1)));   // this warp no longer reads the unit
            if ((hdr.z & 512) && x_ok) {   // last unit of the tile: one coalesced 128-byte store per warp row
#pragma unroll
                for (int k = 0; k < kFRows; ++k) {
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
                    reinterpret_cast<uint32_t*>(P.frame + (int64_t)(ty0 + rbeg + k) * P.stride)[x] = pxl;
                }
            }
        }
    }
}

} // namespace mvb
