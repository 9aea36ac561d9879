// Experiment: issue rates of the integer / float / conversion instructions the compositor could be
// built from, measured per SM on the GPU this runs on (thread-ops per SM clock, 32 warps per SM,
// 8 independent chains per thread), plus mixes that show which instructions share a pipe, plus the
// texture unit's point-fetch rate against plain LDG for a rotated 2-D gather.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kChains = 8;
constexpr int kThreads = 1024;

enum Op {
    IADD3, LOP3, SHF, PRMT, IMAD, IMADHI, IMADWIDE, DP4A, DP2A, FFMA, FFMA2, HFMA2, I2F, F2I, I2F16, RCP,
    VMAX16X2, VIADDMAX16X2, VMAX32, SETPSEL, LDS32, LDS64, LDSU8, POPC, IMADSHL, FMUL, FADD, IABS, LEA,
    MIX_IMAD_LOP3, MIX_IMAD_PRMT, MIX_FFMA_LOP3, MIX_DP_LOP3, MIX_DP_IMAD, MIX_FFMA_IMAD, MIX_I2F_IMAD_LOP3,
    MIX_VMAX_LOP3, MIX_VMAX_IMAD, MIX_SHF_LOP3, MIX_LDS_IMAD_LOP3, MIX_RCP_IMAD_LOP3, MIX_IMADHI_LOP3, N_OPS
};
const char* kNames[] = {
    "IADD3", "LOP3", "SHF", "PRMT", "IMAD", "IMAD.HI", "IMAD.WIDE", "IDP.4A", "IDP.2A", "FFMA", "FFMA2(x2 flops)", "HFMA2", "I2F.U32", "F2I", "I2F.U16", "MUFU.RCP",
    "VIMNMX.U16x2", "VIADDMNMX.S16x2", "IMNMX.U32", "ISETP+SEL (pairs)", "LDS.32", "LDS.64", "LDS.U8", "POPC", "IMAD.SHL(mul pow2)", "FMUL", "FADD", "IABS", "LEA(shl+add)",
    "mix IMAD+LOP3", "mix IMAD+PRMT", "mix FFMA+LOP3", "mix IDP4A+LOP3", "mix IDP4A+IMAD", "mix FFMA+IMAD", "mix I2F+IMAD+LOP3(1:2:2)",
    "mix VIMNMX16x2+LOP3", "mix VIMNMX16x2+IMAD", "mix SHF+LOP3", "mix LDS+IMAD+LOP3(1:2:2)", "mix RCP+IMAD+LOP3(1:2:2)", "mix IMAD.HI+LOP3"};

template <int OP>
__device__ __forceinline__ void body(uint32_t (&a)[kChains], uint32_t b, uint32_t c, const uint32_t* sm)
{
#pragma unroll
    for (int j = 0; j < kChains; j++) {
        uint32_t& x = a[j];
        if constexpr (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
        else if constexpr (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2561;" : "+r"(x) : "r"(b));
        else if constexpr (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == IMADWIDE) {
            uint64_t w = ((uint64_t)x << 32) | b;
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(x), "r"(c));
            x = (uint32_t)w ^ (uint32_t)(w >> 32);
        } else if constexpr (OP == DP4A) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == DP2A) asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(x) : "r"(b));
        else if constexpr (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(x) : "r"(b));
        else if constexpr (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == I2F) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(x));
        else if constexpr (OP == F2I) asm volatile("cvt.rzi.u32.f32 %0, %0;" : "+r"(x));
        else if constexpr (OP == I2F16) {
            asm volatile("{ .reg .u16 t; cvt.u16.u32 t, %0; cvt.rn.f32.u16 %0, t; }" : "+r"(x));
        } else if constexpr (OP == RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+r"(x));
        else if constexpr (OP == VMAX16X2) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
        else if constexpr (OP == VIADDMAX16X2) x = __viaddmax_s16x2(x, b, c);
        else if constexpr (OP == VMAX32) asm volatile("max.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
        else if constexpr (OP == SETPSEL)
            asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(x) : "r"(b), "r"(c));
        else if constexpr (OP == LDS32) x = sm[x & 1023];
        else if constexpr (OP == LDS64) { uint2 v = reinterpret_cast<const uint2*>(sm)[x & 511]; x = v.x ^ v.y; }
        else if constexpr (OP == LDSU8) x = reinterpret_cast<const uint8_t*>(sm)[x & 4095];
        else if constexpr (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x));
        else if constexpr (OP == IMADSHL) asm volatile("mul.lo.u32 %0, %0, 8;" : "+r"(x));
        else if constexpr (OP == IABS) asm volatile("abs.s32 %0, %0;" : "+r"(x));
        else if constexpr (OP == LEA) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(x) : "r"(b));
        else if constexpr (OP == MIX_IMAD_LOP3) {
            if (j & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_IMAD_PRMT) {
            if (j & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("prmt.b32 %0, %0, %1, 0x2561;" : "+r"(x) : "r"(b));
        } else if constexpr (OP == MIX_FFMA_LOP3) {
            if (j & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_DP_LOP3) {
            if (j & 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_DP_IMAD) {
            if (j & 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_FFMA_IMAD) {
            if (j & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_I2F_IMAD_LOP3) { // per 8 chains: 2 I2F? keep ratio 1:2:2 over 5 -> use j%5
            if (j % 5 == 0) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(x));
            else if (j % 5 < 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_VMAX_LOP3) {
            if (j & 1) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_VMAX_IMAD) {
            if (j & 1) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(x) : "r"(b));
            else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_SHF_LOP3) {
            if (j & 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_LDS_IMAD_LOP3) {
            if (j % 5 == 0) x = sm[x & 1023];
            else if (j % 5 < 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_RCP_IMAD_LOP3) {
            if (j % 5 == 0) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+r"(x));
            else if (j % 5 < 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        } else if constexpr (OP == MIX_IMADHI_LOP3) {
            if (j & 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(kThreads) k_rate(uint32_t* out, long long* cyc, int iters, uint32_t b, uint32_t c)
{
    __shared__ uint32_t sm[1024];
    sm[threadIdx.x & 1023] = threadIdx.x * 2654435761u;
    __syncthreads();
    uint32_t a[kChains];
#pragma unroll
    for (int j = 0; j < kChains; j++) a[j] = threadIdx.x * 97u + j * 1315423911u + b;
    if constexpr (OP == FFMA2) {
        uint64_t w[kChains];
#pragma unroll
        for (int j = 0; j < kChains; j++) w[j] = ((uint64_t)a[j] << 32) | (a[j] * 3u);
        const uint64_t bb = ((uint64_t)b << 32) | c, cc = ((uint64_t)c << 32) | b;
        __syncthreads();
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < kChains; j++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[j]) : "l"(bb), "l"(cc));
        }
        const long long t1 = clock64();
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < kChains; j++) s ^= (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32);
        out[blockIdx.x * kThreads + threadIdx.x] = s;
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
        return;
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) body<OP>(a, b, c, sm);
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kChains; j++) s ^= a[j];
    out[blockIdx.x * kThreads + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run_one(uint32_t* out, long long* cyc, int n_sm)
{
    if constexpr (OP < N_OPS) {
        const int iters = 2000;
        k_rate<OP><<<n_sm, kThreads>>>(out, cyc, 10, 0x3f800001u, 0x00010203u);
        k_rate<OP><<<n_sm, kThreads>>>(out, cyc, iters, 0x3f800001u, 0x00010203u);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(n_sm);
        CK(cudaMemcpy(h.data(), cyc, n_sm * sizeof(long long), cudaMemcpyDeviceToHost));
        double avg = 0;
        for (auto v : h) avg += (double)v;
        avg /= n_sm;
        const double ops = (double)kThreads * iters * kChains;
        printf("%-28s %8.1f thread-ops/clk/SM   (%.2f warp-instr/clk/SMSP)\n", kNames[OP], ops / avg, ops / avg / 128.0);
        run_one<OP + 1>(out, cyc, n_sm);
    }
}

// ---- 2-D gather: texture point fetch vs LDG --------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_tex(cudaTextureObject_t tex, uint32_t* out, int w, int h, int reps, float c, float s)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    uint32_t acc = 0;
    for (int r = 0; r < reps; r++) {
        const float fx = c * x - s * y + 200.f + r, fy = s * x + c * y + 100.f;
        const int sx = (int)fx, sy = (int)fy;
        uchar4 p00 = tex2D<uchar4>(tex, sx, sy), p01 = tex2D<uchar4>(tex, sx + 1, sy);
        uchar4 p10 = tex2D<uchar4>(tex, sx, sy + 1), p11 = tex2D<uchar4>(tex, sx + 1, sy + 1);
        acc += p00.x + p01.y + p10.z + p11.w;
    }
    out[y * gridDim.x * 32 + x] = acc;
}
__global__ void __launch_bounds__(256) k_gather_ldg(const uint32_t* __restrict__ img, int pitch_px, uint32_t* out, int w, int h, int reps, float c, float s)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    uint32_t acc = 0;
    for (int r = 0; r < reps; r++) {
        const float fx = c * x - s * y + 200.f + r, fy = s * x + c * y + 100.f;
        const int sx = min(max((int)fx, 0), w - 2), sy = min(max((int)fy, 0), h - 2);
        const uint32_t* p = img + sy * pitch_px + sx;
        acc += (__ldg(p) & 255) + (__ldg(p + 1) >> 8 & 255) + (__ldg(p + pitch_px) >> 16 & 255) + (__ldg(p + pitch_px + 1) >> 24);
    }
    out[y * gridDim.x * 32 + x] = acc;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int n_sm = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz\n", prop.name, n_sm, clk_khz / 1000);
    uint32_t* out;
    long long* cyc;
    CK(cudaMalloc(&out, (size_t)n_sm * kThreads * 4 + (1 << 26)));
    CK(cudaMalloc(&cyc, n_sm * sizeof(long long)));
    run_one<0>(out, cyc, n_sm);

    // gather test: 1920x1080 source, output 1920x1080 grid rotated by ~10 degrees
    const int w = 1920, h = 1080;
    size_t pitch;
    uint8_t* d;
    CK(cudaMallocPitch(&d, &pitch, w * 4, h));
    CK(cudaMemset(d, 0x5a, pitch * h));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = d;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
    rd.res.pitch2D.width = w;
    rd.res.pitch2D.height = h;
    rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    const float c = 0.80f, s = 0.14f; // scale ~1.23 upsample, rotated
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const dim3 grid(1920 / 32, 1080 / 8);
    for (int pass = 0; pass < 2; pass++) {
        const int reps = 8;
        k_gather_tex<<<grid, 256>>>(tex, out, w, h, reps, c, s);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) k_gather_tex<<<grid, 256>>>(tex, out, w, h, reps, c, s);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fetches = 10.0 * reps * 1920 * 1080 * 4;
        if (pass) printf("TEX point uchar4: %.1f G fetch/s = %.2f fetch/clk/SM @max clock\n", fetches / ms / 1e6, fetches / (ms * 1e-3) / (clk_khz * 1e3) / n_sm);
        k_gather_ldg<<<grid, 256>>>((const uint32_t*)d, (int)(pitch / 4), out, w, h, reps, c, s);
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) k_gather_ldg<<<grid, 256>>>((const uint32_t*)d, (int)(pitch / 4), out, w, h, reps, c, s);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
        if (pass) printf("LDG.32 gather:    %.1f G fetch/s = %.2f fetch/clk/SM @max clock\n", fetches / ms / 1e6, fetches / (ms * 1e-3) / (clk_khz * 1e3) / n_sm);
    }
    return 0;
}
