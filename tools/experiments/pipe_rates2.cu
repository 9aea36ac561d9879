// Experiment (part 2): rates of the float instructions a float-state compositor would use, packed
// f32x2 forms and which pipes they share.  Same harness as pipe_rates.cu.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
constexpr int kChains = 8;
constexpr int kThreads = 1024;
enum Op { FFMA_RM, FMNMX, FSETPSEL, IADD3R, RCPADD, I2FP, MIX_I2FP_LOP3, MIX_I2FP_IMAD, MIX_I2FP_FFMA, MIX_FMNMX_LOP3, MIX_FMNMX_FFMA, MIX_PRMT_LEA, MIX_FFMA_LEA, MIX_IDP_FFMA, MIX3, MIX4, N_OPS };
const char* kNames[] = {"FFMA.RM", "FMNMX", "FSETP+FSEL (pairs)", "IADD3 (3 distinct inputs)", "MUFU.RCP + FADD (pairs)", "I2FP.U32", "mix I2FP+LOP3", "mix I2FP+IMAD", "mix I2FP+FFMA", "mix FMNMX+LOP3", "mix FMNMX+FFMA", "mix PRMT+LEA", "mix FFMA+LEA", "mix IDP4A+FFMA", "mix IDP4A+FFMA+PRMT+FFMA (1:1:1:1)", "mix IDP4A+FFMA+LEA+FFMA (1:1:1:1)"};
template <int OP>
__device__ __forceinline__ void body(uint32_t (&a)[kChains], uint32_t b, uint32_t c)
{
#pragma unroll
    for (int j = 0; j < kChains; j++) {
        uint32_t& x = a[j];
        if constexpr (OP == FFMA_RM) { asm volatile("fma.rm.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == FMNMX) { asm volatile("max.f32 %0, %0, %1;" : "+r"(x) : "r"(b)); }
        else if constexpr (OP == FSETPSEL) { asm volatile("{ .reg .pred p; setp.lt.f32 p, %0, %1; selp.b32 %0, %2, %0, p; }" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == IADD3R) { asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == RCPADD) { asm volatile("{ .reg .f32 t; rcp.approx.ftz.f32 t, %0; add.f32 %0, t, %1; }" : "+r"(x) : "r"(b)); }
        else if constexpr (OP == I2FP) { asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(x)); }
        else if constexpr (OP == MIX_I2FP_LOP3) { if (j & 1) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(x)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == MIX_I2FP_IMAD) { if (j & 1) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(x)); else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == MIX_I2FP_FFMA) { if (j & 1) asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(x)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == MIX_FMNMX_LOP3) { if (j & 1) asm volatile("max.f32 %0, %0, %1;" : "+r"(x) : "r"(b)); else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == MIX_FMNMX_FFMA) { if (j & 1) asm volatile("max.f32 %0, %0, %1;" : "+r"(x) : "r"(b)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == MIX_PRMT_LEA) { if (j & 1) asm volatile("prmt.b32 %0, %0, %1, 0x2561;" : "+r"(x) : "r"(b)); else asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(x) : "r"(b)); }
        else if constexpr (OP == MIX_FFMA_LEA) { if (j & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(x) : "r"(b)); }
        else if constexpr (OP == MIX_IDP_FFMA) { if (j & 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == MIX3) { if (j % 4 == 0) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else if (j % 4 == 2) asm volatile("prmt.b32 %0, %0, %1, 0x2561;" : "+r"(x) : "r"(b)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
        else if constexpr (OP == MIX4) { if (j % 4 == 0) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); else if (j % 4 == 2) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(x) : "r"(b)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c)); }
    }
}
template <int OP>
__global__ void __launch_bounds__(kThreads) k_rate(uint32_t* out, long long* cyc, int iters, uint32_t b, uint32_t c)
{
    uint32_t a[kChains];
#pragma unroll
    for (int j = 0; j < kChains; j++) a[j] = threadIdx.x * 97u + j * 1315423911u + b;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) body<OP>(a, b, c);
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kChains; j++) s ^= a[j];
    out[blockIdx.x * kThreads + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// packed f32x2: MODE 0 fma.rn.f32x2, 1 fma.rm? (rn only on x2: see ptxas), 2 add.f32x2, 3 mul.f32x2,
// 4 mix FFMA2+IMAD, 5 mix FFMA2+IDP, 6 mix FFMA2+LOP3, 7 mix FFMA2+FFMA, 8 mix FFMA2+LEA, 9 FFMA2+PRMT+IDP+LEA
template <int MODE>
__global__ void __launch_bounds__(kThreads) k_x2(uint32_t* out, long long* cyc, int iters, uint32_t b, uint32_t c)
{
    uint64_t w[kChains];
    uint32_t a[kChains];
#pragma unroll
    for (int j = 0; j < kChains; j++) { a[j] = threadIdx.x * 97u + j * 1315423911u + b; w[j] = ((uint64_t)a[j] << 32) | (a[j] * 3u); }
    const uint64_t bb = ((uint64_t)b << 32) | c, cc = ((uint64_t)c << 32) | b;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < kChains; j++) {
            if (MODE == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[j]) : "l"(bb), "l"(cc));
            else if (MODE == 1) asm volatile("fma.rm.f32x2 %0, %0, %1, %2;" : "+l"(w[j]) : "l"(bb), "l"(cc));
            else if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[j]) : "l"(bb));
            else if (MODE == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(w[j]) : "l"(bb));
            else {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[j]) : "l"(bb), "l"(cc));
                uint32_t& x = a[j];
                if (MODE == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
                else if (MODE == 5) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
                else if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(b), "r"(c));
                else if (MODE == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
                else if (MODE == 8) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(x) : "r"(b));
                else if (MODE == 9) {
                    if (j % 3 == 0) asm volatile("prmt.b32 %0, %0, %1, 0x2561;" : "+r"(x) : "r"(b));
                    else if (j % 3 == 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(b), "r"(c));
                    else asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(x) : "r"(b));
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < kChains; j++) s ^= (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32) ^ a[j];
    out[blockIdx.x * kThreads + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static double avg_cycles(long long* cyc, int n_sm)
{
    std::vector<long long> h(n_sm);
    CK(cudaMemcpy(h.data(), cyc, n_sm * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (auto v : h) avg += (double)v;
    return avg / n_sm;
}
template <int OP>
void run_one(uint32_t* out, long long* cyc, int n_sm)
{
    if constexpr (OP < N_OPS) {
        const int iters = 2000;
        k_rate<OP><<<n_sm, kThreads>>>(out, cyc, 10, 0x3f800001u, 0x00010203u);
        k_rate<OP><<<n_sm, kThreads>>>(out, cyc, iters, 0x3f800001u, 0x00010203u);
        CK(cudaDeviceSynchronize());
        const double ops = (double)kThreads * iters * kChains;
        const double avg = avg_cycles(cyc, n_sm);
        printf("%-40s %8.1f thread-ops/clk/SM   (%.2f per clk per SMSP)\n", kNames[OP], ops / avg, ops / avg / 128.0);
        run_one<OP + 1>(out, cyc, n_sm);
    }
}
template <int MODE>
void run_x2(uint32_t* out, long long* cyc, int n_sm)
{
    if constexpr (MODE < 10) {
        static const char* nm[] = {"FFMA2.RN", "FFMA2.RM", "FADD2", "FMUL2", "FFMA2+IMAD", "FFMA2+IDP4A", "FFMA2+LOP3", "FFMA2+FFMA", "FFMA2+LEA", "FFMA2 + (PRMT|IDP|LEA)"};
        const int iters = 2000;
        k_x2<MODE><<<n_sm, kThreads>>>(out, cyc, 10, 0x3f800001u, 0x00010203u);
        k_x2<MODE><<<n_sm, kThreads>>>(out, cyc, iters, 0x3f800001u, 0x00010203u);
        CK(cudaDeviceSynchronize());
        const double slots = (double)kThreads * iters * kChains * (MODE >= 4 ? 2 : 1);
        const double avg = avg_cycles(cyc, n_sm);
        printf("%-40s %8.1f thread-instr/clk/SM   (%.2f issue slots per clk per SMSP)\n", nm[MODE], slots / avg, slots / avg / 128.0);
        run_x2<MODE + 1>(out, cyc, n_sm);
    }
}
int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int n_sm = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, n_sm);
    uint32_t* out;
    long long* cyc;
    CK(cudaMalloc(&out, (size_t)n_sm * kThreads * 4));
    CK(cudaMalloc(&cyc, n_sm * sizeof(long long)));
    run_one<0>(out, cyc, n_sm);
    run_x2<0>(out, cyc, n_sm);
    return 0;
}
