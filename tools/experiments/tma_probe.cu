// Experiment: a 2-D tiled TMA box load of RGBA8 pixels (as UINT32 elements) with out-of-bounds
// zero fill, the descriptor passed (a) as a direct __grid_constant__ parameter and (b) inside a
// > 4 KB parameter struct, indexed dynamically.  Prints mismatches against the expected box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu && ./tma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Big { int pad[900]; int n; int pad2[15]; alignas(64) CUtensorMap maps[32]; };

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <class PT>
__device__ void body(const CUtensorMap* map, int bw, int bh, int ox, int oy, uint32_t* out, int* status)
{
    extern __shared__ __align__(1024) uint8_t s_stage[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = smem_addr(&bar), dst = smem_addr(s_stage);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bw * bh * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(map), "r"(ox), "r"(oy), "r"(b) : "memory");
    }
    uint32_t done = 0;
    for (int spins = 0; spins < (1 << 20) && !done; ++spins)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b) : "memory");
    if (threadIdx.x == 0) *status = done ? 1 : -1;
    if (!done) return;
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = reinterpret_cast<uint32_t*>(s_stage)[i];
}

__global__ void k_direct(const __grid_constant__ CUtensorMap map, int bw, int bh, int ox, int oy, uint32_t* out, int* status)
{
    body<int>(&map, bw, bh, ox, oy, out, status);
}
__global__ void k_struct(const __grid_constant__ Big big, int bw, int bh, int ox, int oy, uint32_t* out, int* status)
{
    body<int>(&big.maps[big.n], bw, bh, ox, oy, out, status);
}

int main()
{
    const int w = 4096, h = 4096, bw = 40, bh = 36;
    std::vector<uint32_t> img((size_t)w * h);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint32_t)(i * 2654435761u) | 1u;
    uint32_t *d, *out;
    int* status;
    CK(cudaMalloc(&d, img.size() * 4));
    CK(cudaMalloc(&out, bw * bh * 4));
    CK(cudaMalloc(&status, 4));
    CK(cudaMemcpy(d, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    printf("entry point %p query %d\n", p, (int)q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t strides[1] = {(cuuint64_t)w * 4};
    const cuuint32_t box[2] = {bw, bh};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    static Big big;
    big.n = 5;
    big.maps[5] = map;
    const int origins[5][2] = {{100, 200}, {-4, -1}, {4080, 4090}, {-8, -35}, {4092, 7}};
    for (int variant = 0; variant < 2; variant++)
        for (auto& o : origins) {
            CK(cudaMemset(out, 0xEE, bw * bh * 4));
            CK(cudaMemset(status, 0, 4));
            if (variant == 0) k_direct<<<1, 256, bw * bh * 4>>>(map, bw, bh, o[0], o[1], out, status);
            else k_struct<<<1, 256, bw * bh * 4>>>(big, bw, bh, o[0], o[1], out, status);
            cudaError_t e = cudaDeviceSynchronize();
            int st = 0;
            std::vector<uint32_t> got(bw * bh);
            if (e == cudaSuccess) {
                CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(got.data(), out, bw * bh * 4, cudaMemcpyDeviceToHost));
            }
            long bad = 0;
            for (int y = 0; y < bh; y++)
                for (int x = 0; x < bw; x++) {
                    const int sx = o[0] + x, sy = o[1] + y;
                    const uint32_t want = (sx < 0 || sy < 0 || sx >= w || sy >= h) ? 0u : img[(size_t)sy * w + sx];
                    bad += got[y * bw + x] != want;
                }
            printf("%s origin (%d, %d): sync=%s status=%d mismatches=%ld\n", variant ? "struct" : "direct", o[0], o[1],
                   cudaGetErrorString(e), st, bad);
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
