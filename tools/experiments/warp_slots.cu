// Which scheduler partition (SMSP) do the warps of co-resident CTAs land on?  3 CTAs x 9 warps per SM, as
// k_stack_frames launches them; prints %warpid (hardware warp slot) of every warp of the CTAs on SM 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_slots warp_slots.cu && ./warp_slots
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(288, 3) k(int* out, int regs_dummy)
{
    extern __shared__ int sm[];
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    if ((threadIdx.x & 31) == 0) {
        int* o = out + (blockIdx.x * 9 + (threadIdx.x >> 5)) * 2;
        o[0] = smid; o[1] = warpid;
    }
    // stay resident long enough for all CTAs to be co-resident
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) { }
}
int main()
{
    int n = 148 * 3;
    int* d; cudaMalloc(&d, n * 9 * 2 * sizeof(int));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 51 * 1024);
    k<<<n, 288, 51 * 1024>>>(d, 0);
    cudaDeviceSynchronize();
    int* h = new int[n * 9 * 2];
    cudaMemcpy(h, d, n * 9 * 2 * sizeof(int), cudaMemcpyDeviceToHost);
    for (int sm = 0; sm < 2; sm++) {
        printf("SM %d:\n", sm);
        for (int b = 0; b < n; b++) {
            if (h[b * 18] != sm) continue;
            printf("  CTA %3d warp slots:", b);
            for (int w = 0; w < 9; w++) printf(" %2d", h[(b * 9 + w) * 2 + 1]);
            printf("   producer (warp 8) slot %% 4 = %d\n", h[(b * 9 + 8) * 2 + 1] % 4);
        }
    }
    return 0;
}
