// Experiment: can the texture unit's bilinear filter reproduce cv2's 1/32-px fixed-point bilinear
// exactly?  For every texel position of a random RGBA8 image and every (fx, fy) in 0..31 it
// samples tex2D<float4> (linear filter, border addressing, normalized-float read) at
// (sx + fx/32 + 0.5, sy + fy/32 + 0.5), recovers S = rint(value * 255 * 1024) and compares with the
// exact integer sum of weights (32-fx)(32-fy).. * taps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_exact tex_exact.cu && ./tex_exact
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void k(cudaTextureObject_t tex, const uint8_t* img, int w, int h, int pitch, unsigned long long* bad,
                  unsigned long long* bad_out, float* maxerr)
{
    int sx = blockIdx.x - 2, sy = blockIdx.y - 2; // include the border ring
    int fx = threadIdx.x, fy = threadIdx.y;
    float x = (float)sx + (float)fx * (1.0f / 32.0f) + 0.5f;
    float y = (float)sy + (float)fy * (1.0f / 32.0f) + 0.5f;
    float4 v = tex2D<float4>(tex, x, y);
    float vv[4] = {v.x, v.y, v.z, v.w};
    int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
    for (int c = 0; c < 4; c++) {
        auto tap = [&](int xx, int yy) -> int {
            if (xx < 0 || yy < 0 || xx >= w || yy >= h) return 0;
            return img[yy * pitch + xx * 4 + c];
        };
        int S = w00 * tap(sx, sy) + w01 * tap(sx + 1, sy) + w10 * tap(sx, sy + 1) + w11 * tap(sx + 1, sy + 1);
        float r = vv[c] * 261120.0f;
        int Sr = __float2int_rn(r);
        float err = fabsf(r - (float)S);
        if (Sr != S) atomicAdd(bad, 1ull);
        if (((Sr + 512) >> 10) != ((S + 512) >> 10)) atomicAdd(bad_out, 1ull);
        // atomic max on float via int compare (non-negative)
        atomicMax((int*)maxerr, __float_as_int(err));
    }
}

int main()
{
    const int w = 200, h = 120;
    size_t pitch;
    uint8_t* d;
    CK(cudaMallocPitch(&d, &pitch, w * 4, h));
    std::vector<uint8_t> host(pitch * h);
    srand(1);
    for (auto& b : host) b = rand() & 255;
    // extremes: first rows saturated / zero patterns
    for (int x = 0; x < w * 4; x++) { host[x] = 255; host[pitch + x] = (x & 4) ? 255 : 0; }
    CK(cudaMemcpy(d, host.data(), pitch * h, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = d;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
    rd.res.pitch2D.width = w;
    rd.res.pitch2D.height = h;
    rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    unsigned long long *bad, *bad_out;
    float* maxerr;
    CK(cudaMalloc(&bad, 8)); CK(cudaMalloc(&bad_out, 8)); CK(cudaMalloc(&maxerr, 4));
    CK(cudaMemset(bad, 0, 8)); CK(cudaMemset(bad_out, 0, 8)); CK(cudaMemset(maxerr, 0, 4));
    k<<<dim3(w + 4, h + 4), dim3(32, 32)>>>(tex, d, w, h, (int)pitch, bad, bad_out, maxerr);
    CK(cudaDeviceSynchronize());
    unsigned long long hb, hbo; float me;
    CK(cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hbo, bad_out, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&me, maxerr, 4, cudaMemcpyDeviceToHost));
    unsigned long long total = (unsigned long long)(w + 4) * (h + 4) * 1024 * 4;
    printf("pitch %zu; samples*channels %llu; S mismatches %llu; final-pixel mismatches %llu; max |S_tex - S| %.4f\n",
           pitch, total, hb, hbo, me);
    return 0;
}
