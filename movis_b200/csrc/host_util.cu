// host_util.cu -- error state, device check and host-side fp64 set-up for the C ABI.
#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/movis_b200.h"
#include "host_util.h"

namespace mvb {

static thread_local std::string g_last_error;

int set_error(int code, const char* msg)
{
    g_last_error = msg ? msg : "";
    return code;
}

int require_device()
{
    static thread_local int cached_dev = -1;
    static thread_local int cached_rc = MVB_ERR_NO_DEVICE;
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(MVB_ERR_NO_DEVICE, (std::string("no CUDA device: ") + cudaGetErrorString(e)).c_str());
    }
    if (dev == cached_dev) {
        if (cached_rc != MVB_OK) set_error(cached_rc, "current CUDA device is not compute capability 10.x (sm_100a required)");
        return cached_rc;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(MVB_ERR_CUDA, cudaGetErrorString(e));
    }
    cached_dev = dev;
    cached_rc = (major == 10) ? MVB_OK : MVB_ERR_NO_DEVICE;
    if (cached_rc != MVB_OK) set_error(cached_rc, "current CUDA device is not compute capability 10.x (sm_100a required)");
    return cached_rc;
}

static std::atomic<unsigned long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int check_launch(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return MVB_OK;
    char buf[256];
    std::snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    return set_error(MVB_ERR_CUDA, buf);
}

void invert_affine_cv(const double M[6], double m[6])
{
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    m[0] = A11;
    m[1] = M[1] * (-D);
    m[3] = M[3] * (-D);
    m[4] = A22;
    const double b1 = -m[0] * M[2] - m[1] * M[5];
    const double b2 = -m[3] * M[2] - m[4] * M[5];
    m[2] = b1;
    m[5] = b2;
}

} // namespace mvb

extern "C" const char* mvb_last_error(void) { return mvb::g_last_error.c_str(); }
extern "C" int mvb_abi_version(void) { return MVB_ABI_VERSION; }
extern "C" int mvb_device_check(void) { return mvb::require_device(); }
extern "C" uint64_t mvb_kernel_launches(void) { return mvb::g_launches.load(std::memory_order_relaxed); }
