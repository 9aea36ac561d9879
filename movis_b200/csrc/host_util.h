// host_util.h -- host-side helpers shared by the C-ABI translation units.
#pragma once

#include <cuda_runtime.h>

namespace mvb {

// Records a thread-local message for mvb_last_error() and returns `code`.
int set_error(int code, const char* msg);

// MVB_OK when the current device is an sm_100-class GPU (cached per device), else an error.
int require_device();

// Adds n to the process-wide count of kernels this library has launched (mvb_kernel_launches).
void count_launches(int n);

// cudaGetLastError() after a launch -> MVB_OK / MVB_ERR_CUDA.
int check_launch(const char* what);

// The inversion cv::warpAffine applies to a forward 2x3 matrix (oracle/mvo.c: mvo_invert_affine).
void invert_affine_cv(const double M[6], double m[6]);

} // namespace mvb
