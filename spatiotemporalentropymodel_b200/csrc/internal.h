// Shared host-side helpers of libstemb200 (error string, launch counter, device query).
#pragma once
#include <cuda_runtime.h>

namespace stem {
int set_error(const char* msg);                           // records msg, returns STEMB200_E_INVALID
int set_cuda_error(const char* what, cudaError_t e);      // records "<what>: <cuda error>", returns STEMB200_E_CUDA
void count_launch();
int num_sms();
}  // namespace stem
