// Shared host-side helpers of libstemb200 (error string, launch counter, device query).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace stem {
int set_error(const char* msg);                           // records msg, returns STEMB200_E_INVALID
int set_cuda_error(const char* what, cudaError_t e);      // records "<what>: <cuda error>", returns STEMB200_E_CUDA
void count_launch();
int num_sms();
// un-swizzled 4-D TMA view {c, w, h, n} of an NHWC fp16 tensor with box {box_c, box_w, box_h, 1}; out-of-range box
// elements read as zero (conv_igemm.cu)
int encode_nhwc_plain(CUtensorMap* m, const void* base, int n, int h, int w, int c, int box_c, int box_w, int box_h);
}  // namespace stem
