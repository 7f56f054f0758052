// Host-side pieces of libstemb200: error string, launch counter, device query, and the CDF quantiser the
// reference implements in C++ (compressai/cpp_exts/ops/ops.cpp:24-81).
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/stemb200.h"
#include "internal.h"

namespace stem {
namespace {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

int set_error(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return STEMB200_E_INVALID;
}
int set_cuda_error(const char* what, cudaError_t e) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return STEMB200_E_CUDA;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
      sms = 148;
  }
  return sms;
}
}  // namespace stem

extern "C" const char* stemb200_version(void) { return "stemb200 0.2.0 (sm_100a)"; }
extern "C" const char* stemb200_last_error(void) { return stem::g_err; }
extern "C" uint64_t stemb200_launch_count(void) { return stem::g_launches.load(); }

// Quantise a pmf to a 2^precision-total integer CDF where every symbol keeps a non-zero frequency.
// Same arithmetic as the reference: round each mass to the grid (fp32 product), renormalise with integer
// division, prefix-sum, pin the last entry, then for every zero-width bin steal one count from the
// narrowest bin that still has more than one, shifting the boundaries in between.
extern "C" int stemb200_pmf_to_quantized_cdf_host(const float* pmf, int32_t pmf_len, int32_t precision,
                                                  int32_t* cdf_out) {
  if (!pmf || !cdf_out || pmf_len < 1 || precision < 1 || precision > 30)
    return stem::set_error("pmf_to_quantized_cdf: bad argument");
  const int n = pmf_len + 1;
  std::vector<uint32_t> cdf(n);
  const int scale_i = 1 << precision;
  cdf[0] = 0;
  uint32_t total = 0;
  for (int i = 0; i < pmf_len; ++i) {
    const float scaled = pmf[i] * static_cast<float>(scale_i);
    cdf[i + 1] = static_cast<uint32_t>(std::round(scaled));
    total += cdf[i + 1];
  }
  if (total == 0) return stem::set_error("pmf_to_quantized_cdf: empty pmf");
  uint32_t running = 0;
  for (int i = 0; i < n; ++i) {
    const uint32_t f = static_cast<uint32_t>((static_cast<uint64_t>(scale_i) * cdf[i]) / total);
    running += f;
    cdf[i] = running;
  }
  cdf[n - 1] = static_cast<uint32_t>(scale_i);
  for (int i = 0; i + 1 < n; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best = ~0u;
    int donor = -1;
    for (int j = 0; j + 1 < n; ++j) {
      const uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best) {
        best = f;
        donor = j;
      }
    }
    if (donor < 0) return stem::set_error("pmf_to_quantized_cdf: no donor symbol");
    if (donor < i) {
      for (int j = donor + 1; j <= i; ++j) cdf[j]--;
    } else {
      for (int j = i + 1; j <= donor; ++j) cdf[j]++;
    }
  }
  for (int i = 0; i < n; ++i) cdf_out[i] = static_cast<int32_t>(cdf[i]);
  return 0;
}
