// GaussianConditional arithmetic (entropy_models.py:122-150, 521-526, 570-604; bound_ops.py:50-53), shared by the
// stand-alone entropy kernels (elementwise.cu) and the fused EPM.4 + GaussianConditional epilogue (conv_igemm.cu).
// Every fp32 operation is spelled with a round-to-nearest intrinsic, so the result does not depend on the translation
// unit's -fmad setting: it is the reference's operation order, bit for bit, in both places.
#pragma once
#include <cuda_runtime.h>

namespace stem {

struct GcOut {
  float y_hat, lik;
  int idx, sym;
};

__device__ __forceinline__ float lower_bound(float x, float b) {
  // torch.max(x, bound): NaN propagates
  return (x != x) ? x : fmaxf(x, b);
}

__device__ __forceinline__ GcOut gc_eval(float y, float sigma, float mu, const float* __restrict__ table, int n_scales,
                                         float scale_bound, float lik_bound, bool want_idx) {
  GcOut o;
  const float t = rintf(__fsub_rn(y, mu));  // torch.round: half to even
  o.sym = static_cast<int>(t);
  o.y_hat = __fadd_rn(t, mu);
  const float v = fabsf(__fsub_rn(o.y_hat, mu));  // likelihood is evaluated at the de-quantised value
  const float s = lower_bound(sigma, scale_bound);
  const float c = -0.70710678118654752440f;  // float(-(2 ** -0.5))
  const float upper = __fmul_rn(0.5f, erfcf(__fmul_rn(c, __fdiv_rn(__fsub_rn(0.5f, v), s))));
  const float lower = __fmul_rn(0.5f, erfcf(__fmul_rn(c, __fdiv_rn(__fsub_rn(-0.5f, v), s))));
  o.lik = lower_bound(__fsub_rn(upper, lower), lik_bound);
  o.idx = 0;
  if (want_idx) {
    // idx = (n-1) - #{k < n-1 : s <= table[k]} == first k in [0, n-1) with s <= table[k] (table ascending),
    // n-1 when there is none (also for NaN, where every comparison is false)
    int lo = 0, hi = n_scales - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s <= table[mid]) hi = mid;
      else lo = mid + 1;
    }
    o.idx = lo;
  }
  return o;
}

}  // namespace stem
