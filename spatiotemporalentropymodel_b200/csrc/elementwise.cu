// HBM-bound kernels of the STEM P-frame path: API-boundary layout changes, first-layer im2col staging,
// GaussianConditional (quantise + erfc likelihood + scale-table index + symbols + bit count),
// EntropyBottleneck forward, synthesis tail (pixel-unshuffle + clamp + squared error).
// Compiled with -fmad=false so that the fp32 operation order of the reference (entropy_models.py) is kept.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "../../include/stemb200.h"
#include "gc_math.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace stem {

// ---------------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide fp32 partials -> one fp64 atomicAdd per block
__device__ __forceinline__ void block_accumulate(float v, double* dst, float* red /*>= 32 floats smem*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    float s = lane < nw ? red[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(dst, static_cast<double>(s));
  }
}

// ---------------------------------------------------------------------------------------------------
// frame pixel loads: fp32 frames as they are, 8-bit frames as float(v) / 255 (the IEEE quotient), which is what
// torchvision's ToTensor computes for the PNG frames of stem/evalSTEM.py:185 - bit-identical to uploading fp32
// ---------------------------------------------------------------------------------------------------
// float(v) / 255 for a byte v, correctly rounded, without the division subroutine (range checks + slow path: ~7 % of
// the 8-bit col2im kernel): one Newton step on q0 = v * fp32(1 / 255) - q = fma(fma(-q0, 255, v), r, q0) - equals the
// IEEE quotient for all 256 inputs (checked exhaustively, and by the bit-identity tests against torch's division).
__device__ __forceinline__ float u8_unit(uint32_t v) {
  const float x = static_cast<float>(v);
  const float r = 0.0039215688593685627f;  // fp32(1 / 255)
  const float q0 = __fmul_rn(x, r);
  return __fmaf_rn(__fmaf_rn(-q0, 255.0f, x), r, q0);
}
__device__ __forceinline__ float px_load(const float* p) { return __ldg(p); }
__device__ __forceinline__ float px_load(const uint8_t* p) { return u8_unit(__ldg(p)); }
// Raw pixel pairs: fetched early (before a kernel waits on something else), converted where they are used - converting
// at the load site makes the thread wait for the DRAM round trip right there, in front of the TMA wait instead of
// behind it (col2im with 8-bit frames: 446 us that way, 226 us this way; 243 us with fp32 frames).
struct RawPairF { float a, b; };
struct RawPairU8 { uint32_t v; };  // byte 0 = first pixel, byte 1 = second
__device__ __forceinline__ RawPairF px_raw2(const float* p, bool vec, bool ok0, bool ok1) {
  RawPairF r{0.f, 0.f};
  if (vec && ok0 && ok1) {
    const float2 f = __ldg(reinterpret_cast<const float2*>(p));
    r.a = f.x;
    r.b = f.y;
  } else {
    if (ok0) r.a = __ldg(p);
    if (ok1) r.b = __ldg(p + 1);
  }
  return r;
}
__device__ __forceinline__ RawPairU8 px_raw2(const uint8_t* p, bool vec, bool ok0, bool ok1) {
  RawPairU8 r{0u};
  if (vec && ok0 && ok1) {
    r.v = __ldg(reinterpret_cast<const unsigned short*>(p));
  } else {
    if (ok0) r.v = __ldg(p);
    if (ok1) r.v |= static_cast<uint32_t>(__ldg(p + 1)) << 8;
  }
  return r;
}
__device__ __forceinline__ void px_unpack(const RawPairF& r, float& a, float& b) {
  a = r.a;
  b = r.b;
}
__device__ __forceinline__ void px_unpack(const RawPairU8& r, float& a, float& b) {
  a = u8_unit(r.v & 0xFFu);
  b = u8_unit((r.v >> 8) & 0xFFu);
}
template <typename T> struct RawPairOf;
template <> struct RawPairOf<float> { using type = RawPairF; };
template <> struct RawPairOf<uint8_t> { using type = RawPairU8; };

// ---------------------------------------------------------------------------------------------------
// layout kernels (32x32 smem transposes)
// ---------------------------------------------------------------------------------------------------
// in: [n][c][hw] fp32 -> out: [n][hw][c] fp16
__global__ void nchw_to_nhwc_f16_kernel(const float* __restrict__ in, const float* __restrict__ sub,
                                        __half* __restrict__ out, int c, int hw, int round_first) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = in + static_cast<long long>(n) * c * hw;
  __half* dst = out + static_cast<long long>(n) * c * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i, pp = p0 + threadIdx.x;
    float v = 0.f;
    if (cc < c && pp < hw) {
      const long long o = static_cast<long long>(cc) * hw + pp;
      v = src[o];
      if (sub) v -= sub[static_cast<long long>(n) * c * hw + o];
    }
    tile[i][threadIdx.x] = round_first ? rintf(v) : v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, cc = c0 + threadIdx.x;
    if (cc < c && pp < hw) dst[static_cast<long long>(pp) * c + cc] = __float2half_rn(tile[threadIdx.x][i]);
  }
}

// in: [n][hw][c] (fp16 or fp32) -> out: [n][c][hw] fp32
template <typename TIn>
__global__ void nhwc_to_nchw_f32_kernel(const TIn* __restrict__ in, float* __restrict__ out, int c, int hw) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const TIn* src = in + static_cast<long long>(n) * c * hw;
  float* dst = out + static_cast<long long>(n) * c * hw;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, cc = c0 + threadIdx.x;
    float v = 0.f;
    if (cc < c && pp < hw) v = static_cast<float>(src[static_cast<long long>(pp) * c + cc]);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i, pp = p0 + threadIdx.x;
    if (cc < c && pp < hw) dst[static_cast<long long>(cc) * hw + pp] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------------------------
// first analysis layer operand staging: rows of 80 fp16 per output pixel, k = (r*5 + s)*3 + ch for k < 75
// (5 zero columns pad the row to 160 bytes; the conv kernel's TMA box zero-fills channels 80..127).
// One block = 64 consecutive output pixels of one output row: the 5 x 131 x 3 input patch is staged in smem with
// coalesced reads, then every 16-byte piece of the output rows is written by one thread (fully coalesced).
// ---------------------------------------------------------------------------------------------------
constexpr int kIm2colRow = 80;
constexpr int kIm2colPix = 64;
constexpr int kIm2colInW = 2 * kIm2colPix + 3;  // 131 input columns feed 64 stride-2 outputs of a 5-tap filter

template <typename TIn>
__global__ void __launch_bounds__(256)
im2col_k5s2_c3_kernel(const TIn* __restrict__ x, __half* __restrict__ rows, int h, int w, int h_out, int w_out,
                      int pad_top, int pad_left) {
  __shared__ float s_in[3][5][kIm2colInW + 1];
  __shared__ unsigned short s_off[kIm2colRow];  // k -> offset of (ch, r, s) inside s_in, 0xFFFF for the padding
  const int n = blockIdx.z, oh = blockIdx.y, ow0 = blockIdx.x * kIm2colPix;
  if (threadIdx.x < kIm2colRow) {
    const int k = threadIdx.x;
    unsigned short off = 0xFFFF;
    if (k < 75) {
      const int tap = k / 3, ch = k - tap * 3, r = tap / 5, sx = tap - r * 5;
      off = static_cast<unsigned short>((ch * 5 + r) * (kIm2colInW + 1) + sx);
    }
    s_off[k] = off;
  }
  const TIn* xi = x + static_cast<long long>(n) * 3 * h * w;
  const int ih0 = 2 * oh - 2 - pad_top, iw0 = 2 * ow0 - 2 - pad_left;
  for (int i = threadIdx.x; i < 3 * 5 * kIm2colInW; i += blockDim.x) {
    const int col = i % kIm2colInW, rr = (i / kIm2colInW) % 5, ch = i / (5 * kIm2colInW);
    const int ih = ih0 + rr, iw = iw0 + col;
    float v = 0.f;
    if (ih >= 0 && ih < h && iw >= 0 && iw < w) v = px_load(xi + (static_cast<long long>(ch) * h + ih) * w + iw);
    s_in[ch][rr][col] = v;
  }
  __syncthreads();
  const float* sf = &s_in[0][0][0];
  const int npix = min(kIm2colPix, w_out - ow0);
  uint4* dst = reinterpret_cast<uint4*>(rows + (static_cast<long long>(n) * h_out + oh) * w_out * kIm2colRow +
                                        static_cast<long long>(ow0) * kIm2colRow);
  for (int q = threadIdx.x; q < npix * (kIm2colRow / 8); q += blockDim.x) {
    const int px = q / (kIm2colRow / 8), piece = q - px * (kIm2colRow / 8);
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned short o0 = s_off[piece * 8 + 2 * j], o1 = s_off[piece * 8 + 2 * j + 1];
      const float f0 = o0 == 0xFFFF ? 0.f : sf[o0 + 2 * px];
      const float f1 = o1 == 0xFFFF ? 0.f : sf[o1 + 2 * px];
      const __half2 hh = __floats2half2_rn(f0, f1);
      pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    dst[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---------------------------------------------------------------------------------------------------
// stem_roi staging: im2col for conv(4, 192, k3, s1) on cat[x (3 ch), Qmap (1 ch)] (stem_roi.py:379, :586),
// rows of 40 fp16: k = (r*3 + s)*4 + ch for the 36 real entries, then 4 zeros. Same scheme as above.
// ---------------------------------------------------------------------------------------------------
constexpr int kIm2col3Row = 40;
constexpr int kIm2col3InW = kIm2colPix + 2;

template <typename TIn>
__global__ void __launch_bounds__(256)
im2col_k3s1_c4_kernel(const TIn* __restrict__ x, const float* __restrict__ q, __half* __restrict__ rows, int h,
                      int w) {
  __shared__ float s_in[4][3][kIm2col3InW + 2];
  const int n = blockIdx.z, oh = blockIdx.y, ow0 = blockIdx.x * kIm2colPix;
  for (int i = threadIdx.x; i < 4 * 3 * kIm2col3InW; i += blockDim.x) {
    const int col = i % kIm2col3InW, rr = (i / kIm2col3InW) % 3, ch = i / (3 * kIm2col3InW);
    const int ih = oh - 1 + rr, iw = ow0 - 1 + col;
    float v = 0.f;
    if (ih >= 0 && ih < h && iw >= 0 && iw < w)
      v = ch < 3 ? px_load(x + ((static_cast<long long>(n) * 3 + ch) * h + ih) * w + iw)
                 : __ldg(q + (static_cast<long long>(n) * h + ih) * w + iw);
    s_in[ch][rr][col] = v;
  }
  __syncthreads();
  const int npix = min(kIm2colPix, w - ow0);
  uint4* dst = reinterpret_cast<uint4*>(rows + ((static_cast<long long>(n) * h + oh) * w + ow0) * kIm2col3Row);
  for (int qd = threadIdx.x; qd < npix * (kIm2col3Row / 8); qd += blockDim.x) {
    const int px = qd / (kIm2col3Row / 8), piece = qd - px * (kIm2col3Row / 8);
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = piece * 8 + 2 * j + e;
        const int tap = k >> 2, ch = k & 3, r = tap / 3, sx = tap - r * 3;
        f[e] = k < 36 ? s_in[ch][r][px + sx] : 0.f;
      }
      const __half2 hh = __floats2half2_rn(f[0], f[1]);
      pk[j] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    dst[qd] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// mean pooling by an integer factor, NHWC fp16 (F.adaptive_avg_pool2d with divisible sizes, stem_utils.py:37)
__global__ void avgpool_nhwc_f16_kernel(const __half* __restrict__ in, __half* __restrict__ out, int h_out, int w_out,
                                        int c, int f, long long total) {
  const int c8 = c >> 3;
  const float inv = 1.0f / static_cast<float>(f * f);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int piece = static_cast<int>(i % c8);
    long long t = i / c8;
    const int ow = static_cast<int>(t % w_out);
    t /= w_out;
    const int oh = static_cast<int>(t % h_out);
    const long long n = t / h_out;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) {
        const long long pin = ((n * h_out * f + (oh * f + dy)) * static_cast<long long>(w_out) * f + (ow * f + dx));
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + pin * c) + piece);
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u[k]));
          acc[2 * k] += f2.x;
          acc[2 * k + 1] += f2.y;
        }
      }
    uint32_t pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2 hh = __floats2half2_rn(acc[2 * k] * inv, acc[2 * k + 1] * inv);
      pk[k] = *reinterpret_cast<const uint32_t*>(&hh);
    }
    reinterpret_cast<uint4*>(out + i / c8 * c)[piece] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// quality map staging for the hyper-encoder (stem_roi.py:563-564): (B,1,H,W) fp32 -> mean over f x f blocks ->
// NHWC fp16 with 8 channels (channel 0 = pooled map, 1..7 = 0) so it can be a K-segment of the next conv
__global__ void qmap_pool_kernel(const float* __restrict__ q, __half* __restrict__ out, int h_out, int w_out, int f,
                                 long long total) {
  const float inv = 1.0f / static_cast<float>(f * f);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ow = static_cast<int>(i % w_out);
    long long t = i / w_out;
    const int oh = static_cast<int>(t % h_out);
    const long long n = t / h_out;
    const float* base = q + (n * h_out * f + static_cast<long long>(oh) * f) * (static_cast<long long>(w_out) * f) +
                        static_cast<long long>(ow) * f;
    float acc = 0.f;
    for (int dy = 0; dy < f; ++dy)
      for (int dx = 0; dx < f; ++dx) acc += __ldg(base + static_cast<long long>(dy) * w_out * f + dx);
    const __half2 hh = __floats2half2_rn(acc * inv, 0.f);
    reinterpret_cast<uint4*>(out)[i] = make_uint4(*reinterpret_cast<const uint32_t*>(&hh), 0, 0, 0);
  }
}

__global__ void cast_f16_to_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, long long n8) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
    float4 a, b;
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u[0]));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u[1]));
    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&u[2]));
    const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&u[3]));
    a = make_float4(f0.x, f0.y, f1.x, f1.y);
    b = make_float4(f2.x, f2.y, f3.x, f3.y);
    reinterpret_cast<float4*>(out)[2 * i] = a;
    reinterpret_cast<float4*>(out)[2 * i + 1] = b;
  }
}

// ---------------------------------------------------------------------------------------------------
// latent staging
// ---------------------------------------------------------------------------------------------------
__global__ void latent_stage_kernel(const float* __restrict__ y, const __half* __restrict__ cond,
                                    __half* __restrict__ y16, __half* __restrict__ yq16,
                                    __half* __restrict__ yhat16, long long numel) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = y[i];
    if (y16) y16[i] = __float2half_rn(v);
    const float sub = cond ? __half2float(cond[i]) : 0.f;
    const float q = rintf(v - sub);
    if (yq16) yq16[i] = __float2half_rn(q);
    if (yhat16) yhat16[i] = __float2half_rn(q + sub);
  }
}

// ---------------------------------------------------------------------------------------------------
// GaussianConditional arithmetic (entropy_models.py:122-150, 521-526, 570-604; bound_ops.py:50-53)
// ---------------------------------------------------------------------------------------------------
__global__ void gc_flat_kernel(const float* __restrict__ y, const float* __restrict__ scales,
                               const float* __restrict__ means, long long numel,
                               const float* __restrict__ table_g, int n_scales, float scale_bound,
                               float lik_bound, float* __restrict__ y_hat, float* __restrict__ lik,
                               int* __restrict__ idx, int* __restrict__ sym, double* bits) {
  __shared__ float table[256];
  __shared__ float red[32];
  const bool want_idx = idx != nullptr;
  if (want_idx)
    for (int i = threadIdx.x; i < n_scales; i += blockDim.x) table[i] = table_g[i];
  __syncthreads();
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const GcOut o = gc_eval(y[i], scales[i], means ? means[i] : 0.f, table, n_scales, scale_bound, lik_bound,
                            want_idx);
    if (y_hat) y_hat[i] = o.y_hat;
    if (lik) lik[i] = o.lik;
    if (idx) idx[i] = o.idx;
    if (sym) sym[i] = o.sym;
    acc -= log2f(o.lik);
  }
  if (bits) block_accumulate(acc, bits, red);
}

constexpr int kGcPix = 32, kGcCh = 64, kGcIter = kGcPix / 4;
// One thread's 8 positions of a gc_nhwc_kernel tile (pixel slots 0, 4, ..., 28 of one channel). kFull: all 8 slots
// are inside the frame, so the operand loads are unguarded and issue back to back.
template <bool kIdx, bool kFull>
__device__ __forceinline__ float gc_tile_eval(const float* __restrict__ yp, const __half* __restrict__ cp,
                                              const float* s_cond, const float* __restrict__ sp, int c, int lim,
                                              int y_is_nchw,
                                              int yhat_mode, bool want_idx, const float* table, int n_scales,
                                              float scale_bound, float lik_bound, float* s_yhat, float* s_lik,
                                              int* s_idx, int* s_sym) {
  constexpr int kIter = 8;
  float yv[kIter], cv[kIter], sg[kIter], mu[kIter];
  const int step = 4 * c;
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    const bool ok = kFull || 4 * k < lim;
    yv[k] = y_is_nchw ? s_yhat[4 * k] : (ok ? __ldg(yp) : 0.f);
    cv[k] = s_cond ? s_cond[4 * k] : ((cp && ok) ? __half2float(*cp) : 0.f);
    sg[k] = ok ? __ldg(sp) : 1.f;
    mu[k] = ok ? __ldg(sp + c) : 0.f;
    yp += step;
    if (cp) cp += step;
    sp += 2 * step;
  }
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < kIter; ++k) {
    if (kFull || 4 * k < lim) {
      const float v = __fsub_rn(yv[k], cv[k]);  // _Res: the coded quantity is y_cur - y_conditioned (:852)
      GcOut o = gc_eval(v, sg[k], mu[k], table, n_scales, scale_bound, lik_bound, want_idx);
      // SPM variants return y_hat = round(y [- cond]) [+ cond] (:570, :856-868), not the GC output
      if (yhat_mode == 1) o.y_hat = __fadd_rn(rintf(v), cv[k]);
      s_yhat[4 * k] = o.y_hat;
      s_lik[4 * k] = o.lik;
      if (kIdx) {
        s_idx[4 * k] = o.idx;
        s_sym[4 * k] = o.sym;
      }
      acc -= __log2f(o.lik);  // lik >= 1e-9: no denormals; the bit count is a statistic (0.5 % gate)
    }
  }
  return acc;
}

// NHWC inputs -> NCHW outputs. Block = 32 pixels x 64 channels; reads are 256-byte channel runs, writes are 128-byte
// pixel runs. Thread (ch, q) evaluates pixels q, q + 4, ..., q + 28 of channel ch: all operands of the 8 positions
// are fetched up front through pointers that advance by a constant, then the tile is transposed through smem.
template <bool kIdx>
__global__ void __launch_bounds__(256)
gc_nhwc_kernel(const float* __restrict__ y, int y_is_nchw, const __half* __restrict__ cond,
               const float* __restrict__ cond32_nchw, const float* __restrict__ params, int c, int hw, int yhat_mode,
               const float* __restrict__ table_g, int n_scales, float scale_bound, float lik_bound,
               float* __restrict__ y_hat, float* __restrict__ lik, int* __restrict__ idx, int* __restrict__ sym,
               double* bits) {
  __shared__ float s_yhat[kGcCh][kGcPix + 1];
  __shared__ float s_lik[kGcCh][kGcPix + 1];  // with cond32_nchw: holds the fp32 conditioning tile until it is consumed
  __shared__ int s_idx[kIdx ? kGcCh : 1][kGcPix + 1];
  __shared__ int s_sym[kIdx ? kGcCh : 1][kGcPix + 1];
  __shared__ float table[kIdx ? 256 : 1];
  __shared__ float red[32];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * kGcPix, c0 = blockIdx.y * kGcCh;
  const bool want_idx = kIdx && idx != nullptr;
  if (want_idx) {
    for (int i = threadIdx.x; i < n_scales; i += blockDim.x) table[i] = table_g[i];
    __syncthreads();
  }
  const float* yb = y + static_cast<long long>(n) * hw * c;
  float acc = 0.f;
  if (y_is_nchw) {
    // stage the NCHW tile through smem so the channel-major phase below reads it conflict-free
    const int pi = threadIdx.x & (kGcPix - 1);
    const int pp = p0 + pi;
    for (int ch = threadIdx.x / kGcPix; ch < kGcCh; ch += 256 / kGcPix) {
      const int cc = c0 + ch;
      const bool ok = pp < hw && cc < c;
      s_yhat[ch][pi] = ok ? yb[static_cast<long long>(cc) * hw + pp] : 0.f;
      if (cond32_nchw) s_lik[ch][pi] = ok ? cond32_nchw[(static_cast<long long>(n) * c + cc) * hw + pp] : 0.f;
    }
    __syncthreads();
  }
  {
    const int ch = threadIdx.x & (kGcCh - 1), q = threadIdx.x / kGcCh;
    const int cc = c0 + ch;
    if (cc < c) {
      const long long e0 = (static_cast<long long>(n) * hw + p0 + q) * c + cc;  // NHWC element of the first pixel
      const float* yp = y + e0;
      const __half* cp = cond ? cond + e0 : nullptr;
      const float* sp = params + 2 * e0 - cc;  // sigma of (pixel, cc); mu is c floats further
      const int lim = hw - p0 - q;             // pixel slot 4 k of this thread is inside the frame iff 4 k < lim
      // fp32 NCHW conditioning (drop-in forward of the _Res variant): its tile was staged in s_lik; a thread reads
      // its 8 values before it writes its 8 likelihoods into the same slots
      const float* sc = (y_is_nchw && cond32_nchw) ? &s_lik[ch][q] : nullptr;
      if (lim > 4 * (kGcIter - 1))
        acc = gc_tile_eval<kIdx, true>(yp, cp, sc, sp, c, lim, y_is_nchw, yhat_mode, want_idx, table, n_scales,
                                       scale_bound, lik_bound, &s_yhat[ch][q], &s_lik[ch][q], &s_idx[kIdx ? ch : 0][q],
                                       &s_sym[kIdx ? ch : 0][q]);
      else
        acc = gc_tile_eval<kIdx, false>(yp, cp, sc, sp, c, lim, y_is_nchw, yhat_mode, want_idx, table, n_scales,
                                        scale_bound, lik_bound, &s_yhat[ch][q], &s_lik[ch][q], &s_idx[kIdx ? ch : 0][q],
                                        &s_sym[kIdx ? ch : 0][q]);
    }
  }
  __syncthreads();
  {
    const int pi = threadIdx.x & (kGcPix - 1), chq = threadIdx.x / kGcPix;
    const int pp = p0 + pi;
    if (pp < hw) {
      long long o = (static_cast<long long>(n) * c + c0 + chq) * hw + pp;
      const long long ostep = static_cast<long long>(256 / kGcPix) * hw;
      const bool ch_full = c0 + kGcCh <= c;
#pragma unroll
      for (int k = 0; k < kGcCh / (256 / kGcPix); ++k) {
        const int ch = chq + k * (256 / kGcPix);
        if (ch_full || c0 + ch < c) {
          if (y_hat) y_hat[o] = s_yhat[ch][pi];
          if (lik) lik[o] = s_lik[ch][pi];
          if (kIdx) {
            if (idx) idx[o] = s_idx[ch][pi];
            if (sym) sym[o] = s_sym[ch][pi];
          }
        }
        o += ostep;
      }
    }
  }
  if (bits) block_accumulate(acc, bits + n, red);
}

// ---------------------------------------------------------------------------------------------------
// EntropyBottleneck forward (eval): one thread per channel, 16 pixels per block
// ---------------------------------------------------------------------------------------------------
constexpr int kEbParams = 59;
constexpr int kEbPix = 8;

__device__ __forceinline__ float eb_logits(float x, const float* __restrict__ q) {
  float l[3], m[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float t = q[j] * x + q[3 + j];
    l[j] = t + q[6 + j] * tanhf(t);
  }
  const float* r = q + 9;
#pragma unroll
  for (int layer = 0; layer < 3; ++layer) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float t = r[3 * i] * l[0] + r[3 * i + 1] * l[1] + r[3 * i + 2] * l[2] + r[9 + i];
      m[i] = t + r[12 + i] * tanhf(t);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) l[i] = m[i];
    r += 15;
  }
  return r[0] * l[0] + r[1] * l[1] + r[2] * l[2] + r[3];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void eb_fwd_kernel(const float* __restrict__ z, const float* __restrict__ params, int c, int hw,
                              float lik_bound, __half* __restrict__ zhat16, float* __restrict__ zhat_nchw,
                              float* __restrict__ lik_nchw, double* bits) {
  extern __shared__ float sm[];  // [2][c][kEbPix+1] + red[32]
  float* s_z = sm;
  float* s_l = sm + c * (kEbPix + 1);
  float* red = sm + 2 * c * (kEbPix + 1);
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * kEbPix;
  const int ch = threadIdx.x;
  float acc = 0.f;
  if (ch < c) {
    float q[kEbParams];
#pragma unroll
    for (int i = 0; i < kEbParams; ++i) q[i] = params[ch * kEbParams + i];
    const float med = q[58];
    // 4 positions per round: their (independent) tanh chains interleave, which is what hides the latency here
    for (int pb = 0; pb < kEbPix; pb += 4) {
      float zv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = p0 + pb + u;
        zv[u] = pp < hw ? z[(static_cast<long long>(n) * hw + pp) * c + ch] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pi = pb + u, pp = p0 + pi;
        const float zq = rintf(zv[u] - med) + med;
        const float lo = eb_logits(zq - 0.5f, q);
        const float up = eb_logits(zq + 0.5f, q);
        const float sum = lo + up;
        const float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
        float lk = fabsf(sigmoidf_(sgn * up) - sigmoidf_(sgn * lo));
        lk = lower_bound(lk, lik_bound);
        if (pp < hw) {
          if (zhat16) zhat16[(static_cast<long long>(n) * hw + pp) * c + ch] = __float2half_rn(zq);
          s_z[ch * (kEbPix + 1) + pi] = zq;
          s_l[ch * (kEbPix + 1) + pi] = lk;
          acc -= log2f(lk);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c * kEbPix; i += blockDim.x) {
    const int cc = i / kEbPix, pi = i % kEbPix;
    const int pp = p0 + pi;
    if (pp < hw) {
      const long long o = (static_cast<long long>(n) * c + cc) * hw + pp;
      if (zhat_nchw) zhat_nchw[o] = s_z[cc * (kEbPix + 1) + pi];
      if (lik_nchw) lik_nchw[o] = s_l[cc * (kEbPix + 1) + pi];
    }
  }
  if (bits) block_accumulate(acc, bits + n, red);
}

// ---------------------------------------------------------------------------------------------------
// synthesis tail. The last transposed conv (N -> 3, k5 s2 p2 op1, priors.py:438) runs on the tensor cores as a
// stride-2 conv over 2x2 input "super pixels" with 4x4x3 = 48 outputs each (channel (u*4+v)*3 + c holds
// x_hat[c][4i+u][4j+v]); this kernel un-shuffles that to NCHW, clamps to [0, 1] (priors.py:399) and accumulates
// the squared error against the unpadded frame (evalSTEM.py:29-31,127-129). One thread = one super pixel.
// ---------------------------------------------------------------------------------------------------
constexpr int kTailCh = 64;

template <typename TRef>
__global__ void __launch_bounds__(256)
synthesis_tail_kernel(const float* __restrict__ in, float* __restrict__ xhat, int h4, int w4,
                      const TRef* __restrict__ xref, int h_ref, int w_ref, int pad_top, int pad_left,
                      double* sq_err, int clamp01) {
  __shared__ float red[32];
  const int n = blockIdx.y;
  const long long per = static_cast<long long>(h4) * w4;
  const int H = 4 * h4, W = 4 * w4;
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ii = static_cast<int>(i / w4), jj = static_cast<int>(i - static_cast<long long>(ii) * w4);
    const float4* src = reinterpret_cast<const float4*>(in + (static_cast<long long>(n) * per + i) * kTailCh);
    float v[48];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float4 f = __ldg(src + k);
      v[4 * k] = f.x;
      v[4 * k + 1] = f.y;
      v[4 * k + 2] = f.z;
      v[4 * k + 3] = f.w;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int Y = 4 * ii + u;
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float raw = v[(u * 4 + q) * 3 + ch];
          o[q] = clamp01 ? fminf(fmaxf(raw, 0.f), 1.f) : raw;
        }
        const long long off = ((static_cast<long long>(n) * 3 + ch) * H + Y) * W + 4 * jj;
        *reinterpret_cast<float4*>(xhat + off) = make_float4(o[0], o[1], o[2], o[3]);
        if (xref) {
          const int yr = Y - pad_top;
          if (yr >= 0 && yr < h_ref) {
            const TRef* rrow = xref + ((static_cast<long long>(n) * 3 + ch) * h_ref + yr) * w_ref;
            const int x0 = 4 * jj - pad_left;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (x0 + q >= 0 && x0 + q < w_ref) {
                const float d = px_load(rrow + x0 + q) - o[q];
                acc += d * d;
              }
            }
          }
        }
      }
    }
  }
  if (sq_err) block_accumulate(acc, sq_err + n, red);
}

// ---------------------------------------------------------------------------------------------------
// Last synthesis layer as GEMM + col2im (priors.py:438 deconv(N, 3, k5, s2, p2, op1)): the producing layer's kernel
// left, for every input pixel (i, j), the 75 products sum_ci x[ci][i][j] * w[ci][c][r][s] (fp16, 96 per pixel).
// Output pixel (oh, ow) sums the taps with oh = 2 i - 2 + r, ow = 2 j - 2 + s (2-3 per axis), adds the bias, clamps,
// and accumulates the squared error against the un-padded source frame.
//
// Column order ("quad-grouped", col_index below): the 2x2 output quad (2 i' + a, 2 j' + b) takes from input pixel
// (i' + di, j' + dj), di, dj in {+1, 0, -1}, exactly the taps r in R(di), s in S(dj) with R(+1) = {0, 1},
// R(0) = {2, 3}, R(-1) = {4} (a = r & 1, b = s & 1). The 9 groups are stored contiguously, 8-byte aligned and ordered
// [a][b][c], so one thread sums a whole quad from 13 vector loads of shared memory with compile-time offsets.
// Persistent CTAs (one per SM, 512 threads = 16 x 32 quads = 32 x 64 output pixels per tile): the 18 x 34 contributing
// col rows (88 of the 96 columns) arrive as ONE 4-D TMA box per tile - frame borders are the box's out-of-range zero
// fill - into a two-stage mbarrier ring, so the next tile streams in while this one is summed. Staged pixel pitch
// 176 B = 11 x 16 B: the LDS.128 of 8 consecutive lanes fall into 8 different bank groups.
// ---------------------------------------------------------------------------------------------------
constexpr int kC2iQH = 16, kC2iQW = 32;                    // quads per tile
constexpr int kC2iIH = kC2iQH + 2, kC2iIW = kC2iQW + 2;    // staged input pixels
constexpr int kC2iCols = 88;                               // staged columns (11 chunks of 8)
constexpr int kC2iPitch = kC2iCols * 2;                    // bytes per staged pixel
constexpr int kC2iStageBytes = kC2iIH * kC2iIW * kC2iPitch;
constexpr int kC2iStageStride = (kC2iStageBytes + 127) & ~127;
constexpr int kC2iStages = 2;
constexpr int kC2iThreads = kC2iQH * kC2iQW;
constexpr int kC2iSmem = 128 + kC2iStages * kC2iStageStride;

__host__ __device__ constexpr int c2i_group_base(int di, int dj) {
  return di == 1 ? (dj == 1 ? 0 : dj == 0 ? 12 : 48)
       : di == 0 ? (dj == 1 ? 24 : dj == 0 ? 36 : 56)
                 : (dj == 1 ? 64 : dj == 0 ? 72 : 80);
}
__host__ __device__ constexpr int c2i_tap_d(int r) { return r < 2 ? 1 : r < 4 ? 0 : -1; }
__host__ __device__ constexpr int c2i_col_index(int r, int s, int c) {
  const int di = c2i_tap_d(r), dj = c2i_tap_d(s);
  const int nb = dj == -1 ? 1 : 2;
  return c2i_group_base(di, dj) + (((r & 1) * nb) + (s & 1)) * 3 + c;
}

// adds group (DI, DJ) of the staged pixel `px` to the quad accumulators acc[a][b][c]
template <int DI, int DJ>
__device__ __forceinline__ void c2i_add_group(const unsigned char* px, float (&acc)[2][2][3]) {
  constexpr int NA = DI == -1 ? 1 : 2, NB = DJ == -1 ? 1 : 2, N = NA * NB * 3;
  constexpr int BASE = c2i_group_base(DI, DJ);  // halves, multiple of 4
  uint32_t w[(N + 1) / 2 + 3];
  const unsigned char* p = px + BASE * 2;
  if constexpr (N == 12 && BASE % 8 == 0) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint2 d = *reinterpret_cast<const uint2*>(p + 16);
    w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; w[4] = d.x; w[5] = d.y;
  } else if constexpr (N == 12) {
    const uint2 d = *reinterpret_cast<const uint2*>(p);
    const uint4 q = *reinterpret_cast<const uint4*>(p + 8);
    w[0] = d.x; w[1] = d.y; w[2] = q.x; w[3] = q.y; w[4] = q.z; w[5] = q.w;
  } else if constexpr (N == 6) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    w[0] = q.x; w[1] = q.y; w[2] = q.z;
  } else {
    const uint2 d = *reinterpret_cast<const uint2*>(p);
    w[0] = d.x; w[1] = d.y;
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const __half2 h2 = *reinterpret_cast<const __half2*>(&w[k >> 1]);
    const float f = (k & 1) ? __high2float(h2) : __low2float(h2);
    const int c = k % 3, ab = k / 3, b = ab % NB, a = ab / NB;
    acc[a][b][c] += f;
  }
}

template <typename TRef>
__global__ void __launch_bounds__(kC2iThreads, 1)
synthesis_col2im_kernel(const __grid_constant__ CUtensorMap col_map, const float* __restrict__ bias,
                        float* __restrict__ xhat, int h2, int w2, int tiles_x, int tiles_per_img, int total_tiles,
                        const TRef* __restrict__ xref, int h_ref, int w_ref, int pad_top, int pad_left,
                        double* sq_err, int clamp01, int ref_vec) {
  extern __shared__ uint8_t c2i_smem[];
  __shared__ __align__(8) uint64_t bars[kC2iStages];
  __shared__ float red[32];
  const uint32_t pad128 = (128u - (smem_u32(c2i_smem) & 127u)) & 127u;
  const uint8_t* stage_ptr = c2i_smem + pad128;
  const uint32_t stage_u32 = smem_u32(stage_ptr);
  const int tid = threadIdx.x;
  auto issue = [&](int tile, int s) {  // thread 0: one box = the whole halo tile
    const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const uint32_t bar = smem_u32(&bars[s]);
    mbar_arrive_expect_tx(bar, kC2iStageBytes);
    tma_load_4d(stage_u32 + s * kC2iStageStride, &col_map, bar, 0, tx * kC2iQW - 1, ty * kC2iQH - 1, n);
  };
  if (tid == 0) {
    tma_prefetch_desc(&col_map);
    for (int s = 0; s < kC2iStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int s = 0; s < kC2iStages; ++s) {
      const int tile = blockIdx.x + s * gridDim.x;
      if (tile < total_tiles) issue(tile, s);
    }
  const float b0 = __ldg(bias), b1 = __ldg(bias + 1), b2 = __ldg(bias + 2);
  const int qi = tid >> 5, qj = tid & 31;
  const unsigned char* ctr0 = stage_ptr + ((qi + 1) * kC2iIW + (qj + 1)) * kC2iPitch;
  constexpr int kRow = kC2iIW * kC2iPitch;
  const int H = 2 * h2, W = 2 * w2;
  const long long plane = static_cast<long long>(H) * W;
  const long long rp = static_cast<long long>(h_ref) * w_ref;
  float err = 0.f;
  int err_n = -1;
  int it = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int s = it % kC2iStages;
    const uint32_t ph = (it / kC2iStages) & 1;
    const int n = tile / tiles_per_img, rem = tile - n * tiles_per_img;
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    if (n != err_n) {  // block-uniform: the squared error is kept per frame
      if (err_n >= 0 && sq_err) block_accumulate(err, sq_err + err_n, red);
      err = 0.f;
      err_n = n;
    }
    const int qi_g = ty * kC2iQH + qi, qj_g = tx * kC2iQW + qj;
    const bool valid = qi_g < h2 && qj_g < w2;
    const int oh = 2 * qi_g, ow = 2 * qj_g;
    // reference pixels of this quad: fetched (raw) before the tile is awaited, converted where they are used
    typename RawPairOf<TRef>::type rv[3][2];
    bool rok[2] = {false, false};
    const int xr = ow - pad_left;
    const bool cok0 = xr >= 0 && xr < w_ref, cok1 = xr + 1 >= 0 && xr + 1 < w_ref;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 3; ++c) rv[c][a] = {};
    if (xref && valid) {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int yr = oh + a - pad_top;
        rok[a] = yr >= 0 && yr < h_ref;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (rok[a]) {
            const TRef* rr = xref + (static_cast<long long>(n) * 3 + c) * rp + static_cast<long long>(yr) * w_ref + xr;
            rv[c][a] = px_raw2(rr, ref_vec != 0, cok0, cok1);
          }
        }
      }
    }
    float acc[2][2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        acc[a][b][0] = b0;
        acc[a][b][1] = b1;
        acc[a][b][2] = b2;
      }
    mbar_wait(smem_u32(&bars[s]), ph);
    if (valid) {
      const unsigned char* ctr = ctr0 + s * kC2iStageStride;
      c2i_add_group<1, 1>(ctr + kRow + kC2iPitch, acc);
      c2i_add_group<1, 0>(ctr + kRow, acc);
      c2i_add_group<1, -1>(ctr + kRow - kC2iPitch, acc);
      c2i_add_group<0, 1>(ctr + kC2iPitch, acc);
      c2i_add_group<0, 0>(ctr, acc);
      c2i_add_group<0, -1>(ctr - kC2iPitch, acc);
      c2i_add_group<-1, 1>(ctr - kRow + kC2iPitch, acc);
      c2i_add_group<-1, 0>(ctr - kRow, acc);
      c2i_add_group<-1, -1>(ctr - kRow - kC2iPitch, acc);
    }
    __syncthreads();  // every thread has read stage s: refill it with the tile two rounds ahead
    if (tid == 0) {
      const int next = tile + kC2iStages * static_cast<int>(gridDim.x);
      if (next < total_tiles) issue(next, s);
    }
    if (valid) {
      float* dst = xhat + static_cast<long long>(n) * 3 * plane + static_cast<long long>(oh) * W + ow;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          float v0 = acc[a][0][c], v1 = acc[a][1][c];
          if (clamp01) {
            v0 = fminf(fmaxf(v0, 0.f), 1.f);
            v1 = fminf(fmaxf(v1, 0.f), 1.f);
          }
          *reinterpret_cast<float2*>(dst + c * plane + a * W) = make_float2(v0, v1);
          if (xref && rok[a]) {
            float r0, r1;
            px_unpack(rv[c][a], r0, r1);
            const float d0 = r0 - v0, d1 = r1 - v1;
            if (cok0) err += d0 * d0;
            if (cok1) err += d1 * d1;
          }
        }
    }
  }
  if (err_n >= 0 && sq_err) block_accumulate(err, sq_err + err_n, red);
}

}  // namespace stem

using namespace stem;

#define CHECK_LAUNCH(name)                                         \
  do {                                                             \
    count_launch();                                                \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return set_cuda_error(name, e__);      \
  } while (0)

extern "C" int stemb200_nchw_f32_to_nhwc_f16(const float* in, const float* sub, void* out, int32_t n, int32_t c,
                                             int32_t h, int32_t w, int32_t round_first, void* stream) {
  if (!in || !out || n < 1 || c < 1 || h < 1 || w < 1) return set_error("nchw_to_nhwc: bad argument");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  nchw_to_nhwc_f16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      in, sub, static_cast<__half*>(out), c, hw, round_first);
  CHECK_LAUNCH("nchw_to_nhwc_f16");
  return 0;
}

extern "C" int stemb200_nhwc_f16_to_nchw_f32(const void* in, float* out, int32_t n, int32_t c, int32_t h,
                                             int32_t w, void* stream) {
  if (!in || !out || n < 1 || c < 1 || h < 1 || w < 1) return set_error("nhwc_to_nchw: bad argument");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  nhwc_to_nchw_f32_kernel<__half><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(in), out, c, hw);
  CHECK_LAUNCH("nhwc_f16_to_nchw_f32");
  return 0;
}

extern "C" int stemb200_nhwc_f32_to_nchw_f32(const float* in, float* out, int32_t n, int32_t c, int32_t h,
                                             int32_t w, void* stream) {
  if (!in || !out || n < 1 || c < 1 || h < 1 || w < 1) return set_error("nhwc_to_nchw: bad argument");
  const int hw = h * w;
  dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
  nhwc_to_nchw_f32_kernel<float><<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(in, out, c, hw);
  CHECK_LAUNCH("nhwc_f32_to_nchw_f32");
  return 0;
}

template <typename TIn>
static int im2col_k5s2_c3_impl(const TIn* x_nchw, void* out_rows, int32_t n, int32_t h, int32_t w, int32_t h_pad,
                               int32_t w_pad, int32_t pad_top, int32_t pad_left, void* stream) {
  if (!x_nchw || !out_rows || n < 1 || h < 1 || w < 1 || h_pad < h || w_pad < w || pad_top < 0 || pad_left < 0)
    return set_error("im2col: bad argument");
  const int h_out = (h_pad - 1) / 2 + 1, w_out = (w_pad - 1) / 2 + 1;
  if (h_out > 65535 || n > 65535) return set_error("im2col: frame too large");
  dim3 grid((w_out + kIm2colPix - 1) / kIm2colPix, h_out, n);
  im2col_k5s2_c3_kernel<TIn><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nchw, static_cast<__half*>(out_rows), h, w, h_out, w_out, pad_top, pad_left);
  CHECK_LAUNCH("im2col_k5s2_c3");
  return 0;
}

extern "C" int stemb200_im2col_k5s2_c3(const float* x_nchw, void* out_rows, int32_t n, int32_t h, int32_t w,
                                       int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left,
                                       void* stream) {
  return im2col_k5s2_c3_impl<float>(x_nchw, out_rows, n, h, w, h_pad, w_pad, pad_top, pad_left, stream);
}

extern "C" int stemb200_im2col_k5s2_c3_u8(const uint8_t* x_nchw, void* out_rows, int32_t n, int32_t h, int32_t w,
                                          int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left,
                                          void* stream) {
  return im2col_k5s2_c3_impl<uint8_t>(x_nchw, out_rows, n, h, w, h_pad, w_pad, pad_top, pad_left, stream);
}

// NCHW fp32 frame -> zero-bordered NHWC8 fp16 canvas (operand of the row_taps first layer). One block row = one canvas
// row; one thread = 4 consecutive canvas pixels whose source pixels start at a multiple of 4 in the frame row, so
// the c channel planes are read with one 16-byte load each and the 4 canvas pixels leave as 64 contiguous bytes.
__device__ __forceinline__ uint4 pack_nhwc8(const float (&v)[8]) {
  const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  const __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  uint4 o;
  o.x = *reinterpret_cast<const uint32_t*>(&h0);
  o.y = *reinterpret_cast<const uint32_t*>(&h1);
  o.z = *reinterpret_cast<const uint32_t*>(&h2);
  o.w = *reinterpret_cast<const uint32_t*>(&h3);
  return o;
}

// 4 consecutive pixels of one channel plane, as loaded (conversion happens after every load of the thread is in flight)
struct RawQuadF { float4 v; };
struct RawQuadU8 { uint32_t v; };
__device__ __forceinline__ RawQuadF px_raw4(const float* p) { return {__ldg(reinterpret_cast<const float4*>(p))}; }
__device__ __forceinline__ RawQuadU8 px_raw4(const uint8_t* p) { return {__ldg(reinterpret_cast<const uint32_t*>(p))}; }
__device__ __forceinline__ void px_unpack4(const RawQuadF& r, float (&f)[4]) {
  f[0] = r.v.x, f[1] = r.v.y, f[2] = r.v.z, f[3] = r.v.w;
}
__device__ __forceinline__ void px_unpack4(const RawQuadU8& r, float (&f)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = u8_unit((r.v >> (8 * i)) & 0xFFu);
}
template <typename T> struct RawQuadOf;
template <> struct RawQuadOf<float> { using type = RawQuadF; };
template <> struct RawQuadOf<uint8_t> { using type = RawQuadU8; };

constexpr int kCanvasRowsPerThread = 2;  // two canvas rows per thread: twice the loads in flight (the kernel is latency-bound)

// kCP = channels per canvas pixel: 8 (uint4 per pixel, row_taps conv kernel) or 4 (uint2 per pixel, conv_first.cu)
template <typename TIn, int kCP>
__global__ void __launch_bounds__(128)
frame_to_nhwc8_kernel(const TIn* __restrict__ x, void* __restrict__ canvas, int c, int h, int w, int hc, int wc,
                      int off_top, int off_left, int groups, int vec_ok, int total_rows) {
  const int g = blockIdx.y * 128 + threadIdx.x;
  if (g >= groups) return;
  const int iw0 = 4 * g - ((off_left + 3) & ~3);  // multiple of 4 (may start left of the frame)
  const int cw0 = iw0 + off_left;
  const long long plane = static_cast<long long>(h) * w;
  const bool vec = vec_ok && iw0 >= 0 && iw0 + 3 < w;
  typename RawQuadOf<TIn>::type raw[kCanvasRowsPerThread][3];
  const TIn* px[kCanvasRowsPerThread];
  bool inside[kCanvasRowsPerThread];
#pragma unroll
  for (int r = 0; r < kCanvasRowsPerThread; ++r) {
    const int row = blockIdx.x * kCanvasRowsPerThread + r;  // n * hc + canvas row
    const int n = row / hc, ih = row - n * hc - off_top;
    inside[r] = row < total_rows && ih >= 0 && ih < h;
    px[r] = x + (static_cast<long long>(n) * c * h + (inside[r] ? ih : 0)) * w;
    if (inside[r] && vec && c == 3) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) raw[r][ch] = px_raw4(px[r] + ch * plane + iw0);
    }
  }
#pragma unroll
  for (int r = 0; r < kCanvasRowsPerThread; ++r) {
    const int row = blockIdx.x * kCanvasRowsPerThread + r;
    if (row >= total_rows) break;
    float v[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) v[p][ch] = 0.f;
    if (inside[r]) {
      if (vec && c == 3) {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          float f[4];
          px_unpack4(raw[r][ch], f);
          v[0][ch] = f[0];
          v[1][ch] = f[1];
          v[2][ch] = f[2];
          v[3][ch] = f[3];
        }
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (iw0 + p >= 0 && iw0 + p < w) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              if (ch < c) v[p][ch] = px_load(px[r] + ch * plane + iw0 + p);
          }
      }
    }
    if constexpr (kCP == 8) {
      uint4* dst = static_cast<uint4*>(canvas) + static_cast<long long>(row) * wc;
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (cw0 + p >= 0 && cw0 + p < wc) dst[cw0 + p] = pack_nhwc8(v[p]);
    } else {
      uint2* dst = static_cast<uint2*>(canvas) + static_cast<long long>(row) * wc;
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (cw0 + p >= 0 && cw0 + p < wc) {
          const uint4 u = pack_nhwc8(v[p]);
          dst[cw0 + p] = make_uint2(u.x, u.y);
        }
    }
  }
}

extern "C" int stemb200_synthesis_col_index(int32_t r, int32_t s, int32_t c) {
  if (r < 0 || r > 4 || s < 0 || s > 4 || c < 0 || c > 2) return set_error("synthesis_col_index: bad argument");
  return c2i_col_index(r, s, c);
}

template <typename TRef>
static int synthesis_col2im_impl(const void* col_f16, const float* bias3, float* x_hat_nchw, int32_t n, int32_t h2,
                                 int32_t w2, const TRef* x_ref, int32_t h_ref, int32_t w_ref, int32_t pad_top,
                                 int32_t pad_left, double* sq_err, int32_t clamp01, void* stream) {
  if (!col_f16 || !bias3 || !x_hat_nchw || n < 1 || h2 < 1 || w2 < 1 || n > 65535)
    return set_error("synthesis_col2im: bad argument");
  if (x_ref && (h_ref < 1 || w_ref < 1 || pad_top < 0 || pad_left < 0))
    return set_error("synthesis_col2im: bad reference geometry");
  if (reinterpret_cast<uintptr_t>(x_hat_nchw) & 7) return set_error("synthesis_col2im: x_hat must be 8-byte aligned");
  if (reinterpret_cast<uintptr_t>(col_f16) & 15) return set_error("synthesis_col2im: col must be 16-byte aligned");
  const int tiles_x = (w2 + kC2iQW - 1) / kC2iQW, tiles_y = (h2 + kC2iQH - 1) / kC2iQH;
  const long long total = static_cast<long long>(tiles_x) * tiles_y * n;
  if (total > 0x7fffffffLL - 2LL * 65536) return set_error("synthesis_col2im: frame too large");
  CUtensorMap col_map;
  if (int rc = encode_nhwc_plain(&col_map, col_f16, n, h2, w2, 96, kC2iCols, kC2iIW, kC2iIH)) return rc;
  static bool attr_set = false;  // benign race: the attribute set is idempotent (one flag per instantiation)
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(synthesis_col2im_kernel<TRef>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kC2iSmem);
    if (e != cudaSuccess) return set_cuda_error("synthesis_col2im: smem attribute", e);
    attr_set = true;
  }
  // a pixel pair is loaded with one access when it is aligned to 2 pixels of TRef
  const int ref_vec = x_ref && pad_left % 2 == 0 && w_ref % 2 == 0 &&
                      (reinterpret_cast<uintptr_t>(x_ref) & (2 * sizeof(TRef) - 1)) == 0;
  const int grid = static_cast<int>(std::min<long long>(total, num_sms()));
  synthesis_col2im_kernel<TRef><<<grid, kC2iThreads, kC2iSmem, static_cast<cudaStream_t>(stream)>>>(
      col_map, bias3, x_hat_nchw, h2, w2, tiles_x, tiles_x * tiles_y, static_cast<int>(total), x_ref, h_ref, w_ref,
      pad_top, pad_left, sq_err, clamp01, ref_vec);
  CHECK_LAUNCH("synthesis_col2im");
  return 0;
}

extern "C" int stemb200_synthesis_col2im(const void* col_f16, const float* bias3, float* x_hat_nchw, int32_t n,
                                         int32_t h2, int32_t w2, const float* x_ref, int32_t h_ref, int32_t w_ref,
                                         int32_t pad_top, int32_t pad_left, double* sq_err, int32_t clamp01,
                                         void* stream) {
  return synthesis_col2im_impl<float>(col_f16, bias3, x_hat_nchw, n, h2, w2, x_ref, h_ref, w_ref, pad_top, pad_left,
                                      sq_err, clamp01, stream);
}

extern "C" int stemb200_synthesis_col2im_u8(const void* col_f16, const float* bias3, float* x_hat_nchw, int32_t n,
                                            int32_t h2, int32_t w2, const uint8_t* x_ref, int32_t h_ref,
                                            int32_t w_ref, int32_t pad_top, int32_t pad_left, double* sq_err,
                                            int32_t clamp01, void* stream) {
  return synthesis_col2im_impl<uint8_t>(col_f16, bias3, x_hat_nchw, n, h2, w2, x_ref, h_ref, w_ref, pad_top,
                                        pad_left, sq_err, clamp01, stream);
}

template <typename TIn, int kCP = 8>
static int frame_to_nhwc8_impl(const TIn* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h, int32_t w,
                               int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left, int32_t border,
                               void* stream) {
  if (!x_nchw || !canvas || n < 1 || c < 1 || c > kCP || h < 1 || w < 1 || h_pad < h || w_pad < w || pad_top < 0 ||
      pad_left < 0 || border < 0 || pad_top + h > h_pad || pad_left + w > w_pad)
    return set_error("frame_to_nhwc8: bad argument");
  const int hc = h_pad + 2 * border, wc = w_pad + 2 * border;
  const int off_top = pad_top + border, off_left = pad_left + border;
  if (static_cast<long long>(n) * hc > 0x7fffffffLL) return set_error("frame_to_nhwc8: frame too large");
  // groups of 4 canvas pixels, the first one starting at canvas column off_left - roundup4(off_left) <= 0
  const int groups = (wc + ((off_left + 3) & ~3) - off_left + 3) / 4;
  const int vec_ok = (w % 4 == 0) && (reinterpret_cast<uintptr_t>(x_nchw) % (4 * sizeof(TIn)) == 0);
  const int total_rows = n * hc;
  dim3 grid((total_rows + kCanvasRowsPerThread - 1) / kCanvasRowsPerThread, (groups + 127) / 128);
  frame_to_nhwc8_kernel<TIn, kCP><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nchw, canvas, c, h, w, hc, wc, off_top, off_left, groups, vec_ok, total_rows);
  CHECK_LAUNCH("frame_to_nhwc8");
  return 0;
}

extern "C" int stemb200_frame_to_nhwc8(const float* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h, int32_t w,
                                       int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left,
                                       int32_t border, void* stream) {
  return frame_to_nhwc8_impl<float>(x_nchw, canvas, n, c, h, w, h_pad, w_pad, pad_top, pad_left, border, stream);
}

extern "C" int stemb200_frame_u8_to_nhwc8(const uint8_t* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h,
                                          int32_t w, int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left,
                                          int32_t border, void* stream) {
  return frame_to_nhwc8_impl<uint8_t>(x_nchw, canvas, n, c, h, w, h_pad, w_pad, pad_top, pad_left, border, stream);
}

extern "C" int stemb200_frame_to_nhwc4(const float* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h, int32_t w,
                                       int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left,
                                       int32_t border, void* stream) {
  return frame_to_nhwc8_impl<float, 4>(x_nchw, canvas, n, c, h, w, h_pad, w_pad, pad_top, pad_left, border, stream);
}

extern "C" int stemb200_frame_u8_to_nhwc4(const uint8_t* x_nchw, void* canvas, int32_t n, int32_t c, int32_t h,
                                          int32_t w, int32_t h_pad, int32_t w_pad, int32_t pad_top, int32_t pad_left,
                                          int32_t border, void* stream) {
  return frame_to_nhwc8_impl<uint8_t, 4>(x_nchw, canvas, n, c, h, w, h_pad, w_pad, pad_top, pad_left, border, stream);
}

template <typename TIn>
static int im2col_k3s1_c4_impl(const TIn* x_nchw, const float* q_nchw, void* out_rows, int32_t n, int32_t h, int32_t w,
                               void* stream) {
  if (!x_nchw || !q_nchw || !out_rows || n < 1 || h < 1 || w < 1 || h > 65535 || n > 65535)
    return set_error("im2col_k3s1_c4: bad argument");
  dim3 grid((w + kIm2colPix - 1) / kIm2colPix, h, n);
  im2col_k3s1_c4_kernel<TIn><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x_nchw, q_nchw, static_cast<__half*>(out_rows), h, w);
  CHECK_LAUNCH("im2col_k3s1_c4");
  return 0;
}

extern "C" int stemb200_im2col_k3s1_c4(const float* x_nchw, const float* q_nchw, void* out_rows, int32_t n,
                                       int32_t h, int32_t w, void* stream) {
  return im2col_k3s1_c4_impl<float>(x_nchw, q_nchw, out_rows, n, h, w, stream);
}

extern "C" int stemb200_im2col_k3s1_c4_u8(const uint8_t* x_nchw, const float* q_nchw, void* out_rows, int32_t n,
                                          int32_t h, int32_t w, void* stream) {
  return im2col_k3s1_c4_impl<uint8_t>(x_nchw, q_nchw, out_rows, n, h, w, stream);
}

extern "C" int stemb200_avgpool_nhwc_f16(const void* in, void* out, int32_t n, int32_t h_out, int32_t w_out,
                                         int32_t c, int32_t factor, void* stream) {
  if (!in || !out || n < 1 || h_out < 1 || w_out < 1 || c < 8 || c % 8 || factor < 1)
    return set_error("avgpool_nhwc_f16: bad argument");
  const long long total = static_cast<long long>(n) * h_out * w_out * (c / 8);
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 32));
  avgpool_nhwc_f16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(in), static_cast<__half*>(out), h_out, w_out, c, factor, total);
  CHECK_LAUNCH("avgpool_nhwc_f16");
  return 0;
}

extern "C" int stemb200_qmap_pool(const float* q_nchw, void* out_nhwc8_f16, int32_t n, int32_t h_out, int32_t w_out,
                                  int32_t factor, void* stream) {
  if (!q_nchw || !out_nhwc8_f16 || n < 1 || h_out < 1 || w_out < 1 || factor < 1)
    return set_error("qmap_pool: bad argument");
  const long long total = static_cast<long long>(n) * h_out * w_out;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
  qmap_pool_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      q_nchw, static_cast<__half*>(out_nhwc8_f16), h_out, w_out, factor, total);
  CHECK_LAUNCH("qmap_pool");
  return 0;
}

extern "C" int stemb200_cast_f16_to_f32(const void* in, float* out, int64_t numel, void* stream) {
  if (!in || !out || numel < 8 || numel % 8) return set_error("cast_f16_to_f32: bad argument");
  const long long n8 = numel / 8;
  const int blocks = static_cast<int>(std::min<long long>((n8 + 255) / 256, 148LL * 16));
  cast_f16_to_f32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __half*>(in), out,
                                                                               n8);
  CHECK_LAUNCH("cast_f16_to_f32");
  return 0;
}

extern "C" int stemb200_latent_stage(const float* y_nhwc, const void* cond_f16, void* y_f16, void* yq_f16,
                                     void* yhat_f16, int64_t numel, void* stream) {
  if (!y_nhwc || numel < 1) return set_error("latent_stage: bad argument");
  const int blocks = static_cast<int>(std::min<long long>((numel + 255) / 256, 148LL * 16));
  latent_stage_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y_nhwc, static_cast<const __half*>(cond_f16), static_cast<__half*>(y_f16), static_cast<__half*>(yq_f16),
      static_cast<__half*>(yhat_f16), numel);
  CHECK_LAUNCH("latent_stage");
  return 0;
}

extern "C" int stemb200_gaussian_conditional_flat(const float* y, const float* scales, const float* means,
                                                  int64_t numel, const float* scale_table, int32_t n_scales,
                                                  float scale_bound, float lik_bound, float* y_hat, float* lik,
                                                  int32_t* idx, int32_t* sym, double* bits, void* stream) {
  if (!y || !scales || numel < 1) return set_error("gaussian_conditional_flat: bad argument");
  if (idx && (!scale_table || n_scales < 1 || n_scales > 256))
    return set_error("gaussian_conditional_flat: idx needs a scale table of 1..256 entries");
  const int blocks = static_cast<int>(std::min<long long>((numel + 255) / 256, 148LL * 16));
  gc_flat_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, scales, means, numel, scale_table, n_scales, scale_bound, lik_bound, y_hat, lik, idx, sym, bits);
  CHECK_LAUNCH("gaussian_conditional_flat");
  return 0;
}

extern "C" int stemb200_gaussian_conditional_fwd_cond32(const float* y_nchw, const float* cond_f32_nchw,
                                                        const float* params_nhwc, int32_t n, int32_t c, int32_t h,
                                                        int32_t w, const float* scale_table, int32_t n_scales,
                                                        float scale_bound, float lik_bound, int32_t yhat_mode,
                                                        float* y_hat_nchw, float* lik_nchw, int32_t* idx_nchw,
                                                        int32_t* sym_nchw, double* bits, void* stream) {
  if (!y_nchw || !cond_f32_nchw || !params_nhwc || n < 1 || c < 1 || h < 1 || w < 1)
    return set_error("gaussian_conditional_fwd_cond32: bad argument");
  if (idx_nchw && (!scale_table || n_scales < 1 || n_scales > 256))
    return set_error("gaussian_conditional_fwd_cond32: idx needs a scale table of 1..256 entries");
  const int hw = h * w;
  dim3 grid((hw + kGcPix - 1) / kGcPix, (c + kGcCh - 1) / kGcCh, n);
  if (idx_nchw || sym_nchw)
    gc_nhwc_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        y_nchw, 1, nullptr, cond_f32_nchw, params_nhwc, c, hw, yhat_mode, scale_table, n_scales, scale_bound, lik_bound,
        y_hat_nchw, lik_nchw, idx_nchw, sym_nchw, bits);
  else
    gc_nhwc_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        y_nchw, 1, nullptr, cond_f32_nchw, params_nhwc, c, hw, yhat_mode, scale_table, n_scales, scale_bound, lik_bound,
        y_hat_nchw, lik_nchw, idx_nchw, sym_nchw, bits);
  CHECK_LAUNCH("gaussian_conditional_fwd_cond32");
  return 0;
}

extern "C" int stemb200_gaussian_conditional_fwd(const float* y_nhwc, int32_t y_is_nchw, const void* cond_f16,
                                                 const float* params_nhwc, int32_t n, int32_t c, int32_t h,
                                                 int32_t w, const float* scale_table, int32_t n_scales,
                                                 float scale_bound, float lik_bound, int32_t yhat_mode,
                                                 float* y_hat_nchw, float* lik_nchw, int32_t* idx_nchw,
                                                 int32_t* sym_nchw, double* bits, void* stream) {
  if (!y_nhwc || !params_nhwc || n < 1 || c < 1 || h < 1 || w < 1)
    return set_error("gaussian_conditional_fwd: bad argument");
  if (idx_nchw && (!scale_table || n_scales < 1 || n_scales > 256))
    return set_error("gaussian_conditional_fwd: idx needs a scale table of 1..256 entries");
  const int hw = h * w;
  dim3 grid((hw + kGcPix - 1) / kGcPix, (c + kGcCh - 1) / kGcCh, n);
  if (idx_nchw || sym_nchw)
    gc_nhwc_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        y_nhwc, y_is_nchw, static_cast<const __half*>(cond_f16), nullptr, params_nhwc, c, hw, yhat_mode, scale_table,
        n_scales, scale_bound, lik_bound, y_hat_nchw, lik_nchw, idx_nchw, sym_nchw, bits);
  else
    gc_nhwc_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        y_nhwc, y_is_nchw, static_cast<const __half*>(cond_f16), nullptr, params_nhwc, c, hw, yhat_mode, scale_table,
        n_scales, scale_bound, lik_bound, y_hat_nchw, lik_nchw, idx_nchw, sym_nchw, bits);
  CHECK_LAUNCH("gaussian_conditional_fwd");
  return 0;
}

extern "C" int stemb200_entropy_bottleneck_fwd(const float* z_nhwc, const float* params, int32_t n, int32_t c,
                                               int32_t h, int32_t w, float lik_bound, void* z_hat_nhwc_f16,
                                               float* z_hat_nchw, float* lik_nchw, double* bits, void* stream) {
  if (!z_nhwc || !params || n < 1 || c < 1 || c > 1024 || h < 1 || w < 1)
    return set_error("entropy_bottleneck_fwd: bad argument");
  const int hw = h * w;
  const int threads = ((c + 31) / 32) * 32;
  const size_t smem = (2 * static_cast<size_t>(c) * (kEbPix + 1) + 32) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(eb_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(eb)", e);
  }
  dim3 grid((hw + kEbPix - 1) / kEbPix, n);
  eb_fwd_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(
      z_nhwc, params, c, hw, lik_bound, static_cast<__half*>(z_hat_nhwc_f16), z_hat_nchw, lik_nchw, bits);
  CHECK_LAUNCH("entropy_bottleneck_fwd");
  return 0;
}

template <typename TRef>
static int synthesis_tail_impl(const float* in_nhwc64, float* x_hat_nchw, int32_t n, int32_t h4, int32_t w4,
                               const TRef* x_ref, int32_t h_ref, int32_t w_ref, int32_t pad_top, int32_t pad_left,
                               double* sq_err, int32_t clamp01, void* stream) {
  if (!in_nhwc64 || !x_hat_nchw || n < 1 || h4 < 1 || w4 < 1) return set_error("synthesis_tail: bad argument");
  const long long per = static_cast<long long>(h4) * w4;
  dim3 grid(static_cast<unsigned>(std::min<long long>((per + 255) / 256, 148LL * 16)), n);
  synthesis_tail_kernel<TRef><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in_nhwc64, x_hat_nchw, h4, w4, x_ref, h_ref, w_ref, pad_top, pad_left, sq_err, clamp01);
  CHECK_LAUNCH("synthesis_tail");
  return 0;
}

extern "C" int stemb200_synthesis_tail(const float* in_nhwc64, float* x_hat_nchw, int32_t n, int32_t h4,
                                       int32_t w4, const float* x_ref, int32_t h_ref, int32_t w_ref,
                                       int32_t pad_top, int32_t pad_left, double* sq_err, int32_t clamp01,
                                       void* stream) {
  return synthesis_tail_impl<float>(in_nhwc64, x_hat_nchw, n, h4, w4, x_ref, h_ref, w_ref, pad_top, pad_left, sq_err,
                                    clamp01, stream);
}

extern "C" int stemb200_synthesis_tail_u8(const float* in_nhwc64, float* x_hat_nchw, int32_t n, int32_t h4,
                                          int32_t w4, const uint8_t* x_ref, int32_t h_ref, int32_t w_ref,
                                          int32_t pad_top, int32_t pad_left, double* sq_err, int32_t clamp01,
                                          void* stream) {
  return synthesis_tail_impl<uint8_t>(in_nhwc64, x_hat_nchw, n, h4, w4, x_ref, h_ref, w_ref, pad_top, pad_left,
                                      sq_err, clamp01, stream);
}
