// Autoregressive latent coding of the SPM variants and the I-frame model on the GPU.
//
// The reference codes y one latent position at a time on the CPU (spatiotemporalpriors.py:633-678 / :729-768,
// priors.py:556-600 / :651-684): for every (h, w) in raster order
//     ctx   = masked 5x5 conv of the already coded y_hat around (h, w)          (12 causal taps x C channels)
//     g     = EPM(cat(static priors at (h, w), ctx))                           (three 1x1 layers, LeakyReLU)
//     sigma, mu = chunk(g);  idx = build_indexes(sigma);  sym = round(y - mu);  y_hat(h, w) = sym + mu
// ~16 s (encode) / ~50 s (decode) per 1080p frame. Here one persistent kernel (64 CTAs, weights resident in shared
// memory as fp32, software grid barrier between the four layers) runs the same recurrence:
//   * encode: position (h, w) only needs (h, w-1), (h, w-2) and rows h-1, h-2 up to column w+2, so all positions with
//     w + 3h = t are independent: W + 3(H-1) wavefront steps instead of H*W, every image of the batch in the same step;
//     symbols / indexes come out in raster (h, w, c) order, i.e. the order the reference feeds its rANS encoder, and
//     are coded on the host by stemb200_rans_encode_host (same byte stream format).
//   * decode: the rANS stream fixes raster order, so one position per image per step; the rANS state machine runs
//     inside the kernel (one warp per image, 32-ary CDF search with ballots) so no host round trip is needed.
// The part of the first EPM layer that does not depend on y_hat (temporal prior + hyper prior columns, bias) is a
// batched GEMM done beforehand by the tcgen05 conv kernel ("e0"). Encoder and decoder execute the same per-position
// arithmetic in the same order, so the decoder reproduces the encoder's (idx, mu) bit for bit.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>

#include "../../include/stemb200.h"
#include "internal.h"

namespace stem {
namespace {

constexpr int kArCtas = 64;
constexpr int kArThreads = 256;
constexpr int kArWarps = kArThreads / 32;
constexpr int kArTaps = 12;  // mask 'A' of a 5x5 kernel: rows 0-1 complete, row 2 columns 0-1 (layers.py:39-42)
constexpr uint64_t kRansL = 1ull << 31;
constexpr int kArMaxCdfs = 64;
constexpr int kArLutN = 128;  // buckets of the decoder's symbol lookup (cum >> 9)
constexpr int kArMaxScales = 256;

struct ArParams {
  int batch, h, w, c, l1, l2;
  int rc, r1, r2, rg;  // rows per CTA: 2c/64 (ctx), l1/64, l2/64, c/64 (sigma + mu pairs)
  int kmax;            // staging vector length: max(12c, 2c, l1, l2)
  int n_scales, mode;  // mode 0 = encode, 1 = decode
  float slope;
  float scale_bound;  // lower_bound_scale of build_indexes (entropy_models.py:598-604)
  const float* packed;  // per-CTA weight blocks (stemb200_ar_packed_floats / 64 floats each)
  int block_floats;
  const float* e0;      // [B][h][w][l1]
  const float* target;  // encode: [B][h][w][c]
  const float* table;   // [n_scales]
  float* t_hat;         // [B][h][w][c]
  int32_t* sym;         // [B][h][w][c] (may be null)
  int32_t* idx;         // [B][h][w][c]
  float* params_out;    // [B][h][w][2c] sigma | mu (may be null)
  float *ctx_buf, *h1_buf, *h2_buf, *mu_buf;  // scratch [pmax][2c | l1 | l2 | c]
  unsigned int* sync;   // [0] barrier counter, [1] abort flag, [2] completion flag (zeroed before launch)
  const uint8_t* streams;
  const int64_t* stream_off;  // [B] byte offsets (multiples of 4) into streams
  const int64_t* stream_len;  // [B] byte lengths
  const int32_t* cdf;
  const int32_t* cdf_size;
  const int32_t* cdf_off;
  int cdf_stride, n_cdfs;
  int32_t* status;  // [B] decode status: 0 ok, 1 corrupt stream
  int nstage;       // staging vectors: 8 when encoding (one per warp), 2 when decoding (position i + 1 in flight)
  int cdf16_entries;  // decode: total entries of the compact 16-bit CDF table kept in shared memory
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Sense-free counting barrier over the (co-resident, cooperative launch) grid. Returns false when the kernel must
// abort (a peer timed out): spinning forever would take the GPU down with it.
__device__ bool grid_barrier(unsigned int* sync, unsigned int& epoch, int* s_flag) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    epoch += 1;
    atomicAdd(&sync[0], 1u);
    const unsigned int target = epoch * gridDim.x;
    unsigned long long spins = 0;
    int ok = 1;
    while (ld_acquire_u32(&sync[0]) < target) {
      if ((++spins & 1023ull) == 0) {
        if (ld_acquire_u32(&sync[1]) != 0u) {
          ok = 0;
          break;
        }
        if (spins > (1ull << 31)) {  // several seconds: something is wrong
          atomicExch(&sync[1], 1u);
          ok = 0;
          break;
        }
      }
    }
    __threadfence();
    *s_flag = ok;
  }
  __syncthreads();
  return *s_flag != 0;
}

// Asynchronous global -> shared staging (16-byte cp.async.cg: served from L2, so values written by other CTAs before
// the grid barrier are seen; n_bytes = 0 zero-fills). All copies of a position are in flight together.
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst_smem));
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

constexpr int kArMaxRows = 12;  // rows of one layer owned by a CTA

// Canonical evaluation order of one dot product W[r] . v (K elements), shared by the encoder and the decoder so that
// the decoder reproduces the encoder's (sigma, mu) bit for bit:
//   a_s(lane) = fma chain over e = s*32 + lane + 256 i (i ascending),  s = 0..7
//   b(lane)   = ((((a_0 + a_1) + a_2) + ...) + a_7)
//   value     = xor-shuffle tree over the 32 lanes of b
// The decoder (one position at a time) spreads the eight segments s over the eight warps of the CTA; the encoder
// (many positions per step) lets every warp do all eight segments of its own position.
__device__ __forceinline__ float lane_tree(float a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  return a;
}

// encoder: one warp, RB rows at a time; out[r] valid in every lane
template <int RB>
__device__ __forceinline__ void warp_rows(const float* __restrict__ wrow, int K, const float* __restrict__ v, int lane,
                                          float* out) {
  float b[RB];
#pragma unroll
  for (int s = 0; s < 8; ++s) {
    float a[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) a[r] = 0.f;
    for (int e = s * 32 + lane; e < K; e += kArThreads) {
      const float x = v[e];
#pragma unroll
      for (int r = 0; r < RB; ++r) a[r] = fmaf(wrow[r * K + e], x, a[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) b[r] = s == 0 ? a[r] : b[r] + a[r];
  }
#pragma unroll
  for (int r = 0; r < RB; ++r) out[r] = lane_tree(b[r]);
}
// res[r] = W[r] . v for r < R (lane 0 writes)
__device__ __forceinline__ void warp_matvec(const float* Wm, int R, int K, const float* v, int lane, float* res) {
  int r = 0;
  float o[3];
  while (R - r >= 3) {
    warp_rows<3>(Wm + r * K, K, v, lane, o);
    if (lane == 0) {
      res[r] = o[0];
      res[r + 1] = o[1];
      res[r + 2] = o[2];
    }
    r += 3;
  }
  if (R - r == 2) {
    warp_rows<2>(Wm + r * K, K, v, lane, o);
    if (lane == 0) {
      res[r] = o[0];
      res[r + 1] = o[1];
    }
  } else if (R - r == 1) {
    warp_rows<1>(Wm + r * K, K, v, lane, o);
    if (lane == 0) res[r] = o[0];
  }
  __syncwarp();
}

// decoder: segment s = warp; every thread leaves a_warp(lane) of each row in lp[r][warp][lane]
__device__ __forceinline__ void cta_segments(const float* __restrict__ Wm, int R, int K, const float* __restrict__ v,
                                             int warp, int lane, float (*lp)[kArWarps][32]) {
  float acc[kArMaxRows];
#pragma unroll
  for (int r = 0; r < kArMaxRows; ++r) acc[r] = 0.f;
  for (int e = warp * 32 + lane; e < K; e += kArThreads) {
    const float x = v[e];
#pragma unroll
    for (int r = 0; r < kArMaxRows; ++r)
      if (r < R) acc[r] = fmaf(Wm[r * K + e], x, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kArMaxRows; ++r)
    if (r < R) lp[r][warp][lane] = acc[r];
}
// value of row r (in every lane of the calling warp) once all segments are in lp
__device__ __forceinline__ float cta_row(float (*lp)[kArWarps][32], int r, int lane) {
  float b = lp[r][0][lane];
#pragma unroll
  for (int w = 1; w < kArWarps; ++w) b += lp[r][w][lane];
  return lane_tree(b);
}

// shared-memory accesses on explicit 32-bit addresses (the decoder's symbol loop: no address arithmetic per access)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_v4(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_b32(uint32_t a, int v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

struct Pix {
  int b, hh, ww;
  long long g;  // flat position index (b*h + hh)*w + ww
};

__global__ void __launch_bounds__(kArThreads, 1) ar_codec_kernel(const ArParams p) {
  extern __shared__ float smem[];
  __shared__ int s_flag;
  __shared__ float s_lp[kArMaxRows][kArWarps][32];  // decode: per-lane segment sums of the rows of one position
  __shared__ float s_res[kArWarps][16];             // encode: row values of the warp's position
  __shared__ int s_sym[320];                 // decode: symbols of one position
  __shared__ int4 s_desc[320];               //         per channel {row address, lookup address, n, symbol offset}
  // decode: per table row a 128-bucket lookup (symbol that contains cum = b << 9; entry 128 = the last symbol) and
  // {offset into the compact table, size, symbol offset, -}
  __shared__ uint16_t s_lut[kArMaxCdfs][kArLutN + 2];
  __shared__ int4 s_meta[kArMaxCdfs];
  __shared__ float s_table[kArMaxScales];

  const int C = p.c, C2 = 2 * p.c, L1 = p.l1, L2 = p.l2;
  const int KA = kArTaps * C;
  float* w_ctx = smem;
  float* b_ctx = w_ctx + p.rc * KA;
  float* w0 = b_ctx + p.rc;
  float* w1 = w0 + p.r1 * C2;
  float* b1 = w1 + p.r2 * L1;
  float* w2 = b1 + p.r2;
  float* b2 = w2 + 2 * p.rg * L2;
  float* vec = smem + ((p.block_floats + 3) & ~3);  // two staging vectors of kmax floats
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j = blockIdx.x;

  {
    const float* src = p.packed + static_cast<long long>(j) * p.block_floats;
    for (int i = tid; i < p.block_floats; i += kArThreads) smem[i] = __ldg(src + i);
    for (int i = tid; i < p.n_scales; i += kArThreads) s_table[i] = __ldg(p.table + i);
  }
  __syncthreads();

  // decode: compact 16-bit copy of the CDF rows (sum of the row sizes, ~27 k entries = 54 KB for the 64-level table;
  // the final 65536 of a row wraps to 0 and is special-cased) + the first level of the 32-ary search of every row
  uint16_t* cdf16 = reinterpret_cast<uint16_t*>(vec + p.nstage * p.kmax);
  if (p.mode == 1 && j < p.batch) {
    if (tid == 0) {
      int acc = 0;
      for (int r = 0; r < p.n_cdfs; ++r) {
        const int sz = __ldg(p.cdf_size + r);
        s_meta[r] = make_int4(acc, sz, __ldg(p.cdf_off + r), (sz - 1 + 31) >> 5);
        acc += sz;
      }
    }
    __syncthreads();
    for (int r = 0; r < p.n_cdfs; ++r) {
      const int4 mt = s_meta[r];
      const int32_t* row = p.cdf + static_cast<long long>(r) * p.cdf_stride;
      for (int i = tid; i < mt.y; i += kArThreads) cdf16[mt.x + i] = static_cast<uint16_t>(__ldg(row + i));
    }
    __syncthreads();
    for (int i = tid; i < p.n_cdfs * (kArLutN + 1); i += kArThreads) {
      const int row = i / (kArLutN + 1), b = i - row * (kArLutN + 1);
      const int4 mt = s_meta[row];
      const int n = mt.y - 1;  // symbols s in [0, n): row[s] <= cum < row[s + 1], row[n] = 65536 (stored as 0)
      int lo = 0;
      if (b == kArLutN) {
        lo = n > 0 ? n - 1 : 0;
      } else {
        const uint32_t target = static_cast<uint32_t>(b) << 9;
        int hi = n > 0 ? n - 1 : 0;  // largest s in [0, n) with row[s] <= target (row[0] = 0)
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (static_cast<uint32_t>(cdf16[mt.x + mid]) <= target) lo = mid;
          else hi = mid - 1;
        }
      }
      s_lut[row][b] = static_cast<uint16_t>(lo);
    }
    __syncthreads();
  }
  // rANS decoder state of image j (warp 0 of CTA j < batch), identical in every lane; the stream is read 32 words at
  // a time (lane l holds word wbase + l)
  uint64_t rx = 0;
  int rpos = 0, rwords = 0, wbase = 0;  // in 32-bit words (a stream is < 8 GiB)
  uint32_t wbuf = 0;
  const uint32_t* rstream = nullptr;
  bool rbad = false;
  if (p.mode == 1 && j < p.batch && warp == 0) {
    rstream = reinterpret_cast<const uint32_t*>(p.streams + p.stream_off[j]);
    rwords = static_cast<int>(p.stream_len[j] / 4);
    wbuf = lane < rwords ? __ldg(rstream + lane) : 0u;
    if (rwords >= 2) {
      rx = static_cast<uint64_t>(__shfl_sync(0xffffffffu, wbuf, 0)) |
           (static_cast<uint64_t>(__shfl_sync(0xffffffffu, wbuf, 1)) << 32);
      rpos = 2;
    } else {
      rbad = true;
    }
  }
  auto next_word = [&]() -> uint32_t {  // warp-uniform
    if (rpos - wbase >= 32) {
      wbase += 32;
      wbuf = (wbase + lane) < rwords ? __ldg(rstream + wbase + lane) : 0u;
    }
    const uint32_t wv = __shfl_sync(0xffffffffu, wbuf, rpos - wbase);
    ++rpos;
    return wv;
  };

  const int H = p.h, W = p.w;
  const int n_steps = p.mode == 0 ? (W + 3 * (H - 1)) : H * W;
  unsigned int epoch = 0;
  // per-phase SM-clock totals of CTA 0 (debug aid, workspace words [16, 36)): A, bar, B, bar, C, bar, D, bar, E, bar
  long long prof[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tprev = clock64();
  auto tick = [&](int slot) {
    if (j == 0 && tid == 0) {
      const long long now = clock64();
      prof[slot] += now - tprev;
      tprev = now;
    }
  };

  for (int t = 0; t < n_steps; ++t) {
    // ---- positions of this step
    int lo = 0, nh = 1;
    if (p.mode == 0) {
      lo = t - W + 1 > 0 ? (t - W + 1 + 2) / 3 : 0;
      int hi = t / 3;
      if (hi > H - 1) hi = H - 1;
      nh = hi - lo + 1;
      if (nh < 0) nh = 0;
    }
    const int P = p.batch * nh;
    auto pix = [&](int i) -> Pix {
      Pix q;
      if (p.mode == 0) {
        q.b = i / nh;
        q.hh = lo + (i - q.b * nh);
        q.ww = t - 3 * q.hh;
      } else {
        q.b = i;
        q.hh = t / W;
        q.ww = t - q.hh * W;
      }
      q.g = (static_cast<long long>(q.b) * H + q.hh) * W + q.ww;
      return q;
    };
    if (p.mode == 0) {
      // ================= encode: one warp per position, eight positions in flight, staging vector per warp ==========
      float* myvec = vec + warp * p.kmax;
      // ---- A: context rows of this CTA
      for (int i = warp; i < P; i += kArWarps) {
        const Pix q = pix(i);
        for (int tap = 0; tap < kArTaps; ++tap) {
          const int dh = tap < 5 ? -2 : (tap < 10 ? -1 : 0);
          const int dw = (tap < 5 ? tap : (tap < 10 ? tap - 5 : tap - 10)) - 2;
          const int h2 = q.hh + dh, w2c = q.ww + dw;
          const bool inb = h2 >= 0 && w2c >= 0 && w2c < W;
          const float* src = inb ? p.t_hat + ((static_cast<long long>(q.b) * H + h2) * W + w2c) * C : p.t_hat;
          for (int k = 4 * lane; k < C; k += 128) cp_async16(myvec + tap * C + k, src + k, inb);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        warp_matvec(w_ctx, p.rc, KA, myvec, lane, s_res[warp]);
        if (lane < p.rc) p.ctx_buf[static_cast<long long>(i) * C2 + j * p.rc + lane] = s_res[warp][lane] + b_ctx[lane];
        __syncwarp();
      }
      __syncthreads();
      tick(0);
      if (!grid_barrier(p.sync, epoch, &s_flag)) return;
      tick(1);
      // ---- B: first EPM layer, context columns + the precomputed static part
      for (int i = warp; i < P; i += kArWarps) {
        const Pix q = pix(i);
        const float* src = p.ctx_buf + static_cast<long long>(i) * C2;
        for (int k = 4 * lane; k < C2; k += 128) cp_async16(myvec + k, src + k, true);
        cp_async_commit();
        const float e0v = lane < p.r1 ? __ldg(p.e0 + q.g * L1 + j * p.r1 + lane) : 0.f;  // in flight with the staging
        cp_async_wait<0>();
        __syncwarp();
        warp_matvec(w0, p.r1, C2, myvec, lane, s_res[warp]);
        if (lane < p.r1) {
          float v = s_res[warp][lane] + e0v;
          v = v > 0.f ? v : v * p.slope;
          p.h1_buf[static_cast<long long>(i) * L1 + j * p.r1 + lane] = v;
        }
        __syncwarp();
      }
      __syncthreads();
      tick(2);
      if (!grid_barrier(p.sync, epoch, &s_flag)) return;
      tick(3);
      // ---- C: second EPM layer
      for (int i = warp; i < P; i += kArWarps) {
        const float* src = p.h1_buf + static_cast<long long>(i) * L1;
        for (int k = 4 * lane; k < L1; k += 128) cp_async16(myvec + k, src + k, true);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        warp_matvec(w1, p.r2, L1, myvec, lane, s_res[warp]);
        if (lane < p.r2) {
          float v = s_res[warp][lane] + b1[lane];
          v = v > 0.f ? v : v * p.slope;
          p.h2_buf[static_cast<long long>(i) * L2 + j * p.r2 + lane] = v;
        }
        __syncwarp();
      }
      __syncthreads();
      tick(4);
      if (!grid_barrier(p.sync, epoch, &s_flag)) return;
      tick(5);
      // ---- D: third EPM layer (sigma and mu of this CTA's channels), quantise / index
      for (int i = warp; i < P; i += kArWarps) {
        const Pix q = pix(i);
        const float* src = p.h2_buf + static_cast<long long>(i) * L2;
        for (int k = 4 * lane; k < L2; k += 128) cp_async16(myvec + k, src + k, true);
        cp_async_commit();
        const float tgt = lane < p.rg ? __ldg(p.target + q.g * C + j * p.rg + lane) : 0.f;
        cp_async_wait<0>();
        __syncwarp();
        warp_matvec(w2, 2 * p.rg, L2, myvec, lane, s_res[warp]);
        if (lane < p.rg) {
          const int ch = j * p.rg + lane;
          const float sigma = s_res[warp][lane] + b2[lane];
          const float mu = s_res[warp][p.rg + lane] + b2[p.rg + lane];
          // build_indexes clamps the scale first (torch.max(scales, bound): NaN passes through)
          const float sb = (sigma != sigma) ? sigma : fmaxf(sigma, p.scale_bound);
          int cnt = 0;
          for (int k = 0; k + 1 < p.n_scales; ++k) cnt += (sb <= s_table[k]) ? 1 : 0;
          const long long e = q.g * C + ch;
          p.idx[e] = p.n_scales - 1 - cnt;
          if (p.params_out) {
            p.params_out[q.g * C2 + ch] = sigma;
            p.params_out[q.g * C2 + C + ch] = mu;
          }
          const float sq = rintf(tgt - mu);  // torch.round: half to even (entropy_models.py:141)
          p.t_hat[e] = sq + mu;
          if (p.sym) p.sym[e] = static_cast<int>(sq);
        }
        __syncwarp();
      }
      __syncthreads();
      tick(6);
      if (!grid_barrier(p.sync, epoch, &s_flag)) return;
      tick(7);
      continue;
    }

    // ================= decode: the whole CTA works on one position at a time (P = batch) =================
    // staging of position i of a phase into vec[i & 1] (one cp.async group; i + 1 is in flight during i)
    auto stage_ctx = [&](int i) {
      if (i < P) {
        const Pix q = pix(i);
        float* dst = vec + (i & 1) * p.kmax;
        const int per_tap = C >> 2;
        for (int c = tid; c < kArTaps * per_tap; c += kArThreads) {
          const int tap = c / per_tap, k = (c - tap * per_tap) << 2;
          const int dh = tap < 5 ? -2 : (tap < 10 ? -1 : 0);
          const int dw = (tap < 5 ? tap : (tap < 10 ? tap - 5 : tap - 10)) - 2;
          const int h2 = q.hh + dh, w2c = q.ww + dw;
          const bool inb = h2 >= 0 && w2c >= 0 && w2c < W;
          const float* src = inb ? p.t_hat + ((static_cast<long long>(q.b) * H + h2) * W + w2c) * C + k : p.t_hat;
          cp_async16(dst + tap * C + k, src, inb);
        }
      }
      cp_async_commit();
    };
    auto stage_lin = [&](const float* buf, int K, int i) {
      if (i < P) {
        float* dst = vec + (i & 1) * p.kmax;
        const float* src = buf + static_cast<long long>(i) * K;
        for (int k = tid << 2; k < K; k += kArThreads << 2) cp_async16(dst + k, src + k, true);
      }
      cp_async_commit();
    };

    // ---- A: context rows of this CTA
    stage_ctx(0);
    stage_ctx(1);
    for (int i = 0; i < P; ++i) {
      cp_async_wait<1>();
      __syncthreads();
      cta_segments(w_ctx, p.rc, KA, vec + (i & 1) * p.kmax, warp, lane, s_lp);
      __syncthreads();
      stage_ctx(i + 2);
      for (int r = warp; r < p.rc; r += kArWarps) {
        const float val = cta_row(s_lp, r, lane);
        if (lane == 0) p.ctx_buf[static_cast<long long>(i) * C2 + j * p.rc + r] = val + b_ctx[r];
      }
      __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();
    tick(0);
    if (!grid_barrier(p.sync, epoch, &s_flag)) return;
    tick(1);

    // ---- B: first EPM layer, context columns + the precomputed static part
    stage_lin(p.ctx_buf, C2, 0);
    stage_lin(p.ctx_buf, C2, 1);
    for (int i = 0; i < P; ++i) {
      const long long gpos = pix(i).g;
      cp_async_wait<1>();
      __syncthreads();
      cta_segments(w0, p.r1, C2, vec + (i & 1) * p.kmax, warp, lane, s_lp);
      __syncthreads();
      stage_lin(p.ctx_buf, C2, i + 2);
      for (int r = warp; r < p.r1; r += kArWarps) {
        float v = cta_row(s_lp, r, lane) + __ldg(p.e0 + gpos * L1 + j * p.r1 + r);
        v = v > 0.f ? v : v * p.slope;
        if (lane == 0) p.h1_buf[static_cast<long long>(i) * L1 + j * p.r1 + r] = v;
      }
      __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();
    tick(2);
    if (!grid_barrier(p.sync, epoch, &s_flag)) return;
    tick(3);

    // ---- C: second EPM layer
    stage_lin(p.h1_buf, L1, 0);
    stage_lin(p.h1_buf, L1, 1);
    for (int i = 0; i < P; ++i) {
      cp_async_wait<1>();
      __syncthreads();
      cta_segments(w1, p.r2, L1, vec + (i & 1) * p.kmax, warp, lane, s_lp);
      __syncthreads();
      stage_lin(p.h1_buf, L1, i + 2);
      for (int r = warp; r < p.r2; r += kArWarps) {
        float v = cta_row(s_lp, r, lane) + b1[r];
        v = v > 0.f ? v : v * p.slope;
        if (lane == 0) p.h2_buf[static_cast<long long>(i) * L2 + j * p.r2 + r] = v;
      }
      __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();
    tick(4);
    if (!grid_barrier(p.sync, epoch, &s_flag)) return;
    tick(5);

    // ---- D: third EPM layer (sigma and mu of this CTA's channels), index
    stage_lin(p.h2_buf, L2, 0);
    stage_lin(p.h2_buf, L2, 1);
    for (int i = 0; i < P; ++i) {
      const long long gpos = pix(i).g;
      cp_async_wait<1>();
      __syncthreads();
      cta_segments(w2, 2 * p.rg, L2, vec + (i & 1) * p.kmax, warp, lane, s_lp);
      __syncthreads();
      stage_lin(p.h2_buf, L2, i + 2);
      for (int r = warp; r < p.rg; r += kArWarps) {
        const int ch = j * p.rg + r;
        const float sigma = cta_row(s_lp, r, lane) + b2[r];
        const float mu = cta_row(s_lp, p.rg + r, lane) + b2[p.rg + r];
        // index = (n_scales - 1) - #{k < n_scales - 1 : sigma <= table[k]}, lanes share the table
        const float sb = (sigma != sigma) ? sigma : fmaxf(sigma, p.scale_bound);
        int cnt = 0;
        for (int k = lane; k + 1 < p.n_scales; k += 32) cnt += (sb <= s_table[k]) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) {
          p.idx[gpos * C + ch] = p.n_scales - 1 - cnt;
          if (p.params_out) {
            p.params_out[gpos * C2 + ch] = sigma;
            p.params_out[gpos * C2 + C + ch] = mu;
          }
          p.mu_buf[static_cast<long long>(i) * C + ch] = mu;
        }
      }
      __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();
    tick(6);
    if (!grid_barrier(p.sync, epoch, &s_flag)) return;
    tick(7);

    if (p.mode == 1) {
      // ---- E: rANS decode of this position, image j, by warp 0 (rans_interface.cpp:226-275 semantics)
      if (j < p.batch && warp == 0) {
        const Pix q = pix(j);
        const long long e0i = q.g * C;
        // per channel (state-independent, so all lanes prepare it up front): shared-memory byte addresses of the CDF
        // row and of its bucket lookup, the symbol count and the symbol offset
        const uint32_t cdf_sa = smem_addr(cdf16), lut_sa = smem_addr(&s_lut[0][0]);
        const uint32_t desc_sa = smem_addr(s_desc), sym_sa = smem_addr(s_sym);
        for (int ch = lane; ch < C; ch += 32) {
          int ci = __ldcg(p.idx + e0i + ch);
          ci = ci < 0 ? 0 : (ci >= p.n_cdfs ? p.n_cdfs - 1 : ci);
          const int4 mt = s_meta[ci];
          s_desc[ch] = make_int4(static_cast<int>(cdf_sa + 2u * mt.x), static_cast<int>(lut_sa + 2u * (kArLutN + 2) * ci),
                                 mt.y - 1, mt.z);
        }
        __syncwarp();
        // One symbol = one trip through a dependent chain of the rANS state, and ONE warp walks it, so both the length
        // of the chain and the number of instructions around it count (a single warp issues a dependent instruction
        // every ~5 cycles).  Every lane runs the same scalar sequence on explicit 32-bit shared addresses:
        //   cum -> bucket cum >> 9 -> two lookup entries (first / last symbol that can contain cum) -> the two CDF
        //   boundaries of the first candidate (almost always the symbol; a short scan or a bisection otherwise) ->
        //   64-bit state update -> renormalisation with a stream word that was shuffled out speculatively at the top
        //   of the iteration.  The descriptor of the next channel is fetched while the current one is resolved.
        // History (1080p latent, decompress): warp-parallel 32-ary search with ballots / shuffles on the chain 0.66 ->
        // 0.61 s; bucket lookup alone 0.62 s (the ~100 instructions around it, half of them address arithmetic, were the
        // cost); with per-channel descriptors and explicit shared addresses 0.58 s (0.50 s at 4.8 bits per symbol).  ncu
        // on the benchmark's synthetic latent: 42 % of the symbols need the scan behind the first candidate (128
        // buckets are coarse for its sigma ~ 30 rows; there is no shared memory left for more) and 18 % take the
        // bypass path, which a trained checkpoint uses for < 0.1 %.
        int4 dc = lds_v4(desc_sa);
        int widx = rpos - wbase;  // position of the next stream word inside the 32-word window
        for (int ch = 0; ch < C; ++ch) {
          const int4 dn = lds_v4(desc_sa + 16u * static_cast<uint32_t>(ch + 1 < C ? ch + 1 : ch));
          if (widx >= 32) {  // every 32 words: refill the window
            wbase += 32;
            widx -= 32;
            wbuf = (wbase + lane) < rwords ? __ldg(rstream + wbase + lane) : 0u;
          }
          const uint32_t w_spec = __shfl_sync(0xffffffffu, wbuf, widx);  // used only if this symbol renormalises
          const uint32_t row_sa = static_cast<uint32_t>(dc.x);
          const int n = dc.z;  // candidates s in [0, n): row[s] <= cum < row[s+1], row[n] = 65536 stored as 0
          const uint32_t cum = static_cast<uint32_t>(rx) & 0xFFFFu;
          const uint32_t la = static_cast<uint32_t>(dc.y) + ((cum >> 9) << 1);
          int s_found = static_cast<int>(lds_u16(la));
          const int s_last = static_cast<int>(lds_u16(la + 2));
          uint32_t start = lds_u16(row_sa + 2u * s_found);
          uint32_t nxt = lds_u16(row_sa + 2u * s_found + 2u);
          nxt = nxt ? nxt : 65536u;  // only row[n] wraps to 0 (row[0] = 0 is never a "next" boundary)
          if (cum >= nxt) {
            if (s_last - s_found > 6) {  // a tail bucket with many narrow symbols: bisection
              int lo = s_found + 1, hi = s_last;
              while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (lds_u16(row_sa + 2u * mid) <= cum) lo = mid;
                else hi = mid - 1;
              }
              s_found = lo;
            } else {
              do {
                ++s_found;
              } while (s_found < s_last && lds_u16(row_sa + 2u * s_found + 2u) <= cum);
            }
            start = lds_u16(row_sa + 2u * s_found);
            nxt = lds_u16(row_sa + 2u * s_found + 2u);
            nxt = (nxt && s_found + 1 < n) ? nxt : 65536u;
          }
          const uint32_t freq = nxt - start, off = cum - start;
          rbad = rbad || n < 1 || off >= freq;  // (freq == 0 and cum outside [start, nxt) both show up as off >= freq)
          rx = static_cast<uint64_t>(freq) * (rx >> 16) + off;
          if (rx < kRansL && (wbase + widx) < rwords) {
            rx = (rx << 32) | w_spec;
            ++widx;
          }
          int value = s_found;
          if (s_found == n - 1) {
            // bypass (rare): nibble count (unary in chunks of 15), then the nibbles, least significant first
            rpos = wbase + widx;
            auto get4 = [&]() -> int {
              const int v4 = static_cast<int>(rx & 15u);
              rx >>= 4;
              if (rx < kRansL && rpos < rwords) rx = (rx << 32) | next_word();
              return v4;
            };
            int v4 = get4();
            int n_nib = v4;
            while (v4 == 15 && n_nib < 64) {
              v4 = get4();
              n_nib += v4;
            }
            uint32_t raw = 0;
            for (int k2 = 0; k2 < n_nib; ++k2) {
              const uint32_t nib = static_cast<uint32_t>(get4());
              if (k2 < 8) raw |= nib << (4 * k2);
            }
            value = static_cast<int>(raw >> 1);
            if (raw & 1u) value = -value - 1;
            else value += n - 1;
            widx = rpos - wbase;
          }
          value = rbad ? 0 : value + dc.w;
          if (lane == 0) sts_b32(sym_sa + 4u * static_cast<uint32_t>(ch), value);
          dc = dn;
        }
        rpos = wbase + widx;
        __syncwarp();
        for (int ch = lane; ch < C; ch += 32) {
          const int sv = s_sym[ch];
          const float mu = __ldcg(p.mu_buf + static_cast<long long>(j) * C + ch);
          p.t_hat[e0i + ch] = static_cast<float>(sv) + mu;
          if (p.sym) p.sym[e0i + ch] = sv;
        }
      }
      __syncthreads();
      tick(8);
      if (!grid_barrier(p.sync, epoch, &s_flag)) return;
      tick(9);
    }
  }
  if (j == 0 && tid == 0) {
    long long* out = reinterpret_cast<long long*>(p.sync + 16);
    for (int i = 0; i < 10; ++i) out[i] = prof[i];
  }
  if (p.mode == 1 && j < p.batch && warp == 0 && lane == 0) p.status[j] = rbad ? 1 : 0;
  // normal completion (every early return above is an aborted grid barrier): the host checks this flag
  if (j == 0 && tid == 0) p.sync[2] = 1u;
}

int fill_params(const stemb200_ar_desc* d, ArParams& p) {
  if (!d) return set_error("ar: null descriptor");
  if (d->batch < 1 || d->batch > kArCtas || d->h < 1 || d->w < 1) return set_error("ar: bad shape (batch <= 64)");
  if (d->c < 64 || d->c % 64 || d->c > 320 || d->l1 < 64 || d->l1 % 64 || d->l2 < 64 || d->l2 % 64)
    return set_error("ar: c (<= 256), l1, l2 must be multiples of 64");
  if (d->n_scales < 2 || d->n_scales > kArMaxScales) return set_error("ar: scale table needs 2..256 entries");
  p.batch = d->batch;
  p.h = d->h;
  p.w = d->w;
  p.c = d->c;
  p.l1 = d->l1;
  p.l2 = d->l2;
  p.rc = 2 * d->c / kArCtas;
  p.r1 = d->l1 / kArCtas;
  p.r2 = d->l2 / kArCtas;
  p.rg = d->c / kArCtas;
  if (p.rc > kArMaxRows || p.r1 > kArMaxRows || p.r2 > kArMaxRows || 2 * p.rg > kArMaxRows)
    return set_error("ar: layer too wide (at most 12 rows per CTA: 2c, l1, l2 <= 768)");
  p.kmax = std::max(std::max(kArTaps * d->c, 2 * d->c), std::max(d->l1, d->l2));
  p.n_scales = d->n_scales;
  p.slope = d->slope;
  p.scale_bound = d->scale_bound;
  p.block_floats = static_cast<int>(stemb200_ar_packed_floats(d) / kArCtas);
  return 0;
}

long long pmax_of(const stemb200_ar_desc* d) { return static_cast<long long>(d->batch) * ((d->w + 2) / 3 + 1); }

struct Scratch {
  float *ctx, *h1, *h2, *mu;
  unsigned int* sync;
};

Scratch carve(const stemb200_ar_desc* d, void* ws) {
  const long long pm = pmax_of(d);
  float* f = static_cast<float*>(ws);
  Scratch s;
  s.sync = reinterpret_cast<unsigned int*>(f);
  f += 64;
  s.ctx = f;
  f += pm * 2 * d->c;
  s.h1 = f;
  f += pm * d->l1;
  s.h2 = f;
  f += pm * d->l2;
  s.mu = f;
  return s;
}

int launch_ar(ArParams& p, cudaStream_t st) {
  const size_t smem = (static_cast<size_t>((p.block_floats + 3) & ~3) + static_cast<size_t>(p.nstage) * p.kmax) * 4 +
                      static_cast<size_t>(p.cdf16_entries + 8) * 2;
  static size_t static_smem = [] {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, ar_codec_kernel) == cudaSuccess ? a.sharedSizeBytes : size_t(40 * 1024);
  }();
  if (smem + static_smem > 227 * 1024) {
    char buf[160];
    snprintf(buf, sizeof(buf), "ar: weights + staging + CDF table need %zu + %zu bytes of shared memory (limit 232448)",
             smem, static_smem);
    return set_error(buf);
  }
  static size_t configured = 0;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(ar_codec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(ar_codec)", e);
    configured = smem;
  }
  if (num_sms() < kArCtas) return set_error("ar: needs at least 64 SMs (co-resident persistent grid)");
  cudaError_t e = cudaMemsetAsync(p.sync, 0, 64 * sizeof(float), st);
  if (e != cudaSuccess) return set_cuda_error("ar: memset", e);
  void* args[] = {&p};
  e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(ar_codec_kernel), dim3(kArCtas), dim3(kArThreads), args, smem,
                                  st);
  count_launch();
  if (e != cudaSuccess) return set_cuda_error("ar_codec launch", e);
  return 0;
}

}  // namespace
}  // namespace stem

using namespace stem;

extern "C" int64_t stemb200_ar_packed_floats(const stemb200_ar_desc* d) {
  if (!d || d->c < 64 || d->c % 64 || d->l1 % 64 || d->l2 % 64 || d->l1 < 64 || d->l2 < 64) return STEMB200_E_INVALID;
  const int64_t rc = 2 * d->c / kArCtas, r1 = d->l1 / kArCtas, r2 = d->l2 / kArCtas, rg = d->c / kArCtas;
  const int64_t per = rc * kArTaps * d->c + rc + r1 * 2 * d->c + r2 * d->l1 + r2 + 2 * rg * d->l2 + 2 * rg;
  return per * kArCtas;
}

extern "C" int64_t stemb200_ar_workspace_bytes(const stemb200_ar_desc* d) {
  if (!d || d->batch < 1 || d->w < 1) return STEMB200_E_INVALID;
  const long long pm = pmax_of(d);
  return (64 + pm * (2LL * d->c + d->l1 + d->l2 + d->c)) * 4;
}

extern "C" int stemb200_ar_encode(const stemb200_ar_desc* d, const float* packed, const float* e0,
                                  const float* target, const float* scale_table, float* t_hat, int32_t* symbols,
                                  int32_t* indexes, float* params_out, void* workspace, void* stream) {
  if (!packed || !e0 || !target || !scale_table || !t_hat || !indexes || !workspace)
    return set_error("ar_encode: null argument");
  ArParams p{};
  if (int rc = fill_params(d, p)) return rc;
  const Scratch s = carve(d, workspace);
  p.mode = 0;
  p.nstage = kArWarps;
  p.cdf16_entries = 0;
  p.packed = packed;
  p.e0 = e0;
  p.target = target;
  p.table = scale_table;
  p.t_hat = t_hat;
  p.sym = symbols;
  p.idx = indexes;
  p.params_out = params_out;
  p.ctx_buf = s.ctx;
  p.h1_buf = s.h1;
  p.h2_buf = s.h2;
  p.mu_buf = s.mu;
  p.sync = s.sync;
  return launch_ar(p, static_cast<cudaStream_t>(stream));
}

extern "C" int stemb200_ar_decode(const stemb200_ar_desc* d, const float* packed, const float* e0,
                                  const float* scale_table, const uint8_t* streams, const int64_t* stream_off,
                                  const int64_t* stream_len, const int32_t* cdfs, int32_t n_cdfs, int32_t cdf_stride,
                                  const int32_t* cdf_sizes, const int32_t* offsets, int32_t cdf_total_entries,
                                  float* t_hat, int32_t* symbols, int32_t* indexes, float* params_out,
                                  int32_t* status, void* workspace, void* stream) {
  if (!packed || !e0 || !scale_table || !streams || !stream_off || !stream_len || !cdfs || !cdf_sizes || !offsets ||
      !t_hat || !indexes || !status || !workspace || n_cdfs < 1 || n_cdfs > kArMaxCdfs || cdf_stride < 2 ||
      cdf_total_entries < 2 * n_cdfs || cdf_total_entries > n_cdfs * cdf_stride)
    return set_error("ar_decode: null / bad argument (at most 64 CDF rows)");
  ArParams p{};
  if (int rc = fill_params(d, p)) return rc;
  const Scratch s = carve(d, workspace);
  p.mode = 1;
  p.nstage = 2;
  p.cdf16_entries = cdf_total_entries;
  p.packed = packed;
  p.e0 = e0;
  p.table = scale_table;
  p.t_hat = t_hat;
  p.sym = symbols;
  p.idx = indexes;
  p.params_out = params_out;
  p.ctx_buf = s.ctx;
  p.h1_buf = s.h1;
  p.h2_buf = s.h2;
  p.mu_buf = s.mu;
  p.sync = s.sync;
  p.streams = streams;
  p.stream_off = stream_off;
  p.stream_len = stream_len;
  p.cdf = cdfs;
  p.cdf_size = cdf_sizes;
  p.cdf_off = offsets;
  p.cdf_stride = cdf_stride;
  p.n_cdfs = n_cdfs;
  p.status = status;
  return launch_ar(p, static_cast<cudaStream_t>(stream));
}
