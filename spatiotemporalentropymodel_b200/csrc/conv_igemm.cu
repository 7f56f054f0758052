// Implicit-GEMM convolution / transposed convolution on tcgen05 (sm_100a).
//
//   D[pixel, c_out] = sum_{tap, c_in} A[pixel + tap offset, c_in] * W[c_out, tap, c_in]
//
// * A tiles (128 output pixels x 64 input channels, fp16) are fetched by TMA straight from the NHWC activation
//   tensor: a 4-D box {64 ch, tile_w, tile_h, 1} at (c0, w0+dw, h0+dh, n). Out-of-bounds box elements are
//   zero-filled by the TMA unit, which implements the convolution's zero padding for free.
// * stride-2 convolutions read four "phase" views (even/odd rows x even/odd columns) of the input, each a
//   strided tensor map, so every tap is again a unit-stride box.
// * transposed convolutions (k, s2, p=k/2, op=1) are four independent sub-problems, one per output phase
//   (p, q): a stride-1 conv with the taps kh == p+pad (mod 2), written through a strided output tensor map.
// * torch.cat(..., 1) inputs are K-segments: the K loop walks (tap, source, 64-channel chunk) triples listed
//   in a table carried in the kernel parameters.
// * layers with enough tiles run as 2-CTA clusters in duo mode: one tcgen05 cta_group::2 MMA (M = 256) per K slice for
//   the pair, half a weight tile staged per CTA (ptx.cuh); STEMB200_DUO=0 selects pair mode (multicast weight tiles).
// * accumulators live in TMEM (double buffered, 2 x BLOCK_N columns), MMAs are issued by one thread,
//   the epilogue (8 warps) overlaps the next tile's main loop; persistent CTAs, one per SM.
//
// Reference call sites this replaces: compressai/models/utils.py:112-130, spatiotemporalpriors.py:523-554,
// layers/layers.py:44-47, layers/gdn.py:52-67 (F.conv2d(x**2, gamma, beta) + rsqrt/sqrt + multiply).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/stemb200.h"
#include "gc_math.cuh"
#include "internal.h"
#include "ptx.cuh"

namespace stem {

constexpr int kMaxKSteps = 256;
constexpr int kKChunk = 64;         // fp16 elements per 128-byte swizzle row
constexpr int kAStageBytes = 128 * 128;
constexpr int kOutStageBytes = 128 * 128;
constexpr int kNumThreads = 384;     // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps
constexpr int kNumEpiThreads = 256;
constexpr int kSmemLimit = 232448;  // 227 KB

struct ConvKernelParams {
  alignas(64) CUtensorMap a_map[4];
  alignas(64) CUtensorMap b_map;
  alignas(64) CUtensorMap out_map[4];
  uint32_t ksteps[kMaxKSteps];  // [1:0] map | [5:2] dh+8 | [9:6] dw+8 | [31:10] channel offset
  int sub_kbeg[4], sub_kend[4];
  int sub_p[4], sub_q[4];
  int n_sub;
  int batch, h_out, w_out;  // output grid of one sub-problem
  int tile_h, tile_w, tiles_h, tiles_w;
  int n_tiles_n, total_tiles;
  int c_out;
  float slope, sq_scale, sq_inv;
  int out_f32, direct;
  int os, full_h, full_w;  // output phase stride and full output size (direct store / aux addressing)
  const float* bias;
  const __half* aux;  // EPI 1/2: NHWC fp16 multiplicand / residual with the output's geometry
  void* out;
  // pair mode (csize == 2): the two CTAs of a cluster work on neighbouring pixel tiles of the same N tile and each
  // fetches half of every weight tile, multicast into both (L2 -> SM bytes per k-step 40 KB -> 28 KB at N = 192)
  alignas(64) CUtensorMap b_half_map;  // box {64, BLOCK_N / 2}
  alignas(64) CUtensorMap g_half_map;
  int csize;
  // fused last synthesis layer (conv_gdn_kernel<.., kLast = true>): every output pixel of this layer is multiplied by
  // W6 [96][c_out] (deconv(N, 3, k5 s2) as 75 (+21 zero) per-pixel contributions (r, s, c)), the "col" rows go to HBM
  alignas(64) CUtensorMap w6_map;
  alignas(64) CUtensorMap w6_half_map;  // duo mode: box {64, 48}
  int duo;  // conv_gdn_kernel, csize == 2: cta_group::2 MMAs (see ptx.cuh)
  __half* col_out;  // [batch][full_h][full_w][96] fp16
  int store_act;    // 0: the layer's own activation is not written at all
  int kk_main;  // K = 16 slices issued per main-loop k-step (4; 3 in row_taps mode: 5 taps x 8 channels = 40 <= 48)
  // fused GDN / IGDN (conv_gdn_kernel only)
  alignas(64) CUtensorMap g_map;  // gamma [c_out][c_out] fp16, K-major
  const float* beta;
  int igdn;
};

template <int BLOCK_N, bool kDuo = false>
struct ConvCfg {
  static constexpr int kBStageBytes = BLOCK_N * 128;  // one K chunk of the weights (all BLOCK_N rows)
  // duo mode: a CTA stages only its half of every B tile: smaller ring slots, more k-steps in flight (see GdnCfgT)
#ifdef STEMB200_SHALLOW_DUO_RING
  static constexpr int kBSlotBytes = kBStageBytes;
#else
  static constexpr int kBSlotBytes = kDuo ? kBStageBytes / 2 : kBStageBytes;
#endif
  static constexpr int kStageBytes = kAStageBytes + kBSlotBytes;
  static constexpr int kNumOutBufs = 2;
  static constexpr int kBarrierBytes = 256 + BLOCK_N * 4;  // mbarriers + TMEM slot, then the tile's bias slice
  static constexpr int kFree = kSmemLimit - 1024 - kNumOutBufs * kOutStageBytes - kBarrierBytes;
  static constexpr int kStagesRaw = kFree / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes =
      1024 + kStages * kStageBytes + kNumOutBufs * kOutStageBytes + kBarrierBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32)    ? 32
                                   : (2 * BLOCK_N <= 64)  ? 64
                                   : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256
                                                          : 512;
  static_assert(kStages >= 3, "pipeline too shallow");
  static_assert(2 * BLOCK_N <= 512, "TMEM overflow");
};

struct TileCoord {
  int sub, n_img, h0, w0, n0;
};

// q = work item of this CTA's cluster (pair mode: a pair of pixel tiles), r = rank of the CTA inside the cluster
__device__ __forceinline__ TileCoord decode_tile(const ConvKernelParams& p, int q, int r, int block_n) {
  TileCoord t;
  int nt = q % p.n_tiles_n;
  int u = q / p.n_tiles_n;
  int m;
  if (p.n_sub == 4) {
    // transposed conv: the four output phases of a pixel tile are neighbours in the work order (their input tile is
    // read from L2, not four times from HBM); the phase rotates with the tile so that every CTA sees all four K lengths
    const int g = u >> 2;
    t.sub = ((u & 3) + g) & 3;
    m = g * p.csize + r;
  } else {
    m = u * p.csize + r;
    t.sub = 0;
  }
  int twi = m % p.tiles_w;
  m /= p.tiles_w;
  int thi = m % p.tiles_h;
  m /= p.tiles_h;
  t.n_img = m % p.batch;
  t.h0 = thi * p.tile_h;
  t.w0 = twi * p.tile_w;
  t.n0 = nt * block_n;
  return t;
}

// the MMA warp only needs the sub-problem (K range) of a work item: no divisions on the tensor pipe's critical path
__device__ __forceinline__ int tile_sub(const ConvKernelParams& p, int q) {
  if (p.n_sub != 4) return 0;
  const int u = p.n_tiles_n == 1 ? q : q / p.n_tiles_n;
  return ((u & 3) + (u >> 2)) & 3;
}

// single-MUFU reciprocal square root / square root (rel. error ~2^-22; the result is rounded to fp16 anyway)
__device__ __forceinline__ float approx_rsqrt(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float approx_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// EPI: 0 = LeakyReLU(acc + bias); 1 = SFT: x * (acc_gamma + b_gamma) + (acc_beta + b_beta) then LeakyReLU, with the
// accumulator columns laid out per 128-column block as [gamma(64 ch) | beta(64 ch)] and x read from `aux`
// (stem_utils.py:36-43, the "+1" of (1 + gamma) is folded into b_gamma); 2 = LeakyReLU(acc + bias) + aux (residual).
template <int BLOCK_N, int EPI, bool kDuo = false>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvKernelParams p) {
  using Cfg = ConvCfg<BLOCK_N, kDuo>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base;
  const uint32_t out_base = smem_base + kStages * Cfg::kStageBytes;
  const uint32_t bar_base = out_base + Cfg::kNumOutBufs * kOutStageBytes;
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  const uint32_t bias_smem = bar_base + 256u;  // BLOCK_N floats: bias (or GDN beta) of the current N tile

  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps descriptors,
  // barrier addresses and TMA / MMA operands in uniform registers (no per-instruction elect / broadcast loops)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool pair = p.csize == 2;
  const int crank = pair ? static_cast<int>(cluster_ctarank()) : 0;
  // duo mode (ptx.cuh): one cta_group::2 MMA per k-step for the pair, issued by rank 0; half a B tile per CTA
  constexpr bool duo = kDuo;
  const int q_first = pair ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int q_stride = pair ? static_cast<int>(cluster_count_x()) : static_cast<int>(gridDim.x);
  const int n_items = p.total_tiles / p.csize;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
    if (!p.direct) {
      for (int i = 0; i < p.n_sub; ++i) tma_prefetch_desc(&p.out_map[i]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      // pair mode: the peer's MMAs must also be done with the slot it multicasts into; duo: one (multicast) commit
      mbar_init(empty_bar(s), duo ? 1 : p.csize);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), duo ? 16 : 8);  // one arrival per epilogue warp (duo: of both CTAs, at the leader)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (duo) {
      tmem_alloc2(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();  // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const uint32_t a_tx_bytes = static_cast<uint32_t>(p.tile_h * p.tile_w) * 128u;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride) {
      const TileCoord t = decode_tile(p, tile, crank, BLOCK_N);
      const int kbeg = p.sub_kbeg[t.sub], kend = p.sub_kend[t.sub];
      for (int k = kbeg; k < kend; ++k) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t e = p.ksteps[k];
        const int map = e & 3;
        const int dh = static_cast<int>((e >> 2) & 15u) - 8;
        const int dw = static_cast<int>((e >> 6) & 15u) - 8;
        const int c0 = static_cast<int>(e >> 10);
        const uint32_t a_dst = stage_base + s * Cfg::kStageBytes;
        const uint32_t b_dst = a_dst + kAStageBytes;
        if (leader && duo) {
          const uint32_t lbar = mapa_shared(full_bar(s), 0);
          if (crank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * (a_tx_bytes + Cfg::kBStageBytes / 2));
          tma_load_4d_2sm(a_dst, &p.a_map[map], lbar, c0, t.w0 + dw, t.h0 + dh, t.n_img);
          tma_load_2d_2sm(b_dst, &p.b_half_map, lbar, k * kKChunk, t.n0 + crank * (BLOCK_N / 2));
        } else if (leader) {
          mbar_arrive_expect_tx(full_bar(s), a_tx_bytes + Cfg::kBStageBytes);
          tma_load_4d(a_dst, &p.a_map[map], full_bar(s), c0, t.w0 + dw, t.h0 + dh, t.n_img);
          if (pair)
            tma_load_2d_mc(b_dst + crank * (Cfg::kBStageBytes / 2), &p.b_half_map, full_bar(s), k * kKChunk,
                           t.n0 + crank * (BLOCK_N / 2), 3);
          else
            tma_load_2d(b_dst, &p.b_map, full_bar(s), k * kKChunk, t.n0);
        }
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1 && (!duo || crank == 0)) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
    // (duo mode: only the cluster's rank 0 issues; its MMAs write the accumulators of both CTAs)
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc(/*F16*/ 0u, duo ? 256u : 128u, BLOCK_N);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const int sub = tile_sub(p, tile);
      const int kbeg = p.sub_kbeg[sub], kend = p.sub_kend[sub];
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      if (duo) mbar_wait_cluster(tempty_bar(acc), accph ^ 1u);
      else mbar_wait(tempty_bar(acc), accph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int k = kbeg; k < kend; ++k) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = stage_base + s * Cfg::kStageBytes;
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(a_addr + kAStageBytes);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            // +32 bytes (one K=16 slice) inside the 128-byte swizzle row => +2 in the >>4 address field
            if (duo) mma_f16_ss2(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (k > kbeg || kk > 0) ? 1u : 0u);
            else mma_f16_ss(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (k > kbeg || kk > 0) ? 1u : 0u);
          }
          if (duo) mma_commit2_mc(empty_bar(s), 3);
          else if (pair) mma_commit_mc(empty_bar(s), 3);
          else mma_commit(empty_bar(s));
        }
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (leader) {
        if (duo) mma_commit2_mc(tfull_bar(acc), 3);
        else mma_commit(tfull_bar(acc));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (TMEM -> registers -> smem -> TMA store) =====================
    // 8 warps: warp w reads TMEM lane group w % 4; warps 4-7 take the even 32-column chunks, warps 8-11 the odd
    // ones, so two warps per scheduler hide each other's dependency stalls (one warp per scheduler ran at
    // ~0.2 IPC, profiles/r01_ncu_gdn_fused_epilogue.txt).
    const int ew = warp & 3;
    const int etid = threadIdx.x - 128;  // 0..255
    const int row = etid & 127;          // accumulator row == pixel of the patch
    const int half = etid >> 7;          // which 32-column chunk of each 64-column group
    const int npix = p.tile_h * p.tile_w;
    const uint32_t rsw = static_cast<uint32_t>(row & 7);
    uint32_t group_ctr = 0;
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const TileCoord t = decode_tile(p, tile, crank, BLOCK_N);
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      const int th = row / p.tile_w, tw = row - th * p.tile_w;
      const int oh = t.h0 + th, ow = t.w0 + tw;
      const bool inb = (row < npix) && (oh < p.h_out) && (ow < p.w_out);
      // flat pixel index in the full-resolution output (direct store)
      const long long pix =
          (static_cast<long long>(t.n_img) * p.full_h + (oh * p.os + p.sub_p[t.sub])) * p.full_w +
          (ow * p.os + p.sub_q[t.sub]);

      // stage this tile's bias slice in smem (a global load per column used to be the epilogue's critical
      // path: ~200 cycles of exposed latency each, profiles/r01_ncu_gdn_epilogue.txt)
      named_bar_sync(1, kNumEpiThreads);  // every thread is done with the previous tile's slice
      for (int i = etid; i < BLOCK_N; i += kNumEpiThreads) {
        const float b = __ldg(p.bias + t.n0 + i);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + 4u * i), "f"(b) : "memory");
      }
      named_bar_sync(1, kNumEpiThreads);

      mbar_wait(tfull_bar(acc), accph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(ew * 32) << 16);

      constexpr int kCols = (BLOCK_N >= 32) ? 32 : 16;
      constexpr int kChunks = BLOCK_N / kCols;
      if constexpr (EPI == 1) {
        // ================= SFT epilogue: 128 accumulator columns -> 64 output channels per group =================
        const int c_half = p.c_out >> 1;  // channels of the output / aux tensors
#pragma unroll 1
        for (int g = 0; g < BLOCK_N / 128; ++g) {
          const int cg = 128 * g + 32 * half;  // gamma columns of this thread; beta columns are cg + 64
          const int oc = ((t.n0 + 128 * g) >> 1) + 32 * half;  // first output channel of this thread
          uint4 a4[4];
          if (inb) {
            const uint4* ap = reinterpret_cast<const uint4*>(p.aux + pix * c_half + oc);
#pragma unroll
            for (int j = 0; j < 4; ++j) a4[j] = __ldg(ap + j);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) a4[j] = make_uint4(0, 0, 0, 0);
          }
          uint32_t rg[32], rb[32];
          tmem_ld_32x32(t_row + cg, rg);
          tmem_ld_32x32(t_row + cg + 64, rb);
          tmem_ld_wait();
          uint32_t ho[16];
          const uint32_t* ax = reinterpret_cast<const uint32_t*>(a4);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g0, g1, g2, g3, b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(g0), "=f"(g1), "=f"(g2), "=f"(g3)
                         : "r"(bias_smem + 4u * (cg + 4 * j)));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                         : "r"(bias_smem + 4u * (cg + 64 + 4 * j)));
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&ax[2 * j]));
            const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&ax[2 * j + 1]));
            float o0 = fmaf(x0.x, __uint_as_float(rg[4 * j]) + g0, __uint_as_float(rb[4 * j]) + b0);
            float o1 = fmaf(x0.y, __uint_as_float(rg[4 * j + 1]) + g1, __uint_as_float(rb[4 * j + 1]) + b1);
            float o2 = fmaf(x1.x, __uint_as_float(rg[4 * j + 2]) + g2, __uint_as_float(rb[4 * j + 2]) + b2);
            float o3 = fmaf(x1.y, __uint_as_float(rg[4 * j + 3]) + g3, __uint_as_float(rb[4 * j + 3]) + b3);
            o0 = o0 > 0.f ? o0 : o0 * p.slope;
            o1 = o1 > 0.f ? o1 : o1 * p.slope;
            o2 = o2 > 0.f ? o2 : o2 * p.slope;
            o3 = o3 > 0.f ? o3 : o3 * p.slope;
            ho[2 * j] = pack_half2(o0, o1);
            ho[2 * j + 1] = pack_half2(o2, o3);
          }
          const uint32_t obuf = out_base + (group_ctr & 1u) * kOutStageBytes;
          if (etid == 0) tma_store_wait_read<1>();
          named_bar_sync(1, kNumEpiThreads);
          const uint32_t rbase = obuf + static_cast<uint32_t>(row) * 128u;
          const uint32_t jo = half ? 4u : 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t addr = rbase + (((static_cast<uint32_t>(j) + jo) ^ rsw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(ho[4 * j]), "r"(ho[4 * j + 1]),
                         "r"(ho[4 * j + 2]), "r"(ho[4 * j + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          named_bar_sync(1, kNumEpiThreads);
          if (etid == 0) {
            tma_store_4d(&p.out_map[t.sub], obuf, (t.n0 + 128 * g) >> 1, t.w0, t.h0, t.n_img);
            tma_store_commit();
          }
          ++group_ctr;
        }
      } else {
        // ================= linear (+ residual) epilogue =================
        // TMEM loads are software pipelined: chunk ci+2 is requested before chunk ci is processed. Both halves
        // run the same number of iterations (barriers!); with an odd chunk count (BLOCK_N = 160) the last
        // iteration of half 1 is idle and its stale half-row lands on out-of-range channels, which TMA clips.
        constexpr int kIters = (kChunks + 1) / 2;
        uint32_t rn[kCols];
        if (half < kChunks) {
          if constexpr (kCols == 32) tmem_ld_32x32(t_row + half * kCols, rn);
          else tmem_ld_32x16(t_row + half * kCols, rn);
        }
#pragma unroll 1
        for (int gi = 0; gi < kIters; ++gi) {
          const int ci = 2 * gi + half;
          const bool active = ci < kChunks;
          const int c = ci * kCols;
          const int ch0 = t.n0 + c;
          float v[kCols];
          if (active) {
            uint4 a4[kCols / 8];
            if constexpr (EPI == 2) {
              if (inb) {
                const uint4* ap = reinterpret_cast<const uint4*>(p.aux + pix * p.c_out + ch0);
#pragma unroll
                for (int j = 0; j < kCols / 8; ++j) a4[j] = __ldg(ap + j);
              } else {
#pragma unroll
                for (int j = 0; j < kCols / 8; ++j) a4[j] = make_uint4(0, 0, 0, 0);
              }
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < kCols; ++i) v[i] = __uint_as_float(rn[i]);
            if (ci + 2 < kChunks) {
              if constexpr (kCols == 32) tmem_ld_32x32(t_row + c + 2 * kCols, rn);
              else tmem_ld_32x16(t_row + c + 2 * kCols, rn);
            }
#pragma unroll
            for (int j = 0; j < kCols / 4; ++j) {
              float b0, b1, b2, b3;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                           : "r"(bias_smem + 4u * (c + 4 * j)));
              const float x0 = v[4 * j] + b0, x1 = v[4 * j + 1] + b1, x2 = v[4 * j + 2] + b2, x3 = v[4 * j + 3] + b3;
              v[4 * j] = x0 > 0.f ? x0 : x0 * p.slope;
              v[4 * j + 1] = x1 > 0.f ? x1 : x1 * p.slope;
              v[4 * j + 2] = x2 > 0.f ? x2 : x2 * p.slope;
              v[4 * j + 3] = x3 > 0.f ? x3 : x3 * p.slope;
            }
            if constexpr (EPI == 2) {
              const uint32_t* ax = reinterpret_cast<const uint32_t*>(a4);
#pragma unroll
              for (int i = 0; i < kCols / 2; ++i) {
                const float2 r2 = __half22float2(*reinterpret_cast<const __half2*>(&ax[i]));
                v[2 * i] += r2.x;
                v[2 * i + 1] += r2.y;
              }
            }
          }

          if (p.direct) {
            if (inb && active) {
              if (p.out_f32) {
                float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + pix * p.c_out + ch0);
#pragma unroll
                for (int j = 0; j < kCols / 4; ++j)
                  op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              } else {
                uint4* op = reinterpret_cast<uint4*>(static_cast<__half*>(p.out) + pix * p.c_out + ch0);
#pragma unroll
                for (int j = 0; j < kCols / 8; ++j)
                  op[j] = make_uint4(pack_half2(v[8 * j], v[8 * j + 1]), pack_half2(v[8 * j + 2], v[8 * j + 3]),
                                     pack_half2(v[8 * j + 4], v[8 * j + 5]),
                                     pack_half2(v[8 * j + 6], v[8 * j + 7]));
              }
            }
          } else if constexpr (kCols == 32) {
            // ---- staged TMA store, one 64-column group (both halves) per iteration ----
            //   fp16: the two halves fill pieces 0-3 / 4-7 of one 128-byte row; buffers alternate per group
            //   fp32: each half fills its own buffer (32 columns = one 128-byte row)
            const uint32_t buf = p.out_f32 ? static_cast<uint32_t>(half) : (group_ctr & 1u);
            const uint32_t obuf = out_base + buf * kOutStageBytes;
            if (etid == 0) {
              if (p.out_f32) tma_store_wait_read<0>();
              else tma_store_wait_read<1>();
            }
            named_bar_sync(1, kNumEpiThreads);
            const uint32_t rbase = obuf + static_cast<uint32_t>(row) * 128u;
            if (active) {
              if (p.out_f32) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const uint32_t addr = rbase + ((static_cast<uint32_t>(j) ^ rsw) << 4);
                  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]),
                               "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                               : "memory");
                }
              } else {
                const uint32_t jo = half ? 4u : 0u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t addr = rbase + (((static_cast<uint32_t>(j) + jo) ^ rsw) << 4);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr),
                               "r"(pack_half2(v[8 * j], v[8 * j + 1])),
                               "r"(pack_half2(v[8 * j + 2], v[8 * j + 3])),
                               "r"(pack_half2(v[8 * j + 4], v[8 * j + 5])),
                               "r"(pack_half2(v[8 * j + 6], v[8 * j + 7]))
                               : "memory");
                }
              }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, kNumEpiThreads);
            if (etid == 0) {
              const int cgrp = t.n0 + 64 * gi;
              if (p.out_f32) {
                tma_store_4d(&p.out_map[t.sub], out_base, cgrp, t.w0, t.h0, t.n_img);
                if (2 * gi + 1 < kChunks)
                  tma_store_4d(&p.out_map[t.sub], out_base + kOutStageBytes, cgrp + 32, t.w0, t.h0, t.n_img);
              } else {
                const int omap = (BLOCK_N == 160 && t.n0 + BLOCK_N < p.c_out) ? 1 : t.sub;
                tma_store_4d(&p.out_map[omap], obuf, cgrp, t.w0, t.h0, t.n_img);
              }
              tma_store_commit();
            }
            ++group_ctr;
          }
        }
      }
      // all TMEM reads of this accumulator are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (duo) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
      }
    }
    if (etid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (warp == 2) {
    if (duo) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =====================================================================================================
// conv / deconv with GDN or IGDN fused into the epilogue (C_out == 192 or 128 == one N tile)
//
//   x   = conv(in) + bias                    accumulator A[it & 1] in TMEM (double buffered)
//   n   = beta + gamma . x^2                 second contraction on the tensor core: (x s)^2 (fp16, s = sq_scale a
//                                            power of two) is written by the epilogue warps into smem as a K-major
//                                            operand, gamma chunks stream through the B slots of the same TMA ring;
//                                            the result OVERWRITES A[it & 1] (phase 1 has drained it by then)
//   out = x * rsqrt(n)  (GDN)  |  x * sqrt(n)  (IGDN); x s is parked as packed fp16 in TMEM columns [2N, 2.5N)
//
// Schedule (tile it of this CTA):
//   MMA thread : wait accfree[it&1] -> first half of main(it) -> [wait a2rdy(it-1); gamma(it-1); commit nfull(it-1)]
//                -> second half of main(it) -> commit tfull(it)            (gamma of the last tile trails the loop)
//   epilogue   : wait tfull(it) -> phase 1 (A -> x s stash + (x s)^2 smem) -> arrive a2rdy(it)
//                -> wait nfull(it) -> phase 2 (normalise, arrive accfree[it&1], TMA store)
// so BOTH epilogue phases of tile it overlap the main loops of tiles it+1 / it+2; the tensor pipe only idles when
// the epilogue chain itself (phase 1 + gamma latency + phase 2) is longer than a main loop (K <= ~20 k-steps).
// With n' = s^2 n the normalisation is out = (x s) * rsqrt(s^2 beta + acc) for GDN and
// (x s) * sqrt(beta / s^2 + acc / s^4) for IGDN: one FFMA + one MUFU + one FMUL per element.
// Replaces layers/gdn.py:52-67 + the producing conv (priors.py:421-439) without x or x^2 ever reaching HBM.
// =====================================================================================================
constexpr int kLastN = 96;  // columns of the fused last-layer GEMM (75 real)

template <int kNT, bool kLast = false, bool kDuo = false, bool kTm = false>
struct GdnCfgT {
  static constexpr int kN = kNT;
  static constexpr int kBStageBytes = kN * 128;  // one K chunk of the weights (all kN rows)
  // duo mode: a CTA stages only its half of every B tile, so a ring slot is 28 KB instead of 40 KB and the same shared
  // memory holds more k-steps in flight.  The main loop is bound by bytes in flight / TMA latency (~1 250 cycles from
  // L2 under the step's 14 TB/s of operand traffic): 3 x 40 KB -> one k-step per 610 cycles, 4 -> 460, 6 x 28 KB -> 384
  // (the tensor pipe's own rate).
#ifdef STEMB200_SHALLOW_DUO_RING  // A/B build: the round-1 ring (full-size slots, 4 / 3 stages) also in duo mode
  static constexpr bool kSmallSlots = false;
#else
  static constexpr bool kSmallSlots = kDuo;
#endif
  static constexpr int kBSlotBytes = kSmallSlots ? kBStageBytes / 2 : kBStageBytes;
  static constexpr int kStageBytes = kAStageBytes + kBSlotBytes;
  // kLast: the resident W6 operand (all of it, or this CTA's half in duo mode) shares the budget.
  // kTm (operands in tensor memory): no x^2 / staging buffers in shared memory, their room goes to a fourth stage.
  static constexpr int kStages = kSmallSlots ? ((kLast && !kTm) ? 5 : 6) : ((kLast && !kTm) ? 3 : 4);
  static constexpr int kA2Bytes = kTm ? 0 : (kN / 64) * kAStageBytes;  // x^2 operand (64-channel chunks) / staging
  static constexpr int kW6Bytes = kLast ? (kN / 64) * kLastN * 128 : 0;
  static constexpr int kW6SlotBytes = kSmallSlots ? kW6Bytes / 2 : kW6Bytes;
  static constexpr int kBarrierBytes = 256 + 2 * kN * 4;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kA2Bytes + kW6SlotBytes + kBarrierBytes;
  static constexpr int kTmemCols = 512;
  static constexpr uint32_t kStashCol = 2 * kN;  // x s stash; kTm: the packed fp16 operand ((x s)^2, then the output)
  static_assert(kN % 64 == 0 && 2 * kN + kN / 2 <= 512, "TMEM budget");
  static_assert(!kLast || kN / 2 >= kLastN, "the last-layer accumulator reuses the stash columns");
  static_assert(!kTm || kLast, "tensor-memory operands: fused last layer");
  static_assert(2 * kStages + 9 <= 32, "barrier slots");
  static_assert(kSmemBytes <= kSmemLimit, "smem overflow");
};

constexpr int kGdnEpiWarps = 16;                        // 4 per TMEM lane group, one 16-column quarter each
constexpr int kGdnEpiThreads = kGdnEpiWarps * 32;       // 512
constexpr int kGdnThreads = 128 + kGdnEpiThreads;       // 4 control warps + 16 epilogue warps

template <int kNT, bool kInverse, bool kLast = false, bool kDuo = false, bool kTm = false>
__global__ void __launch_bounds__(kGdnThreads, 1)
conv_gdn_kernel(const __grid_constant__ ConvKernelParams p) {
  using Cfg = GdnCfgT<kNT, kLast, kDuo, kTm>;
  constexpr int kStages = Cfg::kStages;
  constexpr int BLOCK_N = Cfg::kN;
  constexpr int kGChunks = BLOCK_N / kKChunk;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base;
  const uint32_t a2_base = smem_base + kStages * Cfg::kStageBytes;
  const uint32_t w6_base = a2_base + Cfg::kA2Bytes;  // kLast: W6 as K-major chunks [kN/64][96 rows][128 B]
  const uint32_t bar_base = w6_base + Cfg::kW6SlotBytes;
  constexpr uint32_t kW6ChunkBytes = (Cfg::kSmallSlots ? kLastN / 2 : kLastN) * 128;  // duo: this CTA's 48 of the 96 rows
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kStages);
  const uint32_t a2rdy_bar = bar_base + 8u * (2 * kStages + 1);
  const uint32_t nfull_bar = bar_base + 8u * (2 * kStages + 2);
  auto accfree_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 3 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 5);
  const uint32_t w6full_bar = bar_base + 8u * (2 * kStages + 6);  // kLast only
  const uint32_t a3rdy_bar = bar_base + 8u * (2 * kStages + 7);
  const uint32_t d3full_bar = bar_base + 8u * (2 * kStages + 8);
  const uint32_t bias_smem = bar_base + 256u;
  const uint32_t beta_smem = bias_smem + 4u * BLOCK_N;

  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps descriptors,
  // barrier addresses and TMA / MMA operands in uniform registers (no per-instruction elect / broadcast loops)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool pair = p.csize == 2;
  const int crank = pair ? static_cast<int>(cluster_ctarank()) : 0;
  // duo mode: one cta_group::2 MMA (M = 256) per k-step for the pair, issued by rank 0; each CTA stages its own A tile
  // and half of every B tile. The leader's barriers collect the TMA bytes and the epilogue arrivals of both CTAs.
  // (a template parameter: a kernel that contains cta_group::2 instructions can only be launched as a cluster)
  constexpr bool duo = kDuo;
  const int q_first = pair ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int q_stride = pair ? static_cast<int>(cluster_count_x()) : static_cast<int>(gridDim.x);
  const int n_items = p.total_tiles / p.csize;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.g_map);
    for (int i = 0; i < p.n_sub; ++i) tma_prefetch_desc(&p.out_map[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      // pair mode: the peer's MMAs must also be done with the slot it multicasts into; duo: one (multicast) commit
      mbar_init(empty_bar(s), duo ? 1 : p.csize);
    }
    const uint32_t epi_arrivals = duo ? 2 * kGdnEpiWarps : kGdnEpiWarps;
    mbar_init(tfull_bar, 1);
    mbar_init(a2rdy_bar, epi_arrivals);
    mbar_init(nfull_bar, 1);
    mbar_init(accfree_bar(0), epi_arrivals);
    mbar_init(accfree_bar(1), epi_arrivals);
    if constexpr (kLast) {
      mbar_init(w6full_bar, 1);
      mbar_init(a3rdy_bar, epi_arrivals);
      mbar_init(d3full_bar, 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (duo) {
      tmem_alloc2(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  if (warp >= 4) {
    // bias, and the folded normaliser offset: s^2 beta (GDN) or beta / s^2 (IGDN)
    const float kb = kInverse ? p.sq_inv : p.sq_scale * p.sq_scale;
    for (int i = threadIdx.x - 128; i < BLOCK_N; i += kGdnEpiThreads) {
      const float b = __ldg(p.bias + i) * p.sq_scale, g = __ldg(p.beta + i) * kb;  // bias pre-scaled: x s = fma(acc, s, b s)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + 4u * i), "f"(b) : "memory");
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(beta_smem + 4u * i), "f"(g) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();  // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const uint32_t a_tx_bytes = static_cast<uint32_t>(p.tile_h * p.tile_w) * 128u;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    auto load_main = [&](const TileCoord& t, int k) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t e = p.ksteps[k];
      const int map = e & 3;
      const int dh = static_cast<int>((e >> 2) & 15u) - 8;
      const int dw = static_cast<int>((e >> 6) & 15u) - 8;
      const int c0 = static_cast<int>(e >> 10);
      const uint32_t a_dst = stage_base + s * Cfg::kStageBytes;
      if (leader && duo) {
        const uint32_t lbar = mapa_shared(full_bar(s), 0);
        if (crank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * (a_tx_bytes + Cfg::kBStageBytes / 2));
        tma_load_4d_2sm(a_dst, &p.a_map[map], lbar, c0, t.w0 + dw, t.h0 + dh, t.n_img);
        tma_load_2d_2sm(a_dst + kAStageBytes, &p.b_half_map, lbar, k * kKChunk, crank * (BLOCK_N / 2));
      } else if (leader) {
        mbar_arrive_expect_tx(full_bar(s), a_tx_bytes + Cfg::kBStageBytes);
        tma_load_4d(a_dst, &p.a_map[map], full_bar(s), c0, t.w0 + dw, t.h0 + dh, t.n_img);
        if (pair)
          tma_load_2d_mc(a_dst + kAStageBytes + crank * (Cfg::kBStageBytes / 2), &p.b_half_map, full_bar(s),
                         k * kKChunk, crank * (BLOCK_N / 2), 3);
        else
          tma_load_2d(a_dst + kAStageBytes, &p.b_map, full_bar(s), k * kKChunk, 0);
      }
      if (++s == kStages) {
        s = 0;
        ph ^= 1u;
      }
    };
    auto load_gamma = [&]() {  // only the B slot of the stage is used
      for (int kc = 0; kc < kGChunks; ++kc) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t a_dst = stage_base + s * Cfg::kStageBytes;
        if (leader && duo) {
          const uint32_t lbar = mapa_shared(full_bar(s), 0);
          if (crank == 0) mbar_arrive_expect_tx(full_bar(s), Cfg::kBStageBytes);
          tma_load_2d_2sm(a_dst + kAStageBytes, &p.g_half_map, lbar, kc * kKChunk, crank * (BLOCK_N / 2));
        } else if (leader) {
          mbar_arrive_expect_tx(full_bar(s), Cfg::kBStageBytes);
          if (pair)
            tma_load_2d_mc(a_dst + kAStageBytes + crank * (Cfg::kBStageBytes / 2), &p.g_half_map, full_bar(s),
                           kc * kKChunk, crank * (BLOCK_N / 2), 3);
          else
            tma_load_2d(a_dst + kAStageBytes, &p.g_map, full_bar(s), kc * kKChunk, 0);
        }
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    };
    if constexpr (kLast) {
      if (leader && duo) {
        // each CTA holds 48 of W6's 96 rows
        tma_prefetch_desc(&p.w6_half_map);
        const uint32_t lbar = mapa_shared(w6full_bar, 0);
        if (crank == 0) mbar_arrive_expect_tx(w6full_bar, Cfg::kW6Bytes);
        for (int kc = 0; kc < kGChunks; ++kc)
          tma_load_2d_2sm(w6_base + kc * kW6ChunkBytes, &p.w6_half_map, lbar, kc * kKChunk, crank * (kLastN / 2));
      } else if (leader) {
        tma_prefetch_desc(&p.w6_map);
        mbar_arrive_expect_tx(w6full_bar, Cfg::kW6Bytes);
        for (int kc = 0; kc < kGChunks; ++kc)
          tma_load_2d(w6_base + kc * kW6ChunkBytes, &p.w6_map, w6full_bar, kc * kKChunk, 0);
      }
    }
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const TileCoord t = decode_tile(p, tile, crank, BLOCK_N);
      const int kbeg = p.sub_kbeg[t.sub], kend = p.sub_kend[t.sub];
      const int ksplit = kbeg + ((kend - kbeg) >> 1);
      for (int k = kbeg; k < ksplit; ++k) load_main(t, k);
      if (it > 0) load_gamma();
      for (int k = ksplit; k < kend; ++k) load_main(t, k);
    }
    if (it > 0) load_gamma();
  } else if (warp == 1 && (!duo || crank == 0)) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====================
    // (duo mode: only the cluster's rank 0 issues; its MMAs write the accumulators of both CTAs)
    const bool leader = elect_one();
    const uint32_t idesc = duo ? umma_idesc(/*F16*/ 0u, 256u, BLOCK_N) : umma_idesc(/*F16*/ 0u, 128u, BLOCK_N);
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t acc) {
      if (duo) mma_f16_ss2(d, ad, bd, id, acc);
      else mma_f16_ss(d, ad, bd, id, acc);
    };
    auto mma_ts = [&](uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t id, uint32_t acc) {  // A from tensor memory
      if (duo) mma_f16_ts2(d, a_tmem, bd, id, acc);
      else mma_f16_ts(d, a_tmem, bd, id, acc);
    };
    auto release_slot = [&](uint32_t bar) {
      if (duo) mma_commit2_mc(bar, 3);
      else if (pair) mma_commit_mc(bar, 3);
      else mma_commit(bar);
    };
    auto commit_local = [&](uint32_t bar) {  // accumulator-ready barriers: every CTA's epilogue waits on its own copy
      if (duo) mma_commit2_mc(bar, 3);
      else mma_commit(bar);
    };
    int s = 0;
    uint32_t ph = 0;
    // kLast: col(j) = out(j) . W6^T once phase 2 of tile j has left the layer's output in the x^2 buffers; it can
    // become ready at any point of this thread's program, so every wait below polls for it.
    int n3_done = 0;
    bool w6_ready = false;
    auto try_mma3 = [&]() {
      if constexpr (kLast) {
        // one lane's probe decides for the warp (n3_done must stay warp-uniform)
        if (!__shfl_sync(0xffffffffu, static_cast<int>(mbar_try_wait(a3rdy_bar, n3_done & 1)), 0)) return;
        if (duo) (void)mbar_try_wait_cluster(a3rdy_bar, n3_done & 1);  // one cluster-scope acquire (peer's arrivals)
        if (!w6_ready) {
          mbar_wait(w6full_bar, 0);
          w6_ready = true;
        }
        tc_fence_after();
        const uint32_t idesc3 = duo ? umma_idesc(/*F16*/ 0u, 256u, kLastN) : umma_idesc(/*F16*/ 0u, 128u, kLastN);
        if (leader) {
#pragma unroll
          for (int kc = 0; kc < kGChunks; ++kc) {
            const uint64_t adesc = umma_desc_sw128(a2_base + kc * kAStageBytes);
            const uint64_t bdesc = umma_desc_sw128(w6_base + kc * kW6ChunkBytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if constexpr (kTm)  // A = the tile's output in tensor memory, D = the (dead) accumulator of that tile
                mma_ts(tmem_base + (n3_done & 1) * BLOCK_N, tmem_base + Cfg::kStashCol + 32u * kc + 8u * kk,
                       bdesc + 2u * kk, idesc3, (kc > 0 || kk > 0) ? 1u : 0u);
              else
                mma(tmem_base + Cfg::kStashCol, adesc + 2u * kk, bdesc + 2u * kk, idesc3, (kc > 0 || kk > 0) ? 1u : 0u);
            }
          }
          commit_local(d3full_bar);
        }
        ++n3_done;
      }
    };
    // `remote`: the barrier also collects arrivals of the peer CTA (duo mode). The polls stay CTA-scope; one
    // cluster-scope acquire follows once the phase has completed (a cluster-scope acquire per poll stalls this warp)
    auto wait_bar = [&](uint32_t bar, uint32_t parity, bool remote = false) {
      if constexpr (kLast) {
        while (!__shfl_sync(0xffffffffu, static_cast<int>(mbar_try_wait(bar, parity)), 0)) try_mma3();
      } else {
        mbar_wait(bar, parity);
      }
      if (duo && remote) (void)mbar_try_wait_cluster(bar, parity);
    };
    auto mma_main = [&](uint32_t d_tmem, bool first) {
      wait_bar(full_bar(s), ph);
      tc_fence_after();
      const uint32_t a_addr = stage_base + s * Cfg::kStageBytes;
      const uint64_t adesc = umma_desc_sw128(a_addr);
      const uint64_t bdesc = umma_desc_sw128(a_addr + kAStageBytes);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (kk < p.kk_main) mma(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (!first || kk > 0) ? 1u : 0u);
        release_slot(empty_bar(s));
      }
      if (++s == kStages) {
        s = 0;
        ph ^= 1u;
      }
      try_mma3();
    };
    // norm(j) = gamma . (x s)^2 into the accumulator tile j just vacated
    auto mma_gamma = [&](int j) {
      wait_bar(a2rdy_bar, j & 1, true);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (j & 1) * BLOCK_N;
      for (int kc = 0; kc < kGChunks; ++kc) {
        wait_bar(full_bar(s), ph);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(a2_base + kc * kAStageBytes);
        const uint64_t bdesc = umma_desc_sw128(stage_base + s * Cfg::kStageBytes + kAStageBytes);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if constexpr (kTm)  // A = (x s)^2 in tensor memory
              mma_ts(d_tmem, tmem_base + Cfg::kStashCol + 32u * kc + 8u * kk, bdesc + 2u * kk, idesc,
                     (kc > 0 || kk > 0) ? 1u : 0u);
            else
              mma(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
          }
          release_slot(empty_bar(s));
        }
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (leader) commit_local(nfull_bar);
    };
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const int sub = tile_sub(p, tile);
      const int kbeg = p.sub_kbeg[sub], kend = p.sub_kend[sub];
      const int ksplit = kbeg + ((kend - kbeg) >> 1);
      const uint32_t d_tmem = tmem_base + (it & 1) * BLOCK_N;
      if (it >= 2) {  // phase 2 of tile it-2 has finished reading this accumulator
        wait_bar(accfree_bar(it & 1), ((it >> 1) - 1) & 1, true);
        tc_fence_after();
      }
      for (int k = kbeg; k < ksplit; ++k) mma_main(d_tmem, k == kbeg);
      if (it > 0) mma_gamma(it - 1);
      for (int k = ksplit; k < kend; ++k) mma_main(d_tmem, k == kbeg);
      if (leader) commit_local(tfull_bar);
    }
    if (it > 0) mma_gamma(it - 1);
    if constexpr (kLast) {
      while (n3_done < it) try_mma3();
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 16 warps =====================
    // warp e reads TMEM lane group e % 4 (the hardware's warp -> lane-group rule) and the 16-column quarter e / 4
    // of every 64-channel chunk: four warps per scheduler hide each other's TMEM / MUFU / LDS latencies (the
    // 8-warp version ran at ~0.3 IPC per scheduler, profiles/r01_ncu_gdn_v2_epilogue.txt).
    const int e = warp - 4;
    const int etid = threadIdx.x - 128;
    const int row = (e & 3) * 32 + lane;  // accumulator row == pixel of the patch
    const int q = e >> 2;                 // 16-column quarter
    const uint32_t lane_off = static_cast<uint32_t>((e & 3) * 32) << 16;
    const uint32_t rsw = static_cast<uint32_t>(row & 7);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t p0 = ((2u * q) ^ rsw) << 4, p1 = ((2u * q + 1u) ^ rsw) << 4;  // swizzled 16-byte pieces
    const float sc = p.sq_scale;  // power of two: fma(acc, s, b s) rounds exactly like (acc + b) s
    const float ka = kInverse ? p.sq_inv * p.sq_inv : 1.0f;
    // duo mode: "operand ready" / "accumulator free" arrivals of both CTAs go to the leader's barriers
    auto arrive_mma = [&](uint32_t bar) {
      if (duo) mbar_arrive_cluster(mapa_shared(bar, 0));
      else mbar_arrive(bar);
    };
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const TileCoord t = decode_tile(p, tile, crank, BLOCK_N);
      const uint32_t par = it & 1;
      const uint32_t acc_col = tmem_base + lane_off + (it & 1) * BLOCK_N + 16 * q;
      const uint32_t stash_col = tmem_base + lane_off + Cfg::kStashCol + 8 * q;
      mbar_wait(tfull_bar, par);
      tc_fence_after();
      if constexpr (kTm) {
        // ===== operands in tensor memory (fused last layer, activation not stored): x s stays in registers, (x s)^2
        // and then the layer's output are written as packed fp16 into the operand columns, from where the gamma and
        // W6 MMAs read their A operand; no shared-memory buffer, no named barrier, no proxy fence
        uint32_t hx[kGChunks * 8];
        {  // ---- phase 1
          uint32_t r[kGChunks][16];
#pragma unroll
          for (int g = 0; g < kGChunks; ++g) tmem_ld_32x16(acc_col + 64 * g, r[g]);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < kGChunks; ++g) {
            uint32_t hq[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float b0, b1, b2, b3;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                           : "r"(bias_smem + 4u * (64 * g + 16 * q + 4 * j)));
              const __half2 h0 = __floats2half2_rn(fmaf(__uint_as_float(r[g][4 * j]), sc, b0),
                                                   fmaf(__uint_as_float(r[g][4 * j + 1]), sc, b1));
              const __half2 h1 = __floats2half2_rn(fmaf(__uint_as_float(r[g][4 * j + 2]), sc, b2),
                                                   fmaf(__uint_as_float(r[g][4 * j + 3]), sc, b3));
              const __half2 q0 = __hmul2(h0, h0), q1 = __hmul2(h1, h1);
              hx[g * 8 + 2 * j] = *reinterpret_cast<const uint32_t*>(&h0);
              hx[g * 8 + 2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
              hq[2 * j] = *reinterpret_cast<const uint32_t*>(&q0);
              hq[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&q1);
            }
            tmem_st_32x8(stash_col + 32 * g, hq);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_mma(a2rdy_bar);
        }
        // ---- phase 2: out = (x s) * (r)sqrt(.) over the operand columns (the gamma MMAs have consumed (x s)^2)
        mbar_wait(nfull_bar, par);
        tc_fence_after();
#pragma unroll
        for (int g = 0; g < kGChunks; ++g) {
          uint32_t r[16];
          tmem_ld_32x16(acc_col + 64 * g, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                         : "r"(beta_smem + 4u * (64 * g + 16 * q + 4 * j)));
            const float n0 = fmaf(__uint_as_float(r[4 * j]), ka, b0);
            const float n1 = fmaf(__uint_as_float(r[4 * j + 1]), ka, b1);
            const float n2 = fmaf(__uint_as_float(r[4 * j + 2]), ka, b2);
            const float n3 = fmaf(__uint_as_float(r[4 * j + 3]), ka, b3);
            float f0, f1, f2, f3;
            if constexpr (kInverse) {
              f0 = approx_sqrt(n0), f1 = approx_sqrt(n1), f2 = approx_sqrt(n2), f3 = approx_sqrt(n3);
            } else {
              f0 = approx_rsqrt(n0), f1 = approx_rsqrt(n1), f2 = approx_rsqrt(n2), f3 = approx_rsqrt(n3);
            }
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&hx[g * 8 + 2 * j]));
            const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&hx[g * 8 + 2 * j + 1]));
            hx[g * 8 + 2 * j] = pack_half2(x0.x * f0, x0.y * f1);
            hx[g * 8 + 2 * j + 1] = pack_half2(x1.x * f2, x1.y * f3);
          }
          uint32_t ho[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) ho[j] = hx[g * 8 + j];
          tmem_st_32x8(stash_col + 32 * g, ho);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_mma(a3rdy_bar);
        // ---- phase 3: col = out . W6^T, accumulated by the MMA warp in the first 96 columns of this tile's (dead)
        // accumulator; this thread's 24 of the 96 columns go to the col buffer as fp16
        mbar_wait(d3full_bar, par);
        tc_fence_after();
        uint32_t d[24];
        {
          uint32_t d16[16], d8[8];
          const uint32_t d3_col = tmem_base + lane_off + (it & 1) * BLOCK_N + 24 * q;
          tmem_ld_32x16(d3_col, d16);
          tmem_ld_32x8(d3_col + 16, d8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) d[i] = d16[i];
#pragma unroll
          for (int i = 0; i < 8; ++i) d[16 + i] = d8[i];
        }
        // every TMEM read of this tile has completed: hand the accumulator back to the MMA thread
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_mma(accfree_bar(it & 1));
        const int th = row / p.tile_w, tw = row - th * p.tile_w;
        const int oh = t.h0 + th, ow = t.w0 + tw;
        if (row < p.tile_h * p.tile_w && oh < p.h_out && ow < p.w_out) {
          const long long pix =
              (static_cast<long long>(t.n_img) * p.full_h + (oh * p.os + p.sub_p[t.sub])) * p.full_w +
              (ow * p.os + p.sub_q[t.sub]);
          uint4* dst = reinterpret_cast<uint4*>(p.col_out + pix * kLastN + 24 * q);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            dst[i] = make_uint4(pack_half2(__uint_as_float(d[8 * i]), __uint_as_float(d[8 * i + 1])),
                                pack_half2(__uint_as_float(d[8 * i + 2]), __uint_as_float(d[8 * i + 3])),
                                pack_half2(__uint_as_float(d[8 * i + 4]), __uint_as_float(d[8 * i + 5])),
                                pack_half2(__uint_as_float(d[8 * i + 6]), __uint_as_float(d[8 * i + 7])));
        }
        continue;
      }
      // ---- phase 1: x = A + bias; x s -> stash (TMEM, packed fp16), (x s)^2 -> smem operand.
      // Buffer g was the staging of the previous tile's g-th output store (one bulk group each, oldest first).
      {
        uint32_t r[2][16];
        tmem_ld_32x16(acc_col, r[0]);
#pragma unroll
        for (int g = 0; g < kGChunks; ++g) {
          const int c = 64 * g + 16 * q;
          if (etid == 0) {
            if (g == 0) tma_store_wait_read<kGChunks - 1>();
            else if (g + 1 < kGChunks) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          tmem_ld_wait();
          if (g + 1 < kGChunks) tmem_ld_32x16(acc_col + 64 * (g + 1), r[(g + 1) & 1]);
          uint32_t hx[8], hq[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                         : "r"(bias_smem + 4u * (c + 4 * j)));
            const uint32_t* rr = r[g & 1];
            // the value that is normalised is the fp16-rounded x; x s is exact (s is a power of two)
            const __half2 h0 = __floats2half2_rn(fmaf(__uint_as_float(rr[4 * j]), sc, b0),
                                                 fmaf(__uint_as_float(rr[4 * j + 1]), sc, b1));
            const __half2 h1 = __floats2half2_rn(fmaf(__uint_as_float(rr[4 * j + 2]), sc, b2),
                                                 fmaf(__uint_as_float(rr[4 * j + 3]), sc, b3));
            const __half2 q0 = __hmul2(h0, h0), q1 = __hmul2(h1, h1);
            hx[2 * j] = *reinterpret_cast<const uint32_t*>(&h0);
            hx[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
            hq[2 * j] = *reinterpret_cast<const uint32_t*>(&q0);
            hq[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&q1);
          }
          tmem_st_32x8(stash_col + 32 * g, hx);
          named_bar_sync(1, kGdnEpiThreads);  // buffer g is free (thread 0 saw its store drain)
          const uint32_t cbase = a2_base + static_cast<uint32_t>(g) * kAStageBytes + row_off;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + p0), "r"(hq[0]), "r"(hq[1]),
                       "r"(hq[2]), "r"(hq[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + p1), "r"(hq[4]), "r"(hq[5]),
                       "r"(hq[6]), "r"(hq[7])
                       : "memory");
        }
      }
      tmem_st_wait();
      fence_proxy_async_smem();  // (x s)^2 is read by the tensor core through the async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_mma(a2rdy_bar);

      // ---- phase 2: out = (x s) * (r)sqrt(.) into the (now free) x^2 buffers, one TMA store per 64 channels
      mbar_wait(nfull_bar, par);
      tc_fence_after();
      {
        uint32_t r[2][16], hs[2][8];
        tmem_ld_32x16(acc_col, r[0]);
        tmem_ld_32x8(stash_col, hs[0]);
#pragma unroll
        for (int g = 0; g < kGChunks; ++g) {
          const int c = 64 * g + 16 * q;
          tmem_ld_wait();
          if (g + 1 < kGChunks) {
            tmem_ld_32x16(acc_col + 64 * (g + 1), r[(g + 1) & 1]);
            tmem_ld_32x8(stash_col + 32 * (g + 1), hs[(g + 1) & 1]);
          } else {
            // every TMEM read of this tile has completed: hand the accumulator back to the MMA thread
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_mma(accfree_bar(it & 1));
          }
          uint32_t ho[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                         : "r"(beta_smem + 4u * (c + 4 * j)));
            const uint32_t* rr = r[g & 1];
            const uint32_t* hx = hs[g & 1];
            const float n0 = fmaf(__uint_as_float(rr[4 * j]), ka, b0);
            const float n1 = fmaf(__uint_as_float(rr[4 * j + 1]), ka, b1);
            const float n2 = fmaf(__uint_as_float(rr[4 * j + 2]), ka, b2);
            const float n3 = fmaf(__uint_as_float(rr[4 * j + 3]), ka, b3);
            float f0, f1, f2, f3;
            if constexpr (kInverse) {
              f0 = approx_sqrt(n0), f1 = approx_sqrt(n1), f2 = approx_sqrt(n2), f3 = approx_sqrt(n3);
            } else {
              f0 = approx_rsqrt(n0), f1 = approx_rsqrt(n1), f2 = approx_rsqrt(n2), f3 = approx_rsqrt(n3);
            }
            // (a packed-half product (x s) * fp16(f) was measured: no faster, one more rounding - not used)
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&hx[2 * j]));
            const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&hx[2 * j + 1]));
            ho[2 * j] = pack_half2(x0.x * f0, x0.y * f1);
            ho[2 * j + 1] = pack_half2(x1.x * f2, x1.y * f3);
          }
          const uint32_t cbase = a2_base + static_cast<uint32_t>(g) * kAStageBytes + row_off;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + p0), "r"(ho[0]), "r"(ho[1]),
                       "r"(ho[2]), "r"(ho[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + p1), "r"(ho[4]), "r"(ho[5]),
                       "r"(ho[6]), "r"(ho[7])
                       : "memory");
          fence_proxy_async_smem();
          named_bar_sync(1, kGdnEpiThreads);
          if (etid == 0 && (!kLast || p.store_act)) {
            tma_store_4d(&p.out_map[t.sub], a2_base + static_cast<uint32_t>(g) * kAStageBytes, 64 * g, t.w0, t.h0,
                         t.n_img);
            tma_store_commit();
          }
        }
      }
      if constexpr (kLast) {
        // ---- phase 3: the layer's output tile (K-major in the x^2 buffers) times W6 on the tensor core, accumulator in
        // the (dead) stash columns; this thread's 24 of the 96 columns go to the col buffer as fp16
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_mma(a3rdy_bar);
        mbar_wait(d3full_bar, par);
        tc_fence_after();
        uint32_t d[24];
        {
          uint32_t d16[16], d8[8];
          const uint32_t d3_col = tmem_base + lane_off + Cfg::kStashCol + 24 * q;
          tmem_ld_32x16(d3_col, d16);
          tmem_ld_32x8(d3_col + 16, d8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) d[i] = d16[i];
#pragma unroll
          for (int i = 0; i < 8; ++i) d[16 + i] = d8[i];
        }
        const int th = row / p.tile_w, tw = row - th * p.tile_w;
        const int oh = t.h0 + th, ow = t.w0 + tw;
        if (row < p.tile_h * p.tile_w && oh < p.h_out && ow < p.w_out) {
          const long long pix =
              (static_cast<long long>(t.n_img) * p.full_h + (oh * p.os + p.sub_p[t.sub])) * p.full_w +
              (ow * p.os + p.sub_q[t.sub]);
          uint4* dst = reinterpret_cast<uint4*>(p.col_out + pix * kLastN + 24 * q);
#pragma unroll
          for (int i = 0; i < 3; ++i)
            dst[i] = make_uint4(pack_half2(__uint_as_float(d[8 * i]), __uint_as_float(d[8 * i + 1])),
                                pack_half2(__uint_as_float(d[8 * i + 2]), __uint_as_float(d[8 * i + 3])),
                                pack_half2(__uint_as_float(d[8 * i + 4]), __uint_as_float(d[8 * i + 5])),
                                pack_half2(__uint_as_float(d[8 * i + 6]), __uint_as_float(d[8 * i + 7])));
        }
        // the next tile's phase 1 writes its stash into these columns: every warp must be done reading them
        tc_fence_before();
        named_bar_sync(1, kGdnEpiThreads);
      }
    }
    if (etid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();  // no CTA leaves while its peer may still multicast into it
  if (warp == 2) {
    if (duo) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =====================================================================================================
// conv + GDN / IGDN for SHORT main loops (first layer: 5 K steps; im2col GEMMs): ping-pong epilogue.
//
// With a K loop of a few steps the kernel above is bound by its epilogue chain (phase 1 -> gamma MMA latency ->
// phase 2, ~6.7 k cycles per tile, tensor pipe 39 % active on g_a.0). Here the 16 epilogue warps form two groups of 8;
// group G owns the tiles it = G (mod 2) together with accumulator G and its own x^2 / staging buffer, keeps x s in
// REGISTERS between its two phases (96 columns per thread = 48 packed registers, so no TMEM stash), and the groups run
// half a period apart: while one waits for its gamma contraction the other normalises. Same arithmetic, bit for bit,
// as conv_gdn_kernel. Ring: 3 stages (2 x (N/64) x 16 KB of x^2 buffers take the room of the fourth).
// =====================================================================================================
template <int kNT>
struct GdnPpCfgT {
  static constexpr int kN = kNT;
  static constexpr int kBStageBytes = kN * 128;
  static constexpr int kStageBytes = kAStageBytes + kBStageBytes;
  static constexpr int kStages = 3;
  static constexpr int kA2Bytes = (kN / 64) * kAStageBytes;  // per group
  static constexpr int kBarrierBytes = 256 + 2 * kN * 4;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * kA2Bytes + kBarrierBytes;
  static constexpr int kTmemCols = 512;
  static_assert(kN % 64 == 0 && 2 * kN <= 512, "TMEM budget");
  static_assert(kSmemBytes <= kSmemLimit, "smem overflow");
};

template <int kNT, bool kInverse>
__global__ void __launch_bounds__(kGdnThreads, 1)
conv_gdn_pp_kernel(const __grid_constant__ ConvKernelParams p) {
  using Cfg = GdnPpCfgT<kNT>;
  constexpr int kStages = Cfg::kStages;
  constexpr int BLOCK_N = Cfg::kN;
  constexpr int kGChunks = BLOCK_N / kKChunk;
  constexpr int kGroupThreads = 256;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base;
  const uint32_t a2_base0 = smem_base + kStages * Cfg::kStageBytes;
  const uint32_t bar_base = a2_base0 + 2 * Cfg::kA2Bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int g) { return bar_base + 8u * (2 * kStages + g); };
  auto a2rdy_bar = [&](int g) { return bar_base + 8u * (2 * kStages + 2 + g); };
  auto nfull_bar = [&](int g) { return bar_base + 8u * (2 * kStages + 4 + g); };
  auto accfree_bar = [&](int g) { return bar_base + 8u * (2 * kStages + 6 + g); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 8);
  const uint32_t bias_smem = bar_base + 256u;
  const uint32_t beta_smem = bias_smem + 4u * BLOCK_N;

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool pair = p.csize == 2;
  const int crank = pair ? static_cast<int>(cluster_ctarank()) : 0;
  const int q_first = pair ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int q_stride = pair ? static_cast<int>(cluster_count_x()) : static_cast<int>(gridDim.x);
  const int n_items = p.total_tiles / p.csize;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&p.a_map[i]);
    tma_prefetch_desc(&p.b_map);
    tma_prefetch_desc(&p.g_map);
    for (int i = 0; i < p.n_sub; ++i) tma_prefetch_desc(&p.out_map[i]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), p.csize);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(tfull_bar(g), 1);
      mbar_init(a2rdy_bar(g), 8);
      mbar_init(nfull_bar(g), 1);
      mbar_init(accfree_bar(g), 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (warp >= 4) {
    const float kb = kInverse ? p.sq_inv : p.sq_scale * p.sq_scale;
    for (int i = threadIdx.x - 128; i < BLOCK_N; i += kGdnEpiThreads) {
      const float b = __ldg(p.bias + i) * p.sq_scale, g = __ldg(p.beta + i) * kb;  // bias pre-scaled: x s = fma(acc, s, b s)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_smem + 4u * i), "f"(b) : "memory");
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(beta_smem + 4u * i), "f"(g) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const uint32_t a_tx_bytes = static_cast<uint32_t>(p.tile_h * p.tile_w) * 128u;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    auto load_main = [&](const TileCoord& t, int k) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      const uint32_t e = p.ksteps[k];
      const int map = e & 3;
      const int dh = static_cast<int>((e >> 2) & 15u) - 8;
      const int dw = static_cast<int>((e >> 6) & 15u) - 8;
      const int c0 = static_cast<int>(e >> 10);
      const uint32_t a_dst = stage_base + s * Cfg::kStageBytes;
      if (leader) {
        mbar_arrive_expect_tx(full_bar(s), a_tx_bytes + Cfg::kBStageBytes);
        tma_load_4d(a_dst, &p.a_map[map], full_bar(s), c0, t.w0 + dw, t.h0 + dh, t.n_img);
        if (pair)
          tma_load_2d_mc(a_dst + kAStageBytes + crank * (Cfg::kBStageBytes / 2), &p.b_half_map, full_bar(s),
                         k * kKChunk, crank * (BLOCK_N / 2), 3);
        else
          tma_load_2d(a_dst + kAStageBytes, &p.b_map, full_bar(s), k * kKChunk, 0);
      }
      if (++s == kStages) {
        s = 0;
        ph ^= 1u;
      }
    };
    auto load_gamma = [&]() {
      for (int kc = 0; kc < kGChunks; ++kc) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t a_dst = stage_base + s * Cfg::kStageBytes;
        if (leader) {
          mbar_arrive_expect_tx(full_bar(s), Cfg::kBStageBytes);
          if (pair)
            tma_load_2d_mc(a_dst + kAStageBytes + crank * (Cfg::kBStageBytes / 2), &p.g_half_map, full_bar(s),
                           kc * kKChunk, crank * (BLOCK_N / 2), 3);
          else
            tma_load_2d(a_dst + kAStageBytes, &p.g_map, full_bar(s), kc * kKChunk, 0);
        }
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    };
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const TileCoord t = decode_tile(p, tile, crank, BLOCK_N);
      const int kbeg = p.sub_kbeg[t.sub], kend = p.sub_kend[t.sub];
      const int ksplit = kbeg + ((kend - kbeg) >> 1);
      for (int k = kbeg; k < ksplit; ++k) load_main(t, k);
      if (it > 0) load_gamma();
      for (int k = ksplit; k < kend; ++k) load_main(t, k);
    }
    if (it > 0) load_gamma();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc(/*F16*/ 0u, 128u, BLOCK_N);
    int s = 0;
    uint32_t ph = 0;
    auto mma_main = [&](uint32_t d_tmem, bool first) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t a_addr = stage_base + s * Cfg::kStageBytes;
      const uint64_t adesc = umma_desc_sw128(a_addr);
      const uint64_t bdesc = umma_desc_sw128(a_addr + kAStageBytes);
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (kk < p.kk_main) mma_f16_ss(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (!first || kk > 0) ? 1u : 0u);
        if (pair) mma_commit_mc(empty_bar(s), 3);
        else mma_commit(empty_bar(s));
      }
      if (++s == kStages) {
        s = 0;
        ph ^= 1u;
      }
    };
    // norm(j) = gamma . (x s)^2 over the accumulator of tile j, operand in group (j & 1)'s buffer
    auto mma_gamma = [&](int j) {
      const int g = j & 1;
      mbar_wait(a2rdy_bar(g), (j >> 1) & 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + g * BLOCK_N;
      const uint32_t a2 = a2_base0 + g * Cfg::kA2Bytes;
      for (int kc = 0; kc < kGChunks; ++kc) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(a2 + kc * kAStageBytes);
        const uint64_t bdesc = umma_desc_sw128(stage_base + s * Cfg::kStageBytes + kAStageBytes);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_f16_ss(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (kc > 0 || kk > 0) ? 1u : 0u);
          if (pair) mma_commit_mc(empty_bar(s), 3);
          else mma_commit(empty_bar(s));
        }
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (leader) mma_commit(nfull_bar(g));
    };
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const int sub = tile_sub(p, tile);
      const int kbeg = p.sub_kbeg[sub], kend = p.sub_kend[sub];
      const int ksplit = kbeg + ((kend - kbeg) >> 1);
      const int g = it & 1;
      const uint32_t d_tmem = tmem_base + g * BLOCK_N;
      if (it >= 2) {  // phase 2 of tile it-2 (same group) has finished reading this accumulator
        mbar_wait(accfree_bar(g), ((it >> 1) - 1) & 1);
        tc_fence_after();
      }
      for (int k = kbeg; k < ksplit; ++k) mma_main(d_tmem, k == kbeg);
      if (it > 0) mma_gamma(it - 1);
      for (int k = ksplit; k < kend; ++k) mma_main(d_tmem, k == kbeg);
      if (leader) mma_commit(tfull_bar(g));
    }
    if (it > 0) mma_gamma(it - 1);
  } else if (warp >= 4) {
    // ===================== epilogue: two groups of 8 warps =====================
    const int G = (warp - 4) >> 3;           // group = parity of the tiles it owns
    const int wg = (warp - 4) & 7;           // warp inside the group
    const int gtid = wg * 32 + lane;         // 0..255
    const int row = (wg & 3) * 32 + lane;    // accumulator row == pixel of the patch (TMEM lane group = warp % 4)
    const int half = wg >> 2;                // 32-column half of every 64-channel chunk
    const uint32_t lane_off = static_cast<uint32_t>((wg & 3) * 32) << 16;
    const uint32_t rsw = static_cast<uint32_t>(row & 7);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t a2_base = a2_base0 + G * Cfg::kA2Bytes;
    const uint32_t acc_col = tmem_base + lane_off + G * BLOCK_N + 32 * half;
    const uint32_t bar_id = 1 + G;
    const float sc = p.sq_scale;  // power of two: fma(acc, s, b s) rounds exactly like (acc + b) s
    const float ka = kInverse ? p.sq_inv * p.sq_inv : 1.0f;
    int n = 0;
    for (int tile = q_first + G * q_stride; tile < n_items; tile += 2 * q_stride, ++n) {
      const TileCoord t = decode_tile(p, tile, crank, BLOCK_N);
      const uint32_t par = n & 1;
      uint32_t hx[kGChunks * 16];  // x s of this thread's 32 columns per chunk, packed fp16
      mbar_wait(tfull_bar(G), par);
      tc_fence_after();
      // ---- phase 1: x s -> registers, (x s)^2 -> this group's smem operand (its previous tile's stores must be out)
#pragma unroll
      for (int g = 0; g < kGChunks; ++g) {
        if (gtid == 0) {
          if (g == 0) tma_store_wait_read<kGChunks - 1>();
          else if (g + 1 < kGChunks) tma_store_wait_read<1>();
          else tma_store_wait_read<0>();
        }
        named_bar_sync(bar_id, kGroupThreads);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int c = 64 * g + 32 * half + 16 * sub;
          uint32_t r[16];
          tmem_ld_32x16(acc_col + 64 * g + 16 * sub, r);
          tmem_ld_wait();
          uint32_t hq[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                         : "r"(bias_smem + 4u * (c + 4 * j)));
            const __half2 h0 =
                __floats2half2_rn(fmaf(__uint_as_float(r[4 * j]), sc, b0), fmaf(__uint_as_float(r[4 * j + 1]), sc, b1));
            const __half2 h1 = __floats2half2_rn(fmaf(__uint_as_float(r[4 * j + 2]), sc, b2),
                                                 fmaf(__uint_as_float(r[4 * j + 3]), sc, b3));
            const __half2 q0 = __hmul2(h0, h0), q1 = __hmul2(h1, h1);
            hx[g * 16 + sub * 8 + 2 * j] = *reinterpret_cast<const uint32_t*>(&h0);
            hx[g * 16 + sub * 8 + 2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
            hq[2 * j] = *reinterpret_cast<const uint32_t*>(&q0);
            hq[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&q1);
          }
          const uint32_t cbase = a2_base + static_cast<uint32_t>(g) * kAStageBytes + row_off;
          const uint32_t pa = ((4u * half + 2u * sub) ^ rsw) << 4, pb = ((4u * half + 2u * sub + 1u) ^ rsw) << 4;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + pa), "r"(hq[0]), "r"(hq[1]), "r"(hq[2]),
                       "r"(hq[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + pb), "r"(hq[4]), "r"(hq[5]), "r"(hq[6]),
                       "r"(hq[7])
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2rdy_bar(G));

      // ---- phase 2: out = (x s) * (r)sqrt(.) into the group's (now free) x^2 buffers, one TMA store per 64 channels
      mbar_wait(nfull_bar(G), par);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < kGChunks; ++g) {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int c = 64 * g + 32 * half + 16 * sub;
          uint32_t r[16];
          tmem_ld_32x16(acc_col + 64 * g + 16 * sub, r);
          tmem_ld_wait();
          if (g + 1 == kGChunks && sub == 1) {
            // every TMEM read of this tile has completed: hand the accumulator back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accfree_bar(G));
          }
          uint32_t ho[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float b0, b1, b2, b3;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3)
                         : "r"(beta_smem + 4u * (c + 4 * j)));
            const float n0 = fmaf(__uint_as_float(r[4 * j]), ka, b0);
            const float n1 = fmaf(__uint_as_float(r[4 * j + 1]), ka, b1);
            const float n2 = fmaf(__uint_as_float(r[4 * j + 2]), ka, b2);
            const float n3 = fmaf(__uint_as_float(r[4 * j + 3]), ka, b3);
            float f0, f1, f2, f3;
            if constexpr (kInverse) {
              f0 = approx_sqrt(n0), f1 = approx_sqrt(n1), f2 = approx_sqrt(n2), f3 = approx_sqrt(n3);
            } else {
              f0 = approx_rsqrt(n0), f1 = approx_rsqrt(n1), f2 = approx_rsqrt(n2), f3 = approx_rsqrt(n3);
            }
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&hx[g * 16 + sub * 8 + 2 * j]));
            const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&hx[g * 16 + sub * 8 + 2 * j + 1]));
            ho[2 * j] = pack_half2(x0.x * f0, x0.y * f1);
            ho[2 * j + 1] = pack_half2(x1.x * f2, x1.y * f3);
          }
          const uint32_t cbase = a2_base + static_cast<uint32_t>(g) * kAStageBytes + row_off;
          const uint32_t pa = ((4u * half + 2u * sub) ^ rsw) << 4, pb = ((4u * half + 2u * sub + 1u) ^ rsw) << 4;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + pa), "r"(ho[0]), "r"(ho[1]), "r"(ho[2]),
                       "r"(ho[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cbase + pb), "r"(ho[4]), "r"(ho[5]), "r"(ho[6]),
                       "r"(ho[7])
                       : "memory");
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, kGroupThreads);
        if (gtid == 0) {
          tma_store_4d(&p.out_map[t.sub], a2_base + static_cast<uint32_t>(g) * kAStageBytes, 64 * g, t.w0, t.h0, t.n_img);
          tma_store_commit();
        }
      }
    }
    if (gtid == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// =====================================================================================================
// entropy_parameters' last layer fused with GaussianConditional (BASELINE.json north_star item 2: "the means and
// scales never round-trip HBM").
//
//   (sigma | mu) = EPM.4(e2) + bias           1x1 contraction, accumulator in TMEM (double buffered)
//   y_hat, L, -sum log2 L = GaussianConditional(y [- cond], sigma, mu)      in the epilogue, straight from TMEM
//
// The weight rows are interleaved on the host per 64 channels: N tile nt (BLOCK_N = 128 accumulator columns) holds
// sigma of channels [64 nt, 64 nt + 64) in columns [0, 64) and mu of the same channels in columns [64, 128), so a
// thread owns both parameters of its elements. 16 epilogue warps (warp e: TMEM lane group e % 4 = 32 pixels, channel
// quarter e / 4 = 16 channels): the GaussianConditional arithmetic is ~200 instructions per element (two erfcf, two
// IEEE divisions - gc_math.cuh, the reference's formula), 20x the MMA time of the tile, so the epilogue is what the
// kernel runs at and it needs every issue slot it can get. Outputs go to NCHW fp32 (the API layout) directly from
// registers: a warp's 32 pixels are consecutive along the tile's rows. Replaces entropy_models.py:588-596 behind
// spatiotemporalpriors.py:577-579 without sigma / mu (2 x 4 B per element written + read) ever reaching HBM.
// Single-CTA mode only (EPM.4 has an odd tile count per frame batch at 1080p, so it never ran in cluster mode).
// =====================================================================================================
constexpr int kGcFuseEpiWarps = 16;
constexpr int kGcFuseEpiThreads = kGcFuseEpiWarps * 32;
constexpr int kGcFuseThreads = 128 + kGcFuseEpiThreads;
constexpr int kGcFuseN = 128;

struct GcFuseParams {
  const float* y;       // NHWC fp32 [batch][h][w][C]
  const __half* cond;   // NHWC fp16 or null: the coded quantity is y - cond (_Res)
  float* y_hat;         // NCHW fp32 or null
  float* lik;           // NCHW fp32 or null
  double* bits;         // [batch] or null: += sum(-log2 lik)
  int C, hw;
  float scale_bound, lik_bound;
  int yhat_mode;
};

__global__ void __launch_bounds__(kGcFuseThreads, 1)
conv_gc_kernel(const __grid_constant__ ConvKernelParams p, const __grid_constant__ GcFuseParams g) {
  constexpr int BLOCK_N = kGcFuseN;
  using Cfg = ConvCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_base = smem_base;
  const uint32_t bar_base = smem_base + kStages * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int q_first = static_cast<int>(blockIdx.x), q_stride = static_cast<int>(gridDim.x);
  const int n_items = p.total_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.a_map[0]);
    tma_prefetch_desc(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kGcFuseEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  const uint32_t a_tx_bytes = static_cast<uint32_t>(p.tile_h * p.tile_w) * 128u;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride) {
      const TileCoord t = decode_tile(p, tile, 0, BLOCK_N);
      const int kbeg = p.sub_kbeg[0], kend = p.sub_kend[0];
      for (int k = kbeg; k < kend; ++k) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t e = p.ksteps[k];
        const int map = e & 3;
        const int dh = static_cast<int>((e >> 2) & 15u) - 8;
        const int dw = static_cast<int>((e >> 6) & 15u) - 8;
        const int c0 = static_cast<int>(e >> 10);
        const uint32_t a_dst = stage_base + s * Cfg::kStageBytes;
        if (leader) {
          mbar_arrive_expect_tx(full_bar(s), a_tx_bytes + Cfg::kBStageBytes);
          tma_load_4d(a_dst, &p.a_map[map], full_bar(s), c0, t.w0 + dw, t.h0 + dh, t.n_img);
          tma_load_2d(a_dst + kAStageBytes, &p.b_map, full_bar(s), k * kKChunk, t.n0);
        }
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc(/*F16*/ 0u, 128u, BLOCK_N);
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const int kbeg = p.sub_kbeg[0], kend = p.sub_kend[0];
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), accph ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int k = kbeg; k < kend; ++k) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = stage_base + s * Cfg::kStageBytes;
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(a_addr + kAStageBytes);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            mma_f16_ss(d_tmem, adesc + 2u * kk, bdesc + 2u * kk, idesc, (k > kbeg || kk > 0) ? 1u : 0u);
          mma_commit(empty_bar(s));
        }
        if (++s == kStages) {
          s = 0;
          ph ^= 1u;
        }
      }
      if (leader) mma_commit(tfull_bar(acc));
    }
  } else if (warp >= 4) {
    // ===================== epilogue: GaussianConditional on the accumulator =====================
    const int e = warp - 4;
    const int row = (e & 3) * 32 + lane;  // accumulator row == pixel of the patch
    const int q = e >> 2;                 // 16-channel quarter of the tile's 64 channels
    const uint32_t lane_off = static_cast<uint32_t>((e & 3) * 32) << 16;
    const int npix = p.tile_h * p.tile_w;
    const int C = g.C;
    int it = 0;
    for (int tile = q_first; tile < n_items; tile += q_stride, ++it) {
      const TileCoord t = decode_tile(p, tile, 0, BLOCK_N);
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      const int th = row / p.tile_w, tw = row - th * p.tile_w;
      const int oh = t.h0 + th, ow = t.w0 + tw;
      const bool inb = (row < npix) && (oh < p.h_out) && (ow < p.w_out);
      const int pix = oh * p.w_out + ow;               // pixel inside the frame
      const int ch0 = (t.n0 >> 1) + 16 * q;            // first latent channel of this thread
      mbar_wait(tfull_bar(acc), accph);
      tc_fence_after();
      uint32_t rs[16], rm[16];
      const uint32_t t_row = tmem_base + lane_off + acc * BLOCK_N;
      tmem_ld_32x16(t_row + 16 * q, rs);
      tmem_ld_32x16(t_row + 64 + 16 * q, rm);
      tmem_ld_wait();
      // the accumulator is in registers: hand it back to the MMA warp before the long arithmetic starts
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      float bits = 0.f;
      if (inb) {
        const long long e0 = (static_cast<long long>(t.n_img) * g.hw + pix) * C + ch0;
        const long long o0 = (static_cast<long long>(t.n_img) * C + ch0) * g.hw + pix;
        const float4* bsp = reinterpret_cast<const float4*>(p.bias + t.n0 + 16 * q);
        const float4* bmp = reinterpret_cast<const float4*>(p.bias + t.n0 + 64 + 16 * q);
        const float4* yp = reinterpret_cast<const float4*>(g.y + e0);
        const uint2* cp = g.cond ? reinterpret_cast<const uint2*>(g.cond + e0) : nullptr;
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 4 channels per round: one 16-byte load of y, sigma bias, mu bias each
          const float4 y4 = __ldg(yp + j), bs4 = __ldg(bsp + j), bm4 = __ldg(bmp + j);
          float c4[4] = {0.f, 0.f, 0.f, 0.f};
          if (cp) {
            const uint2 u = __ldg(cp + j);
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            c4[0] = f0.x, c4[1] = f0.y, c4[2] = f1.x, c4[3] = f1.y;
          }
          const float yy[4] = {y4.x, y4.y, y4.z, y4.w};
          const float bsv[4] = {bs4.x, bs4.y, bs4.z, bs4.w};
          const float bmv[4] = {bm4.x, bm4.y, bm4.z, bm4.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * j + k;
            const float sigma = __fadd_rn(__uint_as_float(rs[i]), bsv[k]);
            const float mu = __fadd_rn(__uint_as_float(rm[i]), bmv[k]);
            const float v = __fsub_rn(yy[k], c4[k]);
            GcOut o = gc_eval(v, sigma, mu, nullptr, 0, g.scale_bound, g.lik_bound, false);
            if (g.yhat_mode == 1) o.y_hat = __fadd_rn(rintf(v), c4[k]);
            const long long oi = o0 + static_cast<long long>(i) * g.hw;
            if (g.y_hat) g.y_hat[oi] = o.y_hat;
            if (g.lik) g.lik[oi] = o.lik;
            bits -= __log2f(o.lik);
          }
        }
      }
      if (g.bits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
        if (lane == 0) atomicAdd(g.bits + t.n_img, static_cast<double>(bits));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// =====================================================================================================
// weight repack: PyTorch OIHW (or IOHW for ConvTranspose2d) fp32 -> [c_out][K] fp16, K ordered like the
// kernel's k-step table
// =====================================================================================================
struct PackParams {
  uint32_t info[kMaxKSteps];  // [2:0] r | [5:3] s | [12:6] valid channels in this chunk - 1 | [31:13] channel base
  int n_steps;
  int c_out, c_in_total, kh, kw, transposed;
  int row_taps;
};

__global__ void pack_weight_kernel(const __grid_constant__ PackParams pp, const float* __restrict__ w,
                                   __half* __restrict__ out) {
  const long long K = static_cast<long long>(pp.n_steps) * kKChunk;
  const long long total = K * pp.c_out;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int o = static_cast<int>(i / K);
    const int k = static_cast<int>(i - static_cast<long long>(o) * K);
    const uint32_t e = pp.info[k / kKChunk];
    const int r = e & 7, s = (e >> 3) & 7;
    const int kc = k % kKChunk;
    const int ci = static_cast<int>(e >> 13) + kc;
    float v = 0.f;
    if (pp.row_taps) {
      // one K step per kernel row r: k = s * 8 + ch (8 channels per tap, taps s >= kw are zero)
      const int sx = kc >> 3, ch = kc & 7;
      if (sx < pp.kw) v = w[((static_cast<long long>(o) * 8 + ch) * pp.kh + r) * pp.kw + sx];
    } else if (kc <= static_cast<int>((e >> 6) & 127u)) {
      const long long idx = pp.transposed
                                ? ((static_cast<long long>(ci) * pp.c_out + o) * pp.kh + r) * pp.kw + s
                                : ((static_cast<long long>(o) * pp.c_in_total + ci) * pp.kh + r) * pp.kw + s;
      v = w[idx];
    }
    out[i] = __float2half_rn(v);
  }
}

// =====================================================================================================
// host side: geometry -> k-step table, tensor maps, launch
// =====================================================================================================
namespace {

struct Plan {
  int n_maps = 0;
  int map_src[4] = {0, 0, 0, 0};
  int map_ph[4] = {0, 0, 0, 0}, map_pw[4] = {0, 0, 0, 0};
  int in_stride_mul = 1;  // 2 for phase views of a stride-2 conv
  std::vector<uint32_t> ksteps;
  std::vector<uint32_t> pack_info;
  int n_sub = 1;
  int sub_kbeg[4] = {0, 0, 0, 0}, sub_kend[4] = {0, 0, 0, 0};
  int sub_p[4] = {0, 0, 0, 0}, sub_q[4] = {0, 0, 0, 0};
  int h_out = 0, w_out = 0, os = 1, full_h = 0, full_w = 0;
  int c_in_total = 0;
  int block_n = 0;
  bool row_taps = false;
};

int pick_block_n(int c_out) {
  if (c_out % 256 == 0) return 256;
  if (c_out % 192 == 0) return 192;
  if (c_out == 160 || c_out == 320) return 160;
  if (c_out % 128 == 0) return 128;
  if (c_out == 96) return 96;
  if (c_out % 64 == 0) return 64;
  if (c_out == 16) return 16;
  return 0;
}

int build_plan(const stemb200_conv_desc& d, Plan& pl) {
  if (d.batch < 1 || d.h_in < 1 || d.w_in < 1) return set_error("conv: bad shape");
  if (d.n_src < 1 || d.n_src > 3) return set_error("conv: n_src must be 1..3");
  if (d.kh != d.kw || (d.kh != 1 && d.kh != 3 && d.kh != 5)) return set_error("conv: kernel must be 1,3,5");
  if (d.stride != 1 && d.stride != 2) return set_error("conv: stride must be 1 or 2");
  if (d.transposed && d.stride != 2) return set_error("conv: transposed needs stride 2");
  if ((d.stride == 2) && d.n_src != 1) return set_error("conv: strided conv takes one source");
  int src_base[3] = {0, 0, 0};
  pl.c_in_total = 0;
  for (int s = 0; s < d.n_src; ++s) {
    if (d.c_in[s] < 8 || d.c_in[s] % 8) return set_error("conv: c_in must be a multiple of 8");
    src_base[s] = pl.c_in_total;
    pl.c_in_total += d.c_in[s];
  }
  pl.block_n = pick_block_n(d.c_out);
  // two 160-wide tiles need the clipped store map of the plain stride-1 / strided path (setup_params)
  if (d.c_out == 320 && (d.transposed || d.epilogue != STEMB200_EPI_LINEAR)) pl.block_n = 64;
  if (!pl.block_n) return set_error("conv: unsupported c_out");
  if (pl.block_n < 32 && !d.direct_store) return set_error("conv: c_out < 32 needs direct_store");
  if (!d.direct_store && (d.c_out % 64) && d.c_out != 160 && d.c_out != 320)
    return set_error("conv: the TMA-store epilogue needs c_out % 64 == 0 (or c_out == 160)");
  if (d.epilogue < 0 || d.epilogue > 2) return set_error("conv: bad epilogue");
  if (d.epilogue != STEMB200_EPI_LINEAR && d.direct_store)
    return set_error("conv: SFT / residual epilogues store through the TMA path");
  // the residual epilogue may write fp32 (acc + bias + fp16 residual, not rounded: tensors that feed quantisation)
  if (d.epilogue == STEMB200_EPI_SFT && d.out_dtype != STEMB200_DT_F16)
    return set_error("conv: the SFT epilogue writes fp16");
  if (d.epilogue == STEMB200_EPI_SFT) {
    pl.block_n = d.c_out % 256 == 0 ? 256 : (d.c_out % 128 == 0 ? 128 : 0);
    if (!pl.block_n) return set_error("conv: SFT epilogue needs c_out (= 2 x channels) % 128 == 0");
  }
  if (d.epilogue == STEMB200_EPI_ADD && pl.block_n != 64 && pl.block_n != 128 && pl.block_n != 192 &&
      pl.block_n != 256)
    return set_error("conv: residual epilogue supports c_out % 64 == 0");

  const int k = d.kh, pad = k / 2;
  const uint32_t mask = d.tap_mask ? d.tap_mask : ((1u << (k * k)) - 1u);
  auto tap_on = [&](int r, int s) { return (mask >> (r * k + s)) & 1u; };
  auto push_v = [&](int map, int dh, int dw, int c0, int r, int s, int cbase, int valid) {
    pl.ksteps.push_back(static_cast<uint32_t>(map) | (static_cast<uint32_t>(dh + 8) << 2) |
                        (static_cast<uint32_t>(dw + 8) << 6) | (static_cast<uint32_t>(c0) << 10));
    pl.pack_info.push_back(static_cast<uint32_t>(r) | (static_cast<uint32_t>(s) << 3) |
                           (static_cast<uint32_t>(valid - 1) << 6) | (static_cast<uint32_t>(cbase) << 13));
  };

  if (d.row_taps) {
    if (d.transposed || d.stride != 2 || d.n_src != 1 || d.c_in[0] != 8 || d.tap_mask || (d.h_in & 1) || (d.w_in & 1))
      return set_error("conv: row_taps needs a plain stride-2 conv with c_in == 8 and even h_in / w_in");
    pl.row_taps = true;
    pl.n_maps = 2;  // even / odd canvas rows
    pl.h_out = d.h_in / 2;
    pl.w_out = d.w_in / 2;
    for (int r = 0; r < k; ++r) push_v(r & 1, r >> 1, 0, 0, r, 0, 0, kKChunk);
    pl.n_sub = 1;
    pl.sub_kbeg[0] = 0;
    pl.sub_kend[0] = static_cast<int>(pl.ksteps.size());
    pl.os = 1;
    pl.full_h = pl.h_out;
    pl.full_w = pl.w_out;
  } else if (!d.transposed && d.stride == 1) {
    pl.n_maps = d.n_src;
    for (int s = 0; s < d.n_src; ++s) pl.map_src[s] = s;
    pl.h_out = d.h_in;
    pl.w_out = d.w_in;
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s) {
        if (!tap_on(r, s)) continue;
        for (int src = 0; src < d.n_src; ++src)
          for (int c0 = 0; c0 < d.c_in[src]; c0 += kKChunk) push_v(src, r - pad, s - pad, c0, r, s, src_base[src] + c0, std::min(kKChunk, d.c_in[src] - c0));
      }
    pl.n_sub = 1;
    pl.sub_kbeg[0] = 0;
    pl.sub_kend[0] = static_cast<int>(pl.ksteps.size());
    pl.os = 1;
    pl.full_h = pl.h_out;
    pl.full_w = pl.w_out;
  } else if (!d.transposed) {  // stride 2
    pl.n_maps = 4;
    pl.in_stride_mul = 2;
    for (int m = 0; m < 4; ++m) {
      pl.map_ph[m] = m >> 1;
      pl.map_pw[m] = m & 1;
    }
    pl.h_out = (d.h_in + 2 * pad - k) / 2 + 1;
    pl.w_out = (d.w_in + 2 * pad - k) / 2 + 1;
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s) {
        if (!tap_on(r, s)) continue;
        const int a = r - pad, b = s - pad;
        const int ph = a & 1, pw = b & 1;
        const int dh = (a - ph) / 2, dw = (b - pw) / 2;
        for (int c0 = 0; c0 < d.c_in[0]; c0 += kKChunk) push_v(ph * 2 + pw, dh, dw, c0, r, s, c0, std::min(kKChunk, d.c_in[0] - c0));
      }
    pl.n_sub = 1;
    pl.sub_kbeg[0] = 0;
    pl.sub_kend[0] = static_cast<int>(pl.ksteps.size());
    pl.os = 1;
    pl.full_h = pl.h_out;
    pl.full_w = pl.w_out;
  } else {  // transposed, stride 2, padding k/2, output_padding 1: out = 2*in
    pl.n_maps = 1;
    pl.h_out = d.h_in;
    pl.w_out = d.w_in;
    pl.os = 2;
    pl.full_h = 2 * d.h_in;
    pl.full_w = 2 * d.w_in;
    pl.n_sub = 4;
    for (int sp = 0; sp < 4; ++sp) {
      const int p = sp >> 1, q = sp & 1;
      pl.sub_p[sp] = p;
      pl.sub_q[sp] = q;
      pl.sub_kbeg[sp] = static_cast<int>(pl.ksteps.size());
      // oh = 2*ih - pad + r  with oh = 2*h + p  =>  ih = h + (p + pad - r)/2, needs (p + pad - r) even
      for (int r = 0; r < k; ++r) {
        if ((p + pad - r) & 1) continue;
        for (int s = 0; s < k; ++s) {
          if ((q + pad - s) & 1) continue;
          if (!tap_on(r, s)) continue;
          const int dh = (p + pad - r) / 2, dw = (q + pad - s) / 2;
          for (int c0 = 0; c0 < d.c_in[0]; c0 += kKChunk) push_v(0, dh, dw, c0, r, s, c0, std::min(kKChunk, d.c_in[0] - c0));
        }
      }
      pl.sub_kend[sp] = static_cast<int>(pl.ksteps.size());
    }
  }
  if (pl.ksteps.empty() || pl.ksteps.size() > kMaxKSteps) return set_error("conv: K-step table overflow");
  return 0;
}

void pick_tile(int h, int w, int& th, int& tw) {
  double best = -1.0;
  int bth = 1, btw = std::min(w, 128);
  for (int a = 1; a <= 128; ++a) {
    for (int b = 1; b <= 128 / a; ++b) {
      if (b > 256) continue;
      const long long tiles = static_cast<long long>((h + a - 1) / a) * ((w + b - 1) / b);
      const double eff = static_cast<double>(h) * w / (static_cast<double>(tiles) * 128.0);
      // tie-break towards square patches (smaller halo => better L2 reuse across taps)
      const double score = eff - 1e-4 * (static_cast<double>(a + b) / (a * b));
      if (score > best) {
        best = score;
        bth = a;
        btw = b;
      }
    }
  }
  th = bth;
  tw = btw;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// NHWC view {C, W, H, N} with optional phase sub-sampling (mul) and origin (ph, pw)
int encode_nhwc(CUtensorMap* m, const void* base, int elem_bytes, int n, int h, int w, int c, int mul, int ph,
                int pw, int box_c, int box_w, int box_h, int c_extent = 0) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled unavailable");
  const int hv = (h - ph + mul - 1) / mul, wv = (w - pw + mul - 1) / mul;
  if (hv < 1 || wv < 1) return set_error("conv: empty phase view");
  // c_extent < c clips the channel dimension (stores beyond it are dropped) while keeping the pixel pitch c
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(c_extent > 0 ? c_extent : c), static_cast<cuuint64_t>(wv),
                        static_cast<cuuint64_t>(hv), static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * elem_bytes * mul,
                           static_cast<cuuint64_t>(w) * c * elem_bytes * mul,
                           static_cast<cuuint64_t>(h) * w * c * elem_bytes};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w),
                       static_cast<cuuint32_t>(box_h), 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const char* origin = static_cast<const char*>(base) +
                       (static_cast<size_t>(ph) * w + pw) * static_cast<size_t>(c) * elem_bytes;
  CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<char*>(origin), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(nhwc) failed: %d (c=%d w=%d h=%d n=%d box=%d,%d,%d)",
             static_cast<int>(r), c, wv, hv, n, box_c, box_w, box_h);
    return set_error(buf);
  }
  return 0;
}

// row_taps first layer: view {64 contiguous fp16, w_out (stride 2 pixels = 32 B), rows of one parity (stride 2 canvas
// rows), n} over the zero-bordered NHWC8 canvas [n][h + 2b][w + 2b][8]. The 64-element box starting at canvas pixel
// 2*ow holds the 8 taps x 8 channels of one kernel row of output pixel ow (windows of neighbouring pixels overlap).
int encode_row_taps(CUtensorMap* m, const void* base, int n, int h, int w, int border, int parity, int box_w,
                    int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled unavailable");
  const long long hc = h + 2 * border, wc = w + 2 * border;
  const long long rows = (hc - parity + 1) / 2;
  cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(w / 2), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {32, static_cast<cuuint64_t>(wc) * 8 * 2 * 2, static_cast<cuuint64_t>(hc * wc) * 8 * 2};
  cuuint32_t box[4] = {64, static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const char* origin = static_cast<const char*>(base) + static_cast<size_t>(parity) * wc * 8 * 2;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<char*>(origin), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(row_taps) failed: %d (w=%d h=%d n=%d box=%d,%d)",
             static_cast<int>(r), w, h, n, box_w, box_h);
    return set_error(buf);
  }
  return 0;
}

int encode_weight(CUtensorMap* m, const void* base, long long K, int c_out, int block_n) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(c_out)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kKChunk), static_cast<cuuint32_t>(block_n)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(weight) failed: %d", static_cast<int>(r));
    return set_error(buf);
  }
  return 0;
}

}  // namespace

int encode_nhwc_plain(CUtensorMap* m, const void* base, int n, int h, int w, int c, int box_c, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error("cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h),
                        static_cast<cuuint64_t>(n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w) * c * 2,
                           static_cast<cuuint64_t>(h) * w * c * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w),
                       static_cast<cuuint32_t>(box_h), 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(plain nhwc) failed: %d (c=%d w=%d h=%d n=%d box=%d,%d,%d)",
             static_cast<int>(r), c, w, h, n, box_c, box_w, box_h);
    return set_error(buf);
  }
  return 0;
}

namespace {
// STEMB200_DUO=0 in the environment turns duo mode (cta_group::2 MMAs) off: the clusters then run in pair mode
// (weight tiles multicast, one cta_group::1 MMA per CTA) - for A/B measurements
bool duo_mode_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("STEMB200_DUO");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// STEMB200_PAIR=0 in the environment disables pair mode (A/B measurements)
bool pair_mode_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("STEMB200_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// launch `kernel` as clusters of two CTAs; grid = 2 x min(co-resident clusters, work items)
template <typename Kern>
int launch_pairs(Kern kernel, const ConvKernelParams& kp, int threads, size_t smem, cudaStream_t stream,
                 int& cached_clusters) {
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cached_clusters <= 0) {
    cfg.gridDim = dim3(2 * (num_sms() / 2));
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
    if (e != cudaSuccess || n < 1) return set_cuda_error("cudaOccupancyMaxActiveClusters", e);
    cached_clusters = n;
  }
  const int clusters = std::min(cached_clusters, kp.total_tiles / 2);
  cfg.gridDim = dim3(2 * clusters);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, kp);
  count_launch();
  if (e != cudaSuccess) return set_cuda_error("cluster launch", e);
  return 0;
}

template <int BLOCK_N, int EPI>
int launch_conv_duo(const ConvKernelParams& kp, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, true>;
  static bool configured = false;  // benign race: attribute set is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, EPI, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(conv duo)", e);
    configured = true;
  }
  static int clusters = 0;
  return launch_pairs(conv_igemm_kernel<BLOCK_N, EPI, true>, kp, kNumThreads, Cfg::kSmemBytes, stream, clusters);
}

template <int BLOCK_N, int EPI = 0>
int launch_conv(const ConvKernelParams& kp, int grid, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N>;
  if constexpr (BLOCK_N >= 64) {
    if (kp.csize == 2 && kp.duo) return launch_conv_duo<BLOCK_N, EPI>(kp, stream);
  }
  static bool configured = false;  // benign race: attribute set is idempotent
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BLOCK_N, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(conv)", e);
    configured = true;
  }
  if (kp.csize == 2) {
    static int clusters = 0;
    return launch_pairs(conv_igemm_kernel<BLOCK_N, EPI>, kp, kNumThreads, Cfg::kSmemBytes, stream, clusters);
  }
  conv_igemm_kernel<BLOCK_N, EPI><<<grid, kNumThreads, Cfg::kSmemBytes, stream>>>(kp);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("conv_igemm launch", e);
  return 0;
}

}  // namespace
}  // namespace stem

using namespace stem;

extern "C" int64_t stemb200_conv2d_packed_k(const stemb200_conv_desc* d) {
  if (!d) return STEMB200_E_INVALID;
  Plan pl;
  if (int rc = build_plan(*d, pl)) return rc;
  return static_cast<int64_t>(pl.ksteps.size()) * kKChunk;
}

extern "C" int stemb200_conv2d_pack_weight(const stemb200_conv_desc* d, const float* weight_f32,
                                           void* packed_f16, void* stream) {
  if (!d || !weight_f32 || !packed_f16) return set_error("pack_weight: null argument");
  Plan pl;
  if (int rc = build_plan(*d, pl)) return rc;
  PackParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.n_steps = static_cast<int>(pl.pack_info.size());
  for (int i = 0; i < pp.n_steps; ++i) pp.info[i] = pl.pack_info[i];
  pp.c_out = d->c_out;
  pp.c_in_total = pl.c_in_total;
  pp.kh = d->kh;
  pp.kw = d->kw;
  pp.transposed = d->transposed;
  pp.row_taps = pl.row_taps ? 1 : 0;
  const long long total = static_cast<long long>(pp.n_steps) * kKChunk * d->c_out;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  pack_weight_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      pp, weight_f32, static_cast<__half*>(packed_f16));
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("pack_weight launch", e);
  return 0;
}

namespace {
// geometry + pointers -> kernel parameters (tensor maps, k-step table, tiling)
int setup_params(const stemb200_conv_desc* d, const Plan& pl, const void* const* in, const void* packed_weight,
                 const float* bias, const void* aux, void* out, ConvKernelParams& kp) {
  memset(&kp, 0, sizeof(kp));
  int th = d->tile_h, tw = d->tile_w;
  if (th <= 0 || tw <= 0) pick_tile(pl.h_out, pl.w_out, th, tw);
  if (th * tw > 128 || tw > 256 || th > 256) return set_error("conv2d_fwd: tile_h*tile_w must be <= 128");

  for (int m = 0; m < 4; ++m) {
    const int mm = m < pl.n_maps ? m : 0;
    if (pl.row_taps) {
      if (int rc = encode_row_taps(&kp.a_map[m], in[0], d->batch, d->h_in, d->w_in, d->kh / 2, mm, tw, th)) return rc;
      continue;
    }
    const int src = pl.map_src[mm];
    if (int rc = encode_nhwc(&kp.a_map[m], in[src], 2, d->batch, d->h_in, d->w_in, d->c_in[src],
                             pl.in_stride_mul, pl.map_ph[mm], pl.map_pw[mm], kKChunk, tw, th))
      return rc;
  }
  const long long K = static_cast<long long>(pl.ksteps.size()) * kKChunk;
  if (int rc = encode_weight(&kp.b_map, packed_weight, K, d->c_out, pl.block_n)) return rc;
  const int out_bytes = d->out_dtype == STEMB200_DT_F32 ? 4 : 2;
  const int out_box_c = d->out_dtype == STEMB200_DT_F32 ? 32 : 64;
  const int out_c = d->epilogue == STEMB200_EPI_SFT ? d->c_out / 2 : d->c_out;
  if (!d->direct_store) {
    for (int sp = 0; sp < 4; ++sp) {
      const int ss = sp < pl.n_sub ? sp : 0;
      if (int rc = encode_nhwc(&kp.out_map[sp], out, out_bytes, d->batch, pl.full_h, pl.full_w, out_c, pl.os,
                               pl.sub_p[ss], pl.sub_q[ss], out_box_c, tw, th))
        return rc;
    }
  }
  if (!d->direct_store && pl.block_n == 160 && d->c_out == 320) {
    // BLOCK_N = 160 has an odd number of 32-column chunks: the idle half of the last 64-channel store group would
    // spill stale staging data into the next N tile's channels, so the first tile stores through a map clipped at 160
    if (pl.n_sub != 1) return set_error("conv: c_out == 320 supports stride-1 / strided convs only");
    if (int rc = encode_nhwc(&kp.out_map[1], out, out_bytes, d->batch, pl.full_h, pl.full_w, out_c, pl.os,
                             pl.sub_p[0], pl.sub_q[0], out_box_c, tw, th, 160))
      return rc;
  }
  kp.g_map = kp.b_map;
  for (size_t i = 0; i < pl.ksteps.size(); ++i) kp.ksteps[i] = pl.ksteps[i];
  for (int i = 0; i < 4; ++i) {
    kp.sub_kbeg[i] = pl.sub_kbeg[i];
    kp.sub_kend[i] = pl.sub_kend[i];
    kp.sub_p[i] = pl.sub_p[i];
    kp.sub_q[i] = pl.sub_q[i];
  }
  kp.n_sub = pl.n_sub;
  kp.batch = d->batch;
  kp.h_out = pl.h_out;
  kp.w_out = pl.w_out;
  kp.tile_h = th;
  kp.tile_w = tw;
  kp.tiles_h = (pl.h_out + th - 1) / th;
  kp.tiles_w = (pl.w_out + tw - 1) / tw;
  kp.n_tiles_n = d->c_out / pl.block_n;
  const long long total =
      static_cast<long long>(pl.n_sub) * d->batch * kp.tiles_h * kp.tiles_w * kp.n_tiles_n;
  if (total > 0x7fffffffLL) return set_error("conv2d_fwd: too many tiles");
  kp.total_tiles = static_cast<int>(total);
  // pair mode: clusters of two CTAs on neighbouring pixel tiles of the same sub-problem share the weight tiles
  kp.csize = 1;
  const long long tiles_per_sub = static_cast<long long>(d->batch) * kp.tiles_h * kp.tiles_w;
  // (measured +2 % on the MMA-bound 5x5 layers, -2 % on the 160-wide tiles and the epilogue-bound first layer)
  // (the 160-wide tiles lose 2 % in pair mode but gain in duo mode, where each CTA stages 80 of the 160 weight rows)
  // An odd tile count per sub-problem (11 frames: 2805 tiles in g_a.4 / g_s.2) used to keep a layer out of cluster
  // mode. The last cluster then gets a phantom second tile: its index runs one past the end and decode_tile wraps it
  // to tile (0, 0) of image 0, which is simply computed and stored twice with identical values (one tile in thousands;
  // STEMB200_PHANTOM=0 restores the old rule).
  static const bool phantom = [] { const char* e = getenv("STEMB200_PHANTOM"); return !(e && e[0] == '0'); }();
  const bool odd = (tiles_per_sub % 2) != 0;
  if (pair_mode_enabled() && pl.block_n >= 64 && (pl.block_n != 160 || duo_mode_enabled()) && !pl.row_taps &&
      (!odd || (phantom && tiles_per_sub >= 255)) && total >= 2LL * num_sms()) {
    kp.csize = 2;
    if (odd) kp.total_tiles = static_cast<int>(static_cast<long long>(pl.n_sub) * (tiles_per_sub + 1) * kp.n_tiles_n);
    if (int rc = encode_weight(&kp.b_half_map, packed_weight, K, d->c_out, pl.block_n / 2)) return rc;
  }
  kp.kk_main = (pl.row_taps && d->kw * 8 <= 48) ? 3 : 4;
  kp.c_out = d->c_out;
  kp.slope = d->lrelu_slope;
  kp.sq_scale = d->sq_scale;
  kp.sq_inv = d->sq_scale != 0.f ? 1.0f / (d->sq_scale * d->sq_scale) : 1.0f;
  kp.out_f32 = d->out_dtype == STEMB200_DT_F32;
  kp.direct = d->direct_store;
  kp.os = pl.os;
  kp.full_h = pl.full_h;
  kp.full_w = pl.full_w;
  kp.bias = bias;
  kp.aux = static_cast<const __half*>(aux);
  kp.out = out;
  kp.beta = bias;
  kp.igdn = 0;
  return 0;
}
}  // namespace

namespace {
// Less than one wave of work items (small batches: one 1080p latent is 64 pixel tiles): a 256-wide layer leaves more
// than half of the SMs idle. Narrower N tiles over the same packed weights ([c_out][K] does not depend on the tiling)
// create more items; pick the width that minimises rounds x tile width, ties towards the wider tile (fewer re-reads of
// the A operand). 64-wide tiles are not considered: their k-steps are bound by shared-memory reads, not by the MMA.
void adapt_block_n(const stemb200_conv_desc& d, Plan& pl) {
  if (d.epilogue == STEMB200_EPI_SFT || d.direct_store || d.out_dtype != STEMB200_DT_F16) return;
  if (pl.block_n != 256 && pl.block_n != 192) return;
  int th = d.tile_h, tw = d.tile_w;
  if (th <= 0 || tw <= 0) pick_tile(pl.h_out, pl.w_out, th, tw);
  const long long tiles_m = static_cast<long long>(pl.n_sub) * d.batch * ((pl.h_out + th - 1) / th) *
                            ((pl.w_out + tw - 1) / tw);
  const int sms = num_sms();
  if (tiles_m * (d.c_out / pl.block_n) >= sms) return;
  int best = pl.block_n;
  long long best_cost = ((tiles_m * (d.c_out / pl.block_n) + sms - 1) / sms) * pl.block_n;
  for (int bn : {256, 192, 128}) {
    if (d.c_out % bn) continue;
    const long long cost = ((tiles_m * (d.c_out / bn) + sms - 1) / sms) * bn;
    if (cost < best_cost) {
      best_cost = cost;
      best = bn;
    }
  }
  pl.block_n = best;
}
}  // namespace

extern "C" int stemb200_conv2d_fwd(const stemb200_conv_desc* d, const void* const* in,
                                   const void* packed_weight, const float* bias, const void* aux, void* out,
                                   void* stream) {
  if (!d || !in || !packed_weight || !bias || !out) return set_error("conv2d_fwd: null argument");
  Plan pl;
  if (int rc = build_plan(*d, pl)) return rc;
  adapt_block_n(*d, pl);
  if (d->epilogue != STEMB200_EPI_LINEAR && !aux) return set_error("conv2d_fwd: SFT / residual epilogue needs aux");
  for (int s = 0; s < d->n_src; ++s)
    if (!in[s]) return set_error("conv2d_fwd: null input");
  ConvKernelParams kp;
  if (int rc = setup_params(d, pl, in, packed_weight, bias, aux, out, kp)) return rc;
  kp.duo = (kp.csize == 2 && duo_mode_enabled()) ? 1 : 0;
  const int grid = std::min(kp.total_tiles, num_sms());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d->epilogue == STEMB200_EPI_SFT) {
    if (pl.block_n == 256) return launch_conv<256, 1>(kp, grid, st);
    return launch_conv<128, 1>(kp, grid, st);
  }
  if (d->epilogue == STEMB200_EPI_ADD) {
    switch (pl.block_n) {
      case 64: return launch_conv<64, 2>(kp, grid, st);
      case 128: return launch_conv<128, 2>(kp, grid, st);
      case 192: return launch_conv<192, 2>(kp, grid, st);
      default: return launch_conv<256, 2>(kp, grid, st);
    }
  }
  switch (pl.block_n) {
    case 16: return launch_conv<16>(kp, grid, st);
    case 64: return launch_conv<64>(kp, grid, st);
    case 96: return launch_conv<96>(kp, grid, st);
    case 128: return launch_conv<128>(kp, grid, st);
    case 160: return launch_conv<160>(kp, grid, st);
    case 192: return launch_conv<192>(kp, grid, st);
    case 256: return launch_conv<256>(kp, grid, st);
    default: return set_error("conv2d_fwd: no kernel for this c_out");
  }
}

namespace {
template <int kNT, bool kInverse, bool kLast = false, bool kDuo = false, bool kTm = false>
int launch_gdn(const ConvKernelParams& kp, int grid, cudaStream_t stream) {
  using Cfg = GdnCfgT<kNT, kLast, kDuo, kTm>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gdn_kernel<kNT, kInverse, kLast, kDuo, kTm>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(conv_gdn)", e);
    configured = true;
  }
  if (kp.csize == 2) {
    static int clusters = 0;
    return launch_pairs(conv_gdn_kernel<kNT, kInverse, kLast, kDuo, kTm>, kp, kGdnThreads, Cfg::kSmemBytes, stream,
                        clusters);
  }
  if constexpr (kDuo) {
    return set_error("conv_gdn: duo mode needs a cluster launch");
  } else {
    conv_gdn_kernel<kNT, kInverse, kLast, false, kTm><<<grid, kGdnThreads, Cfg::kSmemBytes, stream>>>(kp);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error("conv_gdn launch", e);
    return 0;
  }
}
}  // namespace

namespace {
bool pp_mode_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("STEMB200_PP");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int kNT, bool kInverse>
int launch_gdn_pp(const ConvKernelParams& kp, int grid, cudaStream_t stream) {
  using Cfg = GdnPpCfgT<kNT>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gdn_pp_kernel<kNT, kInverse>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(conv_gdn_pp)", e);
    configured = true;
  }
  if (kp.csize == 2) {
    static int clusters = 0;
    return launch_pairs(conv_gdn_pp_kernel<kNT, kInverse>, kp, kGdnThreads, Cfg::kSmemBytes, stream, clusters);
  }
  conv_gdn_pp_kernel<kNT, kInverse><<<grid, kGdnThreads, Cfg::kSmemBytes, stream>>>(kp);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("conv_gdn_pp launch", e);
  return 0;
}

int gdn_forward(const stemb200_conv_desc* d, const void* const* in, const void* packed_weight, const float* bias,
                const void* packed_gamma, const float* beta, int32_t inverse, void* out, const void* packed_w6,
                void* col_out, void* stream) {
  if (!d || !in || !packed_weight || !bias || !packed_gamma || !beta) return set_error("conv2d_gdn_fwd: null argument");
  const bool last = packed_w6 != nullptr;
  if (!last && !out) return set_error("conv2d_gdn_fwd: null output");
  if (last && (!col_out || d->c_out != 192 || !inverse))
    return set_error("conv2d_gdn_last_fwd: needs col_out, c_out == 192 and IGDN");
  if (d->c_out != 192 && d->c_out != 128) return set_error("conv2d_gdn_fwd: fused GDN needs c_out == 128 or 192");
  if (d->out_dtype != STEMB200_DT_F16 || d->direct_store || d->lrelu_slope != 1.0f || d->sq_scale <= 0.f)
    return set_error("conv2d_gdn_fwd: needs fp16 TMA-store output, no activation, sq_scale > 0");
  stemb200_conv_desc dd = *d;
  dd.epilogue = STEMB200_EPI_LINEAR;
  Plan pl;
  if (int rc = build_plan(dd, pl)) return rc;
  pl.block_n = d->c_out;  // one N tile holds every channel of a pixel
  for (int s = 0; s < d->n_src; ++s)
    if (!in[s]) return set_error("conv2d_gdn_fwd: null input");
  ConvKernelParams kp;
  // without an activation output the store maps are built over the col buffer (never dereferenced)
  if (int rc = setup_params(&dd, pl, in, packed_weight, bias, nullptr, out ? out : col_out, kp)) return rc;
  if (int rc = encode_weight(&kp.g_map, packed_gamma, d->c_out, d->c_out, d->c_out)) return rc;
  if (kp.csize == 2)
    if (int rc = encode_weight(&kp.g_half_map, packed_gamma, d->c_out, d->c_out, d->c_out / 2)) return rc;
  kp.beta = beta;
  kp.igdn = inverse ? 1 : 0;
  const int grid = std::min(kp.total_tiles, num_sms());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // short main loops are epilogue-bound: ping-pong variant (GDN only: the transposed layers have long enough loops)
  int max_k = 0;
  for (int i = 0; i < pl.n_sub; ++i) max_k = std::max(max_k, pl.sub_kend[i] - pl.sub_kbeg[i]);
  const bool use_pp = !last && pp_mode_enabled() && !inverse && max_k <= 8;
  // (the fused-last variant: duo mode since round 2 - with cluster-scope release arrives its three epilogue -> MMA
  // hand-offs per tile cost more than the halved B reads gave back (2.9 -> 3.3 ms); with CTA-scope arrives and five
  // 28 KB ring slots it is the faster one (2.66 -> 2.44 ms, the step 9.55 -> 9.27 ms).  STEMB200_DUO_LAST=0: pair mode)
  static const bool duo_last = [] { const char* e = getenv("STEMB200_DUO_LAST"); return !(e && e[0] == '0'); }();
  kp.duo = (kp.csize == 2 && !use_pp && (!last || duo_last) && duo_mode_enabled()) ? 1 : 0;
  if (last) {
    if (kp.duo)
      if (int rc = encode_weight(&kp.w6_half_map, packed_w6, d->c_out, kLastN, kLastN / 2)) return rc;
    if (int rc = encode_weight(&kp.w6_map, packed_w6, d->c_out, kLastN, kLastN)) return rc;
    kp.col_out = static_cast<__half*>(col_out);
    kp.store_act = out ? 1 : 0;
    // the layer's own activation is normally not stored: (x s)^2 and the output then live in tensor memory as MMA
    // operands, the epilogue needs no shared-memory buffer / named barrier / proxy fence, and the room goes to the ring
    // (STEMB200_LAST_TMEM=0: the kernel with shared-memory operands)
    static const bool tm_off = [] { const char* e = getenv("STEMB200_LAST_TMEM"); return e && e[0] == '0'; }();
    const bool tm = !out && !tm_off;
    if (kp.duo) return tm ? launch_gdn<192, true, true, true, true>(kp, grid, st) : launch_gdn<192, true, true, true>(kp, grid, st);
    return tm ? launch_gdn<192, true, true, false, true>(kp, grid, st) : launch_gdn<192, true, true>(kp, grid, st);
  }
  if (use_pp)
    return d->c_out == 192 ? launch_gdn_pp<192, false>(kp, grid, st) : launch_gdn_pp<128, false>(kp, grid, st);
  if (kp.duo) {
    if (d->c_out == 192)
      return inverse ? launch_gdn<192, true, false, true>(kp, grid, st) : launch_gdn<192, false, false, true>(kp, grid, st);
    return inverse ? launch_gdn<128, true, false, true>(kp, grid, st) : launch_gdn<128, false, false, true>(kp, grid, st);
  }
  if (d->c_out == 192) return inverse ? launch_gdn<192, true>(kp, grid, st) : launch_gdn<192, false>(kp, grid, st);
  return inverse ? launch_gdn<128, true>(kp, grid, st) : launch_gdn<128, false>(kp, grid, st);
}
}  // namespace

extern "C" int stemb200_conv2d_gdn_fwd(const stemb200_conv_desc* d, const void* const* in,
                                       const void* packed_weight, const float* bias, const void* packed_gamma,
                                       const float* beta, int32_t inverse, void* out, void* stream) {
  return gdn_forward(d, in, packed_weight, bias, packed_gamma, beta, inverse, out, nullptr, nullptr, stream);
}

extern "C" int stemb200_conv2d_gdn_last_fwd(const stemb200_conv_desc* d, const void* const* in,
                                            const void* packed_weight, const float* bias, const void* packed_gamma,
                                            const float* beta, const void* packed_w6, void* col_out, void* act_out,
                                            void* stream) {
  if (!packed_w6) return set_error("conv2d_gdn_last_fwd: null W6");
  return gdn_forward(d, in, packed_weight, bias, packed_gamma, beta, 1, act_out, packed_w6, col_out, stream);
}

// entropy_parameters' last layer + GaussianConditional in one kernel (conv_gc_kernel above)
extern "C" int stemb200_conv2d_gc_fwd(const stemb200_conv_desc* d, const void* const* in, const void* packed_weight,
                                      const float* bias, const float* y_nhwc, const void* cond_f16, float scale_bound,
                                      float lik_bound, int32_t yhat_mode, float* y_hat_nchw, float* lik_nchw,
                                      double* bits, void* stream) {
  if (!d || !in || !in[0] || !packed_weight || !bias || !y_nhwc) return set_error("conv2d_gc_fwd: null argument");
  if (d->kh != 1 || d->stride != 1 || d->transposed || d->n_src != 1 || d->c_out % 128 || d->tap_mask ||
      d->epilogue != STEMB200_EPI_LINEAR || d->lrelu_slope != 1.0f || d->direct_store)
    return set_error("conv2d_gc_fwd: needs a plain 1x1 layer with c_out = 2 C, C % 64 == 0, no activation");
  if (reinterpret_cast<uintptr_t>(y_nhwc) & 15 || reinterpret_cast<uintptr_t>(cond_f16) & 15 ||
      reinterpret_cast<uintptr_t>(bias) & 15)
    return set_error("conv2d_gc_fwd: y, cond and bias must be 16-byte aligned");
  stemb200_conv_desc dd = *d;
  dd.out_dtype = STEMB200_DT_F32;
  Plan pl;
  if (int rc = build_plan(dd, pl)) return rc;
  pl.block_n = kGcFuseN;
  ConvKernelParams kp;
  // no activation tensor is written: the store maps are built over y (never dereferenced)
  dd.direct_store = 1;
  if (int rc = setup_params(&dd, pl, in, packed_weight, bias, nullptr, const_cast<float*>(y_nhwc), kp)) return rc;
  kp.csize = 1;
  kp.duo = 0;
  GcFuseParams g;
  g.y = y_nhwc;
  g.cond = static_cast<const __half*>(cond_f16);
  g.y_hat = y_hat_nchw;
  g.lik = lik_nchw;
  g.bits = bits;
  g.C = d->c_out / 2;
  g.hw = d->h_in * d->w_in;
  g.scale_bound = scale_bound;
  g.lik_bound = lik_bound;
  g.yhat_mode = yhat_mode;
  using Cfg = ConvCfg<kGcFuseN>;
  constexpr int kSmem = 1024 + Cfg::kStages * Cfg::kStageBytes + 256;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(conv_gc)", e);
    configured = true;
  }
  const int grid = std::min(kp.total_tiles, num_sms());
  conv_gc_kernel<<<grid, kGcFuseThreads, kSmem, static_cast<cudaStream_t>(stream)>>>(kp, g);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error("conv_gc launch", e);
  return 0;
}
