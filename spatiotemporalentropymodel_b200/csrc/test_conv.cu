// Standalone GPU bring-up test for the tcgen05 implicit-GEMM conv (not part of the product path).
// Checks stemb200_conv2d_fwd against a straightforward host-side evaluation of the PyTorch conv / deconv
// definition at sampled output positions, then times a few 1080p-shaped layers.
//   usage: test_conv.bin [filter-substring] [--time]
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../../include/stemb200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static float h2f(__half h) { return __half2float(h); }

struct Case {
  std::string name;
  stemb200_conv_desc d;
  int samples;
};

static stemb200_conv_desc mk(int n, int h, int w, std::vector<int> cin, int cout, int k, int stride, int transposed,
                             uint32_t mask, int epi, float slope, int odt, int wsq, int direct, int th = 0,
                             int tw = 0) {
  stemb200_conv_desc d;
  memset(&d, 0, sizeof(d));
  d.batch = n;
  d.h_in = h;
  d.w_in = w;
  d.n_src = (int)cin.size();
  for (int i = 0; i < d.n_src; ++i) d.c_in[i] = cin[i];
  d.c_out = cout;
  d.kh = d.kw = k;
  d.stride = stride;
  d.transposed = transposed;
  d.tap_mask = mask;
  d.lrelu_slope = slope;
  d.out_dtype = odt;
  d.sq_scale = 0.125f;
  (void)epi;
  (void)wsq;
  d.tile_h = th;
  d.tile_w = tw;
  d.direct_store = direct;
  return d;
}

static int run_case(const Case& cs, bool timing) {
  const stemb200_conv_desc& d = cs.d;
  const int k = d.kh, pad = k / 2;
  int cin_tot = 0, src_base[3] = {0, 0, 0};
  for (int s = 0; s < d.n_src; ++s) {
    src_base[s] = cin_tot;
    cin_tot += d.c_in[s];
  }
  int ho, wo;
  if (d.transposed) {
    ho = 2 * d.h_in;
    wo = 2 * d.w_in;
  } else {
    ho = (d.h_in + 2 * pad - k) / d.stride + 1;
    wo = (d.w_in + 2 * pad - k) / d.stride + 1;
  }
  std::mt19937 rng(1234);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  // inputs (fp16 NHWC per source)
  std::vector<std::vector<__half>> hin(d.n_src);
  std::vector<void*> din(d.n_src);
  for (int s = 0; s < d.n_src; ++s) {
    size_t ne = (size_t)d.batch * d.h_in * d.w_in * d.c_in[s];
    hin[s].resize(ne);
    for (size_t i = 0; i < ne; ++i)
      hin[s][i] = __float2half_rn(U(rng));
    CK(cudaMalloc(&din[s], ne * 2));
    CK(cudaMemcpy(din[s], hin[s].data(), ne * 2, cudaMemcpyHostToDevice));
  }
  // weights fp32, pre-rounded to fp16-representable values
  const size_t nw = (size_t)d.c_out * cin_tot * k * k;
  std::vector<float> hw(nw);
  const float wscale = 1.0f / sqrtf((float)cin_tot * k * k);
  for (size_t i = 0; i < nw; ++i) hw[i] = h2f(__float2half_rn(U(rng) * wscale * 2.f));
  float* dw;
  CK(cudaMalloc(&dw, nw * 4));
  CK(cudaMemcpy(dw, hw.data(), nw * 4, cudaMemcpyHostToDevice));
  std::vector<float> hb(d.c_out);
  for (auto& b : hb) b = U(rng) * 0.5f;
  float* db;
  CK(cudaMalloc(&db, d.c_out * 4));
  CK(cudaMemcpy(db, hb.data(), d.c_out * 4, cudaMemcpyHostToDevice));

  const int64_t K = stemb200_conv2d_packed_k(&d);
  if (K <= 0) {
    printf("[%s] packed_k failed: %s\n", cs.name.c_str(), stemb200_last_error());
    return 1;
  }
  void* dpk;
  CK(cudaMalloc(&dpk, (size_t)K * d.c_out * 2));
  if (stemb200_conv2d_pack_weight(&d, dw, dpk, 0)) {
    printf("[%s] pack failed: %s\n", cs.name.c_str(), stemb200_last_error());
    return 1;
  }
  const size_t nout = (size_t)d.batch * ho * wo * d.c_out;
  const int obytes = d.out_dtype == STEMB200_DT_F32 ? 4 : 2;
  void* dout;
  CK(cudaMalloc(&dout, nout * obytes));
  CK(cudaMemset(dout, 0xFF, nout * obytes));  // NaN pattern: unwritten outputs are caught
  int rc = stemb200_conv2d_fwd(&d, din.data(), dpk, db, nullptr, dout, 0);
  if (rc) {
    printf("[%s] conv2d_fwd failed: %s\n", cs.name.c_str(), stemb200_last_error());
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%s] kernel error: %s\n", cs.name.c_str(), cudaGetErrorString(e));
    return 2;
  }
  std::vector<char> hout(nout * obytes);
  CK(cudaMemcpy(hout.data(), dout, nout * obytes, cudaMemcpyDeviceToHost));

  auto in_at = [&](int n, int ih, int iw, int ci) -> float {
    if (ih < 0 || ih >= d.h_in || iw < 0 || iw >= d.w_in) return 0.f;
    int s = 0;
    while (s + 1 < d.n_src && ci >= src_base[s + 1]) ++s;
    return h2f(hin[s][(((size_t)n * d.h_in + ih) * d.w_in + iw) * d.c_in[s] + (ci - src_base[s])]);
  };
  const uint32_t mask = d.tap_mask ? d.tap_mask : ((1u << (k * k)) - 1u);
  auto ref_at = [&](int n, int oh, int ow, int co) -> double {
    double acc = 0.0;
    for (int r = 0; r < k; ++r)
      for (int s = 0; s < k; ++s) {
        if (!((mask >> (r * k + s)) & 1u)) continue;
        int ih, iw;
        if (d.transposed) {
          if ((oh + pad - r) & 1) continue;
          if ((ow + pad - s) & 1) continue;
          ih = (oh + pad - r) / 2;
          iw = (ow + pad - s) / 2;
          if (oh + pad - r < 0 || ow + pad - s < 0) continue;
        } else {
          ih = oh * d.stride + r - pad;
          iw = ow * d.stride + s - pad;
        }
        for (int ci = 0; ci < cin_tot; ++ci) {
          const float wv = d.transposed ? hw[(((size_t)ci * d.c_out + co) * k + r) * k + s]
                                        : hw[(((size_t)co * cin_tot + ci) * k + r) * k + s];
          acc += (double)in_at(n, ih, iw, ci) * wv;
        }
      }
    return acc;
  };

  // sampled verification (+ a NaN scan over the whole output = coverage check)
  size_t nan_count = 0;
  for (size_t i = 0; i < nout; ++i) {
    float v = obytes == 4 ? ((float*)hout.data())[i] : h2f(((__half*)hout.data())[i]);
    if (v != v) ++nan_count;
  }
  std::uniform_int_distribution<size_t> pick(0, nout - 1);
  double max_err = 0, max_ref = 0, max_sq_err = 0;
  size_t worst = 0;
  const size_t ns = std::min<size_t>(cs.samples, nout);
  for (size_t t = 0; t < ns; ++t) {
    size_t i = (ns == nout) ? t : pick(rng);
    int co = (int)(i % d.c_out);
    size_t px = i / d.c_out;
    int ow = (int)(px % wo);
    px /= wo;
    int oh = (int)(px % ho);
    int n = (int)(px / ho);
    double acc = ref_at(n, oh, ow, co);
    double ref = acc + hb[co];
    if (ref < 0) ref *= d.lrelu_slope;
    float got = obytes == 4 ? ((float*)hout.data())[i] : h2f(((__half*)hout.data())[i]);
    double err = fabs(got - ref);
    if (!(err <= max_err)) {
      if (err > max_err || err != err) {
        max_err = err;
        worst = i;
      }
    }
    if (fabs(ref) > max_ref) max_ref = fabs(ref);
  }
  const double tol = (obytes == 4 ? 2e-3 : 6e-3) * (max_ref > 1 ? max_ref : 1.0);
  const bool ok = nan_count == 0 && max_err <= tol && max_sq_err <= 4e-3 * (max_ref * d.sq_scale) * (max_ref * d.sq_scale) + 1e-6;
  printf("[%-28s] %s  max_err=%.3e (max|ref|=%.2f) sq_err=%.2e unwritten/NaN=%zu worst_idx=%zu K=%lld out=%dx%dx%dx%d\n",
         cs.name.c_str(), ok ? "PASS" : "FAIL", max_err, max_ref, max_sq_err, nan_count, worst, (long long)K,
         d.batch, ho, wo, d.c_out);

  if (timing) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) stemb200_conv2d_fwd(&d, din.data(), dpk, db, nullptr, dout, 0);
    CK(cudaEventRecord(e0));
    const int iters = 10;
    for (int i = 0; i < iters; ++i) stemb200_conv2d_fwd(&d, din.data(), dpk, db, nullptr, dout, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    int taps = 0;
    for (int i = 0; i < k * k; ++i) taps += (mask >> i) & 1;
    double flops = d.transposed ? 2.0 * d.batch * d.h_in * d.w_in * (double)cin_tot * d.c_out * taps
                                : 2.0 * d.batch * ho * wo * (double)cin_tot * d.c_out * taps;
    printf("    time %.3f ms  %.1f TFLOP/s\n", ms, flops / ms * 1e-9);
  }
  for (auto p : din) cudaFree(p);
  cudaFree(dw);
  cudaFree(db);
  cudaFree(dpk);
  cudaFree(dout);
  return ok ? 0 : 1;
}

// fused conv/deconv + GDN/IGDN: reference evaluated for whole pixels (all 192 channels) at sampled positions
static int run_gdn_case(const std::string& name, stemb200_conv_desc d, int inverse, int n_pixels, bool timing) {
  const int k = d.kh, pad = k / 2, C = d.c_out, cin = d.c_in[0];
  int ho, wo;
  if (d.transposed) { ho = 2 * d.h_in; wo = 2 * d.w_in; }
  else { ho = (d.h_in + 2 * pad - k) / d.stride + 1; wo = (d.w_in + 2 * pad - k) / d.stride + 1; }
  std::mt19937 rng(4321);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  const size_t nin = (size_t)d.batch * d.h_in * d.w_in * cin;
  std::vector<__half> hin(nin);
  for (auto& v : hin) v = __float2half_rn(U(rng));
  void* din;
  if (d.row_taps) {
    // zero-bordered NHWC8 canvas [n][h + 2 pad][w + 2 pad][8] (+ slack), what stemb200_frame_to_nhwc8 writes
    const int hc = d.h_in + 2 * pad, wc = d.w_in + 2 * pad;
    std::vector<__half> canvas((size_t)d.batch * hc * wc * 8 + 64, __float2half_rn(0.f));
    for (int n = 0; n < d.batch; ++n)
      for (int ih = 0; ih < d.h_in; ++ih)
        for (int iw = 0; iw < d.w_in; ++iw)
          for (int ci = 0; ci < 8; ++ci)
            canvas[(((size_t)n * hc + ih + pad) * wc + iw + pad) * 8 + ci] =
                hin[(((size_t)n * d.h_in + ih) * d.w_in + iw) * 8 + ci];
    CK(cudaMalloc(&din, canvas.size() * 2));
    CK(cudaMemcpy(din, canvas.data(), canvas.size() * 2, cudaMemcpyHostToDevice));
  } else {
    CK(cudaMalloc(&din, nin * 2));
    CK(cudaMemcpy(din, hin.data(), nin * 2, cudaMemcpyHostToDevice));
  }
  const size_t nw = (size_t)C * cin * k * k;
  std::vector<float> hw(nw);
  const float wscale = 2.0f / sqrtf((float)cin * k * k / (d.transposed ? 4 : 1));
  for (auto& v : hw) v = h2f(__float2half_rn(U(rng) * wscale));
  float* dw; CK(cudaMalloc(&dw, nw * 4)); CK(cudaMemcpy(dw, hw.data(), nw * 4, cudaMemcpyHostToDevice));
  std::vector<float> hb(C), hbeta(C), hg((size_t)C * C);
  for (auto& v : hb) v = U(rng) * 0.5f;
  for (auto& v : hbeta) v = 1.0f + 0.5f * fabsf(U(rng));
  for (int i = 0; i < C; ++i) for (int j = 0; j < C; ++j)
    hg[(size_t)i * C + j] = h2f(__float2half_rn((i == j ? 0.1f : 0.f) + 0.02f * fabsf(U(rng))));
  float *db, *dbeta, *dg;
  CK(cudaMalloc(&db, C * 4)); CK(cudaMemcpy(db, hb.data(), C * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dbeta, C * 4)); CK(cudaMemcpy(dbeta, hbeta.data(), C * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dg, (size_t)C * C * 4)); CK(cudaMemcpy(dg, hg.data(), (size_t)C * C * 4, cudaMemcpyHostToDevice));
  const int64_t K = stemb200_conv2d_packed_k(&d);
  void *dpk, *dgpk;
  CK(cudaMalloc(&dpk, (size_t)K * C * 2));
  if (stemb200_conv2d_pack_weight(&d, dw, dpk, 0)) { printf("[%s] pack failed: %s\n", name.c_str(), stemb200_last_error()); return 1; }
  stemb200_conv_desc gd = mk(1, 8, 8, {C}, C, 1, 1, 0, 0, 0, 1.f, STEMB200_DT_F16, 0, 0);
  CK(cudaMalloc(&dgpk, (size_t)C * C * 2));
  if (stemb200_conv2d_pack_weight(&gd, dg, dgpk, 0)) { printf("[%s] gamma pack failed\n", name.c_str()); return 1; }
  const size_t nout = (size_t)d.batch * ho * wo * C;
  void* dout; CK(cudaMalloc(&dout, nout * 2)); CK(cudaMemset(dout, 0xFF, nout * 2));
  const void* ins[1] = {din};
  int rc = stemb200_conv2d_gdn_fwd(&d, ins, dpk, db, dgpk, dbeta, inverse, dout, 0);
  if (rc) { printf("[%s] conv2d_gdn_fwd failed: %s\n", name.c_str(), stemb200_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] kernel error: %s\n", name.c_str(), cudaGetErrorString(e)); return 2; }
  std::vector<__half> hout(nout);
  CK(cudaMemcpy(hout.data(), dout, nout * 2, cudaMemcpyDeviceToHost));
  size_t nan_count = 0;
  for (size_t i = 0; i < nout; ++i) { float v = h2f(hout[i]); if (v != v) ++nan_count; }
  auto in_at = [&](int n, int ih, int iw, int ci) -> float {
    if (ih < 0 || ih >= d.h_in || iw < 0 || iw >= d.w_in) return 0.f;
    return h2f(hin[(((size_t)n * d.h_in + ih) * d.w_in + iw) * cin + ci]);
  };
  std::uniform_int_distribution<size_t> pick(0, (size_t)d.batch * ho * wo - 1);
  double max_err = 0, max_ref = 0;
  std::vector<double> x(C);
  for (int t = 0; t < n_pixels; ++t) {
    size_t px = pick(rng);
    if (t == 0) px = 0;
    if (t == 1) px = (size_t)d.batch * ho * wo - 1;
    int ow = (int)(px % wo), oh = (int)((px / wo) % ho), n = (int)(px / ((size_t)wo * ho));
    for (int co = 0; co < C; ++co) {
      double acc = 0;
      for (int r = 0; r < k; ++r) for (int s2 = 0; s2 < k; ++s2) {
        int ih, iw;
        if (d.transposed) {
          if (((oh + pad - r) & 1) || ((ow + pad - s2) & 1) || oh + pad - r < 0 || ow + pad - s2 < 0) continue;
          ih = (oh + pad - r) / 2; iw = (ow + pad - s2) / 2;
        } else { ih = oh * d.stride + r - pad; iw = ow * d.stride + s2 - pad; }
        for (int ci = 0; ci < cin; ++ci) {
          const float wv = d.transposed ? hw[(((size_t)ci * C + co) * k + r) * k + s2] : hw[(((size_t)co * cin + ci) * k + r) * k + s2];
          acc += (double)in_at(n, ih, iw, ci) * wv;
        }
      }
      x[co] = h2f(__float2half_rn((float)(acc + hb[co])));
    }
    for (int co = 0; co < C; ++co) {
      double nrm = hbeta[co];
      for (int j = 0; j < C; ++j) nrm += (double)hg[(size_t)co * C + j] * x[j] * x[j];
      double ref = x[co] * (inverse ? sqrt(nrm) : 1.0 / sqrt(nrm));
      double got = h2f(hout[px * C + co]);
      double err = fabs(got - ref);
      if (!(err <= max_err)) max_err = err;
      if (fabs(ref) > max_ref) max_ref = fabs(ref);
    }
  }
  const bool ok = nan_count == 0 && max_err <= 4e-3 * (max_ref > 1 ? max_ref : 1.0);
  printf("[%-28s] %s  max_err=%.3e (max|ref|=%.2f) unwritten/NaN=%zu out=%dx%dx%dx%d\n", name.c_str(), ok ? "PASS" : "FAIL",
         max_err, max_ref, nan_count, d.batch, ho, wo, C);
  if (timing) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) stemb200_conv2d_gdn_fwd(&d, ins, dpk, db, dgpk, dbeta, inverse, dout, 0);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 10; ++i) stemb200_conv2d_gdn_fwd(&d, ins, dpk, db, dgpk, dbeta, inverse, dout, 0);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 10;
    double px = d.transposed ? (double)d.batch * d.h_in * d.w_in : (double)d.batch * ho * wo;
    double flops = 2.0 * px * cin * C * k * k + 2.0 * d.batch * ho * wo * (double)C * C;
    printf("    time %.3f ms  %.1f TFLOP/s (conv + gamma contraction)\n", ms, flops / ms * 1e-9);
  }
  cudaFree(din); cudaFree(dw); cudaFree(db); cudaFree(dbeta); cudaFree(dg); cudaFree(dpk); cudaFree(dgpk); cudaFree(dout);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  const char* filter = nullptr;
  bool timing = false;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--time")) timing = true;
    else filter = argv[i];
  }
  printf("%s\n", stemb200_version());
  const int F16 = STEMB200_DT_F16, F32 = STEMB200_DT_F32, LIN = 0;
  std::vector<Case> cases = {
      {"a_1x1_64_64_direct_f32", mk(1, 8, 16, {64}, 64, 1, 1, 0, 0, LIN, 1.f, F32, 0, 1), 1 << 30},
      {"b_1x1_64_64_tma_f32", mk(1, 8, 16, {64}, 64, 1, 1, 0, 0, LIN, 1.f, F32, 0, 0), 1 << 30},
      {"c_1x1_64_64_tma_f16", mk(1, 8, 16, {64}, 64, 1, 1, 0, 0, LIN, 1.f, F16, 0, 0), 1 << 30},
      {"d_1x1_192_192_2img", mk(2, 16, 16, {192}, 192, 1, 1, 0, 0, LIN, 0.01f, F16, 0, 0), 20000},
      {"e_3x3_64_64_pad", mk(1, 8, 16, {64}, 64, 3, 1, 0, 0, LIN, 1.f, F32, 0, 1), 1 << 30},
      {"f_5x5_192_256_ragged", mk(2, 20, 36, {192}, 256, 5, 1, 0, 0, LIN, 0.01f, F16, 0, 0), 20000},
      {"g_3x3_cat2_256", mk(1, 20, 36, {192, 192}, 256, 3, 1, 0, 0, LIN, 0.01f, F16, 0, 0), 20000},
      {"h_5x5s2_192_192_sq", mk(2, 40, 72, {192}, 192, 5, 2, 0, 0, LIN, 1.f, F16, 1, 0), 20000},
      {"i_5x5s2_odd", mk(1, 17, 30, {256}, 256, 5, 2, 0, 0, LIN, 1.f, F32, 0, 0), 20000},
      {"l_deconv5_256_256", mk(2, 9, 15, {256}, 256, 5, 2, 1, 0, LIN, 0.01f, F16, 0, 0), 20000},
      {"m_deconv5_192_192_sq", mk(1, 17, 30, {192}, 192, 5, 2, 1, 0, LIN, 1.f, F16, 1, 0), 20000},
      {"n_masked5_192_384", mk(1, 20, 36, {192}, 384, 5, 1, 0, 0xFFFu, LIN, 1.f, F16, 0, 0), 20000},
      {"o_1x1_cat3_768", mk(1, 20, 36, {384, 384, 384}, 768, 1, 1, 0, 0, LIN, 0.01f, F16, 0, 0), 20000},
      {"p_3x3_192_16_direct", mk(1, 20, 36, {192}, 16, 3, 1, 0, 0, LIN, 1.f, F32, 0, 1), 20000},
      {"q_5x5_256_320", mk(1, 20, 36, {256}, 320, 5, 1, 0, 0, LIN, 0.01f, F16, 0, 0), 20000},
      {"r_1x1_576_384_f32", mk(1, 68, 120, {576}, 384, 1, 1, 0, 0, LIN, 1.f, F32, 0, 0), 20000},
      {"s_tile4x30", mk(1, 68, 120, {192}, 192, 3, 1, 0, 0, LIN, 0.01f, F16, 0, 0, 4, 30), 20000},
      {"t_1x1_128_192_plainGEMM", mk(1, 64, 960, {128}, 192, 1, 1, 0, 0, LIN, 1.f, F16, 1, 0), 20000},
  };
  std::vector<Case> perf = {
      {"P_tpm4_5x5_320_384_b11", mk(11, 68, 120, {320}, 384, 5, 1, 0, 0, LIN, 1.f, F16, 0, 0), 4000},
      {"P_tpm0_5x5_192_256_b11", mk(11, 68, 120, {192}, 256, 5, 1, 0, 0, LIN, 0.01f, F16, 0, 0), 4000},
      {"P_tpm2_5x5_256_320_b11", mk(11, 68, 120, {256}, 320, 5, 1, 0, 0, LIN, 0.01f, F16, 0, 0), 4000},
      {"P_ga2_5x5s2_192_192_b2", mk(2, 544, 960, {192}, 192, 5, 2, 0, 0, LIN, 1.f, F16, 1, 0), 4000},
      {"P_gs4_deconv_192_192_b2", mk(2, 272, 480, {192}, 192, 5, 2, 1, 0, LIN, 1.f, F16, 1, 0), 4000},
      {"P_epm0_cat3_768_b11", mk(11, 68, 120, {384, 384, 384}, 768, 1, 1, 0, 0, LIN, 0.01f, F16, 0, 0), 4000},
      {"P_gs6_3x3_192_16_b2", mk(2, 544, 960, {192}, 16, 3, 1, 0, 0, LIN, 1.f, F32, 0, 1), 4000},
  };
  int fails = 0;
  for (auto& c : cases) {
    if (filter && c.name.find(filter) == std::string::npos) continue;
    int r = run_case(c, false);
    if (r == 2) {
      printf("fatal CUDA error, stopping\n");
      return 2;
    }
    fails += r;
  }
  struct GCase { std::string name; stemb200_conv_desc d; int inv; int px; bool perf; };
  std::vector<GCase> gcases = {
      {"G_conv5s2_gdn", mk(2, 40, 72, {192}, 192, 5, 2, 0, 0, LIN, 1.f, F16, 0, 0), 0, 150, false},
      {"G_gemm128_gdn", mk(1, 24, 40, {128}, 192, 1, 1, 0, 0, LIN, 1.f, F16, 0, 0), 0, 300, false},
      {"G_deconv5_igdn", mk(2, 17, 30, {192}, 192, 5, 2, 1, 0, LIN, 1.f, F16, 0, 0), 1, 150, false},
      {"G_rowtaps5s2_gdn", [&] { auto d = mk(2, 40, 72, {8}, 192, 5, 2, 0, 0, LIN, 1.f, F16, 0, 0); d.row_taps = 1; return d; }(), 0, 300, false},
      {"G_rowtaps5s2_ragged", [&] { auto d = mk(1, 34, 60, {8}, 192, 5, 2, 0, 0, LIN, 1.f, F16, 0, 0); d.row_taps = 1; return d; }(), 0, 300, false},
      {"G_P_ga0_rowtaps_gdn_b2", [&] { auto d = mk(2, 1088, 1920, {8}, 192, 5, 2, 0, 0, LIN, 1.f, F16, 0, 0); d.row_taps = 1; return d; }(), 0, 100, true},
      {"G_P_ga2_gdn_b2", mk(2, 544, 960, {192}, 192, 5, 2, 0, 0, LIN, 1.f, F16, 0, 0), 0, 40, true},
      {"G_P_ga0_gemm_gdn_b2", mk(2, 544, 960, {128}, 192, 1, 1, 0, 0, LIN, 1.f, F16, 0, 0), 0, 100, true},
      {"G_P_gs4_deconv_igdn_b2", mk(2, 272, 480, {192}, 192, 5, 2, 1, 0, LIN, 1.f, F16, 0, 0), 1, 40, true},
  };
  for (auto& c : gcases) {
    if (filter && c.name.find(filter) == std::string::npos) continue;
    if (c.perf && !timing) continue;
    int r = run_gdn_case(c.name, c.d, c.inv, c.px, c.perf && timing);
    if (r == 2) { printf("fatal CUDA error, stopping\n"); return 2; }
    fails += r;
  }
  if (timing)
    for (auto& c : perf) {
      if (filter && c.name.find(filter) == std::string::npos) continue;
      int r = run_case(c, true);
      if (r == 2) return 2;
      fails += r;
    }
  printf("%s (%d failing)\n", fails ? "SOME FAILED" : "ALL PASSED", fails);
  return fails ? 1 : 0;
}
