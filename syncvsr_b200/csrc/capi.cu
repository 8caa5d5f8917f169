// extern "C" surface declared in include/svsr.h. Thin argument checking + dispatch; no torch types.
#include "../../include/svsr.h"
#include "common.cuh"
#include "igemm.cuh"
#include "wgrad.cuh"

namespace svsr {
}

using namespace svsr;

extern "C" {

int svsr_abi_version(void) { return 1; }
const char* svsr_last_error(void) { return get_last_error(); }

int svsr_gemm_bf16(const void* a, int lda, const void* b, int ldb, void* out, int ldc, const float* bias,
                   const void* resid, int M, int N, int K, int out_fp32, int resid_fp32, float alpha, void* stream) {
  SVSR_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  IgemmProblem p;
  p.a = a, p.a_N = M, p.a_H = 1, p.a_W = 1, p.a_C = lda, p.a_coff = 0, p.cin = K, p.stride = 1;
  p.ntaps = 1;
  p.o_N = M, p.OH = 1, p.OW = 1;
  p.b = b, p.b_rows = N, p.b_cols = ldb;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = ldc, p.c_off = 0;
  p.o_H = 1, p.o_W = 1;
  p.bias = bias, p.resid = resid, p.resid_fp32 = resid_fp32, p.alpha = alpha;
  return igemm_launch(p, static_cast<cudaStream_t>(stream));
}

int svsr_gemm_bf16_ex(const void* a, int lda, const void* b, int ldb, void* out, int ldc, const float* bias,
                      const void* resid, int M, int N, int K, int out_fp32, int resid_fp32, float alpha,
                      float bias_scale, int relu, const void* relu_mask, float drop_p, uint64_t drop_seed,
                      void* stream) {
  SVSR_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  IgemmProblem p;
  p.a = a, p.a_N = M, p.a_C = lda, p.cin = K, p.ntaps = 1;
  p.o_N = M;
  p.b = b, p.b_rows = N, p.b_cols = ldb;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = ldc;
  p.bias = bias, p.resid = resid, p.resid_fp32 = resid_fp32;
  p.alpha = alpha, p.bias_scale = bias_scale, p.relu = relu, p.relu_mask = relu_mask;
  p.drop_p = drop_p, p.drop_seed = (unsigned long long)drop_seed;
  return igemm_launch(p, static_cast<cudaStream_t>(stream));
}

int svsr_conv2d_fprop(const void* x, const void* w, void* y, const void* resid, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int stride, int pad, int out_fp32, void* stream) {
  SVSR_REQUIRE(R * S <= IGEMM_MAX_TAPS, "conv: %dx%d filter has too many taps", R, S);
  IgemmProblem p;
  p.a = x, p.a_N = N, p.a_H = H, p.a_W = W, p.a_C = Cin, p.a_coff = 0, p.cin = Cin, p.stride = stride;
  p.ntaps = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      int t = r * S + s;
      p.tap_dh[t] = r - pad, p.tap_dw[t] = s - pad, p.tap_kbase[t] = t * Cin;
    }
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  p.o_N = N, p.OH = OH, p.OW = OW;
  p.b = w, p.b_rows = Cout, p.b_cols = R * S * Cin;
  p.out = y, p.out_fp32 = out_fp32, p.ldc = Cout, p.c_off = 0;
  p.o_H = OH, p.o_W = OW;
  p.resid = resid;
  return igemm_launch(p, static_cast<cudaStream_t>(stream));
}

int svsr_conv2d_fprop_bnstats(const void* x, const void* w, void* y, double* bn_stats, int N, int H, int W, int Cin,
                              int Cout, int R, int S, int stride, int pad, void* stream) {
  SVSR_REQUIRE(R * S <= IGEMM_MAX_TAPS && bn_stats, "conv_bnstats: bad arguments");
  IgemmProblem p;
  p.a = x, p.a_N = N, p.a_H = H, p.a_W = W, p.a_C = Cin, p.a_coff = 0, p.cin = Cin, p.stride = stride;
  p.ntaps = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      int t = r * S + s;
      p.tap_dh[t] = r - pad, p.tap_dw[t] = s - pad, p.tap_kbase[t] = t * Cin;
    }
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  p.o_N = N, p.OH = OH, p.OW = OW;
  p.b = w, p.b_rows = Cout, p.b_cols = R * S * Cin;
  p.out = y, p.out_fp32 = 0, p.ldc = Cout, p.c_off = 0;
  p.o_H = OH, p.o_W = OW;
  p.bn_stats = bn_stats;
  return igemm_launch(p, static_cast<cudaStream_t>(stream));
}

int svsr_conv_taps_fprop(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int ntaps,
                         const int* tap_dh, const int* tap_dw, int out_fp32, void* stream) {
  SVSR_REQUIRE(ntaps >= 1 && ntaps <= IGEMM_MAX_TAPS && tap_dh && tap_dw, "conv_taps: bad tap list");
  IgemmProblem p;
  p.a = x, p.a_N = N, p.a_H = H, p.a_W = W, p.a_C = Cin, p.cin = Cin, p.stride = 1;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) p.tap_dh[t] = tap_dh[t], p.tap_dw[t] = tap_dw[t], p.tap_kbase[t] = t * Cin;
  p.o_N = N, p.OH = H, p.OW = W;
  p.b = w, p.b_rows = Cout, p.b_cols = ntaps * Cin;
  p.out = y, p.out_fp32 = out_fp32, p.ldc = Cout, p.o_H = H, p.o_W = W;
  return igemm_launch(p, static_cast<cudaStream_t>(stream));
}

int svsr_conv_taps_fprop_bnstats(const void* x, const void* w, void* y, double* bn_stats, int N, int H, int W, int Cin,
                                 int Cout, int ntaps, const int* tap_dh, const int* tap_dw, void* stream) {
  SVSR_REQUIRE(ntaps >= 1 && ntaps <= IGEMM_MAX_TAPS && tap_dh && tap_dw && bn_stats, "conv_taps_bnstats: bad arguments");
  IgemmProblem p;
  p.a = x, p.a_N = N, p.a_H = H, p.a_W = W, p.a_C = Cin, p.cin = Cin, p.stride = 1;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) p.tap_dh[t] = tap_dh[t], p.tap_dw[t] = tap_dw[t], p.tap_kbase[t] = t * Cin;
  p.o_N = N, p.OH = H, p.OW = W;
  p.b = w, p.b_rows = Cout, p.b_cols = ntaps * Cin;
  p.out = y, p.out_fp32 = 0, p.ldc = Cout, p.o_H = H, p.o_W = W;
  p.bn_stats = bn_stats;
  return igemm_launch(p, static_cast<cudaStream_t>(stream));
}

// Shared by svsr_conv2d_dgrad and its fused variant below.
static int conv2d_dgrad_impl(const void* dy, const void* wd, void* dx, const void* resid, int N, int H, int W, int Cin,
                             int Cout, int R, int S, int stride, int pad, int out_fp32, const void* relu_mask,
                             const IgemmBnBwd* bnb, cudaStream_t stream) {
  SVSR_REQUIRE(R * S <= IGEMM_MAX_TAPS, "dgrad: %dx%d filter has too many taps", R, S);
  SVSR_REQUIRE(stride == 1 || stride == 2, "dgrad: stride must be 1 or 2");
  const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
  for (int a = 0; a < stride; ++a)
    for (int b = 0; b < stride; ++b) {
      IgemmProblem p;
      p.a = dy, p.a_N = N, p.a_H = OH, p.a_W = OW, p.a_C = Cout, p.a_coff = 0, p.cin = Cout, p.stride = 1;
      p.ntaps = 0;
      for (int r = 0; r < R; ++r) {
        if ((a + pad - r) % stride != 0) continue;
        for (int s = 0; s < S; ++s) {
          if ((b + pad - s) % stride != 0) continue;
          const int t = p.ntaps++;
          // dx[h,w] += dy[(h+pad-r)/stride, (w+pad-s)/stride] . W[:, :, r, s]  with h = stride*i + a
          p.tap_dh[t] = (a + pad - r) / stride, p.tap_dw[t] = (b + pad - s) / stride;
          p.tap_kbase[t] = (r * S + s) * Cout;
        }
      }
      SVSR_REQUIRE(p.ntaps > 0 || !bnb || !bnb->n, "dgrad: a pixel class without taps cannot carry fused statistics");
      if (p.ntaps == 0) continue;
      p.o_N = N, p.OH = (H - a + stride - 1) / stride, p.OW = (W - b + stride - 1) / stride;
      if (p.OH <= 0 || p.OW <= 0) continue;
      p.b = wd, p.b_rows = Cin, p.b_cols = R * S * Cout;
      p.out = dx, p.out_fp32 = out_fp32, p.ldc = Cin, p.c_off = 0;
      p.o_H = H, p.o_W = W, p.o_sh = stride, p.o_sw = stride, p.o_oh = a, p.o_ow = b;
      p.resid = resid, p.resid_fp32 = out_fp32;
      p.relu_mask = relu_mask;
      if (bnb) p.bnb = *bnb;
      int rc = igemm_launch(p, stream);
      if (rc) return rc;
    }
  return SVSR_OK;
}

int svsr_conv2d_dgrad(const void* dy, const void* wd, void* dx, const void* resid, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int stride, int pad, int out_fp32, void* stream) {
  return conv2d_dgrad_impl(dy, wd, dx, resid, N, H, W, Cin, Cout, R, S, stride, pad, out_fp32, nullptr, nullptr,
                           static_cast<cudaStream_t>(stream));
}

// the same with the ReLU mask of the tensor the gradient flows INTO applied in the epilogue: dx = (W^T dy + resid) * [mask > 0]
int svsr_conv2d_dgrad_masked(const void* dy, const void* wd, void* dx, const void* resid, int N, int H, int W, int Cin,
                             int Cout, int R, int S, int stride, int pad, const void* relu_mask, void* stream) {
  return conv2d_dgrad_impl(dy, wd, dx, resid, N, H, W, Cin, Cout, R, S, stride, pad, 0, relu_mask, nullptr,
                           static_cast<cudaStream_t>(stream));
}

int svsr_conv2d_dgrad_bnbwd(const void* dy, const void* wd, void* dx, const void* resid, int N, int H, int W, int Cin,
                            int Cout, int R, int S, int stride, int pad, const void* relu_mask, int self_mask,
                            const void* c0, const float* coef0, double* stats0, const void* c1, const float* coef1,
                            double* stats1, void* stream) {
  SVSR_REQUIRE(c0 && coef0 && stats0, "dgrad_bnbwd: the first BatchNorm's buffers are required");
  SVSR_REQUIRE(!(self_mask && relu_mask), "dgrad_bnbwd: self_mask and relu_mask are exclusive");
  IgemmBnBwd b;
  b.n = c1 ? 2 : 1;
  b.c[0] = c0, b.coef[0] = coef0, b.stats[0] = stats0;
  b.c[1] = c1, b.coef[1] = coef1, b.stats[1] = stats1;
  b.self_mask = self_mask;
  SVSR_REQUIRE(!c1 || (coef1 && stats1), "dgrad_bnbwd: the second BatchNorm's buffers are incomplete");
  return conv2d_dgrad_impl(dy, wd, dx, resid, N, H, W, Cin, Cout, R, S, stride, pad, 0, relu_mask, &b,
                           static_cast<cudaStream_t>(stream));
}

int svsr_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int R, int S,
                      int stride, int pad, void* stream) {
  SVSR_REQUIRE(R * S <= WGRAD_MAX_TAPS, "wgrad: %dx%d filter has too many taps", R, S);
  WgradProblem p;
  p.a = x, p.a_N = N, p.a_H = H, p.a_W = W, p.a_C = Cin, p.a_coff = 0, p.a_cin = Cin, p.a_stride = stride;
  p.ntaps = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) p.tap_dh[r * S + s] = r - pad, p.tap_dw[r * S + s] = s - pad;
  p.b = dy, p.b_C = Cout, p.b_coff = 0, p.n_cols = Cout;
  p.k_N = N, p.k_H = (H + 2 * pad - R) / stride + 1, p.k_W = (W + 2 * pad - S) / stride + 1;
  p.out = dw, p.ldo = Cout;
  return wgrad_launch(p, static_cast<cudaStream_t>(stream));
}

int svsr_conv_taps_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int ntaps,
                         const int* tap_dh, const int* tap_dw, void* stream) {
  SVSR_REQUIRE(ntaps >= 1 && ntaps <= WGRAD_MAX_TAPS && tap_dh && tap_dw, "conv_taps_wgrad: bad tap list");
  WgradProblem p;
  p.a = x, p.a_N = N, p.a_H = H, p.a_W = W, p.a_C = Cin, p.a_coff = 0, p.a_cin = Cin, p.a_stride = 1;
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) p.tap_dh[t] = tap_dh[t], p.tap_dw[t] = tap_dw[t];
  p.b = dy, p.b_C = Cout, p.b_coff = 0, p.n_cols = Cout;
  p.k_N = N, p.k_H = H, p.k_W = W;
  p.out = dw, p.ldo = Cout;
  return wgrad_launch(p, static_cast<cudaStream_t>(stream));
}

int svsr_gemm_wgrad(const void* dy, int ldy, const void* x, int ldx, float* dw, int ldw, int M, int N, int K,
                    void* stream) {
  SVSR_REQUIRE(N % 64 == 0, "gemm_wgrad: N=%d must be a multiple of 64", N);
  WgradProblem p;
  p.a = dy, p.a_N = M, p.a_H = 1, p.a_W = 1, p.a_C = ldy, p.a_coff = 0, p.a_cin = N, p.a_stride = 1;
  p.ntaps = 1;
  p.b = x, p.b_C = ldx, p.b_coff = 0, p.n_cols = K;
  p.k_N = M, p.k_H = 1, p.k_W = 1;
  p.out = dw, p.ldo = ldw;
  return wgrad_launch(p, static_cast<cudaStream_t>(stream));
}


}  // extern "C"
