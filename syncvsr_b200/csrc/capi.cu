// extern "C" surface declared in include/svsr.h. Thin argument checking + dispatch; no torch types.
#include "../../include/svsr.h"
#include "common.cuh"
#include "igemm.cuh"

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

}  // extern "C"
