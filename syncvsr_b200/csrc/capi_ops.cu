// extern "C" operator surface for the non-GEMM kernels (declared in include/svsr.h): one entry per torch.nn call of
// the reference path they replace. Scratch buffers are caller-provided; nothing is allocated here.
#include "../../include/svsr.h"
#include "common.cuh"
#include "elementwise.cuh"
#include "encoder.cuh"
#include "heads.cuh"

using namespace svsr;
typedef __nv_bfloat16 bf16;
#define ST(s) static_cast<cudaStream_t>(s)
#define RC(expr)         \
  do {                   \
    int _rc = (expr);    \
    if (_rc) return _rc; \
  } while (0)

extern "C" {

int svsr_stem_patch(const float* videos, void* patches, int B, int T, int H, int W, void* stream) {
  return stem_patch(videos, static_cast<bf16*>(patches), B, T, H, W, ST(stream));
}

int svsr_batchnorm_fwd(const void* x, int64_t rows, int C, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float eps, float momentum, int train, const void* res, const float* res_coef,
                       int relu, void* out, float* coef, double* stats_scratch, void* stream) {
  SVSR_REQUIRE(x && out && coef && stats_scratch, "batchnorm_fwd: null pointer");
  if (train) {
    SVSR_CHECK_CUDA(cudaMemsetAsync(stats_scratch, 0, 2 * C * sizeof(double), ST(stream)));
    RC(bn_stats(static_cast<const bf16*>(x), rows, C, stats_scratch, ST(stream)));
  }
  RC(bn_finalize(stats_scratch, rows, C, gamma, beta, eps, momentum, running_mean, running_var, coef, train ? 1 : -1,
                 ST(stream)));
  return bn_apply(static_cast<const bf16*>(x), coef, static_cast<const bf16*>(res), res_coef, relu,
                  static_cast<bf16*>(out), rows, C, ST(stream));
}

int svsr_batchnorm_bwd(const void* dout, const void* relu_ref, const void* c, int64_t rows, int C, const float* coef,
                       float* dgamma, float* dbeta, void* dc, void* gmask_out, double* stats_scratch,
                       float* kcoef_scratch, void* stream) {
  SVSR_REQUIRE(dout && c && coef && dc && stats_scratch && kcoef_scratch, "batchnorm_bwd: null pointer");
  SVSR_CHECK_CUDA(cudaMemsetAsync(stats_scratch, 0, 2 * C * sizeof(double), ST(stream)));
  RC(bn_bwd_reduce(static_cast<const bf16*>(dout), static_cast<const bf16*>(relu_ref), static_cast<const bf16*>(c),
                   coef, rows, C, stats_scratch, 0, ST(stream)));
  RC(bn_bwd_finalize(stats_scratch, rows, C, dgamma, dbeta, kcoef_scratch, ST(stream)));
  return bn_bwd_apply(static_cast<const bf16*>(dout), static_cast<const bf16*>(relu_ref), static_cast<const bf16*>(c),
                      coef, kcoef_scratch, static_cast<bf16*>(dc), static_cast<bf16*>(gmask_out), rows, C, 0, ST(stream));
}

int svsr_stem_bn_gelu_pool_fwd(const void* y0, const float* coef, void* out, uint8_t* argmax, int N, int IH, int IW,
                               void* stream) {
  return stem_bn_gelu_pool(static_cast<const bf16*>(y0), coef, static_cast<bf16*>(out), argmax, N, IH, IW, ST(stream));
}
int svsr_stem_pool_gelu_bwd(const void* dout, const uint8_t* argmax, const void* y0, const float* coef, void* dz, int N,
                            int IH, int IW, void* stream) {
  return stem_pool_gelu_bwd(static_cast<const bf16*>(dout), argmax, static_cast<const bf16*>(y0), coef,
                            static_cast<bf16*>(dz), N, IH, IW, ST(stream));
}

int svsr_stem_bwd_fused(void* dout, const uint8_t* argmax, const void* y0, const float* coef, float* dgamma,
                        float* dbeta, void* dc, double* stats_scratch, float* kcoef_scratch, int N, int IH, int IW,
                        void* stream) {
  SVSR_REQUIRE(dout && argmax && y0 && coef && dgamma && dbeta && dc && stats_scratch && kcoef_scratch,
               "stem_bwd_fused: null pointer");
  return stem_bwd_fused(static_cast<bf16*>(dout), argmax, static_cast<const bf16*>(y0), coef, dgamma, dbeta,
                        static_cast<bf16*>(dc), stats_scratch, kcoef_scratch, N, IH, IW, ST(stream));
}

int svsr_meanpool_cls_fwd(const void* a, const float* cls, float* x_stream, int B, int T, int HW, int C, void* stream) {
  return meanpool_cls(static_cast<const bf16*>(a), cls, x_stream, B, T, HW, C, ST(stream));
}
int svsr_meanpool_cls_bwd(const float* dx, void* dout, float* dcls, int B, int T, int HW, int C, void* stream) {
  return meanpool_cls_bwd(dx, static_cast<bf16*>(dout), dcls, B, T, HW, C, ST(stream));
}

int svsr_rmsnorm_fwd(const float* x, const float* g, void* y, float* inv, int M, int D, float eps, void* stream) {
  return rmsnorm_fwd(x, g, static_cast<bf16*>(y), inv, M, D, eps, ST(stream));
}
int svsr_rmsnorm_bwd(const void* dy, const float* x, const float* g, const float* inv, float* dx, void* dx_bf16,
                     float* dg, int M, int D, float eps, void* stream) {
  return rmsnorm_bwd(static_cast<const bf16*>(dy), x, g, inv, dx, static_cast<bf16*>(dx_bf16), dg, M, D, eps,
                     ST(stream));
}
int svsr_rotary_table(float* tab, int n, void* stream) { return rotary_table(tab, n, ST(stream)); }
int svsr_attention_fwd(const void* qkv, const float* rot, void* o, int B, int n, int heads, int rotary_v,
                       void* stream) {
  return attention_fwd(static_cast<const bf16*>(qkv), rot, static_cast<bf16*>(o), B, n, heads, rotary_v, ST(stream));
}
int svsr_attention_bwd(const void* qkv, const float* rot, const void* d_o, void* dqkv, int B, int n, int heads,
                       int rotary_v, void* stream) {
  return attention_bwd(static_cast<const bf16*>(qkv), rot, static_cast<const bf16*>(d_o), static_cast<bf16*>(dqkv), B,
                       n, heads, rotary_v, ST(stream));
}
int svsr_geglu_fwd(const void* h, void* u, int M, int F, float p_drop, uint64_t seed, void* stream) {
  SVSR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "geglu: dropout p=%f out of [0,1)", p_drop);
  return geglu_fwd(static_cast<const bf16*>(h), static_cast<bf16*>(u), M, F, p_drop, seed, ST(stream));
}
int svsr_geglu_bwd(const void* h, const void* du, void* dh, int M, int F, float p_drop, uint64_t seed, void* stream) {
  SVSR_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "geglu: dropout p=%f out of [0,1)", p_drop);
  return geglu_bwd(static_cast<const bf16*>(h), static_cast<const bf16*>(du), static_cast<bf16*>(dh), M, F, p_drop, seed,
                   ST(stream));
}

int svsr_audio_ce(const float* logits, int ld, const int64_t* tokens, int64_t tok_stride_b, int B, int T, int A, int G,
                  int V, void* dlogits, double* acc, int* bad_token, float dscale, void* stream) {
  return audio_ce(logits, ld, reinterpret_cast<const long long*>(tokens), tok_stride_b, B, T, A, G, V,
                  static_cast<bf16*>(dlogits), acc, bad_token, dscale, ST(stream));
}
int svsr_category_ce(const float* logits, int ld, const int64_t* labels, const float* soft_labels, int B, int C,
                     float eps, void* dlogits, int ldd, double* acc, float dscale, void* stream) {
  return category_ce(logits, ld, reinterpret_cast<const long long*>(labels), soft_labels, B, C, eps,
                     static_cast<bf16*>(dlogits), ldd, acc, dscale, ST(stream));
}

}  // extern "C"
