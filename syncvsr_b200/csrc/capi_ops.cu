// extern "C" operator surface for the non-GEMM kernels (declared in include/svsr.h): one entry per torch.nn call of
// the reference path they replace. Scratch buffers are caller-provided; nothing is allocated here.
#include "../../include/svsr.h"
#include "common.cuh"
#include "elementwise.cuh"
#include "encoder.cuh"
#include "heads.cuh"
#include "igemm.cuh"
#include "conformer.cuh"
#include "stem_direct.cuh"

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

int svsr_stem_conv_direct(const void* video_bf16, const void* w_packed, void* y0, double* bn_stats, int B, int T, int H,
                          int W, void* stream) {
  SVSR_REQUIRE(video_bf16 && w_packed && y0, "stem_conv_direct: null pointer");
  return stem_direct_fwd(video_bf16, static_cast<const bf16*>(w_packed), static_cast<bf16*>(y0), bn_stats, B, T, H, W, 0.0,
                         ST(stream));
}
int svsr_stem_wgrad_direct(const void* video_bf16, const void* dz, float* out, int ldo, int B, int T, int H, int W,
                           void* stream) {
  SVSR_REQUIRE(video_bf16 && dz && out, "stem_wgrad_direct: null pointer");
  return stem_direct_wgrad(video_bf16, static_cast<const bf16*>(dz), out, ldo, B, T, H, W, 0.0, ST(stream));
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
int svsr_attention_qkv_fwd(const void* xn, int ldx, const void* w, int Kp, const float* rot, void* qkv, void* o, int B,
                           int n, int heads, int rotary_v, void* stream) {
  return attention_qkv_tc_fwd(static_cast<const bf16*>(xn), ldx, static_cast<const bf16*>(w), Kp, rot,
                              static_cast<bf16*>(qkv), static_cast<bf16*>(o), B, n, heads, rotary_v, ST(stream));
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
// ---- fused audio head (igemm.cuh IgemmCe; heads.cuh ce_finalize) ----
static int audio_head_problem(IgemmProblem& p, const void* x, int ldx, const void* w, int ldw, int K, const float* bias,
                              const int64_t* tokens, int64_t tok_stride_b, int B, int T, int A, int G, int V,
                              int* bad_token) {
  SVSR_REQUIRE(x && w && tokens && bad_token, "audio_head: null pointer");
  SVSR_REQUIRE(B > 0 && T > 0 && A > 0 && G > 0 && V > 0 && V % 64 == 0 && K > 0 && K % 64 == 0,
               "audio_head: B=%d T=%d A=%d G=%d V=%d (multiple of 64) K=%d (multiple of 64)", B, T, A, G, V, K);
  SVSR_REQUIRE(tok_stride_b >= (int64_t)T * A * G, "audio_head: audio_tokens has fewer than T*alignment rows per clip");
  p.a = x, p.a_N = B * T, p.a_C = ldx, p.cin = K, p.ntaps = 1;
  p.o_N = B * T;
  p.b = w, p.b_rows = A * G * V, p.b_cols = ldw;
  p.bias = bias;
  p.ldc = A * G * V;
  p.ce.T = T, p.ce.A = A, p.ce.G = G, p.ce.V = V, p.ce.AG = A * G;
  p.ce.tokens = reinterpret_cast<const long long*>(tokens), p.ce.tok_stride_b = tok_stride_b;
  p.ce.bad_token = bad_token;
  return SVSR_OK;
}
int svsr_audio_head_fwd(const void* x, int ldx, const void* w, int ldw, int K, const float* bias, const int64_t* tokens,
                        int64_t tok_stride_b, int B, int T, int A, int G, int V, void* part, float* xt, float* lse,
                        double* loss_sum, int* bad_token, void* stream) {
  SVSR_REQUIRE(part && xt && lse && loss_sum, "audio_head_fwd: null buffer");
  IgemmProblem p;
  int rc = audio_head_problem(p, x, ldx, w, ldw, K, bias, tokens, tok_stride_b, B, T, A, G, V, bad_token);
  if (rc) return rc;
  p.ce.mode = 1, p.ce.part = static_cast<float2*>(part), p.ce.xt = xt;
  rc = igemm_launch(p, ST(stream));
  if (rc) return rc;
  return ce_finalize(static_cast<const float2*>(part), xt, p.ce.tokens, tok_stride_b, B, T, A, G, V, lse, loss_sum,
                     ST(stream));
}
int svsr_audio_head_bwd(const void* x, int ldx, const void* w, int ldw, int K, const float* bias, const int64_t* tokens,
                        int64_t tok_stride_b, int B, int T, int A, int G, int V, const float* lse, float dscale,
                        const float* grad_scale, void* dlogits, int* bad_token, void* stream) {
  SVSR_REQUIRE(lse && dlogits, "audio_head_bwd: null buffer");
  IgemmProblem p;
  int rc = audio_head_problem(p, x, ldx, w, ldw, K, bias, tokens, tok_stride_b, B, T, A, G, V, bad_token);
  if (rc) return rc;
  p.ce.mode = 2, p.ce.lse = lse, p.ce.dscale = dscale, p.ce.grad_scale = grad_scale;
  p.out = dlogits;
  return igemm_launch(p, ST(stream));
}

int svsr_category_ce(const float* logits, int ld, const int64_t* labels, const float* soft_labels, int B, int C,
                     float eps, void* dlogits, int ldd, double* acc, float dscale, void* stream) {
  return category_ce(logits, ld, reinterpret_cast<const long long*>(labels), soft_labels, B, C, eps,
                     static_cast<bf16*>(dlogits), ldd, acc, dscale, ST(stream));
}

int svsr_cutmix_gather(const float* videos_in, float* videos_out, const int* vsrc, int B, int T, int64_t frame_elems,
                       const int64_t* audio_in, int64_t* audio_out, const int* asrc, int Ta, int G, const int64_t* labels,
                       const int* tgt, const float* rate, const uint8_t* mixed, float* soft_labels, int num_labels,
                       const float* wm_in, float* wm_out, int Tw, void* stream) {
  SVSR_REQUIRE(videos_in && videos_out && vsrc && audio_in && audio_out && asrc && labels && tgt && rate && mixed &&
                   soft_labels && wm_in && wm_out,
               "cutmix_gather: null pointer");
  return cutmix_gather(videos_in, videos_out, vsrc, B, T, frame_elems, reinterpret_cast<const long long*>(audio_in),
                       reinterpret_cast<long long*>(audio_out), asrc, Ta, G, reinterpret_cast<const long long*>(labels),
                       tgt, rate, mixed, soft_labels, num_labels, wm_in, wm_out, Tw, ST(stream));
}

// ---- LRS sentence-level operators (csrc/conformer.cu) ----
int svsr_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* stats,
                       int M, int D, float eps, void* stream) {
  return layernorm_fwd(x, gamma, beta, static_cast<bf16*>(y_bf16), y_f32, stats, M, D, eps, ST(stream));
}
int svsr_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma, const float* stats,
                       float* dx, int accumulate, float* dgamma, float* dbeta, int M, int D, void* stream) {
  return layernorm_bwd(static_cast<const bf16*>(dy_bf16), dy_f32, x, gamma, stats, dx, accumulate, dgamma, dbeta, M, D,
                       ST(stream));
}
int svsr_glu_fwd(const void* h, void* u, int64_t M, int C, void* stream) {
  return glu_fwd(static_cast<const bf16*>(h), static_cast<bf16*>(u), M, C, ST(stream));
}
int svsr_glu_bwd(const void* h, const void* du, void* dh, int64_t M, int C, void* stream) {
  return glu_bwd(static_cast<const bf16*>(h), static_cast<const bf16*>(du), static_cast<bf16*>(dh), M, C, ST(stream));
}
int svsr_dwconv1d_fwd(const void* x, const float* w, const float* bias, void* y, int B, int T, int C, int K, int flip,
                      void* stream) {
  return dwconv1d_fwd(static_cast<const bf16*>(x), w, bias, static_cast<bf16*>(y), B, T, C, K, flip, ST(stream));
}
int svsr_dwconv1d_wgrad(const void* x, const void* dy, float* dw, float* dbias, int B, int T, int C, int K,
                        void* stream) {
  return dwconv1d_wgrad(static_cast<const bf16*>(x), static_cast<const bf16*>(dy), dw, dbias, B, T, C, K, ST(stream));
}
int svsr_bn_col_reduce(const void* x, const void* dout, const float* coef, int64_t rows, int C, double* stats, int mode,
                       void* stream) {
  return bn_col_reduce(static_cast<const bf16*>(x), static_cast<const bf16*>(dout), coef, rows, C, stats, mode,
                       ST(stream));
}
static void fill_attn(AttnProblem& a, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                      const void* p, int ldp, const float* bias_u, const float* bias_v, const int* klen, int causal,
                      int B, int H, int Tq, int Tk, float scale, const void* o, int ldo, const float* lse) {
  a.q = static_cast<const bf16*>(q), a.k = static_cast<const bf16*>(k), a.v = static_cast<const bf16*>(v);
  a.ldq = ldq, a.ldk = ldk, a.ldv = ldv;
  a.p = static_cast<const bf16*>(p), a.ldp = ldp;
  a.bias_u = bias_u, a.bias_v = bias_v, a.klen = klen, a.causal = causal;
  a.B = B, a.H = H, a.Tq = Tq, a.Tk = Tk, a.scale = scale;
  a.o = const_cast<bf16*>(static_cast<const bf16*>(o)), a.ldo = ldo, a.lse = const_cast<float*>(lse);
}
int svsr_attention_core_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* p,
                            int ldp, const float* bias_u, const float* bias_v, const int* klen, int causal, int B,
                            int H, int Tq, int Tk, float scale, void* o, int ldo, float* lse, float drop_p,
                            uint64_t drop_seed, void* stream) {
  AttnProblem a;
  fill_attn(a, q, ldq, k, ldk, v, ldv, p, ldp, bias_u, bias_v, klen, causal, B, H, Tq, Tk, scale, o, ldo, lse);
  a.drop_p = drop_p, a.drop_seed = (unsigned long long)drop_seed;
  return attention_core_fwd(a, ST(stream));
}
int64_t svsr_attention_scratch_bytes(int B, int H, int Tq, int Tk) {
  return (int64_t)attention_scratch_bytes(B, H, Tq, Tk);
}
int svsr_attention_core_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* p,
                            int ldp, const float* bias_u, const float* bias_v, const int* klen, int causal, int B,
                            int H, int Tq, int Tk, float scale, const void* o, int ldo, const float* lse,
                            const void* d_o, void* dq, void* dk, void* dv, float* dp, float* dbias_u, float* dbias_v,
                            void* scratch, float drop_p, uint64_t drop_seed, void* stream) {
  AttnProblem a;
  fill_attn(a, q, ldq, k, ldk, v, ldv, p, ldp, bias_u, bias_v, klen, causal, B, H, Tq, Tk, scale, o, ldo, lse);
  a.drop_p = drop_p, a.drop_seed = (unsigned long long)drop_seed;
  AttnGrads g;
  g.d_o = static_cast<const bf16*>(d_o);
  g.dq = static_cast<bf16*>(dq), g.dk = static_cast<bf16*>(dk), g.dv = static_cast<bf16*>(dv);
  g.lddq = ldq, g.lddk = ldk, g.lddv = ldv;
  g.dp = dp, g.dbias_u = dbias_u, g.dbias_v = dbias_v;
  g.scratch = static_cast<float*>(scratch);
  return attention_core_bwd(a, g, ST(stream));
}
int svsr_dropout_mask(uint8_t* out, int64_t n, float p, uint64_t seed, void* stream) {
  return dropout_mask_u8(out, n, p, (unsigned long long)seed, ST(stream));
}
int svsr_dropout_bf16(const void* x, void* y, int64_t n, float p, uint64_t seed, void* stream) {
  return dropout_bf16(static_cast<const bf16*>(x), static_cast<bf16*>(y), n, p, (unsigned long long)seed, ST(stream));
}
int64_t svsr_ctc_scratch_bytes(int B, int T, int Lmax) { return (int64_t)ctc_scratch_bytes(B, T, Lmax); }
int svsr_ctc_loss(const float* logits, int ld, int V, const int64_t* labels, int Lmax, const int* in_len, int B, int T,
                  void* dlogits, double* acc, int slot, float dscale, void* scratch, void* stream) {
  return ctc_loss_fwd_bwd(logits, ld, V, reinterpret_cast<const long long*>(labels), Lmax, in_len, B, T,
                          static_cast<bf16*>(dlogits), acc, slot, dscale, static_cast<float*>(scratch), ST(stream));
}
int svsr_label_smoothing_loss(const float* logits, int ld, int V, const int64_t* target, int rows, float smoothing,
                              void* dlogits, double* acc, int slot, float dscale, void* stream) {
  return label_smoothing_loss(logits, ld, V, reinterpret_cast<const long long*>(target), rows, smoothing,
                              static_cast<bf16*>(dlogits), acc, slot, dscale, ST(stream));
}

}  // extern "C"
