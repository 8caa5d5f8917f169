/* syncvsr_b200 C ABI -- the drop-in boundary of the B200-native SyncVSR hot path.
 *
 * The reference (KAIST-AILab/SyncVSR) is pure Python: its "FFI" for this path is the set of torch.nn
 * calls made by LRW/video/src/lightning.py:49-55,82,107-119,133-191 (stem3d, resnet.layer1-4, encoder,
 * audio_projection, category_classifier, the two cross-entropies) and, for the sentence-level model, by
 * LRS/video/espnet/nets/pytorch_backend/e2e_asr_transformer.py:186-227 and the modules it calls. Each entry point below replaces one
 * (or a fused group) of those library calls; the file:line of the call it replaces is cited per symbol.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (16-byte aligned, contiguous in the stated
 *    layout); the library never allocates or frees device memory;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *  - return value: 0 on success, negative svsr status otherwise; svsr_last_error() returns a
 *    thread-local description. Nothing throws across this boundary;
 *  - activations are NHWC ("channels last"): [images, rows, cols, channels], bf16 unless stated;
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef SVSR_H_
#define SVSR_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVSR_OK 0
#define SVSR_ERR_INVALID (-1)
#define SVSR_ERR_CUDA (-2)
#define SVSR_ERR_UNSUPPORTED (-3)

/* ABI version of this header; bumps on any signature change. */
int svsr_abi_version(void);
const char* svsr_last_error(void);

/* out[M,N] = alpha * a[M,K] . b[N,K]^T (+ bias[N]) (+ resid[M,N]); a,b bf16 (K contiguous, pitches lda/ldb),
 * out/resid bf16 or fp32 with pitch ldc. K % 64 == 0. Replaces nn.Linear forward / input-gradient
 * (lightning.py:82,107,161,168 and every Linear inside the encoder, lightning.py:158). */
int svsr_gemm_bf16(const void* a, int lda, const void* b, int ldb, void* out, int ldc, const float* bias,
                   const void* resid, int M, int N, int K, int out_fp32, int resid_fp32, float alpha, void* stream);

/* svsr_gemm_bf16 with the full epilogue: out = act(alpha * a.b^T + bias_scale * bias + resid) * [relu_mask > 0];
 * relu != 0 applies ReLU (Conformer/decoder FFN w_1, transformer/positionwise_feed_forward.py:28-30); relu_mask (bf16,
 * same geometry as out, or NULL) zeroes the result where mask <= 0 (the ReLU backward fused into the input-gradient
 * GEMM of w_2). alpha/bias_scale carry the macaron 1/2 (encoder_layer.py:90-96) and the x*sqrt(adim) of the
 * positional encoding (embedding.py:212). drop_p > 0: Dropout on the branch value (before the residual is added,
 * encoder_layer.py:94-137), mask = keep(drop_seed, element index in `out`). */
int svsr_gemm_bf16_ex(const void* a, int lda, const void* b, int ldb, void* out, int ldc, const float* bias,
                      const void* resid, int M, int N, int K, int out_fp32, int resid_fp32, float alpha,
                      float bias_scale, int relu, const void* relu_mask, float drop_p, uint64_t drop_seed,
                      void* stream);

/* y[N,OH,OW,Cout] = conv2d(x[N,H,W,Cin], w) with w packed as [Cout, R, S, Cin] bf16, zero padding `pad`,
 * stride 1 or 2, no bias (+ resid). Cin % 64 == 0. Replaces the Conv2d calls inside resnet.layer1-4
 * (lightning.py:114-117; timm/torchvision BasicBlock). */
int svsr_conv2d_fprop(const void* x, const void* w, void* y, const void* resid, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int stride, int pad, int out_fp32, void* stream);

/* svsr_conv2d_fprop that also accumulates (+=) the train-mode BatchNorm statistics of its output in the epilogue:
 * bn_stats fp64 [2][Cout] = per-channel sum and sum of squares of the fp32 accumulators over all output pixels
 * (conv -> BatchNorm pairs of lightning.py:50-51 and of every BasicBlock). Cout <= 512. */
int svsr_conv2d_fprop_bnstats(const void* x, const void* w, void* y, double* bn_stats, int N, int H, int W, int Cin,
                              int Cout, int R, int S, int stride, int pad, void* stream);

/* Same-size correlation with an explicit tap list: y[n,h,w,:] = sum_t x[n, h+dh[t], w+dw[t], :] . w[:, t*Cin:(t+1)*Cin]^T
 * (zero outside the image). The temporal half of the stem's Conv3d runs through this with taps (kt-2, 0) over the
 * [T, OH*OW] patch image (lightning.py:50). */
int svsr_conv_taps_fprop(const void* x, const void* w, void* y, int N, int H, int W, int Cin, int Cout, int ntaps,
                         const int* tap_dh, const int* tap_dw, int out_fp32, void* stream);
/* The same with bf16 output and the fused per-channel statistics of svsr_conv2d_fprop_bnstats (bn_stats fp64
 * [2][Cout], +=): exactly the call the stem makes for Conv3d -> BatchNorm3d (lightning.py:50-51). 5 taps (dh in
 * [-2, 2], dw = 0) over 64 -> 64 columns with W % 16 == 0 run on the temporal-halo kernel (csrc/igemm_stem.cu). */
int svsr_conv_taps_fprop_bnstats(const void* x, const void* w, void* y, double* bn_stats, int N, int H, int W, int Cin,
                                 int Cout, int ntaps, const int* tap_dh, const int* tap_dw, void* stream);

/* dx[N,H,W,Cin] = conv2d input-gradient of dy[N,OH,OW,Cout]; wd is the weight packed for dgrad as
 * [Cin, R*S*Cout] bf16 (column (r*S+s)*Cout+co holds W[co][ci][r][s]). If resid != NULL it is added (it may alias
 * dx: gradient accumulation of the residual branch). For stride 2 the four output-parity classes are issued as
 * four launches; pixels no tap reaches (1x1 stride-2) are left untouched, so dx must be pre-zeroed in that case.
 * Replaces autograd's conv backward-data for resnet.layer1-4 (lightning.py:114-117). */
int svsr_conv2d_dgrad(const void* dy, const void* wd, void* dx, const void* resid, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int stride, int pad, int out_fp32, void* stream);
/* The same input-gradient GEMM with the ReLU mask of the tensor the gradient flows into applied in the epilogue (timm
 * BasicBlock `out = relu(bn2(c2) + shortcut)` via lightning.py:114-117): dx = (W^T dy + resid) * [relu_mask > 0], relu_mask a
 * bf16 tensor of dx's geometry (the block's output). The BatchNorm-backward passes downstream then read one tensor less. */
int svsr_conv2d_dgrad_masked(const void* dy, const void* wd, void* dx, const void* resid, int N, int H, int W, int Cin,
                             int Cout, int R, int S, int stride, int pad, const void* relu_mask, void* stream);
/* The same input-gradient GEMM with the BatchNorm-backward reduction of its CONSUMER fused into the epilogue (timm
 * BasicBlock via lightning.py:114-117): dx = (W^T dy + resid) * mask, mask = [relu_mask > 0] (bf16 tensor of dx's geometry,
 * the ReLU after `out = relu(bn2(c2) + shortcut)`) or -- self_mask -- [c0*scale0 + shift0 > 0] (the ReLU directly after
 * bn1). For each consuming BatchNorm i (one or two: bn2 and downsample.1 share the gradient) with input c_i (bf16, dx's
 * geometry) and coef_i (fp32 [4][Cin]: mean, invstd, scale, shift): stats_i[0..Cin) += sum dx, stats_i[Cin..2Cin) +=
 * sum dx * xhat_i -- exactly what svsr_batchnorm_bwd's reduce pass would compute from (dx, c_i). Cin % 64 == 0. */
int svsr_conv2d_dgrad_bnbwd(const void* dy, const void* wd, void* dx, const void* resid, int N, int H, int W, int Cin,
                            int Cout, int R, int S, int stride, int pad, const void* relu_mask, int self_mask,
                            const void* c0, const float* coef0, double* stats0, const void* c1, const float* coef1,
                            double* stats1, void* stream);

/* dw[R*S*Cin, Cout] (fp32, += accumulate) = weight gradient of conv2d: row (r*S+s)*Cin+ci, column co.
 * x[N,H,W,Cin], dy[N,OH,OW,Cout] bf16. Replaces autograd's conv backward-weight. */
int svsr_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int R, int S,
                      int stride, int pad, void* stream);

/* Weight gradient of svsr_conv_taps_fprop: dw[ntaps*Cin, Cout] (fp32, += accumulate), row t*Cin+ci, column co
 * = sum over pixels of x[n, h+dh[t], w+dw[t], ci] * dy[n, h, w, co]. The stem's Conv3d backward-weight
 * (lightning.py:50) runs through this with taps (kt-2, 0) over the [T, OH*OW] patch image; that shape (5 taps, 64 -> 64
 * columns, W % 16 == 0) is served by the temporal-halo kernel (csrc/wgrad_stem.cu). */
int svsr_conv_taps_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int ntaps,
                         const int* tap_dh, const int* tap_dw, void* stream);

/* dw[N,K] (fp32, pitch ldw, += accumulate) = dy[M,N]^T . x[M,K]; dy, x bf16 with pitches ldy, ldx.
 * Replaces autograd's Linear backward-weight. N % 64 == 0. */
int svsr_gemm_wgrad(const void* dy, int ldy, const void* x, int ldx, float* dw, int ldw, int M, int N, int K,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Non-GEMM operators. bf16 NHWC activations ([rows, C] with rows = N*H*W), fp32 parameters/statistics.
 * --------------------------------------------------------------------------------------------------------- */
/* videos fp32 [B,1,T,H,W] -> 7x7/stride-2/pad-3 patches bf16 [B,T,OH*OW,64] (slot kh*8+kw); the 5-tap temporal
 * part of Conv3d(1,64,(5,7,7),(1,2,2),(2,3,3)) (lightning.py:50) then runs as an implicit GEMM over these. */
int svsr_stem_patch(const float* videos, void* patches, int B, int T, int H, int W, void* stream);
/* The same Conv3d WITHOUT the patch tensor (csrc/stem_direct.cu): the 7x7/s2 window rows are built in shared memory
 * from a bf16 copy of the video (video_bf16: [B,T,H,W] bf16, even W, (H/2)*(W/2) a multiple of 16). Forward: y0 bf16
 * [B,T,OH*OW,64], w_packed bf16 [64,320] (column kt*64 + kh*8 + kw), bn_stats fp64 [2][64] (+=) or NULL; bit-identical to
 * svsr_stem_patch + the 5-tap implicit GEMM. Weight gradient: out fp32 [320, ldo] (row kt*64 + kh*8 + kw) += . */
int svsr_stem_conv_direct(const void* video_bf16, const void* w_packed, void* y0, double* bn_stats, int B, int T, int H,
                          int W, void* stream);
int svsr_stem_wgrad_direct(const void* video_bf16, const void* dz, float* out, int ldo, int B, int T, int H, int W,
                           void* stream);
/* nn.BatchNorm{2,3}d forward (+ optional residual with its own BN coefficients, + ReLU): lightning.py:51 and the
 * bn1/bn2/downsample.1 of every BasicBlock. coef (fp32 [4][C]) receives mean, invstd, scale, shift for backward.
 * train=1: batch statistics + running-stat update; train=0: running statistics. stats_scratch: fp64 [2*C]. */
int svsr_batchnorm_fwd(const void* x, int64_t rows, int C, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float eps, float momentum, int train, const void* res, const float* res_coef,
                       int relu, void* out, float* coef, double* stats_scratch, void* stream);
/* BatchNorm backward: g = dout * (relu_ref > 0 if relu_ref); dgamma += sum g*xhat; dbeta += sum g;
 * dc = scale*(g - mean g - xhat*mean(g*xhat)); gmask_out (optional) = g. kcoef_scratch: fp32 [2*C]. */
int svsr_batchnorm_bwd(const void* dout, const void* relu_ref, const void* c, int64_t rows, int C, const float* coef,
                       float* dgamma, float* dbeta, void* dc, void* gmask_out, double* stats_scratch,
                       float* kcoef_scratch, void* stream);
/* BatchNorm3d-apply + GELU(erf) + MaxPool3d((1,3,3),(1,2,2),(0,1,1)) (lightning.py:51-53) and its backward */
int svsr_stem_bn_gelu_pool_fwd(const void* y0, const float* coef, void* out, uint8_t* argmax, int N, int IH, int IW,
                               void* stream);
int svsr_stem_pool_gelu_bwd(const void* dout, const uint8_t* argmax, const void* y0, const float* coef, void* dz, int N,
                            int IH, int IW, void* stream);
/* The whole stem backward between the trunk gradient and the conv weight gradient in two passes, without storing dz:
 * MaxPool3d scatter * GELU' (lightning.py:52-53) -> BatchNorm3d backward (lightning.py:51). dgamma/dbeta += ;
 * dc = d loss / d y0 (bf16 [N,IH,IW,64]); dout is overwritten with the routed gradient dout * GELU'. stats_scratch: fp64 [128]; kcoef_scratch: fp32 [128]. */
int svsr_stem_bwd_fused(void* dout, const uint8_t* argmax, const void* y0, const float* coef, float* dgamma,
                        float* dbeta, void* dc, double* stats_scratch, float* kcoef_scratch, int N, int IH, int IW,
                        void* stream);
/* hidden.mean((2,3)) + CLS concat (lightning.py:118,148-150): x_stream fp32 [B, T+1, C] */
int svsr_meanpool_cls_fwd(const void* a, const float* cls, float* x_stream, int B, int T, int HW, int C, void* stream);
int svsr_meanpool_cls_bwd(const float* dx, void* dout, float* dcls, int B, int T, int HW, int C, void* stream);
/* x-transformers RMSNorm / rotary attention core / GEGLU */
int svsr_rmsnorm_fwd(const float* x, const float* g, void* y, float* inv, int M, int D, float eps, void* stream);
int svsr_rmsnorm_bwd(const void* dy, const float* x, const float* g, const float* inv, float* dx, void* dx_bf16,
                     float* dg, int M, int D, float eps, void* stream);
int svsr_rotary_table(float* tab, int n, void* stream);
int svsr_attention_fwd(const void* qkv, const float* rot, void* o, int B, int n, int heads, int rotary_v, void* stream);
int svsr_attention_bwd(const void* qkv, const float* rot, const void* d_o, void* dqkv, int B, int n, int heads,
                       int rotary_v, void* stream);
/* The forward of the x-transformers attention sublayer core in ONE kernel (lightning.py:95-105,158):
 * to_q | to_k | to_v projection of the normalised activations (TMA-fed tcgen05 GEMM, one head x four clips per CTA), rotary,
 * softmax, PV. xn bf16 [B*n, ldx]; w bf16 [3*heads*64, Kp] (q | k | v weight rows, Kp % 64 == 0 contracted columns, Kp <= ldx);
 * qkv bf16 [B*n, 3*heads*64] receives the projections (what svsr_attention_bwd reads); o bf16 [B*n, heads*64]. n <= 32. */
int svsr_attention_qkv_fwd(const void* xn, int ldx, const void* w, int Kp, const float* rot, void* qkv, void* o, int B,
                           int n, int heads, int rotary_v, void* stream);
/* GEGLU + Dropout(ff_dropout): u = dropout(h[:, :F] * gelu(h[:, F:])). The keep-mask is a counter-based function of
 * (seed, element index), regenerated identically by the backward; p_drop = 0 disables it. */
int svsr_geglu_fwd(const void* h, void* u, int M, int F, float p_drop, uint64_t seed, void* stream);
int svsr_geglu_bwd(const void* h, const void* du, void* dh, int M, int F, float p_drop, uint64_t seed, void* stream);
/* F.cross_entropy over quantised audio tokens (lightning.py:168-171): logits fp32 [B*T, A*G*V]; row (b,t), group
 * c=a*G+g is scored against tokens[b*tok_stride_b + (t*A+a)*G + g]. acc[0] += sum nll (fp64);
 * dlogits (bf16, optional) = (softmax - onehot) * dscale; *bad_token = 1 if a token is outside [0,V). */
int svsr_audio_ce(const float* logits, int ld, const int64_t* tokens, int64_t tok_stride_b, int B, int T, int A, int G,
                  int V, void* dlogits, double* acc, int* bad_token, float dscale, void* stream);
/* The fused audio head: `audio_projection` + reshape [B,T,A*G,V] + log-softmax + NLL in ONE tcgen05 kernel per direction
 * (LRW/video/src/lightning.py:82,168-171; LRS twin e2e_asr_transformer.py:142/157,198-201; README.md:47-53). The fp32
 * logits live only in TMEM / registers, never in HBM; target indexing tokens[b*tok_stride_b + (t*A+a)*G + g] is int64.
 *  x bf16 [B*T, ldx] (K = hidden columns, multiple of 64), w bf16 [A*G*V, ldw], bias fp32 [A*G*V] or null; V % 64 == 0.
 *  fwd: part = scratch of B*T * (A*G*V/64) float2, xt / lse = fp32 [B*T*A*G]; *loss_sum (fp64) += sum_rows (lse - x[target]);
 *       *bad_token = 1 when a token is outside [0,V) (that row adds nothing; its gradient row is zero).
 *  bwd: recomputes the projection tile by tile and writes dlogits bf16 [B*T, A*G*V] = (softmax - onehot) * dscale
 *       (* *grad_scale, optional device scalar) -- the operand of svsr_gemm_bf16 (dX = dlogits.W) and svsr_gemm_wgrad. */
int svsr_audio_head_fwd(const void* x, int ldx, const void* w, int ldw, int K, const float* bias, const int64_t* tokens,
                        int64_t tok_stride_b, int B, int T, int A, int G, int V, void* part, float* xt, float* lse,
                        double* loss_sum, int* bad_token, void* stream);
int svsr_audio_head_bwd(const void* x, int ldx, const void* w, int ldw, int K, const float* bias, const int64_t* tokens,
                        int64_t tok_stride_b, int B, int T, int A, int G, int V, const float* lse, float dscale,
                        const float* grad_scale, void* dlogits, int* bad_token, void* stream);
/* F.cross_entropy(logits_category, labels, label_smoothing) with int64 or soft fp32 labels (lightning.py:161-165);
 * acc[1] += sum loss, acc[2] += #top1, acc[3] += #top5 (lightning.py:177-183). */
int svsr_category_ce(const float* logits, int ld, const int64_t* labels, const float* soft_labels, int B, int C,
                     float eps, void* dlogits, int ldd, double* acc, float dscale, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * LRW word-level model step executor: the whole of TransformerLightningModule.forward (lightning.py:133-191) and
 * its backward as two calls. Parameters, gradients and BatchNorm buffers live in three flat fp32 arenas owned by
 * the caller (torch tensors); the engine defines the layout (svsr_lrw_param_info) using the reference's
 * state-dict names, so nn.Parameter views into the arena round-trip the reference's checkpoints. The gradient
 * arena is what the data-parallel step all-reduces with one NCCL call.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct svsr_lrw_config {
  int B, T, H, W;                 /* clips per step on this GPU, frames, crop height/width */
  int dim, depth, heads;          /* model.bert.{dim,depth,heads} (yaml:19-21); dim = 512 + data.use_word_boundary */
  int audio_alignment, vq_groups, audio_vocab; /* lightning.py:58-67 (there derived from the codec path) */
  int num_labels;                 /* model.bert.num_labels */
  int rotary_v;                   /* x-transformers 1.9.x rotates v as well (SURVEY Appendix A switch 1) */
  float lambda_audio;             /* optim.lambda_audio */
  float label_smoothing;          /* train.label_smoothing */
  float bn_eps, bn_momentum;      /* torch.nn.BatchNorm defaults 1e-5 / 0.1 */
  float ff_dropout;               /* model.bert.ff_dropout (training only; mask seeded per forward call) */
  /* model.bert.type: 0 = x-transformers Encoder (lightning.py:95-105), 1 = HuggingFace BertModel(BertConfig(**cfg))
   * (lightning.py:90-92): dim = hidden_size (512), depth = num_hidden_layers, heads = num_attention_heads (head 64) */
  int enc_type;
  int bert_intermediate, bert_max_pos; /* intermediate_size, max_position_embeddings */
  float bert_ln_eps, bert_hidden_dropout, bert_attn_dropout; /* layer_norm_eps, hidden_dropout_prob, attention_probs_dropout_prob */
  float emb_dropout, attn_dropout; /* model.bert.emb_dropout (lightning.py:106,150), model.bert.attn_dropout (x-transformers) */
} svsr_lrw_config;

int svsr_lrw_create(const svsr_lrw_config* cfg, void** handle);
int svsr_lrw_destroy(void* handle);
int64_t svsr_lrw_param_count(void* handle);     /* elements of the parameter (= gradient) arena */
int64_t svsr_lrw_decay_count(void* handle);     /* leading elements of the arena that AdamW decays (ndim >= 2) */
int64_t svsr_lrw_buffer_count(void* handle);    /* elements of the BatchNorm running-stat arena */
int64_t svsr_lrw_workspace_bytes(void* handle); /* activation + packed-operand workspace */
int svsr_lrw_num_params(void* handle);
int svsr_lrw_num_buffers(void* handle);
/* i-th tensor of the arena: reference state-dict name, shape[5], element offset, AdamW-decay flag (ndim >= 2) */
int svsr_lrw_param_info(void* handle, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset, int* decay);
int svsr_lrw_buffer_info(void* handle, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset);
int svsr_lrw_bind(void* handle, float* params, float* grads, float* buffers, void* workspace, int64_t workspace_bytes);
/* fp32 master weights -> bf16 tensor-core operand layouts; call after every optimizer step. Apart from the first call
 * of a binding the repack is enqueued by the NEXT forward call, on its stream (SVSR_PACK_OVERLAP=1: the stem's operand
 * first, the rest on the engine's side stream beside the stem kernels -- measured neutral, off by default). */
int svsr_lrw_pack_weights(void* handle, void* stream);
/* videos fp32 [B,1,T,H,W]; tokens int64 [B, >=T*A, G] with batch stride tok_stride_b (elements); labels int64 [B]
 * or soft_labels fp32 [B,num_labels] (CutMix); word_mask fp32 [B,T] when dim = 513 (data.use_word_boundary,
 * lightning.py:145-150: it becomes channel 512 of every frame token), else NULL. train: batch-stat BN + buffer update. skip_mask bit i drops encoder
 * sublayer i (layer_dropout decided on the host like the reference); dropout_seed seeds this step's ff_dropout masks.
 * metrics (device fp32[5]) = loss_total,
 * loss_category, loss_audio, accuracy_top1, accuracy_top5. */
int svsr_lrw_forward(void* handle, const float* videos, const int64_t* tokens, int64_t tok_stride_b,
                     const int64_t* labels, const float* soft_labels, const float* word_mask, int train,
                     uint32_t skip_mask, uint64_t dropout_seed, float* metrics, void* stream);
/* Parity-mode forward: same model, fp32 activations and split-bf16 ([hi|lo|hi].[hi|hi|lo]) tensor-core operands through
 * the same tcgen05 kernels (csrc/precise.cuh) -- fp32-class accuracy for north_star's 1e-3 output tolerance.
 * Forward only (no backward, running BatchNorm buffers untouched); results are read with svsr_lrw_tensor
 * ("last_hidden_state", "logits_audio", "logits_category"). precise_ws: caller-allocated scratch. word_mask: device
 * fp32 [B, T], required by the dim-513 word-boundary variant (it becomes hidden channel 512), NULL otherwise. Covers the
 * x-transformers and the HuggingFace-BERT encoder variants; dropouts are never applied (the deterministic function). */
int64_t svsr_lrw_precise_workspace_bytes(void* handle);
int svsr_lrw_forward_precise(void* handle, void* precise_ws, int64_t precise_ws_bytes, const float* videos,
                             const int64_t* tokens, int64_t tok_stride_b, const int64_t* labels,
                             const float* soft_labels, const float* word_mask, int train, uint32_t skip_mask,
                             float* metrics, void* stream);
/* forward_videos (lightning.py:112-119) only: fills the "inputs_embeds" tensor ([B,T+1,dim] fp32, row 0 = CLS) */
int svsr_lrw_forward_videos(void* handle, const float* videos, int train, void* stream);
/* (*grad_scale) * d loss_total / d params accumulated (+=) into the gradient arena; grad_scale is a DEVICE fp32
 * scalar (the upstream gradient autograd hands to loss_total) or NULL for 1. One backward per forward. */
int svsr_lrw_backward(void* handle, const float* grad_scale, void* stream);
/* The native step never writes the audio logits to HBM (the projection is fused with reshape + log-softmax + NLL,
 * lightning.py:168-171); this call materialises `logits_audio` (fp32 [B*T, A*G*V], svsr_lrw_tensor("logits_audio")) once
 * from the last forward's hidden states, for inspection / parity tests. */
int svsr_lrw_logits_audio(void* handle, void* stream);
/* Device-resident step control. The reference drops x-transformers sublayers with host RNG (layer_dropout,
 * lightning.py:95-105) and seeds its dropouts per step; as kernel ARGUMENTS those make every step a different launch
 * sequence. mode 1: {skip_mask, dropout_seed} are written to two control words in the workspace (on `stream`) and the
 * engine launches every sublayer's kernels each step, predicated on its bit -- svsr_lrw_forward's skip_mask /
 * dropout_seed arguments are ignored until mode 0 -- so ONE captured CUDA graph replays any step of the shipped
 * layer_dropout .2 / ff_dropout .3 config (call this before each replay). mode 0: host-valued control again. */
int svsr_lrw_step_control(void* handle, int mode, uint32_t skip_mask, uint64_t dropout_seed, void* stream);
/* The same backward in three stages, called in order, so that the data-parallel step can overlap communication with
 * compute: stage 0 = loss heads + encoder + mean-pool (completes cls_token, encoder and head gradients, ~160 MB; the
 * decayed part is the range svsr_lrw_early_grad_region returns), stage 1 = resnet.layer4 + layer3 (42 MB), stage 2 =
 * layer2 + layer1 + stem3d (3 MB: the only all-reduce that cannot hide behind compute). */
int svsr_lrw_backward_stage(void* handle, const float* grad_scale, int stage, void* stream);
int svsr_lrw_early_grad_region(void* handle, int64_t* begin, int64_t* end);
/* named activation for parity tests: last_hidden_state, logits_audio, ... dtype 0=f32 1=bf16 2=u8 3=i32 */
int svsr_lrw_tensor(void* handle, const char* name, void** ptr, int64_t* numel, int* dtype);

/* CutMix (LRW/video/src/augment.py:27-118) on the device. The reference mixes clip by clip, in place, with host RNG
 * decisions; syncvsr_b200/augment.py draws the same decisions and resolves the sequential swaps into source tables, so
 * one gather suffices: videos_out[i,0,t] = videos_in[vsrc[i,t],0,t] (frame_elems floats per frame),
 * audio_out[i,a,:] = audio_in[asrc[i,a],a,:] (int64, bit-exact), soft_labels[i] = (1-rate)*onehot(labels[i]) +
 * rate*onehot(labels[tgt[i]]) and wm_out likewise for clips with mixed[i] != 0, plain one-hot / copy otherwise. */
int svsr_cutmix_gather(const float* videos_in, float* videos_out, const int* vsrc, int B, int T, int64_t frame_elems,
                       const int64_t* audio_in, int64_t* audio_out, const int* asrc, int Ta, int G, const int64_t* labels,
                       const int* tgt, const float* rate, const uint8_t* mixed, float* soft_labels, int num_labels,
                       const float* wm_in, float* wm_out, int Tw, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Data path (reference: LRW/video/src/data.py:32-68 Dataset.__getitem__ and the transform
 * pipelines of data.py:156-171, run there per sample on CPU DataLoader workers).
 * --------------------------------------------------------------------------------------------------------- */
#define SVSR_JPEG_DESC_INTS 24
#define SVSR_JPEG_HUFF_BYTES 1536
/* HOST function (no GPU needed): walks the markers of n concatenated baseline JPEG files (frame f = blob[offsets[f],
 * offsets[f+1])) and writes one descriptor per frame (layout in csrc/datapath.cu) plus the pools of distinct
 * quantisation tables (natural order, u16 [qcap][64]) and decode-ready Huffman tables (u8 [hcap][SVSR_JPEG_HUFF_BYTES]).
 * Supports what TurboJPEG.encode writes (data.py:41 decodes it): 8-bit sequential Huffman, 1 or 3 components in one
 * interleaved scan, any sampling factors, restart intervals. Progressive / arithmetic files are refused. */
int svsr_jpeg_parse(const uint8_t* blob, const int64_t* offsets, int n, int32_t* desc, uint16_t* qtabs, int qcap, int* n_q,
                    uint8_t* htabs, int hcap, int* n_h);
/* data.py:41 `jpeg.decode(img, pixel_format=TJPF_GRAY)` for a batch: out u8 [n, H, W] = luminance plane, bit-identical to
 * libjpeg-turbo (integer "islow" IDCT). All frames must share W x H; blocks_w x blocks_h is the luminance block grid
 * (ceil(W / (8*hmax)) * h0 by ceil(H / (8*vmax)) * v0); coef_scratch holds n*blocks_w*blocks_h*64 int16. */
int svsr_jpeg_decode_gray(const uint8_t* blob_dev, const int32_t* desc_dev, int n, const uint16_t* qtabs_dev,
                          const uint8_t* htabs_dev, int16_t* coef_scratch, uint8_t* out, int W, int H, int blocks_w,
                          int blocks_h, void* stream);
/* data.py:157-171 transform pipeline for a batch of decoded clips: frames u8 [B,T,H,W] -> out f32 [B,1,T,OH,OW]:
 * x/255 -> horizontal flip -> crop (top,left,h,w) + antialiased bilinear resize to OH x OW (RandomResizedCrop / Resize /
 * CenterCrop) -> TimeMask: frames [t0,t1) <- mean of the clip (augment.py:120-143) -> (x-mean)/std.
 * xform int32 [B][8] = {flip, top, left, crop_h, crop_w, mask_t0, mask_t1, 0}, drawn on the host in the reference's RNG
 * order (syncvsr_b200/data.py). clip_sum: B doubles of scratch (needed when time_mask != 0). */
int svsr_video_transform(const uint8_t* frames, const int* xform, float* out, double* clip_sum, int B, int T, int H, int W,
                         int OH, int OW, float mean, float stdv, int time_mask, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * LRS sentence-level operators (reference: LRS/video/espnet/nets/pytorch_backend/, paths below relative to it).
 * --------------------------------------------------------------------------------------------------------- */
/* transformer/layer_norm.py:12-33 (eps 1e-12): y = (x-mean)*rstd*gamma+beta over the last dim D (multiple of 128,
 * <= 1024); x fp32 [M,D]; y_bf16 and/or y_f32; stats fp32 [M,2] = mean, rstd (saved for backward). */
int svsr_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y_bf16, float* y_f32, float* stats,
                       int M, int D, float eps, void* stream);
/* dy is bf16 (dy_f32 NULL) or fp32 (may alias dx); dx = or += (accumulate); dgamma/dbeta += . */
int svsr_layernorm_bwd(const void* dy_bf16, const float* dy_f32, const float* x, const float* gamma, const float* stats,
                       float* dx, int accumulate, float* dgamma, float* dbeta, int M, int D, void* stream);
/* transformer/convolution.py:61 GLU over channels: u[M,C] = h[:, :C] * sigmoid(h[:, C:]) and its backward */
int svsr_glu_fwd(const void* h, void* u, int64_t M, int C, void* stream);
int svsr_glu_bwd(const void* h, const void* du, void* dh, int64_t M, int C, void* stream);
/* transformer/convolution.py:40-48,64 depthwise Conv1d along time: x,y bf16 [B,T,C]; w fp32 [C,K] (K odd <= 31);
 * flip=1 applies the reversed kernel (input gradient). wgrad: dw[C,K] +=, dbias[C] += . */
int svsr_dwconv1d_fwd(const void* x, const float* w, const float* bias, void* y, int B, int T, int C, int K, int flip,
                      void* stream);
int svsr_dwconv1d_wgrad(const void* x, const void* dy, float* dw, float* dbias, int B, int T, int C, int K, void* stream);
/* BatchNorm1d column reductions over [rows, C] (C % 64 == 0; convolution.py:49,65): mode 0: stats[0..C) += sum x,
 * stats[C..2C) += sum x^2; mode 1: g = dout * swish'(x*scale+shift), stats += sum g, sum g*xhat (coef fp32 [4][C]). */
int svsr_bn_col_reduce(const void* x, const void* dout, const float* coef, int64_t rows, int C, double* stats, int mode,
                       void* stream);
/* Multi-head attention core, d_k = 64 (transformer/attention.py:38-108 and RelPositionMultiHeadedAttention 192-278 with
 * rel_shift as the index map bd[i,j] = raw[i, j-i+Tk-1]): scores = scale*((q+u).k_j + (q+v).p[j-i+Tk-1]); keys
 * j >= klen[b] and (causal) j > i masked; softmax with masked probabilities zero; . V. q/k/v/o bf16 row-major with
 * pitches ld*, head h in columns [h*64, h*64+64); p bf16 [2Tk-1, ldp] or NULL; bias_u/v fp32 [H,64] or NULL;
 * klen int32 [B] or NULL; lse fp32 [B,H,Tq] (saved for backward). */
int svsr_attention_core_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* p,
                            int ldp, const float* bias_u, const float* bias_v, const int* klen, int causal, int B,
                            int H, int Tq, int Tk, float scale, void* o, int ldo, float* lse, float drop_p,
                            uint64_t drop_seed, void* stream);
/* dq/dk/dv bf16 (pitches as q/k/v); dp fp32 [2Tk-1, H*64] +=; dbias_u/v fp32 [H,64] +=;
 * scratch: svsr_attention_scratch_bytes(B,H,Tq,Tk) bytes. */
int64_t svsr_attention_scratch_bytes(int B, int H, int Tq, int Tk);
int svsr_attention_core_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, const void* p,
                            int ldp, const float* bias_u, const float* bias_v, const int* klen, int causal, int B,
                            int H, int Tq, int Tk, float scale, const void* o, int ldo, const float* lse,
                            const void* d_o, void* dq, void* dk, void* dv, float* dp, float* dbias_u, float* dbias_v,
                            void* scratch, float drop_p, uint64_t drop_seed, void* stream);
/* ctc.py:64-73,83-151: log_softmax + CTCLoss(reduction="sum", zero_infinity=True), blank 0. logits fp32 [B*T, ld]
 * (V valid columns); labels int64 [B,Lmax] padded with -1; in_len int32 [B]. acc[slot] += sum_b nll_b (fp64);
 * dlogits bf16 [B*T, ld] (optional) = dscale * d(sum nll)/dlogits. scratch: svsr_ctc_scratch_bytes(B,T,Lmax). */
int64_t svsr_ctc_scratch_bytes(int B, int T, int Lmax);
int svsr_ctc_loss(const float* logits, int ld, int V, const int64_t* labels, int Lmax, const int* in_len, int B, int T,
                  void* dlogits, double* acc, int slot, float dscale, void* scratch, void* stream);
/* transformer/label_smoothing_loss.py:41-63 (normalize_length=False) + nets_utils.py:303 th_accuracy: logits fp32
 * [rows, ld]; target int64 [rows], -1 ignored. acc[slot] += sum KL, acc[slot+1] += #correct, acc[slot+2] += #scored;
 * dlogits bf16 [rows, ld] (optional) = dscale * (softmax - smoothed one-hot). */
int svsr_label_smoothing_loss(const float* logits, int ld, int V, const int64_t* target, int rows, float smoothing,
                              void* dlogits, double* acc, int slot, float dscale, void* stream);
/* Dropout as a counter-based mask: keep(i) is a pure function of (seed, element index i), so backward regenerates it.
 * svsr_dropout_mask writes the uint8 keep-mask of n elements (tests); svsr_dropout_bf16 is y = x * keep / (1-p). */
int svsr_dropout_mask(uint8_t* out, int64_t n, float p, uint64_t seed, void* stream);
int svsr_dropout_bf16(const void* x, void* y, int64_t n, float p, uint64_t seed, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * LRS sentence-level model step executor: E2E.forward (e2e_asr_transformer.py:186-227) and its backward. Same
 * arena conventions as the LRW executor; parameter names are the reference's state-dict keys
 * (encoder.frontend.*, encoder.embed.0.*, encoder.encoders.N.*, encoder.after_norm.*, decoder.*, ctc.ctc_lo.*,
 * audio_classifier.*). linear_q/k/v of every attention are adjacent so that they run as one GEMM.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct svsr_lrs_config {
  int B, T, H, W;               /* clips per step on this GPU, padded frames per clip, crop height/width */
  int adim, aheads, eunits, elayers; /* model.visual_backbone.{adim,aheads,eunits,elayers} (lrs2.yaml:16-19); d_k = 64 */
  int dlayers, dunits;          /* decoder depth / FFN width (ddim == adim, dheads == aheads) */
  int odim;                     /* vocabulary incl. <blank> = 0 and <sos/eos> = odim-1 */
  int cnn_kernel;               /* cnn_module_kernel (31) */
  int audio_alignment, vq_groups, audio_vocab; /* e2e_asr_transformer.py:130-160; alignment 0 = codec None */
  int Lmax;                     /* longest decoder input (label length + 1) the workspace is sized for */
  float mtlalpha, lsm_weight, audio_weight;    /* lrs2.yaml:15,34,36 */
  float bn_eps, bn_momentum;
  float dropout_rate, attn_dropout_rate;       /* lrs2.yaml:21-22 (training only; masks seeded per forward call) */
} svsr_lrs_config;

int svsr_lrs_create(const svsr_lrs_config* cfg, void** handle);
int svsr_lrs_destroy(void* handle);
int64_t svsr_lrs_param_count(void* handle);
int64_t svsr_lrs_decay_count(void* handle);
int64_t svsr_lrs_buffer_count(void* handle);
int64_t svsr_lrs_workspace_bytes(void* handle);
int svsr_lrs_num_params(void* handle);
int svsr_lrs_num_buffers(void* handle);
int svsr_lrs_param_info(void* handle, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset, int* decay);
int svsr_lrs_buffer_info(void* handle, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset);
int svsr_lrs_bind(void* handle, float* params, float* grads, float* buffers, void* workspace, int64_t workspace_bytes);
int svsr_lrs_pack_weights(void* handle, void* stream);
/* x fp32 [B,T,1,H,W]; lengths int64 [B]; tokens int64 [B, >=T*A, G] (batch stride tok_stride_b) or NULL (no audio
 * loss); label int64 [B,label_len] padded with -1 (sos/eos are added on the device, add_sos_eos.py:12-31).
 * dropout_seed seeds this step's dropout masks (train only). metrics (device fp32[5]) = loss, loss_ctc, loss_att,
 * loss_audio, acc  (e2e_asr_transformer.py:227). */
int svsr_lrs_forward(void* handle, const float* x, const int64_t* lengths, const int64_t* tokens, int64_t tok_stride_b,
                     const int64_t* label, int label_len, int train, uint64_t dropout_seed, float* metrics,
                     void* stream);
/* Encoder.forward only (transformer/encoder.py:257-289; called directly by inference, LRS/video/lightning.py:100):
 * fills "encoder_out" fp32 [B,T,adim]. lengths may be NULL (masks=None). */
int svsr_lrs_encode(void* handle, const float* x, const int64_t* lengths, int train, uint64_t dropout_seed,
                    void* stream);
/* (*grad_scale) * d loss / d params accumulated (+=) into the gradient arena; one backward per (train) forward. */
int svsr_lrs_backward(void* handle, const float* grad_scale, void* stream);
/* The same backward in three stages, called in order: 0 = loss heads + attention decoder, 1 = encoder.after_norm + the
 * Conformer blocks, 2 = embed + visual frontend. When a stage returns, the gradients of ITS parameters are final on
 * `stream`, so a data-parallel caller all-reduces them while the next stage computes (syncvsr_b200/train.py). */
int svsr_lrs_backward_stage(void* handle, const float* grad_scale, int stage, void* stream);
/* Device-resident step seed. mode 1: writes dropout_seed to the engine's control word on `stream`; from then on every
 * dropout site of forward / encode / backward reads the step seed from device memory (the dropout_seed ARGUMENTS are
 * ignored), so ONE captured CUDA graph replays every step of the shipped dropout_rate / transformer_attn_dropout_rate
 * config (lrs2.yaml:21-22,46-47): call this before each replay. Same masks as the host-valued path for the same seed.
 * mode 0: host-valued seeds again. */
int svsr_lrs_step_control(void* handle, int mode, uint64_t dropout_seed, void* stream);
/* named tensors: encoder_out, embed_out, frontend, logits_audio, ctc_logits, pred, ys_in, ys_out, layer<i>.x<k>;
 * dtype 0=f32 1=bf16 2=u8 3=i32 4=i64 */
int svsr_lrs_tensor(void* handle, const char* name, void** ptr, int64_t* numel, int* dtype);
/* LRS twin of svsr_lrw_logits_audio: fills "logits_audio" (fp32 [B*T, A*G*V]) from the last forward's encoder output. */
int svsr_lrs_logits_audio(void* handle, void* stream);

/* Fused global-norm clip + AdamW over the flat arenas (lightning.py:216-221; Trainer gradient_clip_val). The arena
 * is [decayed | non-decayed]: the first n_decay elements get weight decay. grad_div divides gradients first (world
 * size after a SUM all-reduce). scratch: >= 32 bytes device memory; afterwards {fp64 sumsq, fp32 clip coef, fp32 norm}. */
int svsr_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_decay,
                    int64_t n_total, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    float max_norm, float grad_div, void* scratch, void* stream);

/* The same step with per-range Adam step counts: seg_begin = nseg+1 ascending HOST element offsets (multiples of 4,
 * seg_begin[0] = 0, seg_begin[nseg] = n_total), seg_step = nseg HOST step counts; a range with step 0 received no
 * gradient this step (`p.grad is None` in the reference: a sublayer dropped by layer_dropout, lightning.py:95-105) and is
 * left untouched -- no weight decay, no moment update, no step increment -- exactly like torch.optim.AdamW. */
#define SVSR_ADAMW_MAX_SEGMENTS 160
int svsr_adamw_step_segmented(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_decay,
                              int64_t n_total, float lr, float beta1, float beta2, float eps, float weight_decay,
                              const int64_t* seg_begin, const int32_t* seg_step, int nseg, float max_norm,
                              float grad_div, void* scratch, void* stream);

/* Kernels launched by this library so far in this process (bench.py reports the per-region difference). */
long long svsr_launch_count(void);
/* Per-launch CUDA-event timing of the tensor-core kernels: enable (resets), run, synchronise, read.
 * kind 0 = igemm_kernel (conv fprop/dgrad, linear fwd/dgrad), 1 = wgrad_kernel. flops are algorithmic (2*MAC). */
int svsr_prof_enable(int on);
int svsr_prof_read(int kind, double* total_ms, double* total_flops, int* launches);


#ifdef __cplusplus
}
#endif
#endif /* SVSR_H_ */
