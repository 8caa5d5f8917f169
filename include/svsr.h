/* syncvsr_b200 C ABI -- the drop-in boundary of the B200-native SyncVSR hot path.
 *
 * The reference (KAIST-AILab/SyncVSR) is pure Python: its "FFI" for this path is the set of torch.nn
 * calls made by LRW/video/src/lightning.py:49-55,82,107-119,133-191 (stem3d, resnet.layer1-4, encoder,
 * audio_projection, category_classifier, the two cross-entropies). Each entry point below replaces one
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

/* y[N,OH,OW,Cout] = conv2d(x[N,H,W,Cin], w) with w packed as [Cout, R, S, Cin] bf16, zero padding `pad`,
 * stride 1 or 2, no bias (+ resid). Cin % 64 == 0. Replaces the Conv2d calls inside resnet.layer1-4
 * (lightning.py:114-117; timm/torchvision BasicBlock). */
int svsr_conv2d_fprop(const void* x, const void* w, void* y, const void* resid, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int stride, int pad, int out_fp32, void* stream);

/* dx[N,H,W,Cin] = conv2d input-gradient of dy[N,OH,OW,Cout]; wd is the weight packed for dgrad as
 * [Cin, R*S*Cout] bf16 (column (r*S+s)*Cout+co holds W[co][ci][r][s]). If resid != NULL it is added (it may alias
 * dx: gradient accumulation of the residual branch). For stride 2 the four output-parity classes are issued as
 * four launches; pixels no tap reaches (1x1 stride-2) are left untouched, so dx must be pre-zeroed in that case.
 * Replaces autograd's conv backward-data for resnet.layer1-4 (lightning.py:114-117). */
int svsr_conv2d_dgrad(const void* dy, const void* wd, void* dx, const void* resid, int N, int H, int W, int Cin,
                      int Cout, int R, int S, int stride, int pad, int out_fp32, void* stream);

/* dw[R*S*Cin, Cout] (fp32, += accumulate) = weight gradient of conv2d: row (r*S+s)*Cin+ci, column co.
 * x[N,H,W,Cin], dy[N,OH,OW,Cout] bf16. Replaces autograd's conv backward-weight. */
int svsr_conv2d_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int R, int S,
                      int stride, int pad, void* stream);

/* dw[N,K] (fp32, pitch ldw, += accumulate) = dy[M,N]^T . x[M,K]; dy, x bf16 with pitches ldy, ldx.
 * Replaces autograd's Linear backward-weight. N % 64 == 0. */
int svsr_gemm_wgrad(const void* dy, int ldy, const void* x, int ldx, float* dw, int ldw, int M, int N, int K,
                    void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * LRW word-level model step executor: the whole of TransformerLightningModule.forward (lightning.py:133-191) and
 * its backward as two calls. Parameters, gradients and BatchNorm buffers live in three flat fp32 arenas owned by
 * the caller (torch tensors); the engine defines the layout (svsr_lrw_param_info) using the reference's
 * state-dict names, so nn.Parameter views into the arena round-trip the reference's checkpoints. The gradient
 * arena is what the data-parallel step all-reduces with one NCCL call.
 * --------------------------------------------------------------------------------------------------------- */
typedef struct svsr_lrw_config {
  int B, T, H, W;                 /* clips per step on this GPU, frames, crop height/width */
  int dim, depth, heads;          /* model.bert.{dim,depth,heads} (yaml:19-21) */
  int audio_alignment, vq_groups, audio_vocab; /* lightning.py:58-67 (there derived from the codec path) */
  int num_labels;                 /* model.bert.num_labels */
  int rotary_v;                   /* x-transformers 1.9.x rotates v as well (SURVEY Appendix A switch 1) */
  float lambda_audio;             /* optim.lambda_audio */
  float label_smoothing;          /* train.label_smoothing */
  float bn_eps, bn_momentum;      /* torch.nn.BatchNorm defaults 1e-5 / 0.1 */
} svsr_lrw_config;

int svsr_lrw_create(const svsr_lrw_config* cfg, void** handle);
int svsr_lrw_destroy(void* handle);
int64_t svsr_lrw_param_count(void* handle);     /* elements of the parameter (= gradient) arena */
int64_t svsr_lrw_buffer_count(void* handle);    /* elements of the BatchNorm running-stat arena */
int64_t svsr_lrw_workspace_bytes(void* handle); /* activation + packed-operand workspace */
int svsr_lrw_num_params(void* handle);
int svsr_lrw_num_buffers(void* handle);
/* i-th tensor of the arena: reference state-dict name, shape[5], element offset, AdamW-decay flag (ndim >= 2) */
int svsr_lrw_param_info(void* handle, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset, int* decay);
int svsr_lrw_buffer_info(void* handle, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset);
int svsr_lrw_bind(void* handle, float* params, float* grads, float* buffers, void* workspace, int64_t workspace_bytes);
/* fp32 master weights -> bf16 tensor-core operand layouts; call after every optimizer step */
int svsr_lrw_pack_weights(void* handle, void* stream);
/* videos fp32 [B,1,T,H,W]; tokens int64 [B, >=T*A, G] with batch stride tok_stride_b (elements); labels int64 [B]
 * or soft_labels fp32 [B,num_labels] (CutMix). train: batch-stat BN + buffer update. skip_mask bit i drops encoder
 * sublayer i (layer_dropout decided on the host like the reference). metrics (device fp32[5]) = loss_total,
 * loss_category, loss_audio, accuracy_top1, accuracy_top5. */
int svsr_lrw_forward(void* handle, const float* videos, const int64_t* tokens, int64_t tok_stride_b,
                     const int64_t* labels, const float* soft_labels, int train, uint32_t skip_mask, float* metrics,
                     void* stream);
/* forward_videos (lightning.py:112-119) only: fills the "inputs_embeds" tensor ([B,T+1,dim] fp32, row 0 = CLS) */
int svsr_lrw_forward_videos(void* handle, const float* videos, int train, void* stream);
/* (*grad_scale) * d loss_total / d params accumulated (+=) into the gradient arena; grad_scale is a DEVICE fp32
 * scalar (the upstream gradient autograd hands to loss_total) or NULL for 1. One backward per forward. */
int svsr_lrw_backward(void* handle, const float* grad_scale, void* stream);
/* named activation for parity tests: last_hidden_state, logits_audio, ... dtype 0=f32 1=bf16 2=u8 3=i32 */
int svsr_lrw_tensor(void* handle, const char* name, void** ptr, int64_t* numel, int* dtype);

/* Developer hardware probe (see csrc/debug_probe.cu); not part of the product path. */
int svsr_debug_rowshift(const void* a, const void* b, float* out, int shift, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVSR_H_ */
