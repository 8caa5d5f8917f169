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

/* Developer hardware probe (see csrc/debug_probe.cu); not part of the product path. */
int svsr_debug_rowshift(const void* a, const void* b, float* out, int shift, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVSR_H_ */
