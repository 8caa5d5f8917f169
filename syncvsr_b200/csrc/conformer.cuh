// Non-GEMM kernels of the LRS sentence-level path (Conformer encoder, CTC head, attention decoder).
// Reference: /root/reference/LRS/video/espnet/nets/pytorch_backend/ -- transformer/{layer_norm.py:12-33,
// attention.py:38-278, convolution.py:14-83, embedding.py:33-217, label_smoothing_loss.py:41-63,
// add_sos_eos.py:12-31}, ctc.py:64-151. Activations bf16 row-major [rows, channels], residual stream fp32,
// every reduction in fp32 (fp64 across blocks for batch statistics).
#pragma once
#include "common.cuh"

namespace svsr {

// ---- LayerNorm (eps 1e-12): layer_norm.py:12-33 --------------------------------------------------------------
// y = (x - mean) * rstd * gamma + beta; stats[m] = {mean, rstd}. y_bf16 / y_f32 optional (at least one). D % 128 == 0.
int layernorm_fwd(const float* x, const float* gamma, const float* beta, __nv_bfloat16* y_bf16, float* y_f32,
                  float* stats, int M, int D, float eps, cudaStream_t s);
// dy is bf16 (dy_f32 == nullptr) or fp32. dx: accumulate != 0 -> dx += grad, else dx = grad. dgamma/dbeta += .
int layernorm_bwd(const __nv_bfloat16* dy_bf16, const float* dy_f32, const float* x, const float* gamma,
                  const float* stats, float* dx, int accumulate, float* dgamma, float* dbeta, int M, int D,
                  cudaStream_t s);

// ---- GLU over channels (convolution.py:61): u = h[:, :C] * sigmoid(h[:, C:]) -----------------------------------
int glu_fwd(const __nv_bfloat16* h, __nv_bfloat16* u, long long M, int C, cudaStream_t s);
int glu_bwd(const __nv_bfloat16* h, const __nv_bfloat16* du, __nv_bfloat16* dh, long long M, int C, cudaStream_t s);

// ---- depthwise Conv1d along time (convolution.py:40-48,64): x,y [B,T,C] bf16; w fp32 [C,K] (K odd <= 31) ----------
// flip = 1 correlates with the reversed kernel and no bias: the input gradient of the same convolution.
// w_transposed = 1: w is a [K, C] copy (transpose_f32) so that the per-tap weight loads are coalesced
int dwconv1d_fwd(const __nv_bfloat16* x, const float* w, const float* bias, __nv_bfloat16* y, int B, int T, int C, int K,
                 int flip, cudaStream_t s, int w_transposed = 0);
int transpose_f32(const float* in, float* out, int R, int Cc, cudaStream_t s);  // in [R, Cc] -> out [Cc, R]
// dw[C,K] += sum_{b,t} dy[b,t,c] * x[b,t+k-pad,c]; dbias[C] += sum dy
int dwconv1d_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* dbias, int B, int T, int C, int K,
                   cudaStream_t s);

// ---- BatchNorm1d over [rows, C] for any C % 64 == 0 (convolution.py:49,65; statistics include padded frames) ------
// mode 0: stats[0..C) += sum x, stats[C..2C) += sum x^2
// mode 1: g = dout * swish'(x*scale+shift); stats[0..C) += sum g, stats[C..2C) += sum g * xhat   (coef = [4][C])
int bn_col_reduce(const __nv_bfloat16* x, const __nv_bfloat16* dout, const float* coef, long long rows, int C,
                  double* stats, int mode, cudaStream_t s);

// ---- multi-head attention core, d_k = 64 ------------------------------------------------------------------------
// scores[i,j] = scale * ((q_i + u) . k_j  +  (q_i + v) . p[j - i + Tk - 1])   (second term only if p != nullptr:
// RelPositionMultiHeadedAttention, attention.py:192-278 with rel_shift as an index map; Tq == Tk then);
// keys j >= klen[b] (klen != nullptr) and, if causal, j > i are masked; softmax; masked probabilities are zero; . V.
struct AttnProblem {
  const __nv_bfloat16 *q = nullptr, *k = nullptr, *v = nullptr;  // head h lives in columns [h*64, h*64+64)
  int ldq = 0, ldk = 0, ldv = 0;                                 // row pitches (elements); rows are (b, t)
  const __nv_bfloat16* p = nullptr;                              // [2*Tk-1, ldp] or nullptr
  int ldp = 0;
  const float *bias_u = nullptr, *bias_v = nullptr;  // [H, 64] fp32 or nullptr
  const int* klen = nullptr;                         // [B] or nullptr
  int causal = 0;
  float drop_p = 0.f;  // dropout on the attention probabilities (attention.py:81); mask index ((b*H+h)*Tq+i)*Tk+j
  unsigned long long drop_seed = 0;
  const unsigned long long* seed_base = nullptr;  // optional device word added to drop_seed (graph-replayed steps)
  int B = 0, H = 0, Tq = 0, Tk = 0;
  float scale = 0.125f;
  __nv_bfloat16* o = nullptr;  // [B*Tq, ldo]
  int ldo = 0;
  float* lse = nullptr;  // [B, H, Tq] log-sum-exp of the masked scores (saved for backward)
};
int attention_core_fwd(const AttnProblem& a, cudaStream_t s);
// Backward. d_o [B*Tq, ldo] bf16. Outputs dq/dk/dv (bf16, same pitches as q/k/v; dk/dv written for every key row,
// zeros for masked keys), dp fp32 [2*Tk-1, H*64] accumulated over the batch (+=, zero it first), dbias_u/dbias_v
// fp32 [H,64] (+=). scratch: fp32 [2][B,H,Tq,Tk] (probabilities and score gradients).
struct AttnGrads {
  const __nv_bfloat16* d_o = nullptr;
  __nv_bfloat16 *dq = nullptr, *dk = nullptr, *dv = nullptr;
  int lddq = 0, lddk = 0, lddv = 0;
  float* dp = nullptr;
  float *dbias_u = nullptr, *dbias_v = nullptr;
  float* scratch = nullptr;
};
int attention_core_bwd(const AttnProblem& a, const AttnGrads& g, cudaStream_t s);
size_t attention_scratch_bytes(int B, int H, int Tq, int Tk);

// ---- positional encodings (embedding.py:33-88,153-217) ---------------------------------------------------------------
// rel: row r of [2T-1, D] encodes relative position T-1-r (bf16, GEMM operand of linear_pos)
int rel_pos_table(__nv_bfloat16* pe, int T, int D, cudaStream_t s);
// decoder input: x[b,l,:] = emb[tok[b,l]] * sqrt(D) + pe_abs[l]   (fp32 stream) and its weight gradient (+=, atomics)
// seed_base (every dropout entry below): optional device word added to the seed, so that a captured CUDA graph replays
// with a new step seed (engine_lrs.cu, svsr_lrs_step_control)
int embed_posenc_fwd(const long long* tok, const float* emb, float* x, int rows, int L, int D, int V, cudaStream_t s,
                     float drop_p = 0.f, unsigned long long drop_seed = 0, const unsigned long long* seed_base = nullptr);
int embed_bwd(const long long* tok, const float* dx, float* demb, int rows, int D, int V, cudaStream_t s,
              float drop_p = 0.f, unsigned long long drop_seed = 0, const unsigned long long* seed_base = nullptr);

// ---- CTC (ctc.py:64-73,83-151: log_softmax + CTCLoss(reduction=sum, zero_infinity) / batch) ------------------------------
// logits fp32 [B*T, ld] (V valid columns); labels int64 [B, Lmax] padded with -1; in_len int [B].
// acc[slot] += sum_b nll_b (fp64). dlogits bf16 [B*T, ld] = dscale * d(sum nll)/dlogits (zero rows for t >= in_len and
// for infeasible samples). scratch fp32: lse [B*T] + alpha, beta [B,T,S] + nll [B], S = 2*Lmax+1 (see ctc_scratch_bytes).
int ctc_loss_fwd_bwd(const float* logits, int ld, int V, const long long* labels, int Lmax, const int* in_len, int B,
                     int T, __nv_bfloat16* dlogits, double* acc, int slot, float dscale, float* scratch, cudaStream_t s);
size_t ctc_scratch_bytes(int B, int T, int Lmax);

// ---- LabelSmoothingLoss (label_smoothing_loss.py:41-63, normalize_length=False) + th_accuracy (nets_utils.py:303) ----
// logits fp32 [rows, ld]; target int64 [rows] (-1 = ignore). acc[slot] += sum KL (fp64), acc[slot+1] += #correct,
// acc[slot+2] += #scored. dlogits bf16 = dscale * (softmax - smoothed one-hot), zero for ignored rows.
int label_smoothing_loss(const float* logits, int ld, int V, const long long* target, int rows, float smoothing,
                         __nv_bfloat16* dlogits, double* acc, int slot, float dscale, cudaStream_t s);

// out[0..4] = loss, loss_ctc, loss_att, loss_audio, acc (e2e_asr_transformer.py:218-227) from the fp64 accumulators
// acc[0] audio nll sum, acc[1] ctc nll sum, acc[2] KL sum, acc[3] #correct, acc[4] #scored
int lrs_finalize_metrics(const double* acc, float* out, int B, long long audio_rows, float mtlalpha, float audio_weight,
                         int has_audio, cudaStream_t s, const int* bad = nullptr);

// AdaptiveAvgPool2d(1) of the trunk output (resnet.py:126,175-176): feats[n,:] = mean_hw a[n,hw,:] (bf16) and backward
int meanpool_bf16(const __nv_bfloat16* a, __nv_bfloat16* out, long long N, int HW, int C, cudaStream_t s);
int meanpool_bf16_bwd(const __nv_bfloat16* df, __nv_bfloat16* dout, long long N, int HW, int C, cudaStream_t s);

// small helpers
int add_f32(float* dst, const float* src, long long n, cudaStream_t s);                                 // dst += src
// y = bf16(alpha * x * dropout_mask): the branch gradient of `residual + alpha * dropout(branch)`
int cast_scale_f32_bf16(const float* x, __nv_bfloat16* y, long long n, float alpha, cudaStream_t s, float drop_p = 0.f,
                        unsigned long long drop_seed = 0, const unsigned long long* seed_base = nullptr);
// stand-alone Dropout forward (bf16) and its backward accumulated into an fp32 gradient (dst += mask * src)
int dropout_bf16(const __nv_bfloat16* x, __nv_bfloat16* y, long long n, float p, unsigned long long seed, cudaStream_t s,
                 const unsigned long long* seed_base = nullptr);
int dropout_add_bf16_to_f32(float* dst, const __nv_bfloat16* src, long long n, float p, unsigned long long seed,
                            cudaStream_t s, const unsigned long long* seed_base = nullptr);
int dropout_mask_u8(unsigned char* out, long long n, float p, unsigned long long seed, cudaStream_t s);  // tests
int lengths_i64_to_i32(const long long* in, int* out, int n, int maxv, cudaStream_t s);                 // clamp to [0, maxv]

}  // namespace svsr
