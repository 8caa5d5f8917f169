// Loss heads of the LRW step (lightning.py:161-191): per-frame cross-entropy over quantised audio tokens and the
// word-classification cross-entropy, each producing the loss AND the logits gradient in one pass over fp32 logits.
#pragma once
#include "common.cuh"

namespace svsr {

// logits fp32 [B*T, A*G*V] (pitch ld, bias already added). Row (b,t), channel c = a*G+g, class v.
// Target = tokens[b*tok_stride_b + (t*A + a)*G + g]  (int64; the reference's audio_tokens[:, :T*A].flatten()).
// acc[0] += sum of -log p[target] over all B*T*A*G rows (fp64). dlogits (bf16, pitch ld) = (softmax - onehot)*dscale.
// A token outside [0,V) sets *bad_token (device int) and leaves a ZERO gradient row (never a stale one).
int audio_ce(const float* logits, int ld, const long long* tokens, long long tok_stride_b, int B, int T, int A, int G,
             int V, __nv_bfloat16* dlogits, double* acc, int* bad_token, float dscale, cudaStream_t s);

// Fused audio head, forward part 2 (part 1 = the projection GEMM with IgemmCe mode 1, igemm.cuh): merges the per-slot
// (max, sum exp) partials [B*T, A*G*V/64] into lse [B*T*A*G] and adds sum(lse - logit[target]) to acc[0].
int ce_finalize(const float2* part, const float* xt, const long long* tokens, long long tok_stride_b, int B, int T, int A,
                int G, int V, float* lse, double* acc, cudaStream_t s);

// logits fp32 [B, C] (pitch ld). Hard labels (int64 [B]) or soft labels (fp32 [B,C]); label smoothing eps.
// acc[1] += sum loss ; acc[2] += #top1 ; acc[3] += #top5. dlogits (bf16, pitch ldd, columns >= C zeroed up to ldd).
int category_ce(const float* logits, int ld, const long long* labels, const float* soft_labels, int B, int C, float eps,
                __nv_bfloat16* dlogits, int ldd, double* acc, float dscale, cudaStream_t s, int* bad_label = nullptr);

// out[0..4] = loss_total, loss_category, loss_audio, accuracy_top1, accuracy_top5. `bad` (device int, optional): set by
// the two kernels above when a target index was outside its vocabulary -> the losses read NaN.
int finalize_metrics(const double* acc, float* out, float lambda_audio, int B, long long audio_rows, cudaStream_t s,
                     const int* bad = nullptr);

// x[i] *= *scale (device scalar): applies the upstream d(loss_total) to the stored logits gradients
int scale_bf16_by_device_scalar(__nv_bfloat16* x, long long n, const float* scale, cudaStream_t s);

}  // namespace svsr
