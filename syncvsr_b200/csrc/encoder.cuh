// Non-GEMM pieces of the x-transformers encoder sublayers (lightning.py:95-105,158):
// RMSNorm, rotary + softmax attention core for short sequences (n <= 64, dim_head = 64), GEGLU.
// The projections around them run on the tcgen05 GEMM (igemm.cu).
#pragma once
#include "common.cuh"

namespace svsr {

// y[M,D] (bf16) = x / clamp(||x||_2 * D^-1/2, eps) * g ; inv[M] = 1 / clamp(...)
// norm_dim (0 = D): the width the RMS is taken over when the row is zero-padded from norm_dim to D columns (the
// word-boundary variant runs dim 513 on a 576-column pitch; g is then a zero-padded copy)
// ctl (optional, common.cuh StepCtl): device-resident "this sublayer is dropped" predicate / dropout seed
int rmsnorm_fwd(const float* x, const float* g, __nv_bfloat16* y, float* inv, int M, int D, float eps, cudaStream_t s,
                int norm_dim = 0, const StepCtl* ctl = nullptr);
// dx[M,D] (fp32) += d rmsnorm ; dx_bf16 = bf16(dx) ; dg[D] += ...   (dy is the gradient wrt y, bf16)
int rmsnorm_bwd(const __nv_bfloat16* dy, const float* x, const float* g, const float* inv, float* dx,
                __nv_bfloat16* dx_bf16, float* dg, int M, int D, float eps, cudaStream_t s, int norm_dim = 0,
                const StepCtl* ctl = nullptr);  // dropped sublayer: dx_bf16 = bf16(dx), nothing else

// rotary cos/sin table for positions 0..n-1, 16 frequencies (rotary dim 32): tab[pos*32 + i] = cos, [pos*32+16+i] = sin
int rotary_table(float* tab, int n, cudaStream_t s);

// qkv [B*n, 3*heads*64] bf16 (q | k | v) -> o [B*n, heads*64] bf16; rotary on the first 32 dims of q, k and v
// drop_p: Attention(dropout=attn_dropout) on the softmax probabilities (counter-based mask, regenerated in backward)
int attention_fwd(const __nv_bfloat16* qkv, const float* rot, __nv_bfloat16* o, int B, int n, int heads,
                  int rotary_v, cudaStream_t s, float drop_p = 0.f, unsigned long long drop_seed = 0,
                  const StepCtl* ctl = nullptr);
int attention_bwd(const __nv_bfloat16* qkv, const float* rot, const __nv_bfloat16* d_o, __nv_bfloat16* dqkv, int B,
                  int n, int heads, int rotary_v, cudaStream_t s, float drop_p = 0.f, unsigned long long drop_seed = 0,
                  const StepCtl* ctl = nullptr);

// attention_tc.cu: the same two operators on the tensor cores (tcgen05 + TMEM; 128 / 32 (batch, head) pairs per UMMA
// tile); attention_fwd / attention_bwd dispatch to them unless SVSR_ATTN_TC=0.
int attention_tc_fwd(const __nv_bfloat16* qkv, const float* rot, __nv_bfloat16* o, int B, int n, int heads, int rotary_v,
                     cudaStream_t s, float drop_p, unsigned long long drop_seed, const StepCtl* ctl = nullptr);
int attention_tc_bwd(const __nv_bfloat16* qkv, const float* rot, const __nv_bfloat16* d_o, __nv_bfloat16* dqkv, int B, int n,
                     int heads, int rotary_v, cudaStream_t s, float drop_p, unsigned long long drop_seed,
                     const StepCtl* ctl = nullptr);

// attention_tc.cu: the FORWARD of the whole sublayer core in one kernel -- q | k | v projection of the normalised activations
// (TMA-fed tcgen05 GEMM, one head x four clips per CTA), rotary, softmax, PV. xn [B*n, ldx] bf16; w [3*heads*64, Kp] bf16
// K-major (Kp % 64 == 0 contracted columns); qkv [B*n, 3*heads*64] receives the projections (the backward kernel reads them).
int attention_qkv_tc_fwd(const __nv_bfloat16* xn, int ldx, const __nv_bfloat16* w, int Kp, const float* rot,
                         __nv_bfloat16* qkv, __nv_bfloat16* o, int B, int n, int heads, int rotary_v, cudaStream_t s,
                         float drop_p = 0.f, unsigned long long drop_seed = 0, const StepCtl* ctl = nullptr);

// u[M,F] = dropout_p(h[:, :F] * gelu(h[:, F:])) ; dh from du. The dropout mask is a counter-based function of
// (seed, element index): forward and backward regenerate the same mask, nothing is stored. p = 0 disables it.
int geglu_fwd(const __nv_bfloat16* h, __nv_bfloat16* u, int M, int F, float p, unsigned long long seed, cudaStream_t s,
              const StepCtl* ctl = nullptr);
int geglu_bwd(const __nv_bfloat16* h, const __nv_bfloat16* du, __nv_bfloat16* dh, int M, int F, float p,
              unsigned long long seed, cudaStream_t s, const StepCtl* ctl = nullptr);

// ---- HuggingFace BERT encoder pieces (`model.bert.type: huggingface`, lightning.py:90-92,152-156) ----
int gelu_fwd(const __nv_bfloat16* pre, __nv_bfloat16* h, long long n, cudaStream_t s);  // BertIntermediate: erf GELU
int gelu_bwd(const __nv_bfloat16* pre, const __nv_bfloat16* dh, __nv_bfloat16* dpre, long long n, cudaStream_t s);
// BertEmbeddings with inputs_embeds: E = x + position_embeddings[arange(L)] + token_type_embeddings[0]; and the
// gradients of the two tables (dpos[l] += sum over clips, dtt[0] += sum over all tokens)
int bert_embed_fwd(const float* x, const float* pos, const float* tt, float* E, long long M, int L, int D, cudaStream_t s);
int bert_embed_bwd(const float* dE, float* dpos, float* dtt, int B, int L, int D, cudaStream_t s);
// x = dropout(x) in place on fp32 (+ bf16 copy): nn.Dropout after the embedding LayerNorm, and its backward on dx
int dropout_f32_inplace(float* x, __nv_bfloat16* xb, long long n, float p, unsigned long long seed, cudaStream_t s,
                        const StepCtl* ctl = nullptr);

}  // namespace svsr
