// fp32-class "parity mode" of the forward pass. Activations stay fp32 in HBM; every tensor-core operand is split into
// two bf16 terms x = hi + lo and the contraction is K-concatenated as [hi | lo | hi] . [hi | hi | lo]^T
// (= hi*hi + lo*hi + hi*lo, ~16 mantissa bits, fp32 accumulation) so the SAME tcgen05 implicit-GEMM kernel computes it
// with three times the channels. Element-wise kernels here are the fp32-I/O twins of elementwise.cu / encoder.cu.
// Forward only: this mode exists to demonstrate north_star's 1e-3 output tolerance, not for throughput.
#pragma once
#include "common.cuh"

namespace svsr {

// x fp32 [rows, C] at pitch ldx -> bf16 [rows, 3 Cp] = hi | lo | hi (Cp >= C, a multiple of 8; columns C..Cp are zero)
int split3_f32(const float* x, int ldx, __nv_bfloat16* out, long long rows, int C, int Cp, cudaStream_t s);
// conv weight fp32 [Cout,Cin,R,S] -> bf16 [Cout, R*S*3*Cin], per tap [hi(Cin) | hi(Cin) | lo(Cin)]
int pack_conv_weight_split(const float* w, __nv_bfloat16* out, int Cout, int Cin, int RS, cudaStream_t s);
// linear weight fp32 [N,K] -> bf16 [N, 3 Kp] = [hi | hi | lo], columns K..Kp of every third zero
int pack_linear_weight_split(const float* w, __nv_bfloat16* out, int N, int K, int Kp, cudaStream_t s);
// stem weight fp32 [64,1,5,7,7] -> bf16 [64, 5*192], per temporal tap [hi(64 slots) | hi | lo], slot = kh*8+kw
int pack_stem_weight_split(const float* w, __nv_bfloat16* out, cudaStream_t s);

int stem_patch_f32(const float* videos, float* patches, int B, int T, int H, int W, cudaStream_t s);
int bn_apply_f32(const float* x, const float* coef, const float* res, const float* rcoef, int relu, float* out,
                 long long rows, int C, cudaStream_t s);
int stem_bn_gelu_pool_f32(const float* y0, const float* coef, float* out, int N, int IH, int IW, cudaStream_t s);
// x_stream: fp32 [B, T+1, C] at row pitch ld (row 0 of a clip = cls[0..C))
int meanpool_cls_f32(const float* a, const float* cls, float* x_stream, int B, int T, int HW, int C, int ld,
                     cudaStream_t s);
int rmsnorm_fwd_f32(const float* x, const float* g, float* y, int M, int D, int ld, float eps, cudaStream_t s);
// qkv fp32 [B*n, 3*heads*64] = q | k | v; rot: rotary table or NULL (HuggingFace BERT: absolute positions, no rotary)
int attention_fwd_f32(const float* qkv, const float* rot, float* o, int B, int n, int heads, int rotary_v,
                      cudaStream_t s);
int layernorm_fwd_f32(const float* x, const float* g, const float* b, float* y, int M, int D, float eps, cudaStream_t s);
int gelu_fwd_f32(const float* x, float* y, long long n, cudaStream_t s);
int geglu_fwd_f32(const float* h, float* u, int M, int F, cudaStream_t s);
// last [B,T+1,D] at row pitch ld -> cls [B,D], frames [B*T,D] (dense)
int split_last_f32(const float* last, int ld, float* cls, float* frames, int B, int T, int D, cudaStream_t s);

}  // namespace svsr
