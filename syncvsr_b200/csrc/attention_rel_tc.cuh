// Tensor-core (tcgen05 + TMEM) version of the d_k = 64 attention core of the LRS path: Conformer relative-position
// self-attention (espnet/nets/pytorch_backend/transformer/attention.py:192-278), decoder causal self-attention and
// source attention (attention.py:38-118), HF-BERT self-attention. conformer.cu's attention_core_fwd / _bwd dispatch
// here unless SVSR_ATTN_TC=0 (the CUDA-core fp32 kernels stay as the A/B comparator).
#pragma once
#include "common.cuh"

namespace svsr {

// flattened AttnProblem + AttnGrads (conformer.cuh) as the kernels take it
struct AttnK {
  const __nv_bfloat16 *q, *k, *v, *p;
  int ldq, ldk, ldv, ldp;
  const float *bu, *bv;
  const int* klen;
  int causal, B, H, Tq, Tk;
  float scale;
  __nv_bfloat16* o;
  int ldo;
  float* lse;
  // backward
  const __nv_bfloat16* d_o;
  __nv_bfloat16 *dq, *dk, *dv;
  int lddq, lddk, lddv;
  float *dp, *dbu, *dbv;
  float *Pg, *DSg;  // CUDA-core path: fp32 scratch [B,H,Tq,Tk] each
  float drop_p;     // dropout on the attention probabilities (attention.py:81)
  unsigned long long drop_seed;
  const unsigned long long* seed_base;  // optional device word added to drop_seed (graph-replayed steps)
};

int attention_rel_tc_fwd(const AttnK& k, cudaStream_t s);
// scratch: attention_rel_tc_scratch_bytes(B, H, Tq, Tk) bytes (bf16 [2][B,H,Tq,roundup(Tk,64)]: dropped probabilities
// and score gradients handed from the query-side kernel to the key/value-side and relative-position-side kernels)
int attention_rel_tc_bwd(const AttnK& k, void* scratch, cudaStream_t s);
size_t attention_rel_tc_scratch_bytes(int B, int H, int Tq, int Tk);

}  // namespace svsr
