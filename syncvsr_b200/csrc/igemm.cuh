// Implicit-GEMM on tcgen05 tensor cores (sm_100a): D[pixels, N] = sum_taps A_shift(tap)[pixels, Cin] * B[N, tap*Cin]^T
//   * A = NHWC bf16 activation tensor, fetched tap by tap with 4-D tiled TMA (out-of-bounds = zero padding,
//     element strides give strided convolutions) -- no im2col buffer ever exists.
//   * B = bf16 weight matrix [N_total, K_total], K contiguous, fetched with 2-D TMA.
//   * accumulators live in TMEM; a 4-warp epilogue applies bias / residual and stores NHWC bf16 or fp32.
// A dense GEMM is the degenerate case of one tap on a [M,1,1,K] "image".
#pragma once
#include "common.cuh"

namespace svsr {

constexpr int IGEMM_MAX_TAPS = 16;

// Fused cross-entropy epilogue of a dense GEMM whose output columns are A*G softmaxes of V classes per row (the audio
// head `audio_projection` -> reshape [B,T,A*G,V] -> F.cross_entropy, LRW/video/src/lightning.py:168-171; LRS twin
// e2e_asr_transformer.py:198-201). Row r of the GEMM is frame (b, t) = (r / T, r % T); columns [c*V, (c+1)*V) are the
// logits of target tokens[b*tok_stride_b + (t*A + a)*G + g], c = a*G + g. V % 64 == 0. The fp32 logits never leave the SM:
//   mode 1 (forward): every 64-column slot of a row leaves a (max, sum exp) partial in `part` [rows, cols/64] and the
//     chunk that holds the target column writes its logit to `xt` [rows, A*G]; ce_finalize() merges the partials into
//     `lse` [rows, A*G] and the loss sum. Nothing else is stored (IgemmProblem::out may be null).
//   mode 2 (backward): the tile is recomputed and the epilogue emits d logits = (exp(x - lse) - onehot) * dscale
//     (* *grad_scale when given) as the GEMM's bf16 output -- the operand of the input- and weight-gradient GEMMs.
struct IgemmCe {
  int mode = 0;
  int T = 1, A = 1, G = 1, V = 64, AG = 1;
  const long long* tokens = nullptr;
  long long tok_stride_b = 0;
  float2* part = nullptr;
  float* xt = nullptr;
  const float* lse = nullptr;
  int* bad_token = nullptr;          // set to 1 when a token is outside [0, V)
  float dscale = 1.f;
  const float* grad_scale = nullptr;  // optional device scalar (upstream d loss)
};

// BatchNorm-backward statistics fused into the epilogue of the input-gradient GEMM that PRODUCES the gradient flowing
// into the BatchNorm (timm BasicBlock bn1 / bn2 / downsample.1 via lightning.py:114-117): the launch's output is
// g = (acc + resid) * [activation mask] (the mask from IgemmProblem::relu_mask, or -- self_mask -- from the sign of BN 0's
// own output c*scale + shift, i.e. the ReLU that directly follows it), and per output channel the epilogue accumulates
//   stats[i][0..C)  += sum_pixels g          stats[i][C..2C) += sum_pixels g * xhat_i,   xhat_i = (c_i - mean_i) * invstd_i
// for up to two BatchNorms i that consume the same gradient (bn2 and downsample.1 of a strided block). The separate
// reduce pass over (gradient, c, mask) disappears; bn_bwd_finalize / bn_bwd_apply run unchanged on these sums.
struct IgemmBnBwd {
  int n = 0;                                    // 0 = off
  const void* c[2] = {nullptr, nullptr};        // conv outputs (bf16), same geometry / pitch as the GEMM output
  const float* coef[2] = {nullptr, nullptr};    // fp32 [4][C]: mean, invstd, scale, shift (bn_finalize)
  double* stats[2] = {nullptr, nullptr};        // fp64 [2][C] accumulators (+=)
  int self_mask = 0;
};

// Host-side problem description. All channel counts are in elements (bf16).
struct IgemmProblem {
  // ---- A operand (activations), NHWC ----
  const void* a = nullptr;
  int a_N = 0, a_H = 1, a_W = 1;  // images, rows, cols of the tensor A lives in
  int a_C = 0;                    // channels per pixel as laid out in memory (pixel pitch)
  int a_coff = 0;                 // first contracted channel
  int cin = 0;                    // contracted channels per tap (multiple of 64)
  int stride = 1;                 // input coordinate = output coordinate * stride + tap offset
  // ---- taps ----
  int ntaps = 1;
  int tap_dh[IGEMM_MAX_TAPS] = {0};
  int tap_dw[IGEMM_MAX_TAPS] = {0};
  int tap_kbase[IGEMM_MAX_TAPS] = {0};  // column of B where this tap's cin block starts
  // ---- logical output grid that the M dimension enumerates ----
  int o_N = 0, OH = 1, OW = 1;
  // ---- B operand ----
  const void* b = nullptr;
  int b_rows = 0;  // N_total (valid output columns)
  int b_cols = 0;  // K_total (row pitch of B, multiple of 8)
  // ---- output tensor: pixel (n, oh*o_sh+o_oh, ow*o_sw+o_ow) of a [o_N, o_H, o_W, ldc] tensor ----
  void* out = nullptr;
  int out_fp32 = 0;
  int ldc = 0;    // elements per output pixel (pitch)
  int c_off = 0;  // first output channel
  int o_H = 1, o_W = 1, o_sh = 1, o_sw = 1, o_oh = 0, o_ow = 0;
  // ---- epilogue ----
  const float* bias = nullptr;  // [b_rows] fp32 or null
  const void* resid = nullptr;  // same geometry as out (pitch ldc, offset c_off), or null
  int resid_fp32 = 0;
  float alpha = 1.0f;  // out = act(alpha*acc (+bias_scale*bias) (+resid)) (* [relu_mask > 0])
  float bias_scale = 1.0f;
  int relu = 0;                     // apply ReLU last (Conformer FFN w_1: positionwise_feed_forward.py:28-30)
  const void* relu_mask = nullptr;  // bf16, same geometry as out: zero the result where mask <= 0 (ReLU backward fused
                                    // into the input-gradient GEMM of the following Linear)
  // Dropout on the branch value alpha*acc + bias_scale*bias BEFORE the residual is added (EncoderLayer /
  // DecoderLayer `residual + dropout(sublayer(x))`, encoder_layer.py:94-137; also commutes with the ReLU of the FFN's
  // w_1, positionwise_feed_forward.py:30). Mask = dropout_keep(drop_seed, element index in the output tensor).
  float drop_p = 0.f;
  unsigned long long drop_seed = 0;
  // optional fused BatchNorm statistics: fp64 [2][b_rows] accumulators (+=): per-output-channel sum and sum of
  // squares of the fp32 accumulators over all valid pixels (train-mode BN of the conv output, lightning.py:51)
  double* bn_stats = nullptr;
  double algo_flops = 0;  // algorithmic FLOPs of this launch for the profiler (0 = 2*pixels*N*taps*cin)
  IgemmCe ce;             // fused cross-entropy epilogue (dense GEMMs only)
  IgemmBnBwd bnb;         // fused BatchNorm-backward statistics (bf16 outputs with a multiple of 64 channels)
  StepCtl ctl;            // device-resident skip predicate (common.cuh): the launch returns at once when its bit is set
};

int igemm_launch(const IgemmProblem& p, cudaStream_t stream);

// igemm_halo.cu: single-load halo-tile kernel for 3x3 / stride-1 / 64 -> 64 channel problems (resnet.layer1); igemm_launch()
// dispatches to it when igemm_halo_matches() (SVSR_HALO_CONV=0 in the environment keeps the generic kernel).
bool igemm_halo_matches(const IgemmProblem& p);
int igemm_halo_launch(const IgemmProblem& p, cudaStream_t stream);

// igemm_stem.cu: temporal-halo kernel for the 3-D conv stem's 5-tap / 64 -> 64 column contraction over [clips, frames,
// pixels, 64] patch rows (one activation load per 8-frame x 16-pixel tile, resident weights); igemm_launch() dispatches
// to it when igemm_stem_matches() (SVSR_STEM_HALO=0 in the environment keeps the generic kernel).
bool igemm_stem_matches(const IgemmProblem& p);
int igemm_stem_launch(const IgemmProblem& p, cudaStream_t stream);

// Chooses the pixel box (bn, bh, bw) with bn*bh*bw <= 128 that wastes the fewest MMA rows.
void igemm_choose_box(int o_N, int OH, int OW, int* bn, int* bh, int* bw);

}  // namespace svsr
