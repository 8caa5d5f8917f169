// Native step executor for the LRS sentence-level model (reference: E2E.forward,
// LRS/video/espnet/nets/pytorch_backend/e2e_asr_transformer.py:186-227): Conv3dResNet frontend (Swish) ->
// Linear embed + relative positional encoding -> macaron Conformer blocks (rel-pos MHA, convolution module) ->
// after_norm -> audio-token cross-entropy + CTC + attention decoder with label-smoothing loss, and the backward of
// all of it. Same arena / workspace conventions as the LRW engine (engine.cu); parameter names are the reference's
// state-dict keys. Dropout (dropout_rate on every sub-block output, FFN hidden, positional encodings, CTC input;
// transformer_attn_dropout_rate on attention probabilities) uses counter-based masks regenerated in backward.
#include "engine_common.cuh"
#include "conformer.cuh"
#include "heads.cuh"

namespace svsr {

struct LnRef {
  long long g, b;  // param arena offsets
  size_t stats;    // fp32 [rows, 2]
};

struct ConfLayerRef {
  LnRef n_mac, n_mha, n_conv, n_ff, n_fin;
  LinRef mac1, mac2, qkv, out, pos, pw1, pw2, ff1, ff2;
  long long bias_u, bias_v, dw_w, dw_b;
  size_t dw_wT = 0;  // fp32 [K, D] transposed copy of the depthwise kernel (coalesced per-tap loads)
  BnRef bn;
  size_t yn[4], h_mac, h_ff, qkvbuf, pbuf, lse, ctx, hpw1, u, dwo, act;
};

struct DecLayerRef {
  LnRef n1, n2, n3;
  LinRef qkv, out_s, q_c, kv_c, out_c, ff1, ff2;
  size_t yn[3], qkvbuf, lse_s, ctx_s, qc, kvc, lse_c, ctx_c, h;
};

struct LrsEngine : EngineBase {
  svsr_lrs_config cfg;
  int M = 0;       // encoder rows = B*T
  int Md_max = 0;  // decoder rows at Lmax
  int ldv = 0;     // padded pitch of vocabulary-sized logits
  int AGV = 0;
  Frontend fe;
  LinRef embed, aud, ctc, outl;
  long long dec_emb = 0;
  LnRef after, dec_after;
  std::vector<ConfLayerRef> enc;
  std::vector<DecLayerRef> dec;

  size_t feats, pe_rel, klen, xs, enc_f32, enc_b, logits_a, dlogits_a, logits_c, dlogits_c, ctc_scratch;
  size_t ce_part = 0, ce_xt = 0, ce_lse = 0, ce_tok = 0;  // fused audio head (igemm.cuh IgemmCe)
  // audio_classifier + unflatten + log-softmax + NLL in the GEMM epilogue: softmax width % 64 == 0
  bool fused_head() const { return AGV >= 128 && cfg.audio_vocab % 64 == 0; }
  size_t ys_in, ys_out, xd, dec_yn, pred, dpred, acc, bad_token, pe_drop, enc_ctc;
  size_t bn_stats_arena = 0, bn_stats_bytes = 0;
  // backward scratch
  // gradient temporaries read by weight-gradient GEMMs on the side stream are double buffered by unit parity
  size_t dx, dxb[2], gF[2], g3D[2], gD[2][3], attn_scratch, dp, dpb[2], dfeat, ddx, ddxb[2], dkv[2];
  size_t pack_jobs;
  int n_pack_jobs = 0;
  bool pack_table_ready = false;
  int last_L = 0, last_Llab = 0, last_train = 0, last_audio = 0;
  int bwd_stage = 0;
  unsigned long long last_seed = 0;
  float pd = 0.f, pa = 0.f;  // dropout probabilities in effect for the last forward (0 in eval mode)
  // per-site seeds: every Dropout module instance of the reference draws an independent mask
  // Device-resident step seed (svsr_lrs_step_control): with dev_ctl on, the step seed lives in the workspace word `ctl`
  // and every dropout kernel adds it to its site constant on the device (seed_base()), so the launch arguments no longer
  // depend on host RNG and ONE captured CUDA graph replays every step of the shipped dropout config. Same masks as the
  // host-valued path for the same seed.
  size_t ctl = 0;
  bool dev_ctl = false;
  unsigned long long site(int id) const {
    return (dev_ctl ? 0ULL : last_seed) + 0x632BE59BD9B4E019ULL * (unsigned long long)(id + 1);
  }
  const unsigned long long* seed_base() const { return dev_ctl ? ws<unsigned long long>(ctl) : nullptr; }
  StepCtl seed_ctl() const {
    StepCtl c;
    c.seed = seed_base();
    return c;
  }
  static int enc_site(int layer, int k) { return 16 * (layer + 1) + k; }        // k: 0 mac hidden, 1 mac out, 2 attn
  static int dec_site(int layer, int k) { return 16 * (layer + 65) + k; }       // probs, 3 attn out, 4 conv out, ...
  bool fwd_done = false;

  float* xs_buf(int i) const { return ws<float>(xs) + (size_t)i * M * cfg.adim; }
  float* xd_buf(int i) const { return ws<float>(xd) + (size_t)i * Md_max * cfg.adim; }
};

namespace {

void add_ln(EngineBase& e, LnRef& n, const std::string& prefix, int D, int rows, Bump& b) {
  n.g = add_param(e.params, e.pc, prefix + ".weight", {D});
  n.b = add_param(e.params, e.pc, prefix + ".bias", {D});
  n.stats = b.take((size_t)rows * 2 * sizeof(float));
}
// Linear [N,K] (or pointwise Conv1d [N,K,1]: same bytes) with packed operand copies
void add_lin(EngineBase& e, LinRef& l, const std::string& prefix, int N, int K, bool bias, bool conv1d, Bump& b) {
  l.N = N, l.K = K;
  if (conv1d)
    l.w = add_param(e.params, e.pc, prefix + ".weight", {N, K, 1});
  else
    l.w = add_param(e.params, e.pc, prefix + ".weight", {N, K});
  l.b = bias ? add_param(e.params, e.pc, prefix + ".bias", {N}) : -1;
  l.ldt = (N + 63) / 64 * 64;
  l.wb = b.take((size_t)N * K * 2);
  l.wt = b.take((size_t)K * l.ldt * 2);
}
// several Linear layers with identical K registered back to back so that weights (and biases) are adjacent in their
// arena regions and run as ONE GEMM: names[i] are module prefixes; `from`..`to` select the fused sub-range.
void add_fused(EngineBase& e, const std::vector<std::string>& prefixes, int n_each, int K, std::vector<long long>& w_off,
               std::vector<long long>& b_off) {
  w_off.clear(), b_off.clear();
  for (auto& p : prefixes) w_off.push_back(add_param(e.params, e.pc, p + ".weight", {n_each, K}));
  for (auto& p : prefixes) b_off.push_back(add_param(e.params, e.pc, p + ".bias", {n_each}));
}
void fused_ref(LinRef& l, long long w, long long bias, int N, int K, Bump& b) {
  l.N = N, l.K = K, l.w = w, l.b = bias;
  l.ldt = (N + 63) / 64 * 64;
  l.wb = b.take((size_t)N * K * 2);
  l.wt = b.take((size_t)K * l.ldt * 2);
}

int lin_fwd(const EngineBase& e, const bf16* x, int rows, const LinRef& l, void* out, int ldc, int out_fp32,
            const void* resid, float alpha, int relu, cudaStream_t s, float drop_p = 0.f,
            unsigned long long drop_seed = 0, const unsigned long long* seed_base = nullptr) {
  IgemmProblem p;
  p.a = x, p.a_N = rows, p.a_C = l.K, p.cin = l.K, p.ntaps = 1;
  p.o_N = rows;
  p.b = e.ws<bf16>(l.wb), p.b_rows = l.N, p.b_cols = l.K;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = ldc;
  p.bias = l.b >= 0 ? e.P + l.b : nullptr;
  p.resid = resid, p.resid_fp32 = 1;
  p.alpha = alpha, p.bias_scale = alpha, p.relu = relu;
  p.drop_p = drop_p, p.drop_seed = drop_seed;
  p.ctl.seed = seed_base;  // the epilogue's mask seed = drop_seed + *seed_base (common.cuh ctl_seed)
  return igemm_launch(p, s);
}
// out[rows, K] = alpha * dy[rows, N (pitch ldy)] . W (+ resid fp32) (* [relu_mask > 0])
int lin_dgrad(const EngineBase& e, const bf16* dy, int ldy, int rows, const LinRef& l, void* out, int ldc, int out_fp32,
              const void* resid, float alpha, const bf16* relu_mask, cudaStream_t s) {
  IgemmProblem p;
  p.a = dy, p.a_N = rows, p.a_C = ldy, p.cin = l.ldt, p.ntaps = 1;
  p.o_N = rows;
  p.b = e.ws<bf16>(l.wt), p.b_rows = l.K, p.b_cols = l.ldt;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = ldc;
  p.resid = resid, p.resid_fp32 = 1;
  p.alpha = alpha, p.relu_mask = relu_mask;
  return igemm_launch(p, s);
}

// audio_classifier fused with unflatten + log-softmax + NLL (e2e_asr_transformer.py:198-201): mode 1 = forward, mode 2 =
// backward (recompute, emit d logits bf16). Tokens were copied to the workspace ([B, T*A, G] contiguous) by the forward.
int lrs_audio_head_gemm(const LrsEngine& e, int mode, const float* grad_scale, cudaStream_t s) {
  const svsr_lrs_config& c = e.cfg;
  const LinRef& l = e.aud;
  const long long audio_rows = (long long)e.M * c.audio_alignment * c.vq_groups;
  IgemmProblem p;
  p.a = e.ws<bf16>(e.enc_b), p.a_N = e.M, p.a_C = l.K, p.cin = l.K, p.ntaps = 1;
  p.o_N = e.M;
  p.b = e.ws<bf16>(l.wb), p.b_rows = l.N, p.b_cols = l.K;
  p.bias = l.b >= 0 ? e.P + l.b : nullptr;
  p.ldc = e.AGV;
  p.out = mode == 2 ? e.ws<bf16>(e.dlogits_a) : nullptr;
  p.ce.mode = mode;
  p.ce.T = c.T, p.ce.A = c.audio_alignment, p.ce.G = c.vq_groups, p.ce.V = c.audio_vocab;
  p.ce.AG = c.audio_alignment * c.vq_groups;
  p.ce.tokens = e.ws<long long>(e.ce_tok), p.ce.tok_stride_b = (long long)c.T * c.audio_alignment * c.vq_groups;
  p.ce.part = e.ws<float2>(e.ce_part), p.ce.xt = e.ws<float>(e.ce_xt), p.ce.lse = e.ws<float>(e.ce_lse);
  p.ce.bad_token = e.ws<int>(e.bad_token);
  p.ce.dscale = c.audio_weight / (float)audio_rows, p.ce.grad_scale = grad_scale;
  return igemm_launch(p, s);
}

int ln_fwd(const LrsEngine& e, const float* x, const LnRef& n, bf16* yb, float* yf, int rows, cudaStream_t s) {
  return layernorm_fwd(x, e.P + n.g, e.P + n.b, yb, yf, e.ws<float>(n.stats), rows, e.cfg.adim, 1e-12f, s);
}
int ln_bwd(const LrsEngine& e, const bf16* dyb, const float* dyf, const float* x, const LnRef& n, float* dx, int accumulate,
           int rows, cudaStream_t s) {
  return layernorm_bwd(dyb, dyf, x, e.P + n.g, e.ws<float>(n.stats), dx, accumulate, e.G + n.g, e.G + n.b, rows,
                       e.cfg.adim, s);
}

__global__ void add_sos_eos_kernel(const long long* __restrict__ label, int Llab, long long* __restrict__ ys_in,
                                   long long* __restrict__ ys_out, int B, int sos, int eos) {
  // add_sos_eos.py:12-31 on a -1 padded label matrix: ys_in = [sos, y..., eos pad], ys_out = [y..., eos, -1 pad]
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int L = Llab + 1;
  int n = 0;
  ys_in[(long long)b * L] = sos;
  for (int i = 0; i < Llab; ++i) {
    const long long y = label[(long long)b * Llab + i];
    if (y >= 0) {
      ys_in[(long long)b * L + 1 + n] = y;
      ys_out[(long long)b * L + n] = y;
      ++n;
    }
  }
  ys_out[(long long)b * L + n] = eos;
  for (int i = n + 1; i < L; ++i) ys_in[(long long)b * L + i] = eos, ys_out[(long long)b * L + i] = -1;
}

}  // namespace

static int lrs_build(LrsEngine& e, long long nodecay_base) {
  e.params.clear(), e.buffers.clear();
  e.pc = ArenaCount(), e.bc = ArenaCount();
  e.pc.nodecay_base = nodecay_base;
  const svsr_lrs_config& c = e.cfg;
  SVSR_REQUIRE(c.B > 0 && c.T > 0 && c.H > 0 && c.W == c.H, "lrs: bad clip geometry B=%d T=%d H=%d W=%d", c.B, c.T, c.H,
               c.W);
  SVSR_REQUIRE(c.adim % 128 == 0 && c.adim <= 1024 && c.aheads * 64 == c.adim,
               "lrs: adim=%d must be a multiple of 128 (<= 1024) with heads of 64 (aheads=%d)", c.adim, c.aheads);
  SVSR_REQUIRE(c.eunits % 64 == 0 && c.dunits % 64 == 0, "lrs: eunits/dunits must be multiples of 64");
  SVSR_REQUIRE(c.elayers >= 1 && c.dlayers >= 1, "lrs: elayers/dlayers must be >= 1");
  SVSR_REQUIRE(c.cnn_kernel % 2 == 1 && c.cnn_kernel <= 31, "lrs: cnn_module_kernel=%d unsupported", c.cnn_kernel);
  SVSR_REQUIRE(c.odim >= 3 && c.Lmax >= 2, "lrs: odim=%d Lmax=%d", c.odim, c.Lmax);
  e.AGV = c.audio_alignment * c.vq_groups * c.audio_vocab;
  SVSR_REQUIRE(e.AGV % 64 == 0, "lrs: audio logits per frame (%d) must be a multiple of 64", e.AGV);
  SVSR_REQUIRE(c.dropout_rate >= 0.f && c.dropout_rate < 1.f && c.attn_dropout_rate >= 0.f && c.attn_dropout_rate < 1.f,
               "lrs: dropout rates must be in [0,1)");
  const int D = c.adim, F = c.eunits, Fd = c.dunits, H = c.aheads, T = c.T;
  e.M = c.B * c.T;
  e.Md_max = c.B * c.Lmax;
  e.ldv = (c.odim + 63) / 64 * 64;
  const int M = e.M, Md = e.Md_max;
  Bump b;

  // ---- parameters (reference state-dict names) ----
  e.fe.B = c.B, e.fe.T = c.T, e.fe.H = c.H, e.fe.swish = 1;
  e.bn_eps = c.bn_eps, e.bn_momentum = c.bn_momentum;
  RC(frontend_build(e, e.fe, "encoder.frontend.frontend3D.0.weight", "encoder.frontend.frontend3D.1",
                    "encoder.frontend.trunk", b));
  add_lin(e, e.embed, "encoder.embed.0", D, 512, true, false, b);
  e.enc.resize(c.elayers);
  std::vector<long long> wo, bo;
  for (int i = 0; i < c.elayers; ++i) {
    ConfLayerRef& L = e.enc[i];
    const std::string pre = "encoder.encoders." + std::to_string(i);
    L.bias_u = add_param(e.params, e.pc, pre + ".self_attn.pos_bias_u", {H, 64});
    L.bias_v = add_param(e.params, e.pc, pre + ".self_attn.pos_bias_v", {H, 64});
    add_fused(e, {pre + ".self_attn.linear_q", pre + ".self_attn.linear_k", pre + ".self_attn.linear_v"}, D, D, wo, bo);
    fused_ref(L.qkv, wo[0], bo[0], 3 * D, D, b);
    add_lin(e, L.out, pre + ".self_attn.linear_out", D, D, true, false, b);
    add_lin(e, L.pos, pre + ".self_attn.linear_pos", D, D, false, false, b);
    add_lin(e, L.ff1, pre + ".feed_forward.w_1", F, D, true, false, b);
    add_lin(e, L.ff2, pre + ".feed_forward.w_2", D, F, true, false, b);
    add_lin(e, L.pw1, pre + ".conv_module.pointwise_cov1", 2 * D, D, true, true, b);
    L.dw_w = add_param(e.params, e.pc, pre + ".conv_module.depthwise_conv.weight", {D, 1, c.cnn_kernel});
    L.dw_b = add_param(e.params, e.pc, pre + ".conv_module.depthwise_conv.bias", {D});
    add_bn(e, L.bn, pre + ".conv_module.norm", D, b);
    add_lin(e, L.pw2, pre + ".conv_module.pointwise_cov2", D, D, true, true, b);
    add_ln(e, L.n_ff, pre + ".norm_ff", D, M, b);
    add_ln(e, L.n_mha, pre + ".norm_mha", D, M, b);
    add_lin(e, L.mac1, pre + ".feed_forward_macaron.w_1", F, D, true, false, b);
    add_lin(e, L.mac2, pre + ".feed_forward_macaron.w_2", D, F, true, false, b);
    add_ln(e, L.n_mac, pre + ".norm_ff_macaron", D, M, b);
    add_ln(e, L.n_conv, pre + ".norm_conv", D, M, b);
    add_ln(e, L.n_fin, pre + ".norm_final", D, M, b);
    for (int k = 0; k < 4; ++k) L.yn[k] = b.take((size_t)M * D * 2);
    L.h_mac = b.take((size_t)M * F * 2), L.h_ff = b.take((size_t)M * F * 2);
    L.qkvbuf = b.take((size_t)M * 3 * D * 2);
    L.pbuf = b.take((size_t)(2 * T - 1) * D * 2);
    L.lse = b.take((size_t)c.B * H * T * 4);
    L.ctx = b.take((size_t)M * D * 2);
    L.hpw1 = b.take((size_t)M * 2 * D * 2);
    L.u = b.take((size_t)M * D * 2), L.dwo = b.take((size_t)M * D * 2), L.act = b.take((size_t)M * D * 2);
    L.dw_wT = b.take((size_t)c.cnn_kernel * D * 4);
  }
  add_ln(e, e.after, "encoder.after_norm", D, M, b);
  e.dec_emb = add_param(e.params, e.pc, "decoder.embed.0.weight", {c.odim, D});
  e.dec.resize(c.dlayers);
  for (int i = 0; i < c.dlayers; ++i) {
    DecLayerRef& L = e.dec[i];
    const std::string pre = "decoder.decoders." + std::to_string(i);
    add_fused(e, {pre + ".self_attn.linear_q", pre + ".self_attn.linear_k", pre + ".self_attn.linear_v"}, D, D, wo, bo);
    fused_ref(L.qkv, wo[0], bo[0], 3 * D, D, b);
    add_lin(e, L.out_s, pre + ".self_attn.linear_out", D, D, true, false, b);
    add_fused(e, {pre + ".src_attn.linear_q", pre + ".src_attn.linear_k", pre + ".src_attn.linear_v"}, D, D, wo, bo);
    fused_ref(L.q_c, wo[0], bo[0], D, D, b);
    fused_ref(L.kv_c, wo[1], bo[1], 2 * D, D, b);
    add_lin(e, L.out_c, pre + ".src_attn.linear_out", D, D, true, false, b);
    add_lin(e, L.ff1, pre + ".feed_forward.w_1", Fd, D, true, false, b);
    add_lin(e, L.ff2, pre + ".feed_forward.w_2", D, Fd, true, false, b);
    add_ln(e, L.n1, pre + ".norm1", D, Md, b);
    add_ln(e, L.n2, pre + ".norm2", D, Md, b);
    add_ln(e, L.n3, pre + ".norm3", D, Md, b);
    for (int k = 0; k < 3; ++k) L.yn[k] = b.take((size_t)Md * D * 2);
    L.qkvbuf = b.take((size_t)Md * 3 * D * 2);
    L.lse_s = b.take((size_t)c.B * H * c.Lmax * 4), L.lse_c = b.take((size_t)c.B * H * c.Lmax * 4);
    L.ctx_s = b.take((size_t)Md * D * 2), L.ctx_c = b.take((size_t)Md * D * 2);
    L.qc = b.take((size_t)Md * D * 2);
    L.kvc = b.take((size_t)M * 2 * D * 2);
    L.h = b.take((size_t)Md * Fd * 2);
  }
  add_ln(e, e.dec_after, "decoder.after_norm", D, Md, b);
  add_lin(e, e.outl, "decoder.output_layer", c.odim, D, true, false, b);
  add_lin(e, e.ctc, "ctc.ctc_lo", c.odim, D, true, false, b);
  if (e.AGV > 0) add_lin(e, e.aud, "audio_classifier", e.AGV, D, true, false, b);

  // ---- activations ----
  frontend_alloc(e, e.fe, b);
  {  // fp64 BatchNorm1d statistic slots (forward + backward) of the convolution modules
    Bump sb;
    for (auto& L : e.enc) {
      L.bn.stats_f = sb.take(2 * D * sizeof(double));
      L.bn.stats_b = sb.take(2 * D * sizeof(double));
    }
    e.bn_stats_bytes = sb.off;
    e.bn_stats_arena = b.take(sb.off);
    for (auto& L : e.enc) L.bn.stats_f += e.bn_stats_arena, L.bn.stats_b += e.bn_stats_arena;
  }
  e.feats = b.take((size_t)M * 512 * 2);
  e.pe_rel = b.take((size_t)(2 * T - 1) * D * 2);
  e.klen = b.take((size_t)c.B * sizeof(int));
  e.ctl = b.take(16);
  e.xs = b.take((size_t)(5 * c.elayers + 1) * M * D * 4);
  e.enc_f32 = b.take((size_t)M * D * 4);
  e.enc_b = b.take((size_t)M * D * 2);
  const size_t agv = e.AGV > 0 ? e.AGV : 64;
  e.logits_a = b.take((size_t)M * agv * 4);
  e.dlogits_a = b.take((size_t)M * agv * 2);
  {
    const size_t ag = (size_t)(c.audio_alignment * c.vq_groups > 0 ? c.audio_alignment * c.vq_groups : 1);
    e.ce_part = b.take((size_t)M * (agv / 64) * sizeof(float2));
    e.ce_xt = b.take((size_t)M * ag * 4), e.ce_lse = b.take((size_t)M * ag * 4), e.ce_tok = b.take((size_t)M * ag * 8);
  }
  e.logits_c = b.take((size_t)M * e.ldv * 4);
  e.dlogits_c = b.take((size_t)M * e.ldv * 2);
  e.ctc_scratch = b.take(ctc_scratch_bytes(c.B, T, c.Lmax));
  e.ys_in = b.take((size_t)Md * 8), e.ys_out = b.take((size_t)Md * 8);
  e.xd = b.take((size_t)(3 * c.dlayers + 1) * Md * D * 4);
  e.dec_yn = b.take((size_t)Md * D * 2);
  e.pred = b.take((size_t)Md * e.ldv * 4);
  e.dpred = b.take((size_t)Md * e.ldv * 2);
  e.acc = b.take(8 * sizeof(double));
  e.bad_token = b.take(sizeof(int));
  e.pe_drop = b.take((size_t)(2 * T - 1) * D * 2);
  e.enc_ctc = b.take((size_t)M * D * 2);
  // ---- backward scratch ----
  const size_t R = (size_t)(M > Md ? M : Md);
  const size_t Fm = (size_t)(F > Fd ? F : Fd);
  e.dx = b.take((size_t)M * D * 4);
  for (int u = 0; u < 2; ++u) {
    e.dxb[u] = b.take(R * D * 2);
    e.gF[u] = b.take(R * Fm * 2);
    e.g3D[u] = b.take(R * 3 * D * 2);
    for (int i = 0; i < 3; ++i) e.gD[u][i] = b.take(R * D * 2);
    e.dpb[u] = b.take((size_t)(2 * T - 1) * D * 2);
    e.ddxb[u] = b.take((size_t)Md * D * 2);
    e.dkv[u] = b.take((size_t)M * 2 * D * 2);
  }
  const int Tmax = T > c.Lmax ? T : c.Lmax;
  e.attn_scratch = b.take(attention_scratch_bytes(c.B, H, Tmax, Tmax));
  e.dp = b.take((size_t)(2 * T - 1) * D * 4);
  e.dfeat = b.take((size_t)M * 512 * 2);
  e.ddx = b.take((size_t)Md * D * 4);
  e.pack_jobs = b.take(384 * sizeof(PackJob));
  e.ws_bytes = b.off;
  e.decay_count = e.pc.decay;
  e.param_count = e.pc.nodecay_base + e.pc.nodecay;
  e.buffer_count = e.bc.nodecay;
  return SVSR_OK;
}

static int lrs_pack(LrsEngine& e, cudaStream_t s) {
  if (!e.pack_table_ready) {
    std::vector<PackJob> jobs;
    std::vector<const LinRef*> padded;
    frontend_pack_jobs(e, e.fe, jobs);
    auto lin = [&](const LinRef& l) {
      jobs.push_back({e.P + l.w, e.ws<bf16>(l.wb), e.ws<bf16>(l.wt), 1, l.N, l.K, l.K, l.ldt});
      if (l.ldt != l.N) padded.push_back(&l);
    };
    lin(e.embed);
    for (auto& L : e.enc) lin(L.mac1), lin(L.mac2), lin(L.qkv), lin(L.out), lin(L.pos), lin(L.pw1), lin(L.pw2), lin(L.ff1), lin(L.ff2);
    for (auto& L : e.dec) lin(L.qkv), lin(L.out_s), lin(L.q_c), lin(L.kv_c), lin(L.out_c), lin(L.ff1), lin(L.ff2);
    lin(e.outl), lin(e.ctc);
    if (e.AGV > 0) lin(e.aud);
    SVSR_REQUIRE(jobs.size() <= 384, "lrs: too many pack jobs (%zu)", jobs.size());
    e.n_pack_jobs = (int)jobs.size();
    SVSR_CHECK_CUDA(cudaMemcpyAsync(e.ws<PackJob>(e.pack_jobs), jobs.data(), jobs.size() * sizeof(PackJob),
                                    cudaMemcpyHostToDevice, s));
    SVSR_CHECK_CUDA(cudaStreamSynchronize(s));
    for (const LinRef* l : padded)  // padding columns of transposed operands stay zero forever
      SVSR_CHECK_CUDA(cudaMemsetAsync(e.ws<bf16>(l->wt), 0, (size_t)l->K * l->ldt * 2, s));
    RC(rel_pos_table(e.ws<bf16>(e.pe_rel), e.cfg.T, e.cfg.adim, s));
    e.pack_table_ready = true;
  }
  for (auto& L : e.enc) RC(transpose_f32(e.P + L.dw_w, e.ws<float>(L.dw_wT), e.cfg.adim, e.cfg.cnn_kernel, s));
  return pack_all_weights(e.ws<PackJob>(e.pack_jobs), e.n_pack_jobs, s);
}

// ------------------------------------------------------------------------------------------------
// encoder only (Encoder.forward, transformer/encoder.py:257-289): fills enc_f32 / enc_b
// ------------------------------------------------------------------------------------------------
static int lrs_encoder_forward(LrsEngine& e, const float* x, const long long* lengths, int train,
                               unsigned long long seed, cudaStream_t s) {
  const svsr_lrs_config& c = e.cfg;
  e.last_seed = seed;
  e.pd = train ? c.dropout_rate : 0.f, e.pa = train ? c.attn_dropout_rate : 0.f;
  const float pd = e.pd, pa = e.pa;
  const int D = c.adim, F = c.eunits, H = c.aheads, T = c.T, M = e.M;
  SVSR_CHECK_CUDA(cudaMemsetAsync(e.ws<uint8_t>(e.bn_stats_arena), 0, e.bn_stats_bytes, s));
  if (lengths) {
    RC(lengths_i64_to_i32(lengths, e.ws<int>(e.klen), c.B, T, s));
  } else {  // masks = None (inference call, LRS/video/lightning.py:100): every frame is valid
    std::vector<int> full(c.B, T);
    SVSR_CHECK_CUDA(cudaMemcpyAsync(e.ws<int>(e.klen), full.data(), c.B * sizeof(int), cudaMemcpyHostToDevice, s));
    SVSR_CHECK_CUDA(cudaStreamSynchronize(s));
  }
  // ---- Conv3dResNet (conv3d_extractor.py:40-48) ----
  const bf16* f4 = nullptr;
  RC(frontend_forward(e, e.fe, x, train, &f4, s));
  const int HW4 = e.fe.blocks[7].Hout * e.fe.blocks[7].Hout;
  RC(meanpool_bf16(f4, e.ws<bf16>(e.feats), M, HW4, 512, s));
  // ---- embed: Linear + x * sqrt(D) (encoder.py:170-174; embedding.py:212) ----
  RC(lin_fwd(e, e.ws<bf16>(e.feats), M, e.embed, e.xs_buf(0), D, 1, nullptr, sqrtf((float)D), 0, s, pd, e.site(1), e.seed_base()));
  const bf16* pe = e.ws<bf16>(e.pe_rel);
  if (pd > 0.f) {  // the dropped pos_emb tensor is shared by every block (embedding.py:217)
    RC(dropout_bf16(pe, e.ws<bf16>(e.pe_drop), (long long)(2 * T - 1) * D, pd, e.site(2), s, e.seed_base()));
    pe = e.ws<bf16>(e.pe_drop);
  }
  // ---- Conformer blocks (encoder_layer.py:76-150) ----
  for (int i = 0; i < c.elayers; ++i) {
    ConfLayerRef& L = e.enc[i];
    float *x0 = e.xs_buf(5 * i), *x1 = e.xs_buf(5 * i + 1), *x2 = e.xs_buf(5 * i + 2), *x3 = e.xs_buf(5 * i + 3),
          *x4 = e.xs_buf(5 * i + 4), *x5 = e.xs_buf(5 * i + 5);
    // macaron feed-forward, scaled by 1/2
    RC(ln_fwd(e, x0, L.n_mac, e.ws<bf16>(L.yn[0]), nullptr, M, s));
    RC(lin_fwd(e, e.ws<bf16>(L.yn[0]), M, L.mac1, e.ws<bf16>(L.h_mac), F, 0, nullptr, 1.f, 1, s, pd,
               e.site(LrsEngine::enc_site(i, 0)), e.seed_base()));
    RC(lin_fwd(e, e.ws<bf16>(L.h_mac), M, L.mac2, x1, D, 1, x0, 0.5f, 0, s, pd, e.site(LrsEngine::enc_site(i, 1)), e.seed_base()));
    // relative-position multi-head self-attention (attention.py:192-278)
    RC(ln_fwd(e, x1, L.n_mha, e.ws<bf16>(L.yn[1]), nullptr, M, s));
    RC(lin_fwd(e, e.ws<bf16>(L.yn[1]), M, L.qkv, e.ws<bf16>(L.qkvbuf), 3 * D, 0, nullptr, 1.f, 0, s));
    RC(lin_fwd(e, pe, 2 * T - 1, L.pos, e.ws<bf16>(L.pbuf), D, 0, nullptr, 1.f, 0, s));
    {
      AttnProblem a;
      a.q = e.ws<bf16>(L.qkvbuf), a.k = a.q + D, a.v = a.q + 2 * D, a.ldq = a.ldk = a.ldv = 3 * D;
      a.p = e.ws<bf16>(L.pbuf), a.ldp = D;
      a.bias_u = e.P + L.bias_u, a.bias_v = e.P + L.bias_v;
      a.klen = e.ws<int>(e.klen);
      a.B = c.B, a.H = H, a.Tq = T, a.Tk = T, a.scale = 0.125f;
      a.o = e.ws<bf16>(L.ctx), a.ldo = D, a.lse = e.ws<float>(L.lse);
      a.drop_p = pa, a.drop_seed = e.site(LrsEngine::enc_site(i, 2)), a.seed_base = e.seed_base();
      RC(attention_core_fwd(a, s));
    }
    RC(lin_fwd(e, e.ws<bf16>(L.ctx), M, L.out, x2, D, 1, x1, 1.f, 0, s, pd, e.site(LrsEngine::enc_site(i, 3)), e.seed_base()));
    // convolution module (convolution.py:56-75)
    RC(ln_fwd(e, x2, L.n_conv, e.ws<bf16>(L.yn[2]), nullptr, M, s));
    RC(lin_fwd(e, e.ws<bf16>(L.yn[2]), M, L.pw1, e.ws<bf16>(L.hpw1), 2 * D, 0, nullptr, 1.f, 0, s));
    RC(glu_fwd(e.ws<bf16>(L.hpw1), e.ws<bf16>(L.u), M, D, s));
    RC(dwconv1d_fwd(e.ws<bf16>(L.u), e.ws<float>(L.dw_wT), e.P + L.dw_b, e.ws<bf16>(L.dwo), c.B, T, D, c.cnn_kernel, 0, s,
                    1));
    if (train) RC(bn_col_reduce(e.ws<bf16>(L.dwo), nullptr, nullptr, M, D, e.ws<double>(L.bn.stats_f), 0, s));
    RC(bn_fwd(e, e.ws<bf16>(L.dwo), M, L.bn, train, s));
    RC(bn_apply(e.ws<bf16>(L.dwo), e.ws<float>(L.bn.coef), nullptr, nullptr, 2, e.ws<bf16>(L.act), M, D, s));
    RC(lin_fwd(e, e.ws<bf16>(L.act), M, L.pw2, x3, D, 1, x2, 1.f, 0, s, pd, e.site(LrsEngine::enc_site(i, 4)), e.seed_base()));
    // feed-forward, scaled by 1/2, and the block's final LayerNorm
    RC(ln_fwd(e, x3, L.n_ff, e.ws<bf16>(L.yn[3]), nullptr, M, s));
    RC(lin_fwd(e, e.ws<bf16>(L.yn[3]), M, L.ff1, e.ws<bf16>(L.h_ff), F, 0, nullptr, 1.f, 1, s, pd,
               e.site(LrsEngine::enc_site(i, 5)), e.seed_base()));
    RC(lin_fwd(e, e.ws<bf16>(L.h_ff), M, L.ff2, x4, D, 1, x3, 0.5f, 0, s, pd, e.site(LrsEngine::enc_site(i, 6)), e.seed_base()));
    RC(ln_fwd(e, x4, L.n_fin, nullptr, x5, M, s));
  }
  return ln_fwd(e, e.xs_buf(5 * c.elayers), e.after, e.ws<bf16>(e.enc_b), e.ws<float>(e.enc_f32), M, s);
}

static int lrs_forward(LrsEngine& e, const float* x, const long long* lengths, const long long* tokens,
                       long long tok_stride_b, const long long* label, int Llab, int train, unsigned long long seed,
                       float* metrics, cudaStream_t s) {
  const svsr_lrs_config& c = e.cfg;
  const int D = c.adim, Fd = c.dunits, H = c.aheads, T = c.T, M = e.M;
  const int L = Llab + 1, Md = c.B * L;
  SVSR_REQUIRE(Llab >= 1 && L <= c.Lmax, "lrs_forward: label length %d exceeds the engine's Lmax-1 = %d", Llab, c.Lmax - 1);
  SVSR_CHECK_CUDA(cudaMemsetAsync(e.ws<uint8_t>(e.acc), 0, 8 * sizeof(double) + 256, s));
  RC(lrs_encoder_forward(e, x, lengths, train, seed, s));
  const bf16* enc_b = e.ws<bf16>(e.enc_b);
  const float pd = e.pd, pa = e.pa;
  // ---- audio-token cross-entropy (e2e_asr_transformer.py:195-201): padded frames are scored too ----
  const int has_audio = (e.AGV > 0 && tokens) ? 1 : 0;
  const long long audio_rows = (long long)M * c.audio_alignment * c.vq_groups;
  if (has_audio) {
    const long long tsb = (long long)T * c.audio_alignment * c.vq_groups;
    SVSR_CHECK_CUDA(cudaMemcpy2DAsync(e.ws<uint8_t>(e.ce_tok), (size_t)tsb * 8, tokens, (size_t)tok_stride_b * 8,
                                      (size_t)tsb * 8, c.B, cudaMemcpyDeviceToDevice, s));
    if (e.fused_head()) {
      RC(lrs_audio_head_gemm(e, 1, nullptr, s));
      RC(ce_finalize(e.ws<float2>(e.ce_part), e.ws<float>(e.ce_xt), e.ws<long long>(e.ce_tok), tsb, c.B, T,
                     c.audio_alignment, c.vq_groups, c.audio_vocab, e.ws<float>(e.ce_lse), e.ws<double>(e.acc), s));
    } else {
      RC(lin_fwd(e, enc_b, M, e.aud, e.ws<float>(e.logits_a), e.AGV, 1, nullptr, 1.f, 0, s));
      RC(audio_ce(e.ws<float>(e.logits_a), e.AGV, e.ws<long long>(e.ce_tok), tsb, c.B, T, c.audio_alignment, c.vq_groups,
                  c.audio_vocab, e.ws<bf16>(e.dlogits_a), e.ws<double>(e.acc), e.ws<int>(e.bad_token),
                  c.audio_weight / (float)audio_rows, s));
    }
  }
  // ---- CTC (ctc.py:83-151) ----
  const bf16* enc_ctc = enc_b;  // ctc_lo(dropout(hs_pad)), ctc.py:97
  if (pd > 0.f) {
    RC(dropout_bf16(enc_b, e.ws<bf16>(e.enc_ctc), (long long)M * D, pd, e.site(3), s, e.seed_base()));
    enc_ctc = e.ws<bf16>(e.enc_ctc);
  }
  RC(lin_fwd(e, enc_ctc, M, e.ctc, e.ws<float>(e.logits_c), e.ldv, 1, nullptr, 1.f, 0, s));
  RC(ctc_loss_fwd_bwd(e.ws<float>(e.logits_c), e.ldv, c.odim, label, Llab, e.ws<int>(e.klen), c.B, T,
                      e.ws<bf16>(e.dlogits_c), e.ws<double>(e.acc), 1, c.mtlalpha / (float)c.B,
                      e.ws<float>(e.ctc_scratch), s));
  // ---- attention decoder (decoder.py:122-151; decoder_layer.py:58-121) ----
  long long* ys_in = e.ws<long long>(e.ys_in);
  long long* ys_out = e.ws<long long>(e.ys_out);
  add_sos_eos_kernel<<<(c.B + 63) / 64, 64, 0, s>>>(label, Llab, ys_in, ys_out, c.B, c.odim - 1, c.odim - 1);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  RC(embed_posenc_fwd(ys_in, e.P + e.dec_emb, e.xd_buf(0), Md, L, D, c.odim, s, pd, e.site(4), e.seed_base()));
  for (int i = 0; i < c.dlayers; ++i) {
    DecLayerRef& Ld = e.dec[i];
    float *x0 = e.xd_buf(3 * i), *x1 = e.xd_buf(3 * i + 1), *x2 = e.xd_buf(3 * i + 2), *x3 = e.xd_buf(3 * i + 3);
    RC(ln_fwd(e, x0, Ld.n1, e.ws<bf16>(Ld.yn[0]), nullptr, Md, s));
    RC(lin_fwd(e, e.ws<bf16>(Ld.yn[0]), Md, Ld.qkv, e.ws<bf16>(Ld.qkvbuf), 3 * D, 0, nullptr, 1.f, 0, s));
    {
      AttnProblem a;
      a.q = e.ws<bf16>(Ld.qkvbuf), a.k = a.q + D, a.v = a.q + 2 * D, a.ldq = a.ldk = a.ldv = 3 * D;
      a.causal = 1;
      a.B = c.B, a.H = H, a.Tq = L, a.Tk = L, a.scale = 0.125f;
      a.o = e.ws<bf16>(Ld.ctx_s), a.ldo = D, a.lse = e.ws<float>(Ld.lse_s);
      a.drop_p = pa, a.drop_seed = e.site(LrsEngine::dec_site(i, 0)), a.seed_base = e.seed_base();
      RC(attention_core_fwd(a, s));
    }
    RC(lin_fwd(e, e.ws<bf16>(Ld.ctx_s), Md, Ld.out_s, x1, D, 1, x0, 1.f, 0, s, pd, e.site(LrsEngine::dec_site(i, 1)), e.seed_base()));
    RC(ln_fwd(e, x1, Ld.n2, e.ws<bf16>(Ld.yn[1]), nullptr, Md, s));
    RC(lin_fwd(e, e.ws<bf16>(Ld.yn[1]), Md, Ld.q_c, e.ws<bf16>(Ld.qc), D, 0, nullptr, 1.f, 0, s));
    RC(lin_fwd(e, enc_b, M, Ld.kv_c, e.ws<bf16>(Ld.kvc), 2 * D, 0, nullptr, 1.f, 0, s));
    {
      AttnProblem a;
      a.q = e.ws<bf16>(Ld.qc), a.ldq = D;
      a.k = e.ws<bf16>(Ld.kvc), a.v = a.k + D, a.ldk = a.ldv = 2 * D;
      a.klen = e.ws<int>(e.klen);
      a.B = c.B, a.H = H, a.Tq = L, a.Tk = T, a.scale = 0.125f;
      a.o = e.ws<bf16>(Ld.ctx_c), a.ldo = D, a.lse = e.ws<float>(Ld.lse_c);
      a.drop_p = pa, a.drop_seed = e.site(LrsEngine::dec_site(i, 2)), a.seed_base = e.seed_base();
      RC(attention_core_fwd(a, s));
    }
    RC(lin_fwd(e, e.ws<bf16>(Ld.ctx_c), Md, Ld.out_c, x2, D, 1, x1, 1.f, 0, s, pd, e.site(LrsEngine::dec_site(i, 3)), e.seed_base()));
    RC(ln_fwd(e, x2, Ld.n3, e.ws<bf16>(Ld.yn[2]), nullptr, Md, s));
    RC(lin_fwd(e, e.ws<bf16>(Ld.yn[2]), Md, Ld.ff1, e.ws<bf16>(Ld.h), Fd, 0, nullptr, 1.f, 1, s, pd,
               e.site(LrsEngine::dec_site(i, 4)), e.seed_base()));
    RC(lin_fwd(e, e.ws<bf16>(Ld.h), Md, Ld.ff2, x3, D, 1, x2, 1.f, 0, s, pd, e.site(LrsEngine::dec_site(i, 5)), e.seed_base()));
  }
  RC(ln_fwd(e, e.xd_buf(3 * c.dlayers), e.dec_after, e.ws<bf16>(e.dec_yn), nullptr, Md, s));
  RC(lin_fwd(e, e.ws<bf16>(e.dec_yn), Md, e.outl, e.ws<float>(e.pred), e.ldv, 1, nullptr, 1.f, 0, s));
  // ---- label-smoothing loss + accuracy (label_smoothing_loss.py:41-63; nets_utils.py:303) ----
  RC(label_smoothing_loss(e.ws<float>(e.pred), e.ldv, c.odim, ys_out, Md, c.lsm_weight, e.ws<bf16>(e.dpred),
                          e.ws<double>(e.acc), 2, (1.f - c.mtlalpha) / (float)c.B, s));
  RC(lrs_finalize_metrics(e.ws<double>(e.acc), metrics, c.B, audio_rows, c.mtlalpha, c.audio_weight, has_audio, s,
                          e.ws<int>(e.bad_token)));
  e.last_L = L, e.last_Llab = Llab, e.last_train = train, e.last_audio = has_audio;
  e.fwd_done = true;
  return SVSR_OK;
}

// ------------------------------------------------------------------------------------------------
// backward: heads -> decoder -> Conformer blocks -> embed -> frontend. Two streams, as in the LRW engine: `s` carries the
// critical chain (input-gradient GEMMs, attention / norm / conv-module backward), `e.side` every weight-gradient GEMM
// (+ bias column sums). One "unit" per sub-block; the temporaries a unit's side work reads are double buffered by unit
// parity and the chain waits for the side stream's unit k-1 before it starts unit k+1 (SideQueue::end_unit).
// ------------------------------------------------------------------------------------------------
struct LrsScratch {
  bf16 *dxb, *gF, *g3D, *gD[3], *dpb, *ddxb, *dkv;
};
static LrsScratch lrs_scratch(const LrsEngine& e, int unit) {
  const int u = unit & 1;
  LrsScratch t;
  t.dxb = e.ws<bf16>(e.dxb[u]), t.gF = e.ws<bf16>(e.gF[u]), t.g3D = e.ws<bf16>(e.g3D[u]);
  for (int i = 0; i < 3; ++i) t.gD[i] = e.ws<bf16>(e.gD[u][i]);
  t.dpb = e.ws<bf16>(e.dpb[u]), t.ddxb = e.ws<bf16>(e.ddxb[u]), t.dkv = e.ws<bf16>(e.dkv[u]);
  return t;
}

// dyb: the (already cast / masked) branch gradient in t.dxb or t.ddxb
static int ffn_bwd(LrsEngine& e, SideQueue& sq, const LrsScratch& t, const bf16* dyb, int rows, const LinRef& w1,
                   const LinRef& w2, const bf16* h, const bf16* yn, const float* x_in, const LnRef& n, float* dx, int F,
                   cudaStream_t s) {
  // h = dropout(relu(.)) is zero where dropped OR clipped, so the fused `h > 0` mask covers both; surviving units
  // carry the 1/(1-p) of the dropout
  const float hs = e.pd > 0.f ? 1.0f / (1.0f - e.pd) : 1.0f;
  const int D = e.cfg.adim;
  bf16* dh = t.gF;
  bf16* dyn = t.gD[0];
  RC(sq.fork());  // dyb complete
  RC(linear_wgrad(e, dyb, D, h, rows, w2, e.wq));
  RC(lin_dgrad(e, dyb, D, rows, w2, dh, F, 0, nullptr, hs, h, s));  // ReLU backward fused: zero where h <= 0
  RC(sq.fork());  // dh complete
  RC(linear_wgrad(e, dh, F, yn, rows, w1, e.wq));
  RC(lin_dgrad(e, dh, F, rows, w1, dyn, D, 0, nullptr, 1.f, nullptr, s));
  RC(ln_bwd(e, dyn, nullptr, x_in, n, dx, 1, rows, s));
  return sq.end_unit();
}

// stage < 0: the whole backward. Staged (data parallel: the gradient all-reduce of a finished group overlaps the next
// stage): 0 = loss heads + attention decoder, 1 = encoder.after_norm + the Conformer blocks, 2 = embed + visual frontend.
// Every stage joins the weight-gradient stream before it returns, so its parameters' gradients are final.
static int lrs_backward(LrsEngine& e, const float* grad_scale, int stage, cudaStream_t s) {
  if (stage <= 0) {
    SVSR_REQUIRE(e.fwd_done, "lrs backward called before (or twice after) forward");
    SVSR_REQUIRE(e.last_train, "lrs backward needs a train-mode forward (batch-statistics BatchNorm backward)");
    e.fwd_done = false;
    e.bwd_stage = 0;
  } else {
    SVSR_REQUIRE(e.bwd_stage == stage - 1, "lrs backward stage %d called out of order (last finished: %d)", stage, e.bwd_stage);
  }
  if (stage >= 0) e.bwd_stage = stage;
  const svsr_lrs_config& c = e.cfg;
  const int D = c.adim, F = c.eunits, Fd = c.dunits, H = c.aheads, T = c.T, M = e.M;
  const int L = e.last_L, Md = c.B * L;
  float* dx = e.ws<float>(e.dx);
  const bf16* enc_b = e.ws<bf16>(e.enc_b);
  const float pd = e.pd, pa = e.pa;
  SideQueue sq(e, s);
  cudaStream_t w = e.wq;
  if (stage <= 0) {
  if (e.last_audio && e.fused_head()) RC(lrs_audio_head_gemm(e, 2, grad_scale, s));
  if (grad_scale) {
    if (e.last_audio && !e.fused_head())
      RC(scale_bf16_by_device_scalar(e.ws<bf16>(e.dlogits_a), (long long)M * e.AGV, grad_scale, s));
    RC(scale_bf16_by_device_scalar(e.ws<bf16>(e.dlogits_c), (long long)M * e.ldv, grad_scale, s));
    RC(scale_bf16_by_device_scalar(e.ws<bf16>(e.dpred), (long long)Md * e.ldv, grad_scale, s));
  }
  // ---- unit: d loss / d encoder output from the audio and CTC heads ----
  {
    const LrsScratch t = lrs_scratch(e, sq.unit);
    RC(sq.fork());
    if (e.last_audio) {
      RC(linear_wgrad(e, e.ws<bf16>(e.dlogits_a), e.AGV, enc_b, M, e.aud, w));
      RC(lin_dgrad(e, e.ws<bf16>(e.dlogits_a), e.AGV, M, e.aud, dx, D, 1, nullptr, 1.f, nullptr, s));
    } else {
      SVSR_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)M * D * 4, s));
    }
    if (pd > 0.f) {  // through the Dropout in front of ctc_lo
      RC(linear_wgrad(e, e.ws<bf16>(e.dlogits_c), e.ldv, e.ws<bf16>(e.enc_ctc), M, e.ctc, w));
      RC(lin_dgrad(e, e.ws<bf16>(e.dlogits_c), e.ldv, M, e.ctc, t.gD[1], D, 0, nullptr, 1.f, nullptr, s));
      RC(dropout_add_bf16_to_f32(dx, t.gD[1], (long long)M * D, pd, e.site(3), s, e.seed_base()));
    } else {
      RC(linear_wgrad(e, e.ws<bf16>(e.dlogits_c), e.ldv, enc_b, M, e.ctc, w));
      RC(lin_dgrad(e, e.ws<bf16>(e.dlogits_c), e.ldv, M, e.ctc, dx, D, 1, dx, 1.f, nullptr, s));
    }
    RC(sq.end_unit());
  }

  // ---- decoder ----
  {
    float* ddx = e.ws<float>(e.ddx);
    {  // unit: output layer + after_norm
      const LrsScratch t = lrs_scratch(e, sq.unit);
      RC(sq.fork());
      RC(linear_wgrad(e, e.ws<bf16>(e.dpred), e.ldv, e.ws<bf16>(e.dec_yn), Md, e.outl, w));
      RC(lin_dgrad(e, e.ws<bf16>(e.dpred), e.ldv, Md, e.outl, t.gD[0], D, 0, nullptr, 1.f, nullptr, s));
      RC(ln_bwd(e, t.gD[0], nullptr, e.xd_buf(3 * c.dlayers), e.dec_after, ddx, 0, Md, s));
      RC(sq.end_unit());
    }
    for (int i = c.dlayers - 1; i >= 0; --i) {
      DecLayerRef& Ld = e.dec[i];
      float *x0 = e.xd_buf(3 * i), *x1 = e.xd_buf(3 * i + 1), *x2 = e.xd_buf(3 * i + 2);
      {  // unit: feed-forward
        const LrsScratch t = lrs_scratch(e, sq.unit);
        RC(cast_scale_f32_bf16(ddx, t.ddxb, (long long)Md * D, 1.f, s, pd, e.site(LrsEngine::dec_site(i, 5)), e.seed_base()));
        RC(ffn_bwd(e, sq, t, t.ddxb, Md, Ld.ff1, Ld.ff2, e.ws<bf16>(Ld.h), e.ws<bf16>(Ld.yn[2]), x2, Ld.n3, ddx, Fd, s));
      }
      {  // unit: source attention over the encoder output
        const LrsScratch t = lrs_scratch(e, sq.unit);
        bf16 *dyn = t.gD[0], *dctx = t.gD[1], *dq = t.gD[2], *dkv = t.dkv;
        RC(cast_scale_f32_bf16(ddx, t.ddxb, (long long)Md * D, 1.f, s, pd, e.site(LrsEngine::dec_site(i, 3)), e.seed_base()));
        RC(sq.fork());
        RC(linear_wgrad(e, t.ddxb, D, e.ws<bf16>(Ld.ctx_c), Md, Ld.out_c, w));
        RC(lin_dgrad(e, t.ddxb, D, Md, Ld.out_c, dctx, D, 0, nullptr, 1.f, nullptr, s));
        {
          AttnProblem a;
          a.q = e.ws<bf16>(Ld.qc), a.ldq = D;
          a.k = e.ws<bf16>(Ld.kvc), a.v = a.k + D, a.ldk = a.ldv = 2 * D;
          a.klen = e.ws<int>(e.klen);
          a.B = c.B, a.H = H, a.Tq = L, a.Tk = T, a.scale = 0.125f;
          a.o = e.ws<bf16>(Ld.ctx_c), a.ldo = D, a.lse = e.ws<float>(Ld.lse_c);
          a.drop_p = pa, a.drop_seed = e.site(LrsEngine::dec_site(i, 2)), a.seed_base = e.seed_base();
          AttnGrads g;
          g.d_o = dctx, g.dq = dq, g.lddq = D, g.dk = dkv, g.dv = dkv + D, g.lddk = g.lddv = 2 * D;
          g.scratch = e.ws<float>(e.attn_scratch);
          RC(attention_core_bwd(a, g, s));
        }
        RC(sq.fork());
        RC(linear_wgrad(e, dq, D, e.ws<bf16>(Ld.yn[1]), Md, Ld.q_c, w));
        RC(linear_wgrad(e, dkv, 2 * D, enc_b, M, Ld.kv_c, w));
        RC(lin_dgrad(e, dkv, 2 * D, M, Ld.kv_c, dx, D, 1, dx, 1.f, nullptr, s));  // memory gradient accumulates
        RC(lin_dgrad(e, dq, D, Md, Ld.q_c, dyn, D, 0, nullptr, 1.f, nullptr, s));
        RC(ln_bwd(e, dyn, nullptr, x1, Ld.n2, ddx, 1, Md, s));
        RC(sq.end_unit());
      }
      {  // unit: causal self-attention
        const LrsScratch t = lrs_scratch(e, sq.unit);
        bf16 *dyn = t.gD[0], *dctx = t.gD[1], *dqkv = t.g3D;
        RC(cast_scale_f32_bf16(ddx, t.ddxb, (long long)Md * D, 1.f, s, pd, e.site(LrsEngine::dec_site(i, 1)), e.seed_base()));
        RC(sq.fork());
        RC(linear_wgrad(e, t.ddxb, D, e.ws<bf16>(Ld.ctx_s), Md, Ld.out_s, w));
        RC(lin_dgrad(e, t.ddxb, D, Md, Ld.out_s, dctx, D, 0, nullptr, 1.f, nullptr, s));
        {
          AttnProblem a;
          a.q = e.ws<bf16>(Ld.qkvbuf), a.k = a.q + D, a.v = a.q + 2 * D, a.ldq = a.ldk = a.ldv = 3 * D;
          a.causal = 1;
          a.B = c.B, a.H = H, a.Tq = L, a.Tk = L, a.scale = 0.125f;
          a.o = e.ws<bf16>(Ld.ctx_s), a.ldo = D, a.lse = e.ws<float>(Ld.lse_s);
          a.drop_p = pa, a.drop_seed = e.site(LrsEngine::dec_site(i, 0)), a.seed_base = e.seed_base();
          AttnGrads g;
          g.d_o = dctx, g.dq = dqkv, g.dk = dqkv + D, g.dv = dqkv + 2 * D, g.lddq = g.lddk = g.lddv = 3 * D;
          g.scratch = e.ws<float>(e.attn_scratch);
          RC(attention_core_bwd(a, g, s));
        }
        RC(sq.fork());
        RC(linear_wgrad(e, dqkv, 3 * D, e.ws<bf16>(Ld.yn[0]), Md, Ld.qkv, w));
        RC(lin_dgrad(e, dqkv, 3 * D, Md, Ld.qkv, dyn, D, 0, nullptr, 1.f, nullptr, s));
        RC(ln_bwd(e, dyn, nullptr, x0, Ld.n1, ddx, 1, Md, s));
        RC(sq.end_unit());
      }
    }
    RC(embed_bwd(e.ws<long long>(e.ys_in), ddx, e.G + e.dec_emb, Md, D, c.odim, s, pd, e.site(4), e.seed_base()));
  }
  if (stage == 0) return sq.join();
  }  // stage <= 0

  if (stage < 0 || stage == 1) {
  // ---- encoder.after_norm, then the Conformer blocks in reverse ----
  RC(ln_bwd(e, nullptr, dx, e.xs_buf(5 * c.elayers), e.after, dx, 0, M, s));
  for (int i = c.elayers - 1; i >= 0; --i) {
    ConfLayerRef& Lc = e.enc[i];
    float *x0 = e.xs_buf(5 * i), *x1 = e.xs_buf(5 * i + 1), *x2 = e.xs_buf(5 * i + 2), *x3 = e.xs_buf(5 * i + 3),
          *x4 = e.xs_buf(5 * i + 4);
    RC(ln_bwd(e, nullptr, dx, x4, Lc.n_fin, dx, 0, M, s));
    {  // unit: feed-forward (x 1/2)
      const LrsScratch t = lrs_scratch(e, sq.unit);
      RC(cast_scale_f32_bf16(dx, t.dxb, (long long)M * D, 0.5f, s, pd, e.site(LrsEngine::enc_site(i, 6)), e.seed_base()));
      RC(ffn_bwd(e, sq, t, t.dxb, M, Lc.ff1, Lc.ff2, e.ws<bf16>(Lc.h_ff), e.ws<bf16>(Lc.yn[3]), x3, Lc.n_ff, dx, F, s));
    }
    {  // unit: convolution module
      const LrsScratch t = lrs_scratch(e, sq.unit);
      bf16 *dyn = t.gD[0], *t1 = t.gD[1], *t2 = t.gD[2], *g3 = t.g3D;
      RC(cast_scale_f32_bf16(dx, t.dxb, (long long)M * D, 1.f, s, pd, e.site(LrsEngine::enc_site(i, 4)), e.seed_base()));
      RC(sq.fork());
      RC(linear_wgrad(e, t.dxb, D, e.ws<bf16>(Lc.act), M, Lc.pw2, w));
      RC(lin_dgrad(e, t.dxb, D, M, Lc.pw2, t1, D, 0, nullptr, 1.f, nullptr, s));  // t1 = d act
      RC(bn_col_reduce(e.ws<bf16>(Lc.dwo), t1, e.ws<float>(Lc.bn.coef), M, D, e.ws<double>(Lc.bn.stats_b), 1, s));
      RC(bn_bwd_apply(t1, nullptr, e.ws<bf16>(Lc.dwo), e.ws<float>(Lc.bn.coef), nullptr, t2, nullptr, M, D, 2, s, nullptr,
                      nullptr, e.ws<double>(Lc.bn.stats_b), e.G + Lc.bn.gamma, e.G + Lc.bn.beta));  // t2 = d dwo
      RC(sq.fork());
      RC(dwconv1d_wgrad(e.ws<bf16>(Lc.u), t2, e.G + Lc.dw_w, e.G + Lc.dw_b, c.B, T, D, c.cnn_kernel, w));
      RC(dwconv1d_fwd(t2, e.ws<float>(Lc.dw_wT), nullptr, t1, c.B, T, D, c.cnn_kernel, 1, s, 1));  // t1 = d u
      RC(glu_bwd(e.ws<bf16>(Lc.hpw1), t1, g3, M, D, s));                                            // g3 = d hpw1 [M, 2D]
      RC(sq.fork());
      RC(linear_wgrad(e, g3, 2 * D, e.ws<bf16>(Lc.yn[2]), M, Lc.pw1, w));
      RC(lin_dgrad(e, g3, 2 * D, M, Lc.pw1, dyn, D, 0, nullptr, 1.f, nullptr, s));
      RC(ln_bwd(e, dyn, nullptr, x2, Lc.n_conv, dx, 1, M, s));
      RC(sq.end_unit());
    }
    {  // unit: relative-position self-attention
      const LrsScratch t = lrs_scratch(e, sq.unit);
      bf16 *dyn = t.gD[0], *t1 = t.gD[1], *g3 = t.g3D;
      RC(cast_scale_f32_bf16(dx, t.dxb, (long long)M * D, 1.f, s, pd, e.site(LrsEngine::enc_site(i, 3)), e.seed_base()));
      RC(sq.fork());
      RC(linear_wgrad(e, t.dxb, D, e.ws<bf16>(Lc.ctx), M, Lc.out, w));
      RC(lin_dgrad(e, t.dxb, D, M, Lc.out, t1, D, 0, nullptr, 1.f, nullptr, s));  // t1 = d ctx
      {
        AttnProblem a;
        a.q = e.ws<bf16>(Lc.qkvbuf), a.k = a.q + D, a.v = a.q + 2 * D, a.ldq = a.ldk = a.ldv = 3 * D;
        a.p = e.ws<bf16>(Lc.pbuf), a.ldp = D;
        a.bias_u = e.P + Lc.bias_u, a.bias_v = e.P + Lc.bias_v;
        a.klen = e.ws<int>(e.klen);
        a.B = c.B, a.H = H, a.Tq = T, a.Tk = T, a.scale = 0.125f;
        a.o = e.ws<bf16>(Lc.ctx), a.ldo = D, a.lse = e.ws<float>(Lc.lse);
        a.drop_p = pa, a.drop_seed = e.site(LrsEngine::enc_site(i, 2)), a.seed_base = e.seed_base();
        AttnGrads g;
        g.d_o = t1, g.dq = g3, g.dk = g3 + D, g.dv = g3 + 2 * D, g.lddq = g.lddk = g.lddv = 3 * D;
        g.dp = e.ws<float>(e.dp), g.dbias_u = e.G + Lc.bias_u, g.dbias_v = e.G + Lc.bias_v;
        g.scratch = e.ws<float>(e.attn_scratch);
        SVSR_CHECK_CUDA(cudaMemsetAsync(g.dp, 0, (size_t)(2 * T - 1) * D * 4, s));
        RC(attention_core_bwd(a, g, s));
      }
      RC(cast_scale_f32_bf16(e.ws<float>(e.dp), t.dpb, (long long)(2 * T - 1) * D, 1.f, s));
      RC(sq.fork());
      RC(linear_wgrad(e, t.dpb, D, pd > 0.f ? e.ws<bf16>(e.pe_drop) : e.ws<bf16>(e.pe_rel), 2 * T - 1, Lc.pos, w));
      RC(linear_wgrad(e, g3, 3 * D, e.ws<bf16>(Lc.yn[1]), M, Lc.qkv, w));
      RC(lin_dgrad(e, g3, 3 * D, M, Lc.qkv, dyn, D, 0, nullptr, 1.f, nullptr, s));
      RC(ln_bwd(e, dyn, nullptr, x1, Lc.n_mha, dx, 1, M, s));
      RC(sq.end_unit());
    }
    {  // unit: macaron feed-forward (x 1/2)
      const LrsScratch t = lrs_scratch(e, sq.unit);
      RC(cast_scale_f32_bf16(dx, t.dxb, (long long)M * D, 0.5f, s, pd, e.site(LrsEngine::enc_site(i, 1)), e.seed_base()));
      RC(ffn_bwd(e, sq, t, t.dxb, M, Lc.mac1, Lc.mac2, e.ws<bf16>(Lc.h_mac), e.ws<bf16>(Lc.yn[0]), x0, Lc.n_mac, dx, F, s));
    }
  }
  if (stage == 1) return sq.join();
  }  // stage 1

  // ---- unit: embed (x * sqrt(D)) -> average pool; then the frontend ----
  {
    const LrsScratch t = lrs_scratch(e, sq.unit);
    RC(cast_scale_f32_bf16(dx, t.dxb, (long long)M * D, sqrtf((float)D), s, pd, e.site(1), e.seed_base()));
    RC(sq.fork());
    RC(linear_wgrad(e, t.dxb, D, e.ws<bf16>(e.feats), M, e.embed, w));
    RC(lin_dgrad(e, t.dxb, D, M, e.embed, e.ws<bf16>(e.dfeat), 512, 0, nullptr, 1.f, nullptr, s));
    const int HW4 = e.fe.blocks[7].Hout * e.fe.blocks[7].Hout;
    RC(meanpool_bf16_bwd(e.ws<bf16>(e.dfeat), e.ws<bf16>(e.fe.gbuf[0]), M, HW4, 512, s));
    RC(sq.end_unit());
  }
  return frontend_backward(e, e.fe, sq, s);
}

}  // namespace svsr

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace svsr;

extern "C" {

int svsr_lrs_create(const svsr_lrs_config* cfg, void** handle) {
  SVSR_REQUIRE(cfg && handle, "lrs_create: null argument");
  LrsEngine* e = new LrsEngine();
  e->cfg = *cfg;
  int rc = lrs_build(*e, 0);
  if (!rc) rc = lrs_build(*e, e->decay_count);
  if (rc) {
    delete e;
    return rc;
  }
  *handle = e;
  return SVSR_OK;
}
int svsr_lrs_destroy(void* h) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  if (e) engine_base_destroy(*e);
  delete e;
  return SVSR_OK;
}
int64_t svsr_lrs_param_count(void* h) { return static_cast<LrsEngine*>(h)->param_count; }
int64_t svsr_lrs_decay_count(void* h) { return static_cast<LrsEngine*>(h)->decay_count; }
int64_t svsr_lrs_buffer_count(void* h) { return static_cast<LrsEngine*>(h)->buffer_count; }
int64_t svsr_lrs_workspace_bytes(void* h) { return (int64_t)static_cast<LrsEngine*>(h)->ws_bytes; }
int svsr_lrs_num_params(void* h) { return (int)static_cast<LrsEngine*>(h)->params.size(); }
int svsr_lrs_num_buffers(void* h) { return (int)static_cast<LrsEngine*>(h)->buffers.size(); }
int svsr_lrs_param_info(void* h, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset, int* decay) {
  return tensor_info(static_cast<LrsEngine*>(h)->params, i, name, ndim, shape, offset, decay);
}
int svsr_lrs_buffer_info(void* h, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset) {
  return tensor_info(static_cast<LrsEngine*>(h)->buffers, i, name, ndim, shape, offset, nullptr);
}
int svsr_lrs_bind(void* h, float* params, float* grads, float* buffers, void* workspace, int64_t workspace_bytes) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(params && grads && buffers && workspace, "lrs_bind: null pointer");
  SVSR_REQUIRE((size_t)workspace_bytes >= e->ws_bytes, "lrs_bind: workspace too small (%lld < %zu)",
               (long long)workspace_bytes, e->ws_bytes);
  SVSR_REQUIRE(((uintptr_t)workspace & 1023) == 0 && ((uintptr_t)params & 15) == 0 && ((uintptr_t)grads & 15) == 0,
               "lrs_bind: workspace must be 1024-byte aligned, arenas 16-byte aligned");
  e->pack_table_ready = false;
  return engine_base_bind(*e, params, grads, buffers, workspace);
}
int svsr_lrs_pack_weights(void* h, void* stream) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrs: bind() first");
  return lrs_pack(*e, static_cast<cudaStream_t>(stream));
}
int svsr_lrs_forward(void* h, const float* x, const int64_t* lengths, const int64_t* tokens, int64_t tok_stride_b,
                     const int64_t* label, int label_len, int train, uint64_t dropout_seed, float* metrics,
                     void* stream) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrs: bind() first");
  SVSR_REQUIRE(x && lengths && label && metrics, "lrs_forward: null input");
  SVSR_REQUIRE(!tokens || tok_stride_b >= (int64_t)e->cfg.T * e->cfg.audio_alignment * e->cfg.vq_groups,
               "lrs_forward: audio tokens have fewer than T*alignment rows per clip");
  return lrs_forward(*e, x, reinterpret_cast<const long long*>(lengths), reinterpret_cast<const long long*>(tokens),
                     tok_stride_b, reinterpret_cast<const long long*>(label), label_len, train,
                     (unsigned long long)dropout_seed, metrics, static_cast<cudaStream_t>(stream));
}
int svsr_lrs_encode(void* h, const float* x, const int64_t* lengths, int train, uint64_t dropout_seed, void* stream) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrs: bind() first");
  SVSR_REQUIRE(x, "lrs_encode: null input");
  e->fwd_done = false;
  return lrs_encoder_forward(*e, x, reinterpret_cast<const long long*>(lengths), train, (unsigned long long)dropout_seed,
                             static_cast<cudaStream_t>(stream));
}
int svsr_lrs_backward(void* h, const float* grad_scale, void* stream) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrs: bind() first");
  return lrs_backward(*e, grad_scale, -1, static_cast<cudaStream_t>(stream));
}
int svsr_lrs_backward_stage(void* h, const float* grad_scale, int stage, void* stream) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrs: bind() first");
  SVSR_REQUIRE(stage >= 0 && stage <= 2, "lrs_backward_stage: stage must be 0, 1 or 2");
  return lrs_backward(*e, grad_scale, stage, static_cast<cudaStream_t>(stream));
}
int svsr_lrs_step_control(void* h, int mode, uint64_t dropout_seed, void* stream) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrs: bind() first");
  e->dev_ctl = mode != 0;
  if (!e->dev_ctl) return SVSR_OK;
  return set_step_ctl(e->ws<unsigned>(e->ctl + 8), e->ws<unsigned long long>(e->ctl), 0u, (unsigned long long)dropout_seed,
                      static_cast<cudaStream_t>(stream));
}
// The step never writes the audio logits to HBM (fused head); materialise them once, on request (svsr_lrs_tensor
// "logits_audio"), from the last forward's encoder output.
int svsr_lrs_logits_audio(void* h, void* stream) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(e->WS && e->AGV > 0, "lrs_logits_audio: bind() first / no audio head configured");
  return lin_fwd(*e, e->ws<bf16>(e->enc_b), e->M, e->aud, e->ws<float>(e->logits_a), e->AGV, 1, nullptr, 1.f, 0,
                 static_cast<cudaStream_t>(stream));
}
int svsr_lrs_tensor(void* h, const char* name, void** ptr, int64_t* numel, int* dtype) {
  LrsEngine* e = static_cast<LrsEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrs: bind() first");
  const svsr_lrs_config& c = e->cfg;
  const std::string n(name);
  auto set = [&](size_t off, int64_t ne, int dt) {
    *ptr = e->WS + off, *numel = ne, *dtype = dt;
    return SVSR_OK;
  };
  const int64_t MD = (int64_t)e->M * c.adim;
  if (n == "encoder_out") return set(e->enc_f32, MD, 0);
  if (n == "embed_out") return set(e->xs, MD, 0);
  if (n == "frontend") return set(e->feats, (int64_t)e->M * 512, 1);
  if (n == "logits_audio") return set(e->logits_a, (int64_t)e->M * (e->AGV > 0 ? e->AGV : 64), 0);
  if (n == "ctc_logits") return set(e->logits_c, (int64_t)e->M * e->ldv, 0);
  if (n == "pred") return set(e->pred, (int64_t)c.B * e->last_L * e->ldv, 0);
  if (n == "ys_in") return set(e->ys_in, (int64_t)c.B * e->last_L, 4);
  if (n == "ys_out") return set(e->ys_out, (int64_t)c.B * e->last_L, 4);
  if (n == "bad_token") return set(e->bad_token, 1, 3);
  if (n == "stem_out") return set(e->fe.x1, (int64_t)e->fe.B * e->fe.T * e->fe.H1 * e->fe.H1 * 64, 1);
  if (n.rfind("layer", 0) == 0) {  // "layer<i>.x<k>": residual stream after sub-block k of Conformer block i
    int li = -1, k = -1;
    if (sscanf(name, "layer%d.x%d", &li, &k) == 2 && li >= 0 && li < c.elayers && k >= 0 && k <= 5) {
      *ptr = e->xs_buf(5 * li + k), *numel = MD, *dtype = 0;
      return SVSR_OK;
    }
  }
  set_last_error("lrs_tensor: unknown tensor '%s'", name);
  return SVSR_ERR_INVALID;
}

}  // extern "C"
