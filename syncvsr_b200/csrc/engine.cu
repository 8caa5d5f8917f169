// Native step executor for the LRW word-level model (reference: LRW/video/src/lightning.py:36-191).
// Owns the layout of the flat parameter / gradient / buffer arenas, the activation workspace and the bf16 operand
// copies of the weights, and sequences every kernel of forward and backward on one stream (no autograd, no
// per-op Python): stem3d -> resnet.layer1-4 -> mean pool + CLS -> x-transformers encoder -> the two loss heads.
#include "engine_common.cuh"
#include "encoder.cuh"
#include "conformer.cuh"
#include "heads.cuh"
#include "precise.cuh"

namespace svsr {

struct EncLayerRef {
  long long g_a, g_f;
  LinRef qkv, out, ff1, ff2;
  size_t xn_a, inv_a, qkvbuf, obuf, xn_f, inv_f, hbuf, ubuf;
  size_t g_a_pad = 0, g_f_pad = 0;  // word-boundary variant: zero-padded fp32 copies of the RMSNorm weights
};

// `model.bert.type: huggingface` (lightning.py:90-92,152-156): post-LN BERT layer
struct BertLayerRef {
  LinRef qkv, out, inter, outd;
  long long ln1_g, ln1_b, ln2_g, ln2_b;
  size_t qkvbuf, lse, ctx, A, st1, x1, x1b, pre, h, O, st2;
};

struct LrwEngine : EngineBase {
  svsr_lrw_config cfg;
  int M;  // tokens = B*(T+1)
  // Encoder width: D = model.bert.dim (+1 with data.use_word_boundary, lightning.py:145-150). Every D-wide buffer has
  // the pitch Dp = ceil64(D) and every 4D-wide one Fp = ceil64(4D); padding columns are exact zeros (the workspace is
  // cleared once per binding and no kernel ever writes a non-zero there), so the padded GEMMs / norms equal the
  // unpadded arithmetic. D = 512: Dp = D, Fp = F and nothing is padded.
  int D = 0, Dp = 0, F = 0, Fp = 0;
  bool padded = false;
  size_t dg_pad = 0, bias_tmp = 0;
  const float* word_mask = nullptr;  // device fp32 [B, T] of the current forward (word-boundary variant)
  // HuggingFace BERT encoder variant
  std::vector<BertLayerRef> bert;
  long long b_pos = 0, b_tt = 0, b_eln_g = 0, b_eln_b = 0;
  size_t b_E = 0, b_est = 0, b_x = 0, b_xb = 0, b_dA = 0, b_g1 = 0, b_g2 = 0, b_gI[2] = {0, 0}, b_dqkv = 0, b_attn = 0;
  float* bx(int i) const { return ws<float>(b_x) + (size_t)i * M * cfg.dim; }
  bf16* bxb(int i) const { return ws<bf16>(b_xb) + (size_t)i * M * cfg.dim; }
  float* last_hidden() const { return cfg.enc_type == 1 ? bx(cfg.depth) : xs_buf(2 * cfg.depth); }
  unsigned long long bsite(int id) const { return last_seed + 0x632BE59BD9B4E019ULL * (unsigned long long)(id + 1); }
  Frontend fe;  // stem3d + resnet.layer1-4

  // model structure
  long long cls_off;
  std::vector<EncLayerRef> enc;
  LinRef cat, aud;
  int cat_ld;  // padded pitch of category logits

  // workspace offsets
  size_t xs /* (2*depth+1) stream buffers */, lastb_cls, lastb_frames, logits_a, dlogits_a, logits_c, dlogits_c, acc,
      bad_token, rot;
  size_t ce_part, ce_xt, ce_lse, ce_tok;  // fused audio head: per-slot (max, sum) partials, target logits, lse, token copy
  // Device-resident step control (common.cuh StepCtl): {uint32 skip mask | uint64 dropout seed} in the workspace. With
  // dev_ctl on, EVERY sublayer's kernels are launched each step and return at once when their bit is set, and dropout
  // sites read the seed from device memory: the launch sequence no longer depends on host RNG, so the shipped
  // layer_dropout / ff_dropout config replays from one CUDA graph (svsr_lrw_step_control).
  size_t ctl = 0;
  bool dev_ctl = false;
  StepCtl sctl(int sublayer) const {  // predicate + seed of one x-transformers sublayer (-1: seed only)
    StepCtl c;
    if (!dev_ctl) return c;
    c.seed = ws<unsigned long long>(ctl + 8);
    if (sublayer >= 0) c.skip = ws<unsigned>(ctl), c.bit = 1u << sublayer;
    return c;
  }
  size_t pack_jobs;  // device table for the single-launch weight repack
  int n_pack_jobs = 0;
  bool pack_table_ready = false;
  bool pack_deferred = false;  // svsr_lrw_pack_weights() was called: the repack is enqueued by the next forward
  size_t dx, dxb[3], t_du, t_dh[2], t_dyn, t_do, t_dqkv[2];
  // parity-mode (fp32 activations, split-bf16 operands) scratch layout: offsets into a caller-provided buffer
  size_t p_patches, p_y0, p_act[6], p_s3, p_w3, p_xs[2], p_xn, p_qkv, p_o, p_h, p_u, p_lc, p_lf, p_bytes = 0;
  // forward inputs remembered for backward
  uint32_t last_skip = 0;
  unsigned long long last_seed = 0;
  int last_train = 0;
  bool fwd_done = false;
  bool bwd_stage0_done = false, bwd_stage1_done = false;

  float* xs_buf(int i) const { return ws<float>(xs) + (size_t)i * M * Dp; }
  // audio_projection + reshape + log-softmax + NLL in the GEMM epilogue (igemm.cuh, IgemmCe): softmax width % 64 == 0
  bool fused_head() const {
    return cfg.audio_vocab % 64 == 0 && cfg.audio_alignment * cfg.vq_groups * cfg.audio_vocab >= 128;
  }
};

static int engine_build(LrwEngine& e, long long nodecay_base) {
  e.params.clear(), e.buffers.clear();
  e.pc = ArenaCount(), e.bc = ArenaCount();
  e.pc.nodecay_base = nodecay_base;
  const svsr_lrw_config& c = e.cfg;
  SVSR_REQUIRE(c.B > 0 && c.T > 0 && c.H > 0 && c.W == c.H, "lrw: bad clip geometry B=%d T=%d H=%d W=%d", c.B, c.T,
               c.H, c.W);
  SVSR_REQUIRE((c.dim == 512 || c.dim == 513) && c.heads * 64 == 512,
               "lrw: encoder dim must be 512 (or 513 with the word-boundary channel) with 8 heads of 64 (got %d/%d)", c.dim,
               c.heads);
  SVSR_REQUIRE(c.depth >= 1 && c.depth <= 16, "lrw: depth %d out of range", c.depth);
  SVSR_REQUIRE(c.enc_type == 0 || c.enc_type == 1, "lrw: enc_type %d unknown", c.enc_type);
  SVSR_REQUIRE(c.ff_dropout >= 0.f && c.ff_dropout < 1.f, "lrw: ff_dropout %f out of [0,1)", c.ff_dropout);
  SVSR_REQUIRE(c.emb_dropout >= 0.f && c.emb_dropout < 1.f && c.attn_dropout >= 0.f && c.attn_dropout < 1.f,
               "lrw: emb_dropout / attn_dropout out of [0,1)");
  SVSR_REQUIRE(c.T + 1 <= 64, "lrw: sequence length %d too long for the attention core", c.T + 1);
  SVSR_REQUIRE((c.audio_alignment * c.vq_groups * c.audio_vocab) % 64 == 0,
               "lrw: audio logits per frame (%d) must be a multiple of 64",
               c.audio_alignment * c.vq_groups * c.audio_vocab);
  e.M = c.B * (c.T + 1);
  e.D = c.dim, e.Dp = (c.dim + 63) / 64 * 64, e.F = 4 * c.dim, e.Fp = (4 * c.dim + 63) / 64 * 64;
  e.padded = e.Dp != e.D;
  Bump b;

  // ---- parameters (reference state-dict names) + packed operand storage ----
  e.fe.B = c.B, e.fe.T = c.T, e.fe.H = c.H, e.fe.swish = 0;
  e.bn_eps = c.bn_eps, e.bn_momentum = c.bn_momentum;
  RC(frontend_build(e, e.fe, "stem3d.0.weight", "stem3d.1", "resnet", b));
  e.cls_off = add_param(e.params, e.pc, "cls_token", {1, 1, c.dim});
  const int D = e.D, Dp = e.Dp, inner = c.heads * 64, F = e.F, Fp = e.Fp;
  // Linear [N, K] with a K pitch of ceil64(K) in its bf16 operand copy; glu: value / gate halves of the GEGLU
  // projection start at rows 0 / ceil64(N/2) of the padded operand (only when N/2 is not a multiple of 64)
  auto plin = [&](LinRef& l, const std::string& wname, const std::string& bname, int N, int K, bool glu) {
    l.N = N, l.K = K;
    l.w = add_param(e.params, e.pc, wname, {N, K});
    l.b = bname.empty() ? -1 : add_param(e.params, e.pc, bname, {N});
    l.Kp = (K + 63) / 64 * 64;
    l.glu = (glu && e.padded) ? 1 : 0;
    l.Np = l.glu ? 2 * ((N / 2 + 63) / 64 * 64) : N;
    l.ldt = (l.Np + 63) / 64 * 64;
    l.wb = b.take((size_t)l.Np * l.Kp * 2);
    l.wt = b.take((size_t)l.Kp * l.ldt * 2);
    l.bpad = l.glu ? b.take((size_t)l.Np * 4) : 0;
  };
  if (c.enc_type == 1) {
    SVSR_REQUIRE(!e.padded, "lrw: the HuggingFace BERT encoder variant supports hidden_size 512 only");
    SVSR_REQUIRE(c.bert_intermediate > 0 && c.bert_intermediate % 64 == 0 && c.bert_max_pos >= c.T + 1,
                 "lrw: bert intermediate_size %d must be a multiple of 64 and max_position_embeddings %d >= %d",
                 c.bert_intermediate, c.bert_max_pos, c.T + 1);
    SVSR_REQUIRE(c.bert_hidden_dropout >= 0.f && c.bert_hidden_dropout < 1.f && c.bert_attn_dropout >= 0.f &&
                     c.bert_attn_dropout < 1.f, "lrw: bert dropout probabilities must be in [0,1)");
    const int I = c.bert_intermediate;
    const std::string eb = "encoder.embeddings.";
    e.b_pos = add_param(e.params, e.pc, eb + "position_embeddings.weight", {c.bert_max_pos, D});
    e.b_tt = add_param(e.params, e.pc, eb + "token_type_embeddings.weight", {2, D});
    e.b_eln_g = add_param(e.params, e.pc, eb + "LayerNorm.weight", {D});
    e.b_eln_b = add_param(e.params, e.pc, eb + "LayerNorm.bias", {D});
    e.bert.resize(c.depth);
    for (int i = 0; i < c.depth; ++i) {
      BertLayerRef& L = e.bert[i];
      const std::string pre = "encoder.encoder.layer." + std::to_string(i);
      // query | key | value adjacent: one [3D, D] operand (weights in the decayed region, biases in the other)
      L.qkv.N = 3 * D, L.qkv.K = D;
      L.qkv.w = add_param(e.params, e.pc, pre + ".attention.self.query.weight", {D, D});
      add_param(e.params, e.pc, pre + ".attention.self.key.weight", {D, D});
      add_param(e.params, e.pc, pre + ".attention.self.value.weight", {D, D});
      L.qkv.b = add_param(e.params, e.pc, pre + ".attention.self.query.bias", {D});
      add_param(e.params, e.pc, pre + ".attention.self.key.bias", {D});
      add_param(e.params, e.pc, pre + ".attention.self.value.bias", {D});
      L.qkv.ldt = 3 * D, L.qkv.Kp = D, L.qkv.Np = 3 * D;
      L.qkv.wb = b.take((size_t)3 * D * D * 2), L.qkv.wt = b.take((size_t)D * 3 * D * 2);
      plin(L.out, pre + ".attention.output.dense.weight", pre + ".attention.output.dense.bias", D, D, false);
      L.ln1_g = add_param(e.params, e.pc, pre + ".attention.output.LayerNorm.weight", {D});
      L.ln1_b = add_param(e.params, e.pc, pre + ".attention.output.LayerNorm.bias", {D});
      plin(L.inter, pre + ".intermediate.dense.weight", pre + ".intermediate.dense.bias", I, D, false);
      plin(L.outd, pre + ".output.dense.weight", pre + ".output.dense.bias", D, I, false);
      L.ln2_g = add_param(e.params, e.pc, pre + ".output.LayerNorm.weight", {D});
      L.ln2_b = add_param(e.params, e.pc, pre + ".output.LayerNorm.bias", {D});
      L.qkvbuf = b.take((size_t)e.M * 3 * D * 2), L.lse = b.take((size_t)c.B * c.heads * (c.T + 1) * 4);
      L.ctx = b.take((size_t)e.M * D * 2);
      L.A = b.take((size_t)e.M * D * 4), L.st1 = b.take((size_t)e.M * 8);
      L.x1 = b.take((size_t)e.M * D * 4), L.x1b = b.take((size_t)e.M * D * 2);
      L.pre = b.take((size_t)e.M * I * 2), L.h = b.take((size_t)e.M * I * 2);
      L.O = b.take((size_t)e.M * D * 4), L.st2 = b.take((size_t)e.M * 8);
    }
    e.b_E = b.take((size_t)e.M * D * 4), e.b_est = b.take((size_t)e.M * 8);
    e.b_x = b.take((size_t)(c.depth + 1) * e.M * D * 4), e.b_xb = b.take((size_t)(c.depth + 1) * e.M * D * 2);
    e.b_dA = b.take((size_t)e.M * D * 4);
    e.b_g1 = b.take((size_t)e.M * D * 2), e.b_g2 = b.take((size_t)e.M * D * 2);
    for (int k = 0; k < 2; ++k) e.b_gI[k] = b.take((size_t)e.M * I * 2);
    e.b_dqkv = b.take((size_t)e.M * 3 * D * 2);
    e.b_attn = b.take(attention_scratch_bytes(c.B, c.heads, c.T + 1, c.T + 1));
  }
  e.enc.resize(c.enc_type == 1 ? 0 : c.depth);
  for (int i = 0; i < (int)e.enc.size(); ++i) {
    EncLayerRef& L = e.enc[i];
    const std::string a = "encoder.layers." + std::to_string(2 * i), f = "encoder.layers." + std::to_string(2 * i + 1);
    L.g_a = add_param(e.params, e.pc, a + ".0.g", {D});
    // to_q / to_k / to_v are adjacent in the arena so that they form one [3*inner, D] operand and gradient
    L.qkv.N = 3 * inner, L.qkv.K = D, L.qkv.b = -1;
    L.qkv.w = add_param(e.params, e.pc, a + ".1.to_q.weight", {inner, D});
    add_param(e.params, e.pc, a + ".1.to_k.weight", {inner, D});
    add_param(e.params, e.pc, a + ".1.to_v.weight", {inner, D});
    L.qkv.ldt = 3 * inner, L.qkv.Kp = Dp, L.qkv.Np = 3 * inner;
    L.qkv.wb = b.take((size_t)3 * inner * Dp * 2);
    L.qkv.wt = b.take((size_t)Dp * 3 * inner * 2);
    plin(L.out, a + ".1.to_out.weight", "", D, inner, false);
    L.g_f = add_param(e.params, e.pc, f + ".0.g", {D});
    plin(L.ff1, f + ".1.ff.0.proj.weight", f + ".1.ff.0.proj.bias", 2 * F, D, true);
    plin(L.ff2, f + ".1.ff.3.weight", f + ".1.ff.3.bias", D, F, false);
    L.xn_a = b.take((size_t)e.M * Dp * 2), L.inv_a = b.take((size_t)e.M * 4);
    L.qkvbuf = b.take((size_t)e.M * 3 * inner * 2), L.obuf = b.take((size_t)e.M * inner * 2);
    L.xn_f = b.take((size_t)e.M * Dp * 2), L.inv_f = b.take((size_t)e.M * 4);
    L.hbuf = b.take((size_t)e.M * 2 * Fp * 2), L.ubuf = b.take((size_t)e.M * Fp * 2);
    if (e.padded) L.g_a_pad = b.take((size_t)Dp * 4), L.g_f_pad = b.take((size_t)Dp * 4);
  }
  plin(e.cat, "category_classifier.weight", "category_classifier.bias", c.num_labels, D, false);
  const int AGV = c.audio_alignment * c.vq_groups * c.audio_vocab;
  plin(e.aud, "audio_projection.weight", "audio_projection.bias", AGV, D, false);
  e.cat_ld = (c.num_labels + 63) / 64 * 64;

  // ---- activations ----
  frontend_alloc(e, e.fe, b);
  if (e.padded) {  // the padded weight-gradient scratch of the widest Linear can exceed the largest conv's
    size_t need = e.enc.empty() ? 0 : (size_t)e.enc[0].ff1.Np * e.enc[0].ff1.Kp * 4;
    if ((size_t)AGV * Dp * 4 > need) need = (size_t)AGV * Dp * 4;
    if (need > (size_t)9 * 512 * 512 * 4) e.wgrad_tmp = b.take(need);
    e.bias_tmp = b.take(e.enc.empty() ? 256 : (size_t)e.enc[0].ff1.Np * 4);
    e.dg_pad = b.take((size_t)2 * c.depth * Dp * 4);
  }
  const size_t n0 = (size_t)e.N * e.fe.H0 * e.fe.H0 * 64;
  const size_t n1 = (size_t)e.N * e.fe.H1 * e.fe.H1 * 64;
  e.xs = b.take((size_t)(2 * c.depth + 1) * e.M * Dp * 4);
  e.lastb_cls = b.take((size_t)c.B * Dp * 2);
  e.lastb_frames = b.take((size_t)e.N * Dp * 2);
  e.logits_a = b.take((size_t)e.N * AGV * 4);
  e.dlogits_a = b.take((size_t)e.N * AGV * 2);
  e.logits_c = b.take((size_t)c.B * e.cat_ld * 4);
  e.dlogits_c = b.take((size_t)c.B * e.cat_ld * 2);
  e.acc = b.take(8 * sizeof(double));
  e.bad_token = b.take(sizeof(int));
  e.rot = b.take((size_t)(c.T + 1) * 32 * 4);
  e.ce_part = b.take((size_t)e.N * (AGV / 64) * sizeof(float2));
  e.ce_xt = b.take((size_t)e.N * c.audio_alignment * c.vq_groups * 4);
  e.ce_lse = b.take((size_t)e.N * c.audio_alignment * c.vq_groups * 4);
  e.ce_tok = b.take((size_t)e.N * c.audio_alignment * c.vq_groups * 8);
  e.ctl = b.take(16);
  // ---- backward scratch ----
  e.dx = b.take((size_t)e.M * Dp * 4);
  for (int i = 0; i < 3; ++i) e.dxb[i] = b.take((size_t)e.M * Dp * 2);
  e.t_du = b.take((size_t)e.M * Fp * 2);
  for (int i = 0; i < 2; ++i) e.t_dh[i] = b.take((size_t)e.M * 2 * Fp * 2);
  e.t_dyn = b.take((size_t)e.M * Dp * 2);
  e.t_do = b.take((size_t)e.M * inner * 2);
  for (int i = 0; i < 2; ++i) e.t_dqkv[i] = b.take((size_t)e.M * 3 * inner * 2);
  e.pack_jobs = b.take(192 * sizeof(PackJob));
  e.ws_bytes = b.off;
  {  // parity-mode scratch (only allocated by the caller when forward_precise is used). D-wide fp32 rows have pitch
     // Dq = ceil8(D); split operands are zero-padded to Kd = ceil64(D) / Kf = ceil64(4D) columns per third
    Bump pb;
    const size_t AGVp = (size_t)AGV;
    const size_t Dq = (size_t)(D + 7) / 8 * 8, Kd = (size_t)Dp, Kf = (size_t)Fp;
    e.p_patches = pb.take(n0 * 4);
    e.p_y0 = pb.take(n0 * 4);
    for (int i = 0; i < 6; ++i) e.p_act[i] = pb.take(n1 * 4);
    const size_t I = c.enc_type == 1 ? (size_t)c.bert_intermediate : 0;  // HuggingFace BERT: intermediate_size
    size_t s3 = n0 * 3 * 2;                                   // stem patches, 192 channels
    if ((size_t)e.M * 3 * Kf * 2 > s3) s3 = (size_t)e.M * 3 * Kf * 2;
    if ((size_t)e.M * 3 * I * 2 > s3) s3 = (size_t)e.M * 3 * I * 2;
    if ((size_t)e.N * 3 * Kd * 2 > s3) s3 = (size_t)e.N * 3 * Kd * 2;
    e.p_s3 = pb.take(s3);
    size_t w3 = (size_t)512 * 9 * 1536 * 2;
    if ((size_t)2 * F * 3 * Kd * 2 > w3) w3 = (size_t)2 * F * 3 * Kd * 2;
    if ((size_t)D * 3 * Kf * 2 > w3) w3 = (size_t)D * 3 * Kf * 2;
    if (I * 3 * Kd * 2 > w3) w3 = I * 3 * Kd * 2;
    if (AGVp * 3 * Kd * 2 > w3) w3 = AGVp * 3 * Kd * 2;
    e.p_w3 = pb.take(w3);
    for (int i = 0; i < 2; ++i) e.p_xs[i] = pb.take((size_t)e.M * Dq * 4);
    e.p_xn = pb.take((size_t)e.M * Dq * 4);
    e.p_qkv = pb.take((size_t)e.M * 3 * inner * 4);
    e.p_o = pb.take((size_t)e.M * inner * 4);
    e.p_h = pb.take((size_t)e.M * (2 * (size_t)F > I ? 2 * (size_t)F : I) * 4);
    e.p_u = pb.take((size_t)e.M * ((size_t)F > I ? (size_t)F : I) * 4);
    e.p_lc = pb.take((size_t)c.B * D * 4);
    e.p_lf = pb.take((size_t)e.N * D * 4);
    e.p_bytes = pb.off;
  }
  e.decay_count = e.pc.decay;
  e.param_count = e.pc.nodecay_base + e.pc.nodecay;
  e.buffer_count = e.bc.nodecay;
  return SVSR_OK;
}

// ------------------------------------------------------------------------------------------------
static int engine_pack(LrwEngine& e, cudaStream_t s) {
  if (!e.pack_table_ready) {  // build the job table once per binding (pointers are static afterwards)
    // one clear of the whole workspace per binding: every padding row / column of the packed operands (category head
    // 500 -> 512, word-boundary 513 -> 576, ...) and of the activation pitches is zero from here on
    SVSR_CHECK_CUDA(cudaMemsetAsync(e.WS, 0, e.ws_bytes, s));
    std::vector<PackJob> jobs;
    frontend_pack_jobs(e, e.fe, jobs);
    auto lin = [&](const LinRef& l) {
      jobs.push_back({e.P + l.w, e.ws<bf16>(l.wb), e.ws<bf16>(l.wt), l.glu ? 4 : 1, l.N, l.K, l.Kp, l.ldt});
      if (l.glu) jobs.push_back({e.P + l.b, reinterpret_cast<bf16*>(e.ws<float>(l.bpad)), nullptr, 3, l.N, 1, 0, 0});
    };
    for (auto& L : e.enc) {
      lin(L.qkv), lin(L.out), lin(L.ff1), lin(L.ff2);
      if (e.padded) {
        jobs.push_back({e.P + L.g_a, reinterpret_cast<bf16*>(e.ws<float>(L.g_a_pad)), nullptr, 3, e.D, 0, 0, 0});
        jobs.push_back({e.P + L.g_f, reinterpret_cast<bf16*>(e.ws<float>(L.g_f_pad)), nullptr, 3, e.D, 0, 0, 0});
      }
    }
    for (auto& L : e.bert) lin(L.qkv), lin(L.out), lin(L.inter), lin(L.outd);
    lin(e.cat), lin(e.aud);
    SVSR_REQUIRE(jobs.size() <= 192, "lrw: too many pack jobs (%zu)", jobs.size());
    e.n_pack_jobs = (int)jobs.size();
    SVSR_CHECK_CUDA(cudaMemcpyAsync(e.ws<PackJob>(e.pack_jobs), jobs.data(), jobs.size() * sizeof(PackJob),
                                    cudaMemcpyHostToDevice, s));
    SVSR_CHECK_CUDA(cudaStreamSynchronize(s));  // `jobs` is a stack vector
    RC(rotary_table(e.ws<float>(e.rot), e.cfg.T + 1, s));
    e.pack_table_ready = true;
  }
  return pack_all_weights(e.ws<PackJob>(e.pack_jobs), e.n_pack_jobs, s);
}

// SVSR_PACK_OVERLAP=1 (off by default): the 253 MB -> 2 x 126 MB repack (~270 us) is only needed from resnet.layer1
// on, so job 0 (the stem's [64, 320] operand) goes on the caller's stream and the rest on the side stream beside the
// stem's patch gather / temporal conv / BN+GELU+pool (frontend_forward joins before the first BasicBlock). Measured: no
// gain (11.68 vs 11.65 ms per step) -- the stem kernels are HBM-bound themselves, the repack only competes with them
// for the same bandwidth -- hence opt-in. Fork and join are enqueued inside ONE forward call, so a stream capture always
// holds both.
static int engine_pack_deferred(LrwEngine& e, cudaStream_t s, bool may_overlap) {
  if (!e.pack_deferred) return SVSR_OK;
  e.pack_deferred = false;
  const char* ov = getenv("SVSR_PACK_OVERLAP");
  const char* one = getenv("SVSR_SINGLE_STREAM");
  if (!may_overlap || !e.pack_table_ready || e.n_pack_jobs < 2 || !(ov && ov[0] == '1') || (one && one[0] == '1'))
    return engine_pack(e, s);
  const PackJob* jobs = e.ws<PackJob>(e.pack_jobs);
  RC(pack_all_weights(jobs, 1, s));
  SVSR_CHECK_CUDA(cudaEventRecord(e.ev_pack_fork, s));
  SVSR_CHECK_CUDA(cudaStreamWaitEvent(e.side, e.ev_pack_fork, 0));
  RC(pack_all_weights(jobs + 1, e.n_pack_jobs - 1, e.side));
  SVSR_CHECK_CUDA(cudaEventRecord(e.ev_pack_done, e.side));
  e.pack_pending = true;
  return SVSR_OK;
}

// Linear forward / input gradient / weight gradient on the (possibly K- or N-padded) bf16 operand copies
static int lw_fwd(const LrwEngine& e, const bf16* x, int ldx, int M, const LinRef& l, void* out, int ldc, int out_fp32,
                  const void* resid, cudaStream_t s, const StepCtl& ctl = StepCtl()) {
  IgemmProblem p;
  p.ctl = ctl;
  p.a = x, p.a_N = M, p.a_C = ldx, p.cin = l.Kp, p.ntaps = 1;
  p.o_N = M;
  p.b = e.ws<bf16>(l.wb), p.b_rows = l.Np, p.b_cols = l.Kp;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = ldc;
  p.bias = l.b < 0 ? nullptr : (l.glu ? e.ws<float>(l.bpad) : e.P + l.b);
  p.resid = resid, p.resid_fp32 = 1;
  return igemm_launch(p, s);
}
static int lw_dgrad(const LrwEngine& e, const bf16* dy, int ldy, int M, const LinRef& l, void* out, int ldc,
                    cudaStream_t s, const StepCtl& ctl = StepCtl()) {
  IgemmProblem p;
  p.ctl = ctl;
  p.a = dy, p.a_N = M, p.a_C = ldy, p.cin = l.ldt, p.ntaps = 1;
  p.o_N = M;
  p.b = e.ws<bf16>(l.wt), p.b_rows = l.K, p.b_cols = l.ldt;
  p.out = out, p.out_fp32 = 0, p.ldc = ldc;
  return igemm_launch(p, s);
}
static int lw_wgrad(const LrwEngine& e, const bf16* dy, int ldy, const bf16* x, int ldx, int M, const LinRef& l,
                    cudaStream_t s, const StepCtl& ctl = StepCtl()) {
  WgradProblem p;
  p.ctl = ctl;
  p.a = dy, p.a_N = M, p.a_C = ldy, p.a_cin = l.ldt, p.ntaps = 1;
  p.b = x, p.b_C = ldx, p.n_cols = l.Kp;
  p.k_N = M;
  if (l.Kp == l.K && !l.glu) {  // rows of the arena matrix are 16-byte aligned: accumulate in place
    p.out = e.G + l.w, p.ldo = l.K, p.m_valid = l.N;
    RC(wgrad_launch(p, s));
    if (l.b >= 0) RC(colsum_bf16(dy, ldy, e.G + l.b, M, l.N, s, &ctl));
    return SVSR_OK;
  }
  // word-boundary widths (K = 513 / 2052): gradient into a zero-padded scratch, then added into the arena layout
  float* tmp = e.ws<float>(e.wgrad_tmp);
  SVSR_CHECK_CUDA(cudaMemsetAsync(tmp, 0, (size_t)l.Np * l.Kp * 4, s));
  p.out = tmp, p.ldo = l.Kp, p.m_valid = l.Np;
  RC(wgrad_launch(p, s));
  RC(unpack_linear_wgrad(tmp, e.G + l.w, l.N, l.K, l.Kp, l.glu, s));
  if (l.b >= 0) {
    if (!l.glu) return colsum_bf16(dy, ldy, e.G + l.b, M, l.N, s, &ctl);
    float* bt = e.ws<float>(e.bias_tmp);
    SVSR_CHECK_CUDA(cudaMemsetAsync(bt, 0, (size_t)l.Np * 4, s));
    RC(colsum_bf16(dy, ldy, bt, M, l.Np, s, &ctl));
    RC(unpack_linear_wgrad(bt, e.G + l.b, l.N, 1, 1, 1, s));
  }
  return SVSR_OK;
}

// audio_projection fused with reshape + log-softmax + NLL (lightning.py:168-171): mode 1 = forward (per-slot partials +
// target logits, no logits in HBM), mode 2 = backward (tile recomputed, d logits emitted in bf16). The audio tokens of
// the step were copied into the workspace by the forward ([B, T*A, G] contiguous), so backward needs no caller memory.
static int audio_head_gemm(const LrwEngine& e, int mode, const float* grad_scale, cudaStream_t s) {
  const svsr_lrw_config& c = e.cfg;
  const LinRef& l = e.aud;
  const int AGV = c.audio_alignment * c.vq_groups * c.audio_vocab;
  const long long audio_rows = (long long)e.N * c.audio_alignment * c.vq_groups;
  IgemmProblem p;
  p.a = e.ws<bf16>(e.lastb_frames), p.a_N = e.N, p.a_C = e.Dp, p.cin = l.Kp, p.ntaps = 1;
  p.o_N = e.N;
  p.b = e.ws<bf16>(l.wb), p.b_rows = l.Np, p.b_cols = l.Kp;
  p.bias = e.P + l.b;
  p.ldc = AGV;
  p.out = mode == 2 ? e.ws<bf16>(e.dlogits_a) : nullptr;
  p.ce.mode = mode;
  p.ce.T = c.T, p.ce.A = c.audio_alignment, p.ce.G = c.vq_groups, p.ce.V = c.audio_vocab;
  p.ce.AG = c.audio_alignment * c.vq_groups;
  p.ce.tokens = e.ws<long long>(e.ce_tok), p.ce.tok_stride_b = (long long)c.T * c.audio_alignment * c.vq_groups;
  p.ce.part = e.ws<float2>(e.ce_part), p.ce.xt = e.ws<float>(e.ce_xt), p.ce.lse = e.ws<float>(e.ce_lse);
  p.ce.bad_token = e.ws<int>(e.bad_token);
  p.ce.dscale = c.lambda_audio / (float)audio_rows, p.ce.grad_scale = grad_scale;
  return igemm_launch(p, s);
}

// ------------------------------------------------------------------------------------------------
// HuggingFace BERT encoder variant (BertModel(inputs_embeds=...).last_hidden_state, lightning.py:152-156): embeddings
// (+ position, + token type 0, LayerNorm, dropout) and post-LN layers (self-attention, GELU FFN). Single stream.
// ------------------------------------------------------------------------------------------------
static int blin_fwd(const LrwEngine& e, const bf16* x, int M, const LinRef& l, void* out, int out_fp32, const float* resid,
                    float drop_p, unsigned long long drop_seed, cudaStream_t s) {
  IgemmProblem p;
  p.a = x, p.a_N = M, p.a_C = l.K, p.cin = l.K, p.ntaps = 1;
  p.o_N = M;
  p.b = e.ws<bf16>(l.wb), p.b_rows = l.N, p.b_cols = l.K;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = l.N;
  p.bias = e.P + l.b;
  p.resid = resid, p.resid_fp32 = 1;
  p.drop_p = drop_p, p.drop_seed = drop_seed;
  return igemm_launch(p, s);
}
static int blin_dgrad(const LrwEngine& e, const bf16* dy, int M, const LinRef& l, void* out, int out_fp32,
                      const float* resid, cudaStream_t s) {
  IgemmProblem p;
  p.a = dy, p.a_N = M, p.a_C = l.N, p.cin = l.ldt, p.ntaps = 1;
  p.o_N = M;
  p.b = e.ws<bf16>(l.wt), p.b_rows = l.K, p.b_cols = l.ldt;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = l.K;
  p.resid = resid, p.resid_fp32 = 1;
  return igemm_launch(p, s);
}
static int bert_forward(LrwEngine& e, int train, cudaStream_t s) {
  const svsr_lrw_config& c = e.cfg;
  const int D = e.D, M = e.M, L1 = c.T + 1, I = c.bert_intermediate;
  const float pd = train ? c.bert_hidden_dropout : 0.f, pa = train ? c.bert_attn_dropout : 0.f;
  RC(bert_embed_fwd(e.xs_buf(0), e.P + e.b_pos, e.P + e.b_tt, e.ws<float>(e.b_E), M, L1, D, s));
  RC(layernorm_fwd(e.ws<float>(e.b_E), e.P + e.b_eln_g, e.P + e.b_eln_b, e.bxb(0), e.bx(0), e.ws<float>(e.b_est), M, D,
                   c.bert_ln_eps, s));
  if (pd > 0.f) RC(dropout_f32_inplace(e.bx(0), e.bxb(0), (long long)M * D, pd, e.bsite(1), s));
  for (int i = 0; i < c.depth; ++i) {
    BertLayerRef& Lb = e.bert[i];
    RC(blin_fwd(e, e.bxb(i), M, Lb.qkv, e.ws<bf16>(Lb.qkvbuf), 0, nullptr, 0.f, 0, s));
    {
      AttnProblem a;
      a.q = e.ws<bf16>(Lb.qkvbuf), a.k = a.q + D, a.v = a.q + 2 * D, a.ldq = a.ldk = a.ldv = 3 * D;
      a.B = c.B, a.H = c.heads, a.Tq = L1, a.Tk = L1, a.scale = 0.125f;
      a.o = e.ws<bf16>(Lb.ctx), a.ldo = D, a.lse = e.ws<float>(Lb.lse);
      a.drop_p = pa, a.drop_seed = e.bsite(16 * (i + 1));
      RC(attention_core_fwd(a, s));
    }
    // BertSelfOutput: LayerNorm(dropout(dense(ctx)) + x)
    RC(blin_fwd(e, e.ws<bf16>(Lb.ctx), M, Lb.out, e.ws<float>(Lb.A), 1, e.bx(i), pd, e.bsite(16 * (i + 1) + 1), s));
    RC(layernorm_fwd(e.ws<float>(Lb.A), e.P + Lb.ln1_g, e.P + Lb.ln1_b, e.ws<bf16>(Lb.x1b), e.ws<float>(Lb.x1),
                     e.ws<float>(Lb.st1), M, D, c.bert_ln_eps, s));
    // BertIntermediate (erf GELU) + BertOutput: LayerNorm(dropout(dense(h)) + x1)
    RC(blin_fwd(e, e.ws<bf16>(Lb.x1b), M, Lb.inter, e.ws<bf16>(Lb.pre), 0, nullptr, 0.f, 0, s));
    RC(gelu_fwd(e.ws<bf16>(Lb.pre), e.ws<bf16>(Lb.h), (long long)M * I, s));
    RC(blin_fwd(e, e.ws<bf16>(Lb.h), M, Lb.outd, e.ws<float>(Lb.O), 1, e.ws<float>(Lb.x1), pd, e.bsite(16 * (i + 1) + 2), s));
    RC(layernorm_fwd(e.ws<float>(Lb.O), e.P + Lb.ln2_g, e.P + Lb.ln2_b, e.bxb(i + 1), e.bx(i + 1), e.ws<float>(Lb.st2), M, D,
                     c.bert_ln_eps, s));
  }
  return SVSR_OK;
}
// dx (fp32 [M, D]) holds d loss / d last_hidden_state on entry and d loss / d inputs_embeds on exit
static int bert_backward(LrwEngine& e, float* dx, cudaStream_t s) {
  const svsr_lrw_config& c = e.cfg;
  const int D = e.D, M = e.M, L1 = c.T + 1, I = c.bert_intermediate;
  const float pd = e.last_train ? c.bert_hidden_dropout : 0.f, pa = e.last_train ? c.bert_attn_dropout : 0.f;
  float* dA = e.ws<float>(e.b_dA);
  bf16 *g1 = e.ws<bf16>(e.b_g1), *g2 = e.ws<bf16>(e.b_g2), *dh = e.ws<bf16>(e.b_gI[0]), *dpre = e.ws<bf16>(e.b_gI[1]);
  bf16* dqkv = e.ws<bf16>(e.b_dqkv);
  for (int i = c.depth - 1; i >= 0; --i) {
    BertLayerRef& Lb = e.bert[i];
    // output.LayerNorm -> dO (fp32, in dA); FFN branch gradient = dropout mask * dO
    RC(layernorm_bwd(nullptr, dx, e.ws<float>(Lb.O), e.P + Lb.ln2_g, e.ws<float>(Lb.st2), dA, 0, e.G + Lb.ln2_g,
                     e.G + Lb.ln2_b, M, D, s));
    RC(cast_scale_f32_bf16(dA, g1, (long long)M * D, 1.f, s, pd, e.bsite(16 * (i + 1) + 2)));
    RC(lw_wgrad(e, g1, D, e.ws<bf16>(Lb.h), I, M, Lb.outd, s));
    RC(blin_dgrad(e, g1, M, Lb.outd, dh, 0, nullptr, s));
    RC(gelu_bwd(e.ws<bf16>(Lb.pre), dh, dpre, (long long)M * I, s));
    RC(lw_wgrad(e, dpre, I, e.ws<bf16>(Lb.x1b), D, M, Lb.inter, s));
    RC(blin_dgrad(e, dpre, M, Lb.inter, dx, 1, dA, s));  // dx := d x1 = dO (residual) + dpre . W_inter
    // attention.output.LayerNorm -> dA; attention branch gradient = dropout mask * dA
    RC(layernorm_bwd(nullptr, dx, e.ws<float>(Lb.A), e.P + Lb.ln1_g, e.ws<float>(Lb.st1), dA, 0, e.G + Lb.ln1_g,
                     e.G + Lb.ln1_b, M, D, s));
    RC(cast_scale_f32_bf16(dA, g1, (long long)M * D, 1.f, s, pd, e.bsite(16 * (i + 1) + 1)));
    RC(lw_wgrad(e, g1, D, e.ws<bf16>(Lb.ctx), D, M, Lb.out, s));
    RC(blin_dgrad(e, g1, M, Lb.out, g2, 0, nullptr, s));  // g2 = d ctx
    {
      AttnProblem a;
      a.q = e.ws<bf16>(Lb.qkvbuf), a.k = a.q + D, a.v = a.q + 2 * D, a.ldq = a.ldk = a.ldv = 3 * D;
      a.B = c.B, a.H = c.heads, a.Tq = L1, a.Tk = L1, a.scale = 0.125f;
      a.o = e.ws<bf16>(Lb.ctx), a.ldo = D, a.lse = e.ws<float>(Lb.lse);
      a.drop_p = pa, a.drop_seed = e.bsite(16 * (i + 1));
      AttnGrads g;
      g.d_o = g2, g.dq = dqkv, g.dk = dqkv + D, g.dv = dqkv + 2 * D, g.lddq = g.lddk = g.lddv = 3 * D;
      g.scratch = e.ws<float>(e.b_attn);
      RC(attention_core_bwd(a, g, s));
    }
    RC(lw_wgrad(e, dqkv, 3 * D, e.bxb(i), D, M, Lb.qkv, s));
    RC(blin_dgrad(e, dqkv, M, Lb.qkv, dx, 1, dA, s));  // dx := d x_i = dA (residual) + dqkv . W_qkv
  }
  // embeddings: dropout, LayerNorm, the two learned tables
  if (pd > 0.f) RC(dropout_f32_inplace(dx, nullptr, (long long)M * D, pd, e.bsite(1), s));
  RC(layernorm_bwd(nullptr, dx, e.ws<float>(e.b_E), e.P + e.b_eln_g, e.ws<float>(e.b_est), dx, 0, e.G + e.b_eln_g,
                   e.G + e.b_eln_b, M, D, s));
  return bert_embed_bwd(dx, e.G + e.b_pos, e.G + e.b_tt, c.B, L1, D, s);
}

static int engine_forward(LrwEngine& e, const float* videos, const long long* tokens, long long tok_stride_b,
                          const long long* labels, const float* soft_labels, int train, uint32_t skip_mask,
                          unsigned long long dropout_seed, float* metrics, int videos_only, cudaStream_t s) {
  const svsr_lrw_config& c = e.cfg;
  const int D = e.D, Dp = e.Dp, inner = c.heads * 64, Fp = e.Fp;
  SVSR_CHECK_CUDA(cudaMemsetAsync(e.ws<uint8_t>(e.acc), 0, 8 * sizeof(double) + 256, s));  // acc + bad_token
  RC(engine_pack_deferred(e, s, true));
  // ---- stem3d + resnet.layer1-4 (lightning.py:49-54,112-117) ----
  const bf16* x = nullptr;
  RC(frontend_forward(e, e.fe, videos, train, &x, s));
  // ---- mean((2,3)) [+ word-boundary channel] + CLS concat (lightning.py:118,145-150) ----
  const int HW4 = e.fe.blocks[7].Hout * e.fe.blocks[7].Hout;
  RC(meanpool_cls(x, e.P + e.cls_off, e.xs_buf(0), c.B, c.T, HW4, 512, s, Dp));
  if (videos_only) return SVSR_OK;  // forward_videos (lightning.py:112-119) ends before the word-boundary concat
  if (e.padded) {
    SVSR_REQUIRE(e.word_mask, "lrw: dim 513 (data.use_word_boundary) needs the word_mask input");
    RC(wb_column(e.xs_buf(0), e.P + e.cls_off, e.word_mask, c.B, c.T, Dp, 512, s));
  }

  // emb_dropout_bert on cat(cls_tokens, inputs_embeds) (lightning.py:150)
  const bool dc = e.dev_ctl && c.enc_type == 0;
  const unsigned long long seed0 = dc ? 0ULL : dropout_seed;  // device control: the kernels add *seed to the site constant
  if (train && c.emb_dropout > 0.f) {
    const StepCtl cs = e.sctl(-1);
    RC(dropout_f32_inplace(e.xs_buf(0), nullptr, (long long)e.M * Dp, c.emb_dropout, seed0 + 0x3000ULL, s, &cs));
  }
  // ---- encoder (lightning.py:152-158) ----
  e.last_seed = dropout_seed;
  if (c.enc_type == 1) RC(bert_forward(e, train, s));
  for (int i = 0; i < (int)e.enc.size(); ++i) {
    EncLayerRef& L = e.enc[i];
    float* xa = e.xs_buf(2 * i);
    float* xf = e.xs_buf(2 * i + 1);
    float* xo = e.xs_buf(2 * i + 2);
    const float* g_a = e.padded ? e.ws<float>(L.g_a_pad) : e.P + L.g_a;
    const float* g_f = e.padded ? e.ws<float>(L.g_f_pad) : e.P + L.g_f;
    const StepCtl ca = e.sctl(2 * i), cf = e.sctl(2 * i + 1);
    if (!dc && (skip_mask & (1u << (2 * i)))) {
      SVSR_CHECK_CUDA(cudaMemcpyAsync(xf, xa, (size_t)e.M * Dp * 4, cudaMemcpyDeviceToDevice, s));
    } else {
      RC(rmsnorm_fwd(xa, g_a, e.ws<bf16>(L.xn_a), e.ws<float>(L.inv_a), e.M, Dp, 1e-8f, s, D, &ca));
      // q | k | v projection + rotary + softmax + PV in ONE kernel (attention_tc.cu) when the sequence fits 32 token slots;
      // SVSR_ATTN_QKV_FUSED=0 (or SVSR_ATTN_TC=0) keeps the projection GEMM and the attention core as two launches
      static const bool fuse_qkv = [] {
        const char* a = getenv("SVSR_ATTN_QKV_FUSED");
        const char* t = getenv("SVSR_ATTN_TC");
        return !(a && a[0] == '0') && !(t && t[0] == '0');
      }();
      if (fuse_qkv && c.T + 1 <= 32 && L.qkv.b < 0 && L.qkv.Kp % 64 == 0 && L.qkv.Kp <= Dp && L.qkv.Np >= 3 * inner) {
        RC(attention_qkv_tc_fwd(e.ws<bf16>(L.xn_a), Dp, e.ws<bf16>(L.qkv.wb), L.qkv.Kp, e.ws<float>(e.rot),
                                e.ws<bf16>(L.qkvbuf), e.ws<bf16>(L.obuf), c.B, c.T + 1, c.heads, c.rotary_v, s,
                                train ? c.attn_dropout : 0.f, seed0 + 0x2000ULL * (unsigned long long)(i + 1), &ca));
      } else {
        RC(lw_fwd(e, e.ws<bf16>(L.xn_a), Dp, e.M, L.qkv, e.ws<bf16>(L.qkvbuf), 3 * inner, 0, nullptr, s, ca));
        RC(attention_fwd(e.ws<bf16>(L.qkvbuf), e.ws<float>(e.rot), e.ws<bf16>(L.obuf), c.B, c.T + 1, c.heads, c.rotary_v,
                         s, train ? c.attn_dropout : 0.f, seed0 + 0x2000ULL * (unsigned long long)(i + 1), &ca));
      }
      RC(lw_fwd(e, e.ws<bf16>(L.obuf), inner, e.M, L.out, xf, Dp, 1, xa, s, ca));
      if (dc) RC(copy_if_skipped(xf, xa, (long long)e.M * Dp, ca, s));
    }
    if (!dc && (skip_mask & (1u << (2 * i + 1)))) {
      SVSR_CHECK_CUDA(cudaMemcpyAsync(xo, xf, (size_t)e.M * Dp * 4, cudaMemcpyDeviceToDevice, s));
    } else {
      RC(rmsnorm_fwd(xf, g_f, e.ws<bf16>(L.xn_f), e.ws<float>(L.inv_f), e.M, Dp, 1e-8f, s, D, &cf));
      RC(lw_fwd(e, e.ws<bf16>(L.xn_f), Dp, e.M, L.ff1, e.ws<bf16>(L.hbuf), 2 * Fp, 0, nullptr, s, cf));
      RC(geglu_fwd(e.ws<bf16>(L.hbuf), e.ws<bf16>(L.ubuf), e.M, Fp, train ? c.ff_dropout : 0.f,
                   seed0 + 0x1000ULL * (unsigned long long)i, s, &cf));
      RC(lw_fwd(e, e.ws<bf16>(L.ubuf), Fp, e.M, L.ff2, xo, Dp, 1, xf, s, cf));
      if (dc) RC(copy_if_skipped(xo, xf, (long long)e.M * Dp, cf, s));
    }
  }
  const float* last = e.last_hidden();
  RC(split_cast_last(last, e.ws<bf16>(e.lastb_cls), e.ws<bf16>(e.lastb_frames), c.B, c.T, Dp, s));

  // ---- heads + losses (lightning.py:161-174) ----
  const int AGV = c.audio_alignment * c.vq_groups * c.audio_vocab;
  RC(lw_fwd(e, e.ws<bf16>(e.lastb_cls), Dp, c.B, e.cat, e.ws<float>(e.logits_c), e.cat_ld, 1, nullptr, s));
  const long long audio_rows = (long long)e.N * c.audio_alignment * c.vq_groups;
  {  // audio_tokens[:, :T*A] (lightning.py:147) -> workspace copy [B, T*A, G] that backward can still read
    const size_t rowb = (size_t)c.T * c.audio_alignment * c.vq_groups * 8;
    SVSR_CHECK_CUDA(cudaMemcpy2DAsync(e.ws<uint8_t>(e.ce_tok), rowb, tokens, (size_t)tok_stride_b * 8, rowb, c.B,
                                      cudaMemcpyDeviceToDevice, s));
  }
  if (e.fused_head()) {
    RC(audio_head_gemm(e, 1, nullptr, s));
    RC(ce_finalize(e.ws<float2>(e.ce_part), e.ws<float>(e.ce_xt), e.ws<long long>(e.ce_tok),
                   (long long)c.T * c.audio_alignment * c.vq_groups, c.B, c.T, c.audio_alignment, c.vq_groups,
                   c.audio_vocab, e.ws<float>(e.ce_lse), e.ws<double>(e.acc), s));
  } else {  // a vocabulary that is not a multiple of 64: projection, then the separate log-softmax + NLL kernel
    RC(lw_fwd(e, e.ws<bf16>(e.lastb_frames), Dp, e.N, e.aud, e.ws<float>(e.logits_a), AGV, 1, nullptr, s));
    RC(audio_ce(e.ws<float>(e.logits_a), AGV, e.ws<long long>(e.ce_tok), (long long)c.T * c.audio_alignment * c.vq_groups,
                c.B, c.T, c.audio_alignment, c.vq_groups, c.audio_vocab, e.ws<bf16>(e.dlogits_a), e.ws<double>(e.acc),
                e.ws<int>(e.bad_token), c.lambda_audio / (float)audio_rows, s));
  }
  RC(category_ce(e.ws<float>(e.logits_c), e.cat_ld, labels, soft_labels, c.B, c.num_labels, c.label_smoothing,
                 e.ws<bf16>(e.dlogits_c), e.cat_ld, e.ws<double>(e.acc), 1.0f / (float)c.B, s, e.ws<int>(e.bad_token)));
  RC(finalize_metrics(e.ws<double>(e.acc), metrics, c.lambda_audio, c.B, audio_rows, s, e.ws<int>(e.bad_token)));
  e.last_skip = skip_mask;
  e.last_seed = dropout_seed;
  e.last_train = train;
  e.fwd_done = true;
  return SVSR_OK;
}

// ------------------------------------------------------------------------------------------------
// Parity-mode forward (see precise.cuh): fp32 activations, split-bf16 tensor-core operands, forward only.
// ------------------------------------------------------------------------------------------------
// out[rows, N] (pitch ldc) = x[rows, K] (pitch ldx) . w[N, K]^T (+ bias) (+ resid, pitch ldc). K is zero-padded to a
// multiple of 64 per split third (K = 513 / 2052 of the word-boundary variant), which adds exact zeros to the sums.
static int precise_gemm(const LrwEngine& e, uint8_t* PW, const float* x, int ldx, long long rows, int K, const float* w,
                        int N, const float* bias, const float* resid, float* out, int ldc, cudaStream_t s) {
  bf16* S3 = reinterpret_cast<bf16*>(PW + e.p_s3);
  bf16* W3 = reinterpret_cast<bf16*>(PW + e.p_w3);
  const int Kp = (K + 63) / 64 * 64;
  RC(split3_f32(x, ldx, S3, rows, K, Kp, s));
  RC(pack_linear_weight_split(w, W3, N, K, Kp, s));
  IgemmProblem p;
  p.a = S3, p.a_N = (int)rows, p.a_C = 3 * Kp, p.cin = 3 * Kp, p.ntaps = 1;
  p.o_N = (int)rows;
  p.b = W3, p.b_rows = N, p.b_cols = 3 * Kp;
  p.out = out, p.out_fp32 = 1, p.ldc = ldc;
  p.bias = bias, p.resid = resid, p.resid_fp32 = 1;
  return igemm_launch(p, s);
}
static int precise_conv(const LrwEngine& e, uint8_t* PW, const float* x, int Hin, const ConvRef& c, float* y,
                        double* bn_stats, cudaStream_t s) {
  bf16* S3 = reinterpret_cast<bf16*>(PW + e.p_s3);
  bf16* W3 = reinterpret_cast<bf16*>(PW + e.p_w3);
  RC(split3_f32(x, c.cin, S3, (long long)e.N * Hin * Hin, c.cin, c.cin, s));
  RC(pack_conv_weight_split(e.P + c.w, W3, c.cout, c.cin, c.R * c.R, s));
  IgemmProblem p;
  p.a = S3, p.a_N = e.N, p.a_H = Hin, p.a_W = Hin, p.a_C = 3 * c.cin, p.cin = 3 * c.cin, p.stride = c.stride;
  p.ntaps = c.R * c.R;
  for (int r = 0; r < c.R; ++r)
    for (int q = 0; q < c.R; ++q) {
      const int t = r * c.R + q;
      p.tap_dh[t] = r - c.pad, p.tap_dw[t] = q - c.pad, p.tap_kbase[t] = t * 3 * c.cin;
    }
  const int Ho = conv_out(Hin, c.R, c.stride, c.pad);
  p.o_N = e.N, p.OH = Ho, p.OW = Ho;
  p.b = W3, p.b_rows = c.cout, p.b_cols = c.R * c.R * 3 * c.cin;
  p.out = y, p.out_fp32 = 1, p.ldc = c.cout, p.o_H = Ho, p.o_W = Ho;
  p.bn_stats = bn_stats;
  return igemm_launch(p, s);
}
static int precise_bn_coef(const LrwEngine& e, long long rows, const BnRef& bn, int train, cudaStream_t s) {
  // batch statistics were accumulated by the conv epilogue; running buffers are NOT updated in parity mode
  return bn_finalize(e.ws<double>(bn.stats_f), rows, bn.C, e.P + bn.gamma, e.P + bn.beta, e.cfg.bn_eps,
                     e.cfg.bn_momentum, e.BUF + bn.rmean, e.BUF + bn.rvar, e.ws<float>(bn.coef), train ? 0 : -1, s);
}

static int engine_forward_precise(LrwEngine& e, uint8_t* PW, const float* videos, const long long* tokens,
                                  long long tok_stride_b, const long long* labels, const float* soft_labels,
                                  const float* word_mask, int train, uint32_t skip_mask, float* metrics,
                                  cudaStream_t s) {
  const svsr_lrw_config& c = e.cfg;
  SVSR_REQUIRE(!e.padded || word_mask, "lrw: dim 513 (data.use_word_boundary) needs the word_mask input");
  RC(engine_pack_deferred(e, s, false));
  const int D = c.dim, inner = c.heads * 64, F = 4 * D;
  const int Dq = (D + 7) / 8 * 8;  // pitch of the D-wide fp32 rows (== D for dim 512)
  auto PF = [&](size_t off) { return reinterpret_cast<float*>(PW + off); };
  SVSR_CHECK_CUDA(cudaMemsetAsync(e.ws<uint8_t>(e.fe.stats_arena), 0, e.fe.stats_arena_bytes, s));
  SVSR_CHECK_CUDA(cudaMemsetAsync(e.ws<uint8_t>(e.acc), 0, 8 * sizeof(double) + 256, s));
  RC(rotary_table(e.ws<float>(e.rot), c.T + 1, s));
  // ---- stem ----
  RC(stem_patch_f32(videos, PF(e.p_patches), c.B, c.T, c.H, c.W, s));
  {
    bf16* S3 = reinterpret_cast<bf16*>(PW + e.p_s3);
    bf16* W3 = reinterpret_cast<bf16*>(PW + e.p_w3);
    RC(split3_f32(PF(e.p_patches), 64, S3, (long long)e.N * e.fe.H0 * e.fe.H0, 64, 64, s));
    RC(pack_stem_weight_split(e.P + e.fe.stem_conv.w, W3, s));
    IgemmProblem p;
    p.a = S3, p.a_N = c.B, p.a_H = c.T, p.a_W = e.fe.H0 * e.fe.H0, p.a_C = 192, p.cin = 192;
    p.ntaps = 5;
    for (int kt = 0; kt < 5; ++kt) p.tap_dh[kt] = kt - 2, p.tap_dw[kt] = 0, p.tap_kbase[kt] = kt * 192;
    p.o_N = c.B, p.OH = c.T, p.OW = e.fe.H0 * e.fe.H0;
    p.b = W3, p.b_rows = 64, p.b_cols = 960;
    p.out = PF(e.p_y0), p.out_fp32 = 1, p.ldc = 64, p.o_H = c.T, p.o_W = e.fe.H0 * e.fe.H0;
    p.bn_stats = train ? e.ws<double>(e.fe.stem_bn.stats_f) : nullptr;
    RC(igemm_launch(p, s));
  }
  RC(precise_bn_coef(e, (long long)e.N * e.fe.H0 * e.fe.H0, e.fe.stem_bn, train, s));
  float* x = PF(e.p_act[0]);
  float* nxt = PF(e.p_act[1]);
  float *C1 = PF(e.p_act[2]), *A1 = PF(e.p_act[3]), *C2 = PF(e.p_act[4]), *CDS = PF(e.p_act[5]);
  RC(stem_bn_gelu_pool_f32(PF(e.p_y0), e.ws<float>(e.fe.stem_bn.coef), x, e.N, e.fe.H0, e.fe.H0, s));
  // ---- trunk ----
  for (auto& blk : e.fe.blocks) {
    const long long rows = (long long)e.N * blk.Hout * blk.Hout;
    RC(precise_conv(e, PW, x, blk.Hin, blk.conv1, C1, train ? e.ws<double>(blk.bn1.stats_f) : nullptr, s));
    RC(precise_bn_coef(e, rows, blk.bn1, train, s));
    RC(bn_apply_f32(C1, e.ws<float>(blk.bn1.coef), nullptr, nullptr, 1, A1, rows, blk.cout, s));
    RC(precise_conv(e, PW, A1, blk.Hout, blk.conv2, C2, train ? e.ws<double>(blk.bn2.stats_f) : nullptr, s));
    RC(precise_bn_coef(e, rows, blk.bn2, train, s));
    if (blk.ds) {
      RC(precise_conv(e, PW, x, blk.Hin, blk.convds, CDS, train ? e.ws<double>(blk.bnds.stats_f) : nullptr, s));
      RC(precise_bn_coef(e, rows, blk.bnds, train, s));
      RC(bn_apply_f32(C2, e.ws<float>(blk.bn2.coef), CDS, e.ws<float>(blk.bnds.coef), 1, nxt, rows, blk.cout, s));
    } else {
      RC(bn_apply_f32(C2, e.ws<float>(blk.bn2.coef), x, nullptr, 1, nxt, rows, blk.cout, s));
    }
    float* t = x;
    x = nxt, nxt = t;
  }
  const int HW4 = e.fe.blocks[7].Hout * e.fe.blocks[7].Hout;
  float* xa = PF(e.p_xs[0]);
  float* xb = PF(e.p_xs[1]);
  RC(meanpool_cls_f32(x, e.P + e.cls_off, xa, c.B, c.T, HW4, 512, Dq, s));
  if (e.padded) RC(wb_column(xa, e.P + e.cls_off, word_mask, c.B, c.T, Dq, 512, s));  // channel 512 = word_mask / cls[512]
  // the product path's tensors ("inputs_embeds", "last_hidden_state") have pitch Dp; their padding columns stay zero
  auto publish = [&](float* dst, const float* src) {
    return cudaMemcpy2DAsync(dst, (size_t)e.Dp * 4, src, (size_t)Dq * 4, (size_t)D * 4, (size_t)e.M,
                             cudaMemcpyDeviceToDevice, s);
  };
  SVSR_CHECK_CUDA(publish(e.xs_buf(0), xa));
  // ---- encoder (sublayer outputs ping-pong between xa and xb) ----
  for (int i = 0; i < (int)e.enc.size(); ++i) {
    EncLayerRef& L = e.enc[i];
    if (!(skip_mask & (1u << (2 * i)))) {
      RC(rmsnorm_fwd_f32(xa, e.P + L.g_a, PF(e.p_xn), e.M, D, Dq, 1e-8f, s));
      RC(precise_gemm(e, PW, PF(e.p_xn), Dq, e.M, D, e.P + L.qkv.w, 3 * inner, nullptr, nullptr, PF(e.p_qkv), 3 * inner, s));
      RC(attention_fwd_f32(PF(e.p_qkv), e.ws<float>(e.rot), PF(e.p_o), c.B, c.T + 1, c.heads, c.rotary_v, s));
      RC(precise_gemm(e, PW, PF(e.p_o), inner, e.M, inner, e.P + L.out.w, D, nullptr, xa, xb, Dq, s));
      float* t = xa;
      xa = xb, xb = t;
    }
    if (!(skip_mask & (1u << (2 * i + 1)))) {
      RC(rmsnorm_fwd_f32(xa, e.P + L.g_f, PF(e.p_xn), e.M, D, Dq, 1e-8f, s));
      RC(precise_gemm(e, PW, PF(e.p_xn), Dq, e.M, D, e.P + L.ff1.w, 2 * F, e.P + L.ff1.b, nullptr, PF(e.p_h), 2 * F, s));
      RC(geglu_fwd_f32(PF(e.p_h), PF(e.p_u), e.M, F, s));
      RC(precise_gemm(e, PW, PF(e.p_u), F, e.M, F, e.P + L.ff2.w, D, e.P + L.ff2.b, xa, xb, Dq, s));
      float* t = xa;
      xa = xb, xb = t;
    }
  }
  if (c.enc_type == 1) {
    // transformers.BertModel(inputs_embeds=...) (lightning.py:90-92,152-156): embeddings (+ position, + token type 0,
    // LayerNorm), then post-LayerNorm layers; dropouts are not applied in parity mode (it is the deterministic function)
    const int I = c.bert_intermediate, L1 = c.T + 1;
    float* x1 = PF(e.p_xn);
    RC(bert_embed_fwd(xa, e.P + e.b_pos, e.P + e.b_tt, x1, e.M, L1, D, s));
    RC(layernorm_fwd_f32(x1, e.P + e.b_eln_g, e.P + e.b_eln_b, xa, e.M, D, c.bert_ln_eps, s));
    for (int i = 0; i < c.depth; ++i) {
      BertLayerRef& Lb = e.bert[i];
      RC(precise_gemm(e, PW, xa, D, e.M, D, e.P + Lb.qkv.w, 3 * D, e.P + Lb.qkv.b, nullptr, PF(e.p_qkv), 3 * D, s));
      RC(attention_fwd_f32(PF(e.p_qkv), nullptr, PF(e.p_o), c.B, L1, c.heads, 0, s));
      RC(precise_gemm(e, PW, PF(e.p_o), D, e.M, D, e.P + Lb.out.w, D, e.P + Lb.out.b, xa, xb, D, s));
      RC(layernorm_fwd_f32(xb, e.P + Lb.ln1_g, e.P + Lb.ln1_b, x1, e.M, D, c.bert_ln_eps, s));
      RC(precise_gemm(e, PW, x1, D, e.M, D, e.P + Lb.inter.w, I, e.P + Lb.inter.b, nullptr, PF(e.p_h), I, s));
      RC(gelu_fwd_f32(PF(e.p_h), PF(e.p_u), (long long)e.M * I, s));
      RC(precise_gemm(e, PW, PF(e.p_u), I, e.M, I, e.P + Lb.outd.w, D, e.P + Lb.outd.b, x1, xb, D, s));
      RC(layernorm_fwd_f32(xb, e.P + Lb.ln2_g, e.P + Lb.ln2_b, xa, e.M, D, c.bert_ln_eps, s));
    }
  }
  SVSR_CHECK_CUDA(publish(e.last_hidden(), xa));
  // ---- heads ----
  const int AGV = c.audio_alignment * c.vq_groups * c.audio_vocab;
  RC(split_last_f32(xa, Dq, PF(e.p_lc), PF(e.p_lf), c.B, c.T, D, s));
  RC(precise_gemm(e, PW, PF(e.p_lc), D, c.B, D, e.P + e.cat.w, c.num_labels, e.P + e.cat.b, nullptr,
                  e.ws<float>(e.logits_c), e.cat_ld, s));
  RC(precise_gemm(e, PW, PF(e.p_lf), D, e.N, D, e.P + e.aud.w, AGV, e.P + e.aud.b, nullptr, e.ws<float>(e.logits_a), AGV, s));
  const long long audio_rows = (long long)e.N * c.audio_alignment * c.vq_groups;
  RC(category_ce(e.ws<float>(e.logits_c), e.cat_ld, labels, soft_labels, c.B, c.num_labels, c.label_smoothing, nullptr,
                 e.cat_ld, e.ws<double>(e.acc), 0.f, s));
  RC(audio_ce(e.ws<float>(e.logits_a), AGV, tokens, tok_stride_b, c.B, c.T, c.audio_alignment, c.vq_groups,
              c.audio_vocab, nullptr, e.ws<double>(e.acc), e.ws<int>(e.bad_token), 0.f, s));
  RC(finalize_metrics(e.ws<double>(e.acc), metrics, c.lambda_audio, c.B, audio_rows, s));
  e.fwd_done = false;  // no backward exists for this mode
  return SVSR_OK;
}

// stage 0: loss heads + encoder + mean-pool/CLS (completes the gradients of cls_token, encoder and head weights);
// stage 1: resnet.layer4 + layer3; stage 2: layer2 + layer1 + stem3d. stage < 0: everything. Each stage joins the side
// stream before returning, so a data-parallel caller can all-reduce its parameters' gradients while the next one runs.
static int engine_backward(LrwEngine& e, const float* grad_scale, int stage, cudaStream_t s) {
  if (stage <= 0) {
    SVSR_REQUIRE(e.fwd_done, "lrw backward called before (or twice after) forward");
    e.fwd_done = false;
    e.bwd_stage0_done = false;
  } else {
    SVSR_REQUIRE(e.bwd_stage0_done, "lrw backward stage 1 called before stage 0");
  }
  const svsr_lrw_config& c = e.cfg;
  const int D = e.D, Dp = e.Dp, inner = c.heads * 64, Fp = e.Fp;
  const int AGV = c.audio_alignment * c.vq_groups * c.audio_vocab;
  float* dx = e.ws<float>(e.dx);
  SideQueue sq(e, s);
  cudaStream_t w = e.wq;  // weight-gradient stream (set by SideQueue)
  bf16* T0 = e.ws<bf16>(e.fe.gbuf[0]);  // d loss / d (last block output): input of frontend_backward
  if (stage <= 0) {
  // audio head: recompute the logits tile by tile and emit d logits (bf16), scaled by the upstream d(loss_total)
  if (e.fused_head())
    RC(audio_head_gemm(e, 2, grad_scale, s));
  else if (grad_scale)
    RC(scale_bf16_by_device_scalar(e.ws<bf16>(e.dlogits_a), (long long)e.N * AGV, grad_scale, s));
  if (grad_scale)  // every other gradient is linear in the stored category-logits gradient
    RC(scale_bf16_by_device_scalar(e.ws<bf16>(e.dlogits_c), (long long)c.B * e.cat_ld, grad_scale, s));
  if (e.padded) SVSR_CHECK_CUDA(cudaMemsetAsync(e.ws<float>(e.dg_pad), 0, (size_t)2 * c.depth * Dp * 4, s));
  // ---- heads: d last_hidden_state (fp32 stream gradient), weight/bias gradients ----
  RC(sq.fork());
  RC(lw_wgrad(e, e.ws<bf16>(e.dlogits_c), e.cat_ld, e.ws<bf16>(e.lastb_cls), Dp, c.B, e.cat, w));
  RC(lw_wgrad(e, e.ws<bf16>(e.dlogits_a), AGV, e.ws<bf16>(e.lastb_frames), Dp, e.N, e.aud, w));
  {
    IgemmProblem p;  // CLS rows: dx[b, 0, :] = dlogits_c[b] . Wc
    p.a = e.ws<bf16>(e.dlogits_c), p.a_N = c.B, p.a_C = e.cat_ld, p.cin = e.cat.ldt, p.ntaps = 1;
    p.o_N = c.B, p.OH = 1, p.OW = 1;
    p.b = e.ws<bf16>(e.cat.wt), p.b_rows = D, p.b_cols = e.cat.ldt;
    p.out = dx, p.out_fp32 = 1, p.ldc = Dp, p.o_H = 1, p.o_W = c.T + 1, p.o_ow = 0;
    RC(igemm_launch(p, s));
    IgemmProblem q;  // frame rows: dx[b, 1+t, :] = dlogits_a[b, t] . Wa
    q.a = e.ws<bf16>(e.dlogits_a), q.a_N = c.B, q.a_H = 1, q.a_W = c.T, q.a_C = AGV, q.cin = AGV, q.ntaps = 1;
    q.o_N = c.B, q.OH = 1, q.OW = c.T;
    q.b = e.ws<bf16>(e.aud.wt), q.b_rows = D, q.b_cols = e.aud.ldt;
    q.out = dx, q.out_fp32 = 1, q.ldc = Dp, q.o_H = 1, q.o_W = c.T + 1, q.o_ow = 1;
    RC(igemm_launch(q, s));
  }
  int xb = 0;  // which dxb buffer holds the current bf16 copy of the stream gradient
  RC(cast_f32_to_bf16(dx, e.ws<bf16>(e.dxb[xb]), (long long)e.M * Dp, s));
  RC(sq.end_unit());

  if (c.enc_type == 1) {  // HuggingFace BERT variant: single stream (weight gradients included)
    RC(sq.join());
    e.wq = s;
    RC(bert_backward(e, dx, s));
    e.wq = sq.w0;
  }
  // ---- encoder, reversed: one unit per sublayer ----
  for (int i = (int)e.enc.size() - 1; i >= 0; --i) {
    EncLayerRef& L = e.enc[i];
    const float* g_a = e.padded ? e.ws<float>(L.g_a_pad) : e.P + L.g_a;
    const float* g_f = e.padded ? e.ws<float>(L.g_f_pad) : e.P + L.g_f;
    float* dg_a = e.padded ? e.ws<float>(e.dg_pad) + (size_t)(2 * i) * Dp : e.G + L.g_a;
    float* dg_f = e.padded ? e.ws<float>(e.dg_pad) + (size_t)(2 * i + 1) * Dp : e.G + L.g_f;
    const bool dc = e.dev_ctl;
    const unsigned long long seed0 = dc ? 0ULL : e.last_seed;
    const StepCtl ca = e.sctl(2 * i), cf = e.sctl(2 * i + 1);
    if (dc || !(e.last_skip & (1u << (2 * i + 1)))) {
      bf16* dxb = e.ws<bf16>(e.dxb[xb]);
      bf16* dxb_next = e.ws<bf16>(e.dxb[(xb + 1) % 3]);  // 3-deep: unit k-1's side work may still read its copy
      bf16* du = e.ws<bf16>(e.t_du);
      bf16* dh = e.ws<bf16>(e.t_dh[i & 1]);
      bf16* dyn = e.ws<bf16>(e.t_dyn);
      RC(sq.fork());  // dxb complete
      RC(lw_wgrad(e, dxb, Dp, e.ws<bf16>(L.ubuf), Fp, e.M, L.ff2, w, cf));
      RC(lw_dgrad(e, dxb, Dp, e.M, L.ff2, du, Fp, s, cf));
      RC(geglu_bwd(e.ws<bf16>(L.hbuf), du, dh, e.M, Fp, e.last_train ? c.ff_dropout : 0.f,
                   seed0 + 0x1000ULL * (unsigned long long)i, s, &cf));
      RC(sq.fork());  // dh complete
      RC(lw_wgrad(e, dh, 2 * Fp, e.ws<bf16>(L.xn_f), Dp, e.M, L.ff1, w, cf));
      RC(lw_dgrad(e, dh, 2 * Fp, e.M, L.ff1, dyn, Dp, s, cf));
      // (a dropped sublayer only hands the stream gradient's bf16 copy on: rmsnorm_bwd under `cf`)
      RC(rmsnorm_bwd(dyn, e.xs_buf(2 * i + 1), g_f, e.ws<float>(L.inv_f), dx, dxb_next, dg_f, e.M, Dp, 1e-8f, s, D, &cf));
      xb = (xb + 1) % 3;
      RC(sq.end_unit());
    }
    if (dc || !(e.last_skip & (1u << (2 * i)))) {
      bf16* dxb = e.ws<bf16>(e.dxb[xb]);
      bf16* dxb_next = e.ws<bf16>(e.dxb[(xb + 1) % 3]);  // 3-deep: unit k-1's side work may still read its copy
      bf16* d_o = e.ws<bf16>(e.t_do);
      bf16* dqkv = e.ws<bf16>(e.t_dqkv[i & 1]);
      bf16* dyn = e.ws<bf16>(e.t_dyn);
      RC(sq.fork());
      RC(lw_wgrad(e, dxb, Dp, e.ws<bf16>(L.obuf), inner, e.M, L.out, w, ca));
      RC(lw_dgrad(e, dxb, Dp, e.M, L.out, d_o, inner, s, ca));
      RC(attention_bwd(e.ws<bf16>(L.qkvbuf), e.ws<float>(e.rot), d_o, dqkv, c.B, c.T + 1, c.heads, c.rotary_v, s,
                       e.last_train ? c.attn_dropout : 0.f, seed0 + 0x2000ULL * (unsigned long long)(i + 1), &ca));
      RC(sq.fork());
      RC(lw_wgrad(e, dqkv, 3 * inner, e.ws<bf16>(L.xn_a), Dp, e.M, L.qkv, w, ca));
      RC(lw_dgrad(e, dqkv, 3 * inner, e.M, L.qkv, dyn, Dp, s, ca));
      RC(rmsnorm_bwd(dyn, e.xs_buf(2 * i), g_a, e.ws<float>(L.inv_a), dx, dxb_next, dg_a, e.M, Dp, 1e-8f, s, D, &ca));
      xb = (xb + 1) % 3;
      RC(sq.end_unit());
    }
  }
  if (e.padded)  // RMSNorm weight gradients: padded scratch -> arena
    for (int i = 0; i < (int)e.enc.size(); ++i) {
      RC(unpack_linear_wgrad(e.ws<float>(e.dg_pad) + (size_t)(2 * i) * Dp, e.G + e.enc[i].g_a, D, 1, 1, 0, s));
      RC(unpack_linear_wgrad(e.ws<float>(e.dg_pad) + (size_t)(2 * i + 1) * Dp, e.G + e.enc[i].g_f, D, 1, 1, 0, s));
    }

  // ---- emb_dropout, mean pool / CLS ----
  if (e.last_train && c.emb_dropout > 0.f) {
    const StepCtl cs = e.sctl(-1);
    RC(dropout_f32_inplace(dx, nullptr, (long long)e.M * Dp, c.emb_dropout, (e.dev_ctl ? 0ULL : e.last_seed) + 0x3000ULL, s,
                           &cs));
  }
  const int HW4 = e.fe.blocks[7].Hout * e.fe.blocks[7].Hout;
  RC(meanpool_cls_bwd(dx, T0, e.G + e.cls_off, c.B, c.T, HW4, 512, s, Dp));
  if (e.padded) RC(wb_column_bwd(dx, e.G + e.cls_off, c.B, c.T, Dp, 512, s));
  e.bwd_stage0_done = true;
  if (stage == 0) return sq.join();
  }  // stage <= 0

  // ---- resnet trunk + stem ----
  if (stage == 1) {  // layer4 + layer3: 10.5 M of the trunk's 11.2 M parameters
    e.bwd_stage1_done = true;
    return frontend_backward(e, e.fe, sq, s, 7, 4, false);
  }
  if (stage == 2) {
    SVSR_REQUIRE(e.bwd_stage1_done, "lrw backward stage 2 called before stage 1");
    e.bwd_stage1_done = false;
    return frontend_backward(e, e.fe, sq, s, 3, 0, true);
  }
  return frontend_backward(e, e.fe, sq, s);
}

}  // namespace svsr

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace svsr;

extern "C" {

int svsr_lrw_create(const svsr_lrw_config* cfg, void** handle) {
  SVSR_REQUIRE(cfg && handle, "lrw_create: null argument");
  LrwEngine* e = new LrwEngine();
  e->cfg = *cfg;
  int rc = engine_build(*e, 0);  // sizing pass: how large is the decayed region?
  if (!rc) rc = engine_build(*e, e->decay_count);
  if (rc) {
    delete e;
    return rc;
  }
  *handle = e;
  return SVSR_OK;
}
int svsr_lrw_destroy(void* h) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  if (e) engine_base_destroy(*e);
  delete e;
  return SVSR_OK;
}
int64_t svsr_lrw_param_count(void* h) { return static_cast<LrwEngine*>(h)->param_count; }
int64_t svsr_lrw_decay_count(void* h) { return static_cast<LrwEngine*>(h)->decay_count; }
int64_t svsr_lrw_buffer_count(void* h) { return static_cast<LrwEngine*>(h)->buffer_count; }
int64_t svsr_lrw_workspace_bytes(void* h) { return (int64_t)static_cast<LrwEngine*>(h)->ws_bytes; }
int svsr_lrw_num_params(void* h) { return (int)static_cast<LrwEngine*>(h)->params.size(); }
int svsr_lrw_num_buffers(void* h) { return (int)static_cast<LrwEngine*>(h)->buffers.size(); }

int svsr_lrw_param_info(void* h, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset, int* decay) {
  return tensor_info(static_cast<LrwEngine*>(h)->params, i, name, ndim, shape, offset, decay);
}
int svsr_lrw_buffer_info(void* h, int i, const char** name, int* ndim, int64_t* shape, int64_t* offset) {
  return tensor_info(static_cast<LrwEngine*>(h)->buffers, i, name, ndim, shape, offset, nullptr);
}
int svsr_lrw_bind(void* h, float* params, float* grads, float* buffers, void* workspace, int64_t workspace_bytes) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(params && grads && buffers && workspace, "lrw_bind: null pointer");
  SVSR_REQUIRE((size_t)workspace_bytes >= e->ws_bytes, "lrw_bind: workspace too small (%lld < %zu)",
               (long long)workspace_bytes, e->ws_bytes);
  SVSR_REQUIRE(((uintptr_t)workspace & 1023) == 0 && ((uintptr_t)params & 15) == 0 && ((uintptr_t)grads & 15) == 0,
               "lrw_bind: workspace must be 1024-byte aligned, arenas 16-byte aligned");
  e->pack_table_ready = false;
  e->pack_deferred = false, e->pack_pending = false;
  e->dev_ctl = false;
  return engine_base_bind(*e, params, grads, buffers, workspace);
}
int svsr_lrw_pack_weights(void* h, void* stream) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  if (!e->pack_table_ready) return engine_pack(*e, static_cast<cudaStream_t>(stream));  // first call of a binding
  e->pack_deferred = true;  // enqueued by the next forward, overlapped with its stem (engine_pack_deferred)
  return SVSR_OK;
}
int svsr_lrw_forward(void* h, const float* videos, const int64_t* tokens, int64_t tok_stride_b, const int64_t* labels,
                     const float* soft_labels, const float* word_mask, int train, uint32_t skip_mask,
                     uint64_t dropout_seed, float* metrics, void* stream) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  e->word_mask = word_mask;
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  SVSR_REQUIRE(videos && tokens && metrics, "lrw_forward: null input");
  SVSR_REQUIRE(tok_stride_b >= (int64_t)e->cfg.T * e->cfg.audio_alignment * e->cfg.vq_groups,
               "lrw_forward: audio_tokens has fewer than T*alignment rows per clip");
  return engine_forward(*e, videos, reinterpret_cast<const long long*>(tokens), tok_stride_b,
                        reinterpret_cast<const long long*>(labels), soft_labels, train, skip_mask, dropout_seed, metrics, 0,
                        static_cast<cudaStream_t>(stream));
}
int64_t svsr_lrw_precise_workspace_bytes(void* h) { return (int64_t)static_cast<LrwEngine*>(h)->p_bytes; }
int svsr_lrw_forward_precise(void* h, void* precise_ws, int64_t precise_ws_bytes, const float* videos,
                             const int64_t* tokens, int64_t tok_stride_b, const int64_t* labels,
                             const float* soft_labels, const float* word_mask, int train, uint32_t skip_mask,
                             float* metrics, void* stream) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  SVSR_REQUIRE(precise_ws && (size_t)precise_ws_bytes >= e->p_bytes && ((uintptr_t)precise_ws & 1023) == 0,
               "lrw_forward_precise: scratch missing, too small or not 1024-byte aligned");
  SVSR_REQUIRE(videos && tokens && metrics, "lrw_forward_precise: null input");
  SVSR_REQUIRE(tok_stride_b >= (int64_t)e->cfg.T * e->cfg.audio_alignment * e->cfg.vq_groups,
               "lrw_forward_precise: audio_tokens has fewer than T*alignment rows per clip");
  return engine_forward_precise(*e, static_cast<uint8_t*>(precise_ws), videos, reinterpret_cast<const long long*>(tokens),
                                tok_stride_b, reinterpret_cast<const long long*>(labels), soft_labels, word_mask, train,
                                skip_mask, metrics, static_cast<cudaStream_t>(stream));
}
int svsr_lrw_forward_videos(void* h, const float* videos, int train, void* stream) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  SVSR_REQUIRE(videos, "lrw_forward_videos: null input");
  return engine_forward(*e, videos, nullptr, 0, nullptr, nullptr, train, 0, 0, nullptr, 1,
                        static_cast<cudaStream_t>(stream));
}
int svsr_lrw_backward(void* h, const float* grad_scale, void* stream) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  return engine_backward(*e, grad_scale, -1, static_cast<cudaStream_t>(stream));
}
int svsr_lrw_backward_stage(void* h, const float* grad_scale, int stage, void* stream) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  SVSR_REQUIRE(stage >= 0 && stage <= 2, "lrw_backward_stage: stage must be 0, 1 or 2");
  return engine_backward(*e, grad_scale, stage, static_cast<cudaStream_t>(stream));
}
// The step never writes the audio logits to HBM (fused head); this materialises them once, on request, from the last
// forward's hidden states: fp32 [B*T, A*G*V] readable through svsr_lrw_tensor("logits_audio").
int svsr_lrw_logits_audio(void* h, void* stream) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  const svsr_lrw_config& c = e->cfg;
  const int AGV = c.audio_alignment * c.vq_groups * c.audio_vocab;
  return lw_fwd(*e, e->ws<bf16>(e->lastb_frames), e->Dp, e->N, e->aud, e->ws<float>(e->logits_a), AGV, 1, nullptr,
                static_cast<cudaStream_t>(stream));
}
// Device-resident step control (StepCtl): mode 1 writes {skip_mask, dropout_seed} to the engine's control words on
// `stream` and switches the engine to predicated launches (every sublayer's kernels are enqueued each step; a dropped
// sublayer's return at once) -- the forward's skip_mask / dropout_seed ARGUMENTS are then ignored, so a captured CUDA
// graph replays any step of the layer_dropout / ff_dropout config. mode 0 returns to host-valued control.
int svsr_lrw_step_control(void* h, int mode, uint32_t skip_mask, uint64_t dropout_seed, void* stream) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  SVSR_REQUIRE(mode == 0 || e->cfg.enc_type == 0, "lrw_step_control: the HuggingFace encoder variant is host controlled");
  e->dev_ctl = mode != 0;
  if (!e->dev_ctl) return SVSR_OK;
  return set_step_ctl(e->ws<unsigned>(e->ctl), e->ws<unsigned long long>(e->ctl + 8), skip_mask,
                      (unsigned long long)dropout_seed, static_cast<cudaStream_t>(stream));
}
int svsr_lrw_early_grad_region(void* h, int64_t* begin, int64_t* end) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  *begin = e->cls_off, *end = e->decay_count;
  return SVSR_OK;
}
int svsr_lrw_tensor(void* h, const char* name, void** ptr, int64_t* numel, int* dtype) {
  LrwEngine* e = static_cast<LrwEngine*>(h);
  SVSR_REQUIRE(e->WS, "lrw: bind() first");
  const svsr_lrw_config& c = e->cfg;
  const std::string n(name);
  const int AGV = c.audio_alignment * c.vq_groups * c.audio_vocab;
  auto set = [&](size_t off, int64_t ne, int dt) {
    *ptr = e->WS + off, *numel = ne, *dtype = dt;
    return SVSR_OK;
  };
  if (n == "last_hidden_state") {
    *ptr = e->last_hidden(), *numel = (int64_t)e->M * e->Dp, *dtype = 0;  // row pitch Dp = ceil64(dim)
    return SVSR_OK;
  }
  if (n == "inputs_embeds") {
    *ptr = e->xs_buf(0), *numel = (int64_t)e->M * e->Dp, *dtype = 0;
    return SVSR_OK;
  }
  if (n == "logits_audio") return set(e->logits_a, (int64_t)e->N * AGV, 0);  // filled by svsr_lrw_logits_audio()
  if (n == "logits_category") return set(e->logits_c, (int64_t)c.B * e->cat_ld, 0);
  if (n == "stem_conv") return set(e->fe.y0, (int64_t)e->N * e->fe.H0 * e->fe.H0 * 64, 1);
  if (n == "stem_out") return set(e->fe.x1, (int64_t)e->N * e->fe.H1 * e->fe.H1 * 64, 1);
  if (n == "bad_token") return set(e->bad_token, 1, 3);
  if (n.rfind("block", 0) == 0 && n.size() >= 6) {  // "block<i>.out|c1|a1|c2"
    const int bi = n[5] - '0';
    SVSR_REQUIRE(bi >= 0 && bi < 8 && n.size() > 7, "lrw_tensor: bad block tensor %s", name);
    const BlockRef& blk = e->fe.blocks[bi];
    const int64_t ne = (int64_t)e->N * blk.Hout * blk.Hout * blk.cout;
    const std::string f = n.substr(7);
    if (f == "out") return set(blk.out, ne, 1);
    if (f == "c1") return set(blk.c1, ne, 1);
    if (f == "a1") return set(blk.a1, ne, 1);
    if (f == "c2") return set(blk.c2, ne, 1);
  }
  set_last_error("lrw_tensor: unknown tensor '%s'", name);
  return SVSR_ERR_INVALID;
}

}  // extern "C"
