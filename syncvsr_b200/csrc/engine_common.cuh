// Shared pieces of the native step executors (LRW word-level: engine.cu, LRS sentence-level: engine_lrs.cu):
// flat parameter / gradient / buffer arenas with the reference's state-dict names, bump-allocated workspace, the
// launch helpers for Linear / Conv2d / BatchNorm forward+backward, the two-stream backward scheduler and the visual
// frontend (3-D conv stem + ResNet-18 trunk) that both models run:
//   LRW  stem3d + resnet.layer1-4, GELU / ReLU            LRW/video/src/lightning.py:49-55,112-119
//   LRS  frontend3D + trunk, Swish everywhere             LRS/video/espnet/nets/pytorch_backend/backbones/conv3d_extractor.py:19-48,
//                                                         backbones/modules/resnet.py:45-177
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/svsr.h"
#include "common.cuh"
#include "elementwise.cuh"
#include "igemm.cuh"
#include "wgrad.cuh"
#include "stem_direct.cuh"

namespace svsr {

#define RC(expr)              \
  do {                        \
    int _rc = (expr);         \
    if (_rc) return _rc;      \
  } while (0)

typedef __nv_bfloat16 bf16;

struct ParamInfo {
  std::string name;
  int ndim;
  long long shape[5];
  long long offset;  // elements into the fp32 arena
  long long numel;
  int decay;  // AdamW weight decay applies (ndim >= 2; lightning.py:217-219)
};

struct BnRef {
  long long gamma, beta;  // param arena offsets
  long long rmean, rvar;  // buffer arena offsets
  int C;
  size_t coef, kcoef, stats_f, stats_b;  // workspace offsets
};

struct ConvRef {
  long long w;  // param arena offset
  int cin, cout, R, stride, pad;
  size_t wf, wd;  // packed bf16 operands (fprop / dgrad)
};

struct BlockRef {
  int cin, cout, stride, Hin, Hout;
  bool ds;
  ConvRef conv1, conv2, convds;
  BnRef bn1, bn2, bnds;
  size_t c1, a1, c2, out, cds;  // saved activations (bf16)
};

struct LinRef {
  long long w, b;  // param arena offsets (b < 0: no bias)
  int N, K;
  size_t wb, wt;  // bf16 [N, K] and transposed [K, ldt]
  int ldt;
  // LRW word-boundary variant (engine.cu): K pitch of wb, rows of the padded operand, GEGLU row remap, padded bias
  int Kp = 0, Np = 0, glu = 0;
  size_t bpad = 0;
};

// The parameter arena is [decayed (ndim >= 2) | non-decayed]; `nodecay_base` is where the second region starts
// (found by a first sizing pass of engine_build). Buffers use one region.
struct ArenaCount {
  long long decay = 0, nodecay = 0, nodecay_base = 0;
};

struct EngineBase {
  std::vector<ParamInfo> params;
  std::vector<ParamInfo> buffers;
  long long param_count = 0, buffer_count = 0, decay_count = 0;
  ArenaCount pc, bc;
  size_t ws_bytes = 0;
  int N = 0;  // frames = B*T (the image batch of the 2-D trunk)
  float bn_eps = 1e-5f, bn_momentum = 0.1f;  // torch.nn.BatchNorm defaults

  // bound storage
  float* P = nullptr;
  float* G = nullptr;
  float* BUF = nullptr;
  uint8_t* WS = nullptr;

  size_t wgrad_tmp = 0;  // fp32 scratch of the conv weight-gradient GEMM (largest [R*S*Cin, Cout])
  // weight-gradient side stream (backward): forked from / joined to the caller's stream with events
  cudaStream_t side = nullptr;
  cudaStream_t wq = nullptr;  // where weight-gradient work goes for the current backward: `side`, or the caller's
                              // stream when SVSR_SINGLE_STREAM=1 (hazard check: both orders must give the same gradients)
  cudaEvent_t ev_fork[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_done[4] = {nullptr, nullptr, nullptr, nullptr};
  int fork_idx = 0;
  // weight repack left running on the side stream while the stem computes (engine.cu: engine_pack_overlapped)
  cudaEvent_t ev_pack_fork = nullptr, ev_pack_done = nullptr;
  bool pack_pending = false;

  template <class T>
  T* ws(size_t off) const {
    return reinterpret_cast<T*>(WS + off);
  }
};

// The visual frontend shared by both models: stem (Conv3d as 7x7/s2 patch gather + 5-tap temporal implicit GEMM,
// BatchNorm3d, GELU|Swish, 3x3/s2 max-pool) and the eight BasicBlocks of ResNet-18 (ReLU|Swish).
struct Frontend {
  int B = 0, T = 0, H = 0;  // clips, frames per clip, crop size
  int H0 = 0, H1 = 0;       // stem conv output size, pooled size (layer1 input)
  int swish = 0;            // 0: GELU stem + ReLU trunk (LRW); 1: Swish everywhere (LRS)
  ConvRef stem_conv;
  BnRef stem_bn;
  BlockRef blocks[8];
  size_t patches = 0, y0 = 0, x1 = 0, argmax = 0, gbuf[9] = {0}, stem_dz = 0;
  // No patch tensor (default; SVSR_STEM_DIRECT=0 restores it): the stem's forward and weight-gradient kernels build the
  // 7x7/s2 window rows in shared memory from `vid`, a bf16 copy of the clip batch kept for backward (stem_direct.cu)
  bool direct = false;
  size_t vid = 0;
  size_t stats_arena = 0, stats_arena_bytes = 0;  // fp64 BN statistic accumulators (forward + backward slot per BN)
};

namespace {

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  }
};

long long add_param(std::vector<ParamInfo>& v, ArenaCount& count, const std::string& name,
                    std::initializer_list<long long> shape) {
  ParamInfo p;
  p.name = name;
  p.ndim = (int)shape.size();
  p.numel = 1;
  int i = 0;
  for (long long s : shape) p.shape[i++] = s, p.numel *= s;
  for (; i < 5; ++i) p.shape[i] = 1;
  p.decay = p.ndim >= 2;
  const long long padded = (p.numel + 3) & ~3LL;  // keep every tensor 16-byte aligned inside the arena
  if (p.decay) {
    p.offset = count.decay;
    count.decay += padded;
  } else {
    p.offset = count.nodecay_base + count.nodecay;
    count.nodecay += padded;
  }
  v.push_back(p);
  return p.offset;
}

void add_bn(EngineBase& e, BnRef& bn, const std::string& prefix, int C, Bump& b) {
  bn.C = C;
  bn.gamma = add_param(e.params, e.pc, prefix + ".weight", {C});
  bn.beta = add_param(e.params, e.pc, prefix + ".bias", {C});
  bn.rmean = add_param(e.buffers, e.bc, prefix + ".running_mean", {C});
  bn.rvar = add_param(e.buffers, e.bc, prefix + ".running_var", {C});
  bn.coef = b.take(4 * C * sizeof(float));
  bn.kcoef = b.take(2 * C * sizeof(float));
}

void add_conv(EngineBase& e, ConvRef& c, const std::string& name, int cin, int cout, int R, int stride, int pad,
              Bump& b) {
  c.cin = cin, c.cout = cout, c.R = R, c.stride = stride, c.pad = pad;
  c.w = add_param(e.params, e.pc, name, {cout, cin, R, R});
  c.wf = b.take((size_t)cout * R * R * cin * 2);
  c.wd = b.take((size_t)cin * R * R * cout * 2);
}

void add_linear(EngineBase& e, LinRef& l, const std::string& wname, const std::string& bname, int N, int K, Bump& b) {
  l.N = N, l.K = K;
  l.w = add_param(e.params, e.pc, wname, {N, K});
  l.b = bname.empty() ? -1 : add_param(e.params, e.pc, bname, {N});
  l.ldt = (N + 63) / 64 * 64;
  l.wb = b.take((size_t)N * K * 2);
  l.wt = b.take((size_t)K * l.ldt * 2);
}

int conv_out(int h, int k, int s, int p) { return (h + 2 * p - k) / s + 1; }


// ------------------------------------------------------------------------------------------------
// small launch helpers
// ------------------------------------------------------------------------------------------------
static int linear_fwd(const EngineBase& e, const bf16* x, int M, const LinRef& l, void* out, int ldc, int out_fp32,
                      const void* resid, int resid_fp32, cudaStream_t s) {
  IgemmProblem p;
  p.a = x, p.a_N = M, p.a_C = l.K, p.cin = l.K, p.ntaps = 1;
  p.o_N = M;
  p.b = e.ws<bf16>(l.wb), p.b_rows = l.N, p.b_cols = l.K;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = ldc;
  p.bias = l.b >= 0 ? e.P + l.b : nullptr;
  p.resid = resid, p.resid_fp32 = resid_fp32;
  return igemm_launch(p, s);
}
// dx[M, K] = dy[M, N(ld = ldy)] . W   (uses the transposed operand copy)
static int linear_dgrad(const EngineBase& e, const bf16* dy, int ldy, int M, const LinRef& l, void* out, int ldc,
                        int out_fp32, cudaStream_t s) {
  IgemmProblem p;
  p.a = dy, p.a_N = M, p.a_C = ldy, p.cin = l.ldt, p.ntaps = 1;
  p.o_N = M;
  p.b = e.ws<bf16>(l.wt), p.b_rows = l.K, p.b_cols = l.ldt;
  p.out = out, p.out_fp32 = out_fp32, p.ldc = ldc;
  return igemm_launch(p, s);
}
static int linear_wgrad(const EngineBase& e, const bf16* dy, int ldy, const bf16* x, int M, const LinRef& l,
                        cudaStream_t s) {
  WgradProblem p;
  p.a = dy, p.a_N = M, p.a_C = ldy, p.a_cin = l.ldt, p.ntaps = 1;
  p.b = x, p.b_C = l.K, p.n_cols = l.K;
  p.k_N = M;
  p.out = e.G + l.w, p.ldo = l.K, p.m_valid = l.N;
  RC(wgrad_launch(p, s));
  if (l.b >= 0) RC(colsum_bf16(dy, ldy, e.G + l.b, M, l.N, s));
  return SVSR_OK;
}

static int conv_fwd(const EngineBase& e, const bf16* x, int Hin, const ConvRef& c, bf16* y, double* bn_stats,
                    cudaStream_t s) {
  IgemmProblem p;
  p.a = x, p.a_N = e.N, p.a_H = Hin, p.a_W = Hin, p.a_C = c.cin, p.cin = c.cin, p.stride = c.stride;
  p.ntaps = c.R * c.R;
  for (int r = 0; r < c.R; ++r)
    for (int q = 0; q < c.R; ++q) {
      const int t = r * c.R + q;
      p.tap_dh[t] = r - c.pad, p.tap_dw[t] = q - c.pad, p.tap_kbase[t] = t * c.cin;
    }
  const int Ho = conv_out(Hin, c.R, c.stride, c.pad);
  p.o_N = e.N, p.OH = Ho, p.OW = Ho;
  p.b = e.ws<bf16>(c.wf), p.b_rows = c.cout, p.b_cols = c.R * c.R * c.cin;
  p.out = y, p.ldc = c.cout, p.o_H = Ho, p.o_W = Ho;
  p.bn_stats = bn_stats;
  return igemm_launch(p, s);
}
// relu_mask (optional): the ReLU output the gradient flows into (bf16, dx's geometry) -- applied in the GEMM epilogue
static int conv_dgrad(const EngineBase& e, const bf16* dy, int Hin, const ConvRef& c, bf16* dx, const bf16* resid,
                      cudaStream_t s, const bf16* relu_mask = nullptr) {
  if (relu_mask)
    return svsr_conv2d_dgrad_masked(dy, e.ws<bf16>(c.wd), dx, resid, e.N, Hin, Hin, c.cin, c.cout, c.R, c.R, c.stride,
                                    c.pad, relu_mask, s);
  return svsr_conv2d_dgrad(dy, e.ws<bf16>(c.wd), dx, resid, e.N, Hin, Hin, c.cin, c.cout, c.R, c.R, c.stride, c.pad, 0,
                           s);
}
// Input gradient with the BatchNorm-backward reduction of its consumer(s) fused into the epilogue (igemm.cuh,
// IgemmBnBwd): dx = (W^T dy + resid) * mask; mask from `relu_mask` (> 0) or, self_mask, from bn_a's own output sign.
static int conv_dgrad_bnb(const EngineBase& e, const bf16* dy, int Hin, const ConvRef& c, bf16* dx, const bf16* resid,
                          const bf16* relu_mask, int self_mask, const bf16* ca, const BnRef& bna, const bf16* cb,
                          const BnRef* bnb, cudaStream_t s) {
  return svsr_conv2d_dgrad_bnbwd(dy, e.ws<bf16>(c.wd), dx, resid, e.N, Hin, Hin, c.cin, c.cout, c.R, c.R, c.stride, c.pad,
                                 relu_mask, self_mask, ca, e.ws<float>(bna.coef), e.ws<double>(bna.stats_b), cb,
                                 bnb ? e.ws<float>(bnb->coef) : nullptr, bnb ? e.ws<double>(bnb->stats_b) : nullptr, s);
}
// BatchNorm backward whose reduction already happened in the producing GEMM's epilogue: finalize + apply on the
// pre-masked gradient g (dc = scale * (g - k1 - xhat * k2))
static int bn_bwd_prereduced(const EngineBase& e, const bf16* g, const bf16* c, long long rows, const BnRef& bn, bf16* dc,
                             cudaStream_t s) {
  RC(bn_bwd_finalize(e.ws<double>(bn.stats_b), rows, bn.C, e.G + bn.gamma, e.G + bn.beta, e.ws<float>(bn.kcoef), s));
  return bn_bwd_apply(g, nullptr, c, e.ws<float>(bn.coef), e.ws<float>(bn.kcoef), dc, nullptr, rows, bn.C, 0, s);
}

static int conv_wgrad(const EngineBase& e, const bf16* x, int Hin, const bf16* dy, const ConvRef& c, cudaStream_t s) {
  float* tmp = e.ws<float>(e.wgrad_tmp);
  const size_t n = (size_t)c.R * c.R * c.cin * c.cout;
  SVSR_CHECK_CUDA(cudaMemsetAsync(tmp, 0, n * 4, s));
  WgradProblem p;
  p.a = x, p.a_N = e.N, p.a_H = Hin, p.a_W = Hin, p.a_C = c.cin, p.a_cin = c.cin, p.a_stride = c.stride;
  p.ntaps = c.R * c.R;
  for (int r = 0; r < c.R; ++r)
    for (int q = 0; q < c.R; ++q) p.tap_dh[r * c.R + q] = r - c.pad, p.tap_dw[r * c.R + q] = q - c.pad;
  const int Ho = conv_out(Hin, c.R, c.stride, c.pad);
  p.b = dy, p.b_C = c.cout, p.n_cols = c.cout;
  p.k_N = e.N, p.k_H = Ho, p.k_W = Ho;
  p.out = tmp, p.ldo = c.cout;
  RC(wgrad_launch(p, s));
  return unpack_conv_wgrad(tmp, e.G + c.w, c.cout, c.cin, c.R, c.R, s);
}

// batch statistics were accumulated by the producing conv's epilogue (IgemmProblem::bn_stats)
static int bn_fwd(const EngineBase& e, const bf16* x, long long rows, const BnRef& bn, int train, cudaStream_t s) {
  (void)x;
  return bn_finalize(e.ws<double>(bn.stats_f), rows, bn.C, e.P + bn.gamma, e.P + bn.beta, e.bn_eps,
                     e.bn_momentum, e.BUF + bn.rmean, e.BUF + bn.rvar, e.ws<float>(bn.coef), train ? 1 : -1, s);
}
// full BN backward: returns dc (may alias nothing); optionally emits the relu-masked upstream gradient
static int bn_bwd(const EngineBase& e, const bf16* dout, const bf16* relu_ref, const bf16* c, long long rows,
                  const BnRef& bn, bf16* dc, bf16* gmask_out, cudaStream_t s, int self_mask = 0,
                  const bf16* sw_res = nullptr, const float* sw_rcoef = nullptr) {
  RC(bn_bwd_reduce(dout, relu_ref, c, e.ws<float>(bn.coef), rows, bn.C, e.ws<double>(bn.stats_b), self_mask, s, sw_res,
                   sw_rcoef));
  // (bn_bwd_finalize is folded into the apply launch: one kernel less per BatchNorm on the backward chain)
  return bn_bwd_apply(dout, relu_ref, c, e.ws<float>(bn.coef), nullptr, dc, gmask_out, rows, bn.C, self_mask, s, sw_res,
                      sw_rcoef, e.ws<double>(bn.stats_b), e.G + bn.gamma, e.G + bn.beta);
}

// Backward runs on two streams: `s` carries the critical chain (dgrad GEMMs, BatchNorm / attention / norm backward),
// `e.side` carries every weight-gradient GEMM (+ bias column sums, gradient unpack). The side work only reads
// tensors the chain has finished (fork event) and the chain never overwrites a tensor the side stream may still be
// reading: the hazard buffers (dc*, dxb, dh, dqkv) are double buffered and the chain waits for the side stream's
// unit k-2 before starting unit k (bounded lag). HBM-bound BN kernels thereby overlap tensor-bound wgrad kernels.
// The caller's stream waits for a repack that was left running on the side stream (no-op otherwise).
static int pack_join(EngineBase& e, cudaStream_t s) {
  if (e.pack_pending) {
    SVSR_CHECK_CUDA(cudaStreamWaitEvent(s, e.ev_pack_done, 0));
    e.pack_pending = false;
  }
  return SVSR_OK;
}

struct SideQueue {
  EngineBase& e;
  cudaStream_t s;
  int unit = 0;
  int rc = SVSR_OK;
  cudaStream_t w0;  // the weight-gradient stream chosen for this backward
  SideQueue(EngineBase& e_, cudaStream_t s_) : e(e_), s(s_) {
    const char* one = getenv("SVSR_SINGLE_STREAM");
    e.wq = w0 = (one && one[0] == '1') ? s_ : e_.side;
  }
  // everything enqueued on `s` so far becomes visible to the side stream
  int fork() {
    cudaEvent_t ev = e.ev_fork[e.fork_idx++ & 3];
    SVSR_CHECK_CUDA(cudaEventRecord(ev, s));
    SVSR_CHECK_CUDA(cudaStreamWaitEvent(e.wq, ev, 0));
    return SVSR_OK;
  }
  // close unit `unit` on the side stream and make the chain wait for unit-1 (so unit-2's buffers are reusable next)
  int end_unit() {
    SVSR_CHECK_CUDA(cudaEventRecord(e.ev_done[unit & 3], e.wq));
    if (unit >= 1) SVSR_CHECK_CUDA(cudaStreamWaitEvent(s, e.ev_done[(unit - 1) & 3], 0));
    ++unit;
    return SVSR_OK;
  }
  int join() {
    cudaEvent_t ev = e.ev_done[unit & 3];
    SVSR_CHECK_CUDA(cudaEventRecord(ev, e.wq));
    SVSR_CHECK_CUDA(cudaStreamWaitEvent(s, ev, 0));
    return SVSR_OK;
  }
};


// ------------------------------------------------------------------------------------------------
// visual frontend: parameters, workspace, forward, backward
// ------------------------------------------------------------------------------------------------
// stem_w / stem_bn_prefix / trunk_prefix are the reference's state-dict names, e.g. "stem3d.0.weight", "stem3d.1",
// "resnet" (LRW) or "encoder.frontend.frontend3D.0.weight", "encoder.frontend.frontend3D.1", "encoder.frontend.trunk".
static int frontend_build(EngineBase& e, Frontend& f, const std::string& stem_w, const std::string& stem_bn_prefix,
                          const std::string& trunk_prefix, Bump& b) {
  e.N = f.B * f.T;
  f.H0 = conv_out(f.H, 7, 2, 3);
  f.H1 = conv_out(f.H0, 3, 2, 1);
  f.stem_conv.cin = 1, f.stem_conv.cout = 64;
  f.stem_conv.w = add_param(e.params, e.pc, stem_w, {64, 1, 5, 7, 7});
  f.stem_conv.wf = b.take(64 * 320 * 2);
  add_bn(e, f.stem_bn, stem_bn_prefix, 64, b);
  int cin = 64, h = f.H1;
  const int widths[4] = {64, 128, 256, 512};
  for (int li = 0; li < 4; ++li) {
    for (int bi = 0; bi < 2; ++bi) {
      BlockRef& blk = f.blocks[li * 2 + bi];
      const std::string pre = trunk_prefix + ".layer" + std::to_string(li + 1) + "." + std::to_string(bi);
      blk.cin = bi == 0 ? cin : widths[li];
      blk.cout = widths[li];
      blk.stride = (bi == 0 && li > 0) ? 2 : 1;
      blk.Hin = h;
      blk.Hout = conv_out(h, 3, blk.stride, 1);
      blk.ds = (bi == 0 && li > 0);
      add_conv(e, blk.conv1, pre + ".conv1.weight", blk.cin, blk.cout, 3, blk.stride, 1, b);
      add_bn(e, blk.bn1, pre + ".bn1", blk.cout, b);
      add_conv(e, blk.conv2, pre + ".conv2.weight", blk.cout, blk.cout, 3, 1, 1, b);
      add_bn(e, blk.bn2, pre + ".bn2", blk.cout, b);
      if (blk.ds) {
        add_conv(e, blk.convds, pre + ".downsample.0.weight", blk.cin, blk.cout, 1, blk.stride, 0, b);
        add_bn(e, blk.bnds, pre + ".downsample.1", blk.cout, b);
      }
      h = blk.Hout;
      const size_t act = (size_t)e.N * blk.Hout * blk.Hout * blk.cout * 2;
      blk.c1 = b.take(act), blk.a1 = b.take(act), blk.c2 = b.take(act), blk.out = b.take(act);
      blk.cds = blk.ds ? b.take(act) : 0;
    }
    cin = widths[li];
  }
  return SVSR_OK;
}

// activations, BN statistic slots and backward scratch of the frontend
static void frontend_alloc(EngineBase& e, Frontend& f, Bump& b) {
  const size_t n0 = (size_t)e.N * f.H0 * f.H0 * 64;
  {
    const char* d = getenv("SVSR_STEM_DIRECT");
    f.direct = !(d && d[0] == '0') && stem_direct_supported(f.H, f.H);
  }
  if (f.direct)
    f.vid = b.take((size_t)e.N * f.H * f.H * 2);
  else
    f.patches = b.take(n0 * 2);
  f.y0 = b.take(n0 * 2);
  const size_t n1 = (size_t)e.N * f.H1 * f.H1 * 64;
  f.x1 = b.take(n1 * 2);
  f.argmax = b.take(n1);
  {
    Bump sb;
    auto slot = [&](BnRef& bn) {
      bn.stats_f = sb.take(2 * bn.C * sizeof(double));
      bn.stats_b = sb.take(2 * bn.C * sizeof(double));
    };
    slot(f.stem_bn);
    for (auto& blk : f.blocks) {
      slot(blk.bn1), slot(blk.bn2);
      if (blk.ds) slot(blk.bnds);
    }
    f.stats_arena_bytes = sb.off;
    f.stats_arena = b.take(sb.off);
    auto fix = [&](BnRef& bn) { bn.stats_f += f.stats_arena, bn.stats_b += f.stats_arena; };
    fix(f.stem_bn);
    for (auto& blk : f.blocks) {
      fix(blk.bn1), fix(blk.bn2);
      if (blk.ds) fix(blk.bnds);
    }
  }
  for (int i = 0; i < 9; ++i) f.gbuf[i] = b.take(n1 * 2);
  f.stem_dz = b.take(n0 * 2);
  e.wgrad_tmp = b.take((size_t)9 * 512 * 512 * 4);
}

static void frontend_pack_jobs(const EngineBase& e, const Frontend& f, std::vector<PackJob>& jobs) {
  jobs.push_back({e.P + f.stem_conv.w, e.ws<bf16>(f.stem_conv.wf), nullptr, 2, 0, 0, 0, 0});
  auto conv = [&](const ConvRef& c) {
    jobs.push_back({e.P + c.w, e.ws<bf16>(c.wf), e.ws<bf16>(c.wd), 0, c.cout, c.cin, c.R * c.R, 0});
  };
  for (auto& blk : f.blocks) {
    conv(blk.conv1), conv(blk.conv2);
    if (blk.ds) conv(blk.convds);
  }
}

// videos fp32 [B,T,H,W] (LRW [B,1,T,H,W] and LRS [B,T,1,H,W] are the same bytes). Returns the last block's output
// (bf16 [N, H4, H4, 512]) in *out. Zeroes the frontend's BN statistic slots first.
static int frontend_forward(EngineBase& e, Frontend& f, const float* videos, int train, const bf16** out,
                            cudaStream_t s) {
  const int act = f.swish ? 2 : 1;
  SVSR_CHECK_CUDA(cudaMemsetAsync(e.ws<uint8_t>(f.stats_arena), 0, f.stats_arena_bytes, s));
  // ---- stem: 7x7/s2 patch gather, 5-tap temporal implicit GEMM, BN3d + GELU|Swish + max-pool ----
  if (f.direct) {
    RC(cast_f32_to_bf16(videos, e.ws<bf16>(f.vid), (long long)e.N * f.H * f.H, s));
    RC(stem_direct_fwd(e.ws<bf16>(f.vid), e.ws<bf16>(f.stem_conv.wf), e.ws<bf16>(f.y0),
                       train ? e.ws<double>(f.stem_bn.stats_f) : nullptr, f.B, f.T, f.H, f.H,
                       2.0 * e.N * f.H0 * f.H0 * 64.0 * 245.0, s));
  } else {
    RC(stem_patch(videos, e.ws<bf16>(f.patches), f.B, f.T, f.H, f.H, s));
    IgemmProblem p;
    p.a = e.ws<bf16>(f.patches), p.a_N = f.B, p.a_H = f.T, p.a_W = f.H0 * f.H0, p.a_C = 64, p.cin = 64;
    p.ntaps = 5;
    for (int kt = 0; kt < 5; ++kt) p.tap_dh[kt] = kt - 2, p.tap_dw[kt] = 0, p.tap_kbase[kt] = kt * 64;
    p.o_N = f.B, p.OH = f.T, p.OW = f.H0 * f.H0;
    p.b = e.ws<bf16>(f.stem_conv.wf), p.b_rows = 64, p.b_cols = 320;
    p.out = e.ws<bf16>(f.y0), p.ldc = 64, p.o_H = f.T, p.o_W = f.H0 * f.H0;
    p.algo_flops = 2.0 * e.N * f.H0 * f.H0 * 64.0 * 245.0;
    p.bn_stats = train ? e.ws<double>(f.stem_bn.stats_f) : nullptr;
    RC(igemm_launch(p, s));
  }
  RC(bn_fwd(e, e.ws<bf16>(f.y0), (long long)e.N * f.H0 * f.H0, f.stem_bn, train, s));
  RC(stem_bn_gelu_pool(e.ws<bf16>(f.y0), e.ws<float>(f.stem_bn.coef), e.ws<bf16>(f.x1), e.ws<uint8_t>(f.argmax), e.N,
                       f.H0, f.H0, s, f.swish));
  // ---- the eight BasicBlocks ----
  RC(pack_join(e, s));  // trunk / encoder operand copies repacked beside the stem
  const bf16* x = e.ws<bf16>(f.x1);
  for (auto& blk : f.blocks) {
    const long long rows = (long long)e.N * blk.Hout * blk.Hout;
    RC(conv_fwd(e, x, blk.Hin, blk.conv1, e.ws<bf16>(blk.c1), train ? e.ws<double>(blk.bn1.stats_f) : nullptr, s));
    RC(bn_fwd(e, e.ws<bf16>(blk.c1), rows, blk.bn1, train, s));
    RC(bn_apply(e.ws<bf16>(blk.c1), e.ws<float>(blk.bn1.coef), nullptr, nullptr, act, e.ws<bf16>(blk.a1), rows,
                blk.cout, s));
    RC(conv_fwd(e, e.ws<bf16>(blk.a1), blk.Hout, blk.conv2, e.ws<bf16>(blk.c2),
                train ? e.ws<double>(blk.bn2.stats_f) : nullptr, s));
    RC(bn_fwd(e, e.ws<bf16>(blk.c2), rows, blk.bn2, train, s));
    if (blk.ds) {
      RC(conv_fwd(e, x, blk.Hin, blk.convds, e.ws<bf16>(blk.cds), train ? e.ws<double>(blk.bnds.stats_f) : nullptr,
                  s));
      RC(bn_fwd(e, e.ws<bf16>(blk.cds), rows, blk.bnds, train, s));
      RC(bn_apply(e.ws<bf16>(blk.c2), e.ws<float>(blk.bn2.coef), e.ws<bf16>(blk.cds), e.ws<float>(blk.bnds.coef), act,
                  e.ws<bf16>(blk.out), rows, blk.cout, s));
    } else {
      RC(bn_apply(e.ws<bf16>(blk.c2), e.ws<float>(blk.bn2.coef), x, nullptr, act, e.ws<bf16>(blk.out), rows, blk.cout,
                  s));
    }
    x = e.ws<bf16>(blk.out);
  }
  *out = x;
  return SVSR_OK;
}

// Backward of the frontend. On entry gbuf[0] holds d loss / d (last block output) (bf16, NHWC); weight-gradient
// GEMMs go to the side stream through `sq`. Ends with sq.join().
// Blocks bi_hi .. bi_lo (7 = layer4.1 ... 0 = layer1.0), then -- with_stem -- the stem. A caller that splits the trunk
// (data parallel: all-reduce layer3-4's gradients while layer1-2 + the stem still compute) calls it with an EVEN number
// of blocks per call, so the T0 / T4 ping-pong is back in place for the next call.
static int frontend_backward(EngineBase& e, Frontend& f, SideQueue& sq, cudaStream_t s, int bi_hi = 7, int bi_lo = 0,
                             bool with_stem = true) {
  cudaStream_t w = e.wq;
  bf16* T0 = e.ws<bf16>(f.gbuf[0]);  // dOut of the current block, later da1
  bf16* T2 = e.ws<bf16>(f.gbuf[1]);  // activation-masked upstream gradient (identity shortcut branch)
  bf16* T4 = e.ws<bf16>(f.gbuf[2]);  // dX of the current block
  // SVSR_BN_BWD_FUSED=1 (off by default) moves the BatchNorm-backward reductions into the epilogues of the input-gradient
  // GEMMs that produce the gradient (igemm.cuh, IgemmBnBwd). Measured at the bench geometry (profiles/r2_bn_bwd_fused_ab.md):
  // the 19 reduce launches (1.36 ms) disappear, but the transposed warp reductions double the epilogue-bound dgrad
  // kernels (1.67 -> 3.1 ms) and the step does not move (11.07 vs 11.01 ms) -- and the trunk backward is bounded by the
  // dgrad + wgrad tensor kernels time-sharing the SMs anyway (3.55 of its 3.98 ms), so the reduce passes run in the
  // weight-gradient kernels' shadow. Kept as a tested option.
  const char* fuse = getenv("SVSR_BN_BWD_FUSED");
  const bool fused = !f.swish && fuse && fuse[0] == '1' && bi_hi == 7 && bi_lo == 0;
  if (fused) {
    // ---- ReLU trunk (LRW), reversed, BatchNorm-backward reductions fused into the producing input-gradient GEMMs ----
    // G  = relu-masked gradient w.r.t. the block's output (produced masked by the NEXT block's conv1 dgrad, which also
    //      left sum g / sum g*xhat of bn2 (and downsample.1) in their stats slots); the last block's comes from the
    //      mean pool and goes through the standalone reduce.
    // DA = conv2's input gradient, masked by bn1's own ReLU in the conv2-dgrad epilogue (+ bn1's sums).
    bf16* G = T0;
    bf16* DA = T2;
    bf16* NX = T4;
    for (int bi = 7; bi >= 0; --bi) {
      BlockRef& blk = f.blocks[bi];
      bf16* DC2 = e.ws<bf16>(f.gbuf[3 + (bi & 1)]);
      bf16* DC1 = e.ws<bf16>(f.gbuf[5 + (bi & 1)]);
      bf16* DCD = e.ws<bf16>(f.gbuf[7 + (bi & 1)]);
      const bf16* xin = bi == 0 ? e.ws<bf16>(f.x1) : e.ws<bf16>(f.blocks[bi - 1].out);
      const long long rows = (long long)e.N * blk.Hout * blk.Hout;
      if (bi == 7) {  // G = d loss / d out, not yet masked: the classic three-kernel BatchNorm backward, which also masks
        RC(bn_bwd(e, G, e.ws<bf16>(blk.out), e.ws<bf16>(blk.c2), rows, blk.bn2, DC2, blk.ds ? nullptr : DA, s));
        if (blk.ds) RC(bn_bwd(e, G, e.ws<bf16>(blk.out), e.ws<bf16>(blk.cds), rows, blk.bnds, DCD, nullptr, s));
        if (!blk.ds) {  // the masked gradient (identity-shortcut term) was written to DA: make it G
          bf16* t = G;
          G = DA, DA = t;
        }
      } else {
        RC(bn_bwd_prereduced(e, G, e.ws<bf16>(blk.c2), rows, blk.bn2, DC2, s));
        if (blk.ds && blk.cout <= 256) RC(bn_bwd_prereduced(e, G, e.ws<bf16>(blk.cds), rows, blk.bnds, DCD, s));
        // (two fused reductions over 512 channels do not fit the epilogue's shared-memory slices: layer4.0's
        //  downsample BatchNorm -- a 17 MB tensor -- keeps the standalone reduce, on the already masked gradient)
        if (blk.ds && blk.cout > 256) RC(bn_bwd(e, G, nullptr, e.ws<bf16>(blk.cds), rows, blk.bnds, DCD, nullptr, s, 0));
      }
      RC(sq.fork());  // dc2 (and dcds) complete
      RC(conv_wgrad(e, e.ws<bf16>(blk.a1), blk.Hout, DC2, blk.conv2, w));
      if (blk.ds) RC(conv_wgrad(e, xin, blk.Hin, DCD, blk.convds, w));
      // DA := relu'(bn1) * conv2^T dc2, with bn1's backward sums
      RC(conv_dgrad_bnb(e, DC2, blk.Hout, blk.conv2, DA, nullptr, nullptr, 1, e.ws<bf16>(blk.c1), blk.bn1, nullptr,
                        nullptr, s));
      RC(bn_bwd_prereduced(e, DA, e.ws<bf16>(blk.c1), rows, blk.bn1, DC1, s));
      RC(sq.fork());  // dc1 complete
      RC(conv_wgrad(e, xin, blk.Hin, DC1, blk.conv1, w));
      // NX := gradient w.r.t. the block's input = conv1^T dc1 + shortcut term; for bi > 0 it is the previous block's
      // output gradient: mask it with that block's ReLU (xin > 0) and leave the sums of its bn2 (and downsample.1)
      const bf16* shortcut = G;
      if (blk.ds) {
        SVSR_CHECK_CUDA(cudaMemsetAsync(NX, 0, (size_t)e.N * blk.Hin * blk.Hin * blk.cin * 2, s));
        RC(conv_dgrad(e, DCD, blk.Hin, blk.convds, NX, nullptr, s));
        shortcut = NX;
      }
      if (bi > 0) {
        BlockRef& pb = f.blocks[bi - 1];
        const bool two = pb.ds && pb.cout <= 256;
        RC(conv_dgrad_bnb(e, DC1, blk.Hin, blk.conv1, NX, shortcut, xin, 0, e.ws<bf16>(pb.c2), pb.bn2,
                          two ? e.ws<bf16>(pb.cds) : nullptr, two ? &pb.bnds : nullptr, s));
      } else {
        RC(conv_dgrad(e, DC1, blk.Hin, blk.conv1, NX, shortcut, s));
      }
      bf16* t = G;
      G = NX, NX = t;
      RC(sq.end_unit());
    }
    T0 = G;
  } else {
    // ---- trunk, reversed: one unit per block; dc2 / dc1 / dcds double buffered by block parity ----
    // ReLU trunk: the gradient flowing into a block's output is masked by that output's ReLU where it is PRODUCED -- in the
    // epilogue of the next block's conv1 input-gradient GEMM (conv_dgrad's relu_mask) -- so the two or four BatchNorm-backward
    // passes that consume it read no mask tensor and write no masked copy (SVSR_RELU_MASK_IN_DGRAD=0: the passes mask).
    static const bool mask_in_dgrad = [] {
      const char* v = getenv("SVSR_RELU_MASK_IN_DGRAD");
      return !(v && v[0] == '0');
    }();
    bool g_masked = !f.swish && mask_in_dgrad && bi_hi < 7;  // (a later stage starts on a gradient the earlier one masked)
    for (int bi = bi_hi; bi >= bi_lo; --bi) {
      BlockRef& blk = f.blocks[bi];
      bf16* DC2 = e.ws<bf16>(f.gbuf[3 + (bi & 1)]);
      bf16* DC1 = e.ws<bf16>(f.gbuf[5 + (bi & 1)]);
      bf16* DCD = e.ws<bf16>(f.gbuf[7 + (bi & 1)]);
      const bf16* xin = bi == 0 ? e.ws<bf16>(f.x1) : e.ws<bf16>(f.blocks[bi - 1].out);
      const long long rows = (long long)e.N * blk.Hout * blk.Hout;
      const bf16* out = e.ws<bf16>(blk.out);
      if (!f.swish) {
        const bf16* ref = g_masked ? nullptr : out;
        RC(bn_bwd(e, T0, ref, e.ws<bf16>(blk.c2), rows, blk.bn2, DC2, (blk.ds || g_masked) ? nullptr : T2, s));
        if (blk.ds) RC(bn_bwd(e, T0, ref, e.ws<bf16>(blk.cds), rows, blk.bnds, DCD, nullptr, s));
      } else {
        // Swish(bn2(c2) + shortcut): the pre-activation is rebuilt from c2 and the shortcut operand (resnet.py:104-105)
        if (blk.ds) {
          RC(bn_bwd(e, T0, nullptr, e.ws<bf16>(blk.c2), rows, blk.bn2, DC2, nullptr, s, 2, e.ws<bf16>(blk.cds),
                    e.ws<float>(blk.bnds.coef)));
          RC(bn_bwd(e, T0, nullptr, e.ws<bf16>(blk.cds), rows, blk.bnds, DCD, nullptr, s, 2, e.ws<bf16>(blk.c2),
                    e.ws<float>(blk.bn2.coef)));
        } else {
          RC(bn_bwd(e, T0, nullptr, e.ws<bf16>(blk.c2), rows, blk.bn2, DC2, T2, s, 2, xin, nullptr));
        }
      }
      RC(sq.fork());  // dc2 (and dcds) complete
      RC(conv_wgrad(e, e.ws<bf16>(blk.a1), blk.Hout, DC2, blk.conv2, w));
      if (blk.ds) RC(conv_wgrad(e, xin, blk.Hin, DCD, blk.convds, w));
      // da1: a pre-masked upstream gradient is itself the identity-path residual below, so it must survive -> da1 goes to T2
      bf16* DA1 = g_masked ? T2 : T0;
      const bf16* ident = g_masked ? T0 : T2;
      RC(conv_dgrad(e, DC2, blk.Hout, blk.conv2, DA1, nullptr, s));
      // bn1 is followed directly by its activation: mask / derivative recomputed from c1 (a1 is not read)
      RC(bn_bwd(e, DA1, nullptr, e.ws<bf16>(blk.c1), rows, blk.bn1, DC1, nullptr, s, f.swish ? 2 : 1));
      RC(sq.fork());  // dc1 complete
      RC(conv_wgrad(e, xin, blk.Hin, DC1, blk.conv1, w));
      // the gradient w.r.t. this block's input = the previous block's ReLU output (block 0: the stem's GELU output, no mask)
      const bf16* next_mask = (!f.swish && mask_in_dgrad && bi > 0) ? xin : nullptr;
      if (blk.ds) {
        SVSR_CHECK_CUDA(cudaMemsetAsync(T4, 0, (size_t)e.N * blk.Hin * blk.Hin * blk.cin * 2, s));
        RC(conv_dgrad(e, DCD, blk.Hin, blk.convds, T4, nullptr, s));
        RC(conv_dgrad(e, DC1, blk.Hin, blk.conv1, T4, T4, s, next_mask));
      } else {
        RC(conv_dgrad(e, DC1, blk.Hin, blk.conv1, T4, ident, s, next_mask));
      }
      g_masked = next_mask != nullptr;
      bf16* t = T0;
      T0 = T4, T4 = t;
      RC(sq.end_unit());
    }
  }
  if (!with_stem) return sq.join();
  // ---- stem ----
  bf16* dz = e.ws<bf16>(f.stem_dz);
  RC(stem_bwd_fused(T0, e.ws<uint8_t>(f.argmax), e.ws<bf16>(f.y0), e.ws<float>(f.stem_bn.coef), e.G + f.stem_bn.gamma,
                    e.G + f.stem_bn.beta, dz, e.ws<double>(f.stem_bn.stats_b), e.ws<float>(f.stem_bn.kcoef), e.N, f.H0,
                    f.H0, s, f.swish));
  RC(sq.fork());
  {
    float* tmp = e.ws<float>(e.wgrad_tmp);
    SVSR_CHECK_CUDA(cudaMemsetAsync(tmp, 0, 320 * 64 * 4, w));
    if (f.direct) {
      RC(stem_direct_wgrad(e.ws<bf16>(f.vid), dz, tmp, 64, f.B, f.T, f.H, f.H, 2.0 * e.N * f.H0 * f.H0 * 64.0 * 245.0, w));
    } else {
      WgradProblem p;
      p.a = e.ws<bf16>(f.patches), p.a_N = f.B, p.a_H = f.T, p.a_W = f.H0 * f.H0, p.a_C = 64, p.a_cin = 64;
      p.ntaps = 5;
      for (int kt = 0; kt < 5; ++kt) p.tap_dh[kt] = kt - 2, p.tap_dw[kt] = 0;
      p.b = dz, p.b_C = 64, p.n_cols = 64;
      p.k_N = f.B, p.k_H = f.T, p.k_W = f.H0 * f.H0;
      p.out = tmp, p.ldo = 64;
      p.algo_flops = 2.0 * e.N * f.H0 * f.H0 * 64.0 * 245.0;
      RC(wgrad_launch(p, w));
    }
    RC(unpack_stem_wgrad(tmp, e.G + f.stem_conv.w, w));
  }
  return sq.join();
}

// side stream + events, created at the first bind (which happens on the GPU box)
static int engine_base_bind(EngineBase& e, float* params, float* grads, float* buffers, void* workspace) {
  e.P = params, e.G = grads, e.BUF = buffers, e.WS = static_cast<uint8_t*>(workspace);
  if (!e.side) {
    SVSR_CHECK_CUDA(cudaStreamCreateWithFlags(&e.side, cudaStreamNonBlocking));
    for (int i = 0; i < 4; ++i) {
      SVSR_CHECK_CUDA(cudaEventCreateWithFlags(&e.ev_fork[i], cudaEventDisableTiming));
      SVSR_CHECK_CUDA(cudaEventCreateWithFlags(&e.ev_done[i], cudaEventDisableTiming));
    }
    SVSR_CHECK_CUDA(cudaEventCreateWithFlags(&e.ev_pack_fork, cudaEventDisableTiming));
    SVSR_CHECK_CUDA(cudaEventCreateWithFlags(&e.ev_pack_done, cudaEventDisableTiming));
  }
  return SVSR_OK;
}
static void engine_base_destroy(EngineBase& e) {
  if (e.side) {
    cudaStreamDestroy(e.side);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(e.ev_fork[i]), cudaEventDestroy(e.ev_done[i]);
    cudaEventDestroy(e.ev_pack_fork), cudaEventDestroy(e.ev_pack_done);
    e.side = nullptr;
  }
}

static int tensor_info(const std::vector<ParamInfo>& v, int i, const char** name, int* ndim, int64_t* shape,
                       int64_t* offset, int* decay) {
  SVSR_REQUIRE(i >= 0 && i < (int)v.size(), "tensor index %d out of range", i);
  *name = v[i].name.c_str();
  *ndim = v[i].ndim;
  for (int k = 0; k < 5; ++k) shape[k] = v[i].shape[k];
  *offset = v[i].offset;
  if (decay) *decay = v[i].decay;
  return SVSR_OK;
}

}  // namespace

}  // namespace svsr
