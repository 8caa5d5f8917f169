#include "elementwise.cuh"

namespace svsr {

namespace {

struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  F8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.v[0] = a.x, r.v[1] = a.y, r.v[2] = b.x, r.v[3] = b.y, r.v[4] = c.x, r.v[5] = c.y, r.v[6] = d.x, r.v[7] = d.y;
  return r;
}
// raw 16-byte load / late unpack: unrolled streaming loops keep 4 registers per in-flight load instead of 8
__device__ __forceinline__ uint4 ldraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ F8 unpack8(const uint4& u) {
  F8 r;
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  r.v[0] = a.x, r.v[1] = a.y, r.v[2] = b.x, r.v[3] = b.y, r.v[4] = c.x, r.v[5] = c.y, r.v[6] = d.x, r.v[7] = d.y;
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& r) {
  uint4 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]), u.y = pack_bf16x2(r.v[2], r.v[3]);
  u.z = pack_bf16x2(r.v[4], r.v[5]), u.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ F8 ldf8(const float* p) {
  F8 r;
  float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
  return r;
}
// erf via Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of every tensor these kernels
// write): 2 MUFU + ~8 FMA instead of libdevice's branchy erff -- the stem's BN+GELU+pool passes are erf-bound otherwise.
// MUFU.RCP only (1 ulp): __frcp_rn expands to a Newton fix-up plus a slow-path CALL per element
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = fast_rcp(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float y = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(y, x);
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.0f + fast_erf(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// Swish / SiLU (LRS frontend: conv3d_extractor.py:36, modules/resnet.py relu_type="swish"; convolution.py:78-83)
__device__ __forceinline__ float sigmoid_f(float x) { return fast_rcp(1.0f + __expf(-x)); }
__device__ __forceinline__ float swish_f(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float swish_grad_f(float x) {
  const float sg = sigmoid_f(x);
  return sg * fmaf(x, 1.0f - sg, 1.0f);
}

inline unsigned grid_for(long long work_items, int per_block, int max_blocks = 148 * 8) {
  long long b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (unsigned)b;
}

// Reduce 16 per-thread partials (8 channels x 2 quantities) across the row slots of a 256-thread block and add
// them to the fp64 accumulators stats[q*C + channel].
__device__ __forceinline__ void block_channel_reduce(const float (&acc)[16], int cg, int C, double* stats) {
  __shared__ float sred[256 * 16];
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) sred[i * 256 + tid] = acc[i];
  __syncthreads();
  const int slots = 256 / cg;
  // thread (g = tid % cg) sums quantity i over slots; spread the 16 quantities over the slot threads
  for (int i = tid / cg; i < 16; i += slots) {
    const int g = tid % cg;
    float s = 0.f;
    for (int sl = 0; sl < slots; ++sl) s += sred[i * 256 + sl * cg + g];
    const int q = i >> 3, ch = g * 8 + (i & 7);
    atomicAdd(&stats[q * C + ch], (double)s);
  }
}

// -------------------------------------------------------------------------------------------------
__global__ void stem_patch_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ P, int B, int T, int H,
                                  int W, int OH, int OW) {
  const long long total = (long long)B * T * OH * OW * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int kh = (int)(i & 7);
    long long pix = i >> 3;
    const int ow = (int)(pix % OW);
    long long t1 = pix / OW;
    const int oh = (int)(t1 % OH);
    const long long bt = t1 / OH;
    F8 r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = 0.f;
    const int ih = 2 * oh + kh - 3;
    if (kh < 7 && ih >= 0 && ih < H) {
      const float* row = x + (bt * H + ih) * (long long)W;
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        const int iw = 2 * ow + kw - 3;
        if (iw >= 0 && iw < W) r.v[kw] = __ldg(row + iw);
      }
    }
    st8(P + pix * 64 + kh * 8, r);
  }
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int C,
                                                       double* stats) {
  const int cg = C >> 3;
  const int g = threadIdx.x % cg, slot = threadIdx.x / cg, rpb = 256 / cg;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (long long r = (long long)blockIdx.x * rpb + slot; r < rows; r += (long long)gridDim.x * rpb) {
    F8 v = ld8(x + r * C + g * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v.v[k], acc[8 + k] += v.v[k] * v.v[k];
  }
  block_channel_reduce(acc, cg, C, stats);
}

// coef layout: [0] mean, [1] invstd, [2] scale = gamma*invstd, [3] shift = beta - mean*scale  (each [C])
__global__ void bn_finalize_kernel(const double* __restrict__ stats, long long rows, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* coef,
                                   int update_running) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = (double)rows;
  double mean, var;
  if (update_running < 0) {  // eval mode: normalise with the running statistics
    mean = running_mean[c], var = running_var[c];
  } else {
    mean = stats[c] / n;
    var = stats[C + c] / n - mean * mean;
    if (var < 0) var = 0;
  }
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * invstd;
  coef[c] = (float)mean;
  coef[C + c] = invstd;
  coef[2 * C + c] = sc;
  coef[3 * C + c] = beta[c] - (float)mean * sc;
  if (update_running > 0) {
    const double unbiased = rows > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// Each thread owns one 8-channel group (grid stride is a multiple of C/8: coefficients load once, no index division)
// and walks rows four at a time with every load issued before the first use (this pass is pure HBM streaming; one
// 16-byte load per thread in flight left it at half the bandwidth).
__global__ void __launch_bounds__(256, 2)
bn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ coef,
                const __nv_bfloat16* __restrict__ res, const float* __restrict__ rcoef, int relu,
                __nv_bfloat16* __restrict__ out, long long rows, int C) {
  constexpr int U = 6;
  const int cg = C >> 3, rpb = 256 / cg;  // rows per block pass; threads beyond rpb*cg idle (C = 768: 192 of 256 work)
  if (threadIdx.x >= rpb * cg) return;
  const int g = threadIdx.x % cg;
  const F8 sc = ldf8(coef + 2 * C + g * 8), sh = ldf8(coef + 3 * C + g * 8);
  const long long rstep = (long long)gridDim.x * rpb;
  for (long long r = (long long)blockIdx.x * rpb + threadIdx.x / cg; r < rows; r += U * rstep) {
    uint4 xr[U], rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * rstep;
      xr[u] = ldraw(x + (ru < rows ? ru : r) * C + g * 8);
    }
    if (res) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long ru = r + u * rstep;
        rr[u] = ldraw(res + (ru < rows ? ru : r) * C + g * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ru = r + u * rstep;
      if (ru < rows) {
        F8 v = unpack8(xr[u]);
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] = v.v[k] * sc.v[k] + sh.v[k];
        if (res) {
          F8 rv = unpack8(rr[u]);
          if (rcoef) {  // downsample branch only (3 of 16 launches): coefficients re-read from L1, not kept live
            const F8 rs = ldf8(rcoef + 2 * C + g * 8), rh = ldf8(rcoef + 3 * C + g * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) rv.v[k] = rv.v[k] * rs.v[k] + rh.v[k];
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) v.v[k] += rv.v[k];
        }
        if (relu == 1) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v.v[k] = fmaxf(v.v[k], 0.f);
        } else if (relu == 2) {
#pragma unroll
          for (int k = 0; k < 8; ++k) v.v[k] = swish_f(v.v[k]);
        }
        st8(out + ru * C + g * 8, v);
      }
    }
  }
}

template <int MODE>  // 0: plain / ReLU mask from relu_ref, 1: ReLU mask from this BN's own output, 2: Swish follows
__global__ void __launch_bounds__(256, MODE == 2 ? 2 : 3)
bn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ relu_ref,
                     const __nv_bfloat16* __restrict__ c, const float* __restrict__ coef, long long rows, int C,
                     double* stats, int self_mask, const __nv_bfloat16* __restrict__ sw_res,
                     const float* __restrict__ sw_rcoef) {
  const int cg = C >> 3;
  const int g = threadIdx.x % cg, slot = threadIdx.x / cg, rpb = 256 / cg;
  const F8 mean = ldf8(coef + g * 8);
  F8 scl, shf;  // only the self-masking modes keep these live across the loop
  if (MODE != 0) scl = ldf8(coef + 2 * C + g * 8), shf = ldf8(coef + 3 * C + g * 8);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  // two rows per iteration, every load issued before the first use (this pass is latency-bound otherwise)
  const long long rstep = (long long)gridDim.x * rpb;
  for (long long r = (long long)blockIdx.x * rpb + slot; r < rows; r += 2 * rstep) {
    const bool two = r + rstep < rows;
    const long long off0 = r * C + g * 8, off1 = (two ? r + rstep : r) * C + g * 8;
    F8 gv[2] = {ld8(dout + off0), ld8(dout + off1)};
    const F8 cv[2] = {ld8(c + off0), ld8(c + off1)};
    if (MODE == 0 && relu_ref) {
      const F8 o[2] = {ld8(relu_ref + off0), ld8(relu_ref + off1)};
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[u].v[k] = o[u].v[k] > 0.f ? gv[u].v[k] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      if (MODE == 1) {  // ReLU directly follows this BN: its mask is the sign of the BN output, no extra tensor read
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[u].v[k] = (cv[u].v[k] * scl.v[k] + shf.v[k]) > 0.f ? gv[u].v[k] : 0.f;
      } else if (MODE == 2) {  // Swish follows (this BN output [+ residual branch]): g = dout * swish'(pre-activation)
        F8 pre;
#pragma unroll
        for (int k = 0; k < 8; ++k) pre.v[k] = cv[u].v[k] * scl.v[k] + shf.v[k];
        if (sw_res) {
          F8 rv = ld8(sw_res + (u ? off1 : off0));
          if (sw_rcoef) {
            const F8 rs = ldf8(sw_rcoef + 2 * C + g * 8), rh = ldf8(sw_rcoef + 3 * C + g * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) rv.v[k] = rv.v[k] * rs.v[k] + rh.v[k];
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) pre.v[k] += rv.v[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[u].v[k] *= swish_grad_f(pre.v[k]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[k] += gv[u].v[k];
        acc[8 + k] += gv[u].v[k] * (cv[u].v[k] - mean.v[k]);  // x invstd once, after the loop
      }
    }
  }
  {
    const F8 invstd = ldf8(coef + C + g * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[8 + k] *= invstd.v[k];
  }
  block_channel_reduce(acc, cg, C, stats);
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ stats, long long rows, int C, float* dgamma,
                                       float* dbeta, float* kcoef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sg = stats[c], sgx = stats[C + c];
  dbeta[c] += (float)sg;
  dgamma[c] += (float)sgx;
  kcoef[c] = (float)(sg / (double)rows);
  kcoef[C + c] = (float)(sgx / (double)rows);
}

// dc = sc * (g - k1 - xhat * k2) with g = dout masked by the activation that follows the BN, regrouped as
// dc = sc * g + cB * c + cD (cB = -sc*invstd*k2, cD = sc*(invstd*k2*mean - k1)): three coefficient vectors stay in
// registers instead of five. Each thread owns one 8-channel group (the grid stride is a multiple of C/8, so no index
// division in the loop) and walks rows two at a time with all loads issued before the first use.
template <int MODE>  // 0: plain / ReLU mask from relu_ref, 1: ReLU mask from this BN's own output, 2: Swish follows
__global__ void __launch_bounds__(256, MODE == 2 ? 2 : 3)
bn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ relu_ref,
                    const __nv_bfloat16* __restrict__ c, const float* __restrict__ coef,
                    const float* __restrict__ kcoef, __nv_bfloat16* __restrict__ dc,
                    __nv_bfloat16* __restrict__ gmask_out, long long rows, int C,
                    const __nv_bfloat16* __restrict__ sw_res, const float* __restrict__ sw_rcoef,
                    const double* __restrict__ fstats, float* dgamma, float* dbeta) {
  const int cg = C >> 3, rpb = 256 / cg;
  if (threadIdx.x >= rpb * cg) return;
  const int g = threadIdx.x % cg;
  F8 sc = ldf8(coef + 2 * C + g * 8), cB, cD, shf;
  {
    const F8 mean = ldf8(coef + g * 8), invstd = ldf8(coef + C + g * 8);
    F8 k1, k2;
    if (fstats) {
      // bn_bwd_finalize folded in: k1 = sum g / rows, k2 = sum g*xhat / rows straight from the fp64 reduction; the first
      // row lane of block 0 accumulates d gamma / d beta (one writer per channel)
      const double inv_rows = 1.0 / (double)rows;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double sg = fstats[g * 8 + k], sgx = fstats[C + g * 8 + k];
        k1.v[k] = (float)(sg * inv_rows), k2.v[k] = (float)(sgx * inv_rows);
        if (blockIdx.x == 0 && threadIdx.x < cg) dbeta[g * 8 + k] += (float)sg, dgamma[g * 8 + k] += (float)sgx;
      }
    } else {
      k1 = ldf8(kcoef + g * 8), k2 = ldf8(kcoef + C + g * 8);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = sc.v[k] * invstd.v[k] * k2.v[k];
      cB.v[k] = -t, cD.v[k] = t * mean.v[k] - sc.v[k] * k1.v[k];
    }
  }
  if (MODE != 0) shf = ldf8(coef + 3 * C + g * 8);
  const long long rstep = (long long)gridDim.x * rpb;
  for (long long r = (long long)blockIdx.x * rpb + threadIdx.x / cg; r < rows; r += 2 * rstep) {
    const bool two = r + rstep < rows;
    const long long offs[2] = {r * C + g * 8, (two ? r + rstep : r) * C + g * 8};
    F8 gv[2] = {ld8(dout + offs[0]), ld8(dout + offs[1])};
    const F8 cv[2] = {ld8(c + offs[0]), ld8(c + offs[1])};
    if (MODE == 0 && relu_ref) {
      const F8 o[2] = {ld8(relu_ref + offs[0]), ld8(relu_ref + offs[1])};
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[u].v[k] = o[u].v[k] > 0.f ? gv[u].v[k] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u == 1 && !two) break;
      const long long off = offs[u];
      if (MODE == 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[u].v[k] = (cv[u].v[k] * sc.v[k] + shf.v[k]) > 0.f ? gv[u].v[k] : 0.f;
      } else if (MODE == 2) {
        F8 pre;
#pragma unroll
        for (int k = 0; k < 8; ++k) pre.v[k] = cv[u].v[k] * sc.v[k] + shf.v[k];
        if (sw_res) {
          F8 rv = ld8(sw_res + off);
          if (sw_rcoef) {
            const F8 rs = ldf8(sw_rcoef + 2 * C + g * 8), rh = ldf8(sw_rcoef + 3 * C + g * 8);
#pragma unroll
            for (int k = 0; k < 8; ++k) rv.v[k] = rv.v[k] * rs.v[k] + rh.v[k];
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) pre.v[k] += rv.v[k];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) gv[u].v[k] *= swish_grad_f(pre.v[k]);
      }
      if (gmask_out) st8(gmask_out + off, gv[u]);
      F8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = sc.v[k] * gv[u].v[k] + (cB.v[k] * cv[u].v[k] + cD.v[k]);
      st8(dc + off, o);
    }
  }
}

// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
stem_bn_gelu_pool_kernel(const __nv_bfloat16* __restrict__ y0, const float* __restrict__ coef,
                         __nv_bfloat16* __restrict__ out, uint8_t* __restrict__ argmax, int N, int IH, int IW, int OH,
                         int OW, int swish) {
  constexpr int C = 64, cg = 8;
  const unsigned total = (unsigned)N * OH * OW * cg;  // < 2^31 (checked by the launcher): 32-bit index arithmetic
  const int g = threadIdx.x & 7;                       // the grid stride is a multiple of 8
  const F8 sc = ldf8(coef + 2 * C + g * 8), sh = ldf8(coef + 3 * C + g * 8);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned pix = i >> 3;
    const unsigned ow = pix % (unsigned)OW, t1 = pix / (unsigned)OW;
    const unsigned oh = t1 % (unsigned)OH, n = t1 / (unsigned)OH;
    // GELU (and Swish) are unimodal -- decreasing below z ~ -0.75 (-1.278), increasing above, negative for z < 0 -- so the
    // window maximum of act(z) is act(max z) whenever max z >= 0: one activation per output instead of nine. An
    // all-negative window has two candidates (its largest or its smallest z).
    // Both extrema are tracked in the ONE pass over the window: nearly every warp holds at least one all-negative
    // (channel, window) pair, so a second pass over the nine pixels used to run warp-wide.
    // The window extrema are found on the RAW bf16 conv outputs with packed bf16x2 max / min (z = y*scale + shift is
    // monotone in y, so the extreme z sit at the extreme y: which one depends on the sign of the scale) -- two channels per
    // instruction and no fp32 transform of the 72 window elements; the first position attaining each extreme (torch's
    // max_pool tie rule) comes from a second sweep in descending position order with packed equality masks.
    float zmax[8], zmin[8];
    int imax[8], imin[8];
    const __nv_bfloat16* img = y0 + (size_t)n * IH * IW * C + g * 8;
    uint4 win[9];
    unsigned vmask = 0;  // bit p set: window position p lies inside the image
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = 2 * (int)oh + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = 2 * (int)ow + kw - 1;
        const bool ok = ih >= 0 && ih < IH && iw >= 0 && iw < IW;
        win[kh * 3 + kw] = ok ? *reinterpret_cast<const uint4*>(img + (size_t)(ih * IW + iw) * C) : make_uint4(0u, 0u, 0u, 0u);
        vmask |= ok ? (1u << (kh * 3 + kw)) : 0u;
      }
    }
    const int first_valid = __ffs(vmask) - 1;  // >= 0: the window centre always exists
    uint32_t mx[4], mn[4], ix[4], in_[4];
    mx[0] = mn[0] = win[4].x, mx[1] = mn[1] = win[4].y, mx[2] = mn[2] = win[4].z, mx[3] = mn[3] = win[4].w;  // the centre
#pragma unroll
    for (int pos = 0; pos < 9; ++pos) {
      if (!(vmask & (1u << pos))) continue;
      const uint32_t w[4] = {win[pos].x, win[pos].y, win[pos].z, win[pos].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
        const __nv_bfloat162 a = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&mx[q]), v);
        const __nv_bfloat162 b = __hmin2(*reinterpret_cast<const __nv_bfloat162*>(&mn[q]), v);
        mx[q] = *reinterpret_cast<const uint32_t*>(&a), mn[q] = *reinterpret_cast<const uint32_t*>(&b);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) ix[q] = in_[q] = 0u;
#pragma unroll
    for (int pos = 8; pos >= 0; --pos) {  // descending: the LAST assignment = the FIRST position with equality
      if (!(vmask & (1u << pos))) continue;
      const uint32_t w[4] = {win[pos].x, win[pos].y, win[pos].z, win[pos].w};
      const uint32_t pp = (uint32_t)pos * 0x00010001u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
        const uint32_t ex = __heq2_mask(v, *reinterpret_cast<const __nv_bfloat162*>(&mx[q]));
        const uint32_t en = __heq2_mask(v, *reinterpret_cast<const __nv_bfloat162*>(&mn[q]));
        ix[q] = (pp & ex) | (ix[q] & ~ex);
        in_[q] = (pp & en) | (in_[q] & ~en);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int q = k >> 1, hi = k & 1;
      const float ymx = __uint_as_float(hi ? (mx[q] & 0xffff0000u) : (mx[q] << 16));
      const float ymn = __uint_as_float(hi ? (mn[q] & 0xffff0000u) : (mn[q] << 16));
      const int pmx = (int)((ix[q] >> (16 * hi)) & 0xffu), pmn = (int)((in_[q] >> (16 * hi)) & 0xffu);
      const float za = fmaf(ymx, sc.v[k], sh.v[k]), zb = fmaf(ymn, sc.v[k], sh.v[k]);
      const bool up = sc.v[k] > 0.f;
      zmax[k] = up ? za : zb, imax[k] = up ? pmx : pmn;
      zmin[k] = up ? zb : za, imin[k] = up ? pmn : pmx;
      if (sc.v[k] == 0.f) imax[k] = imin[k] = first_valid;  // constant channel: every position ties, the first one wins
    }
    float best[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      best[k] = swish ? swish_f(zmax[k]) : gelu_f(zmax[k]);
      if (zmax[k] < 0.f) {  // all-negative window: the largest or the smallest z carries the maximum
        const float bmin = swish ? swish_f(zmin[k]) : gelu_f(zmin[k]);
        // first maximum wins (torch max_pool semantics): on a tie the earlier window position
        if ((bmin > best[k]) || (bmin == best[k] && imin[k] < imax[k])) best[k] = bmin, imax[k] = imin[k];
      }
    }
    F8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = best[k];
    st8(out + (size_t)pix * C + g * 8, o);
    uint2 packed;
    packed.x = (uint32_t)imax[0] | ((uint32_t)imax[1] << 8) | ((uint32_t)imax[2] << 16) | ((uint32_t)imax[3] << 24);
    packed.y = (uint32_t)imax[4] | ((uint32_t)imax[5] << 8) | ((uint32_t)imax[6] << 16) | ((uint32_t)imax[7] << 24);
    *reinterpret_cast<uint2*>(argmax + (size_t)pix * C + g * 8) = packed;
  }
}

__global__ void __launch_bounds__(256)
stem_pool_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ argmax,
                          const __nv_bfloat16* __restrict__ y0, const float* __restrict__ coef,
                          __nv_bfloat16* __restrict__ dz, int N, int IH, int IW, int OH, int OW) {
  constexpr int C = 64, cg = 8;
  const long long total = (long long)N * IH * IW * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long pix = i / cg;
    const int iw = (int)(pix % IW);
    long long t1 = pix / IW;
    const int ih = (int)(t1 % IH);
    const long long n = t1 / IH;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    // windows (oh, ow) that contain (ih, iw): 2*oh-1 <= ih <= 2*oh+1
    const int oh_lo = ih >> 1, oh_hi = (ih + 1) >> 1;
    const int ow_lo = iw >> 1, ow_hi = (iw + 1) >> 1;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      if (oh >= OH) continue;
      const int kh = ih - (2 * oh - 1);
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        if (ow >= OW) continue;
        const int pos = kh * 3 + (iw - (2 * ow - 1));
        const long long o = ((n * OH + oh) * OW + ow) * C + g * 8;
        const uint2 am = *reinterpret_cast<const uint2*>(argmax + o);
        const F8 d = ld8(dout + o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t a = ((k < 4 ? am.x : am.y) >> (8 * (k & 3))) & 0xff;
          if ((int)a == pos) acc[k] += d.v[k];
        }
      }
    }
    const F8 v = ld8(y0 + pix * C + g * 8);
    const F8 sc = ldf8(coef + 2 * C + g * 8), sh = ldf8(coef + 3 * C + g * 8);
    F8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = acc[k] == 0.f ? 0.f : acc[k] * gelu_grad_f(v.v[k] * sc.v[k] + sh.v[k]);
    st8(dz + pix * C + g * 8, o);
  }
}

// Fused stem backward (max-pool scatter * GELU' -> BatchNorm3d backward) without ever materialising dz:
//   dz(n,ih,iw,c) = gelu'(bn(y0)) * (sum of dout over the <= 4 pooling windows that contain (ih,iw) and selected it)
//   REDUCE (one thread per POOLED element group): every pooled output routes its gradient to exactly one input, so
//     sum dz = sum_windows dout * gelu'(z_sel) and sum dz*xhat likewise -- a gather of the selected y0 values, 4x fewer
//     elements than the input grid and no window search;
//   APPLY (one thread per input pixel x 8 channels): dc = scale * (dz - k1 - xhat * k2), window membership resolved
//     with byte-wise SIMD compares (vcmpeq4 + prmt build bf16x2 lane masks).
// Replaces stem_pool_gelu_bwd + bn_bwd_reduce + bn_bwd_apply (3.4 GB of traffic, instruction-bound) by 1.7 GB.
__device__ __forceinline__ float gelu_grad_shared_exp(float z) {
  const float ax = fabsf(z) * 0.70710678118654752f;
  const float t = fast_rcp(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = exp2f(ax * ax * -1.4426950408889634f);  // = exp(-z^2/2): shared by erf and by the density term
  const float half_erf = copysignf(fmaf(-0.5f * p * t, e, 0.5f), z);
  return fmaf(z * 0.3989422804014327f, e, 0.5f + half_erf);
}

// x / d for x * d < 2^32 with m = 2^32 / d + 1 (host-computed): one IMAD.HI instead of a ~20-instruction division
__device__ __forceinline__ unsigned fastdiv(unsigned x, unsigned m) { return __umulhi(x, m); }

// REDUCE: also overwrites dout in place with dzp = dout * gelu'(z_sel) (bf16), the per-window routed gradient the
// APPLY pass scatters -- GELU' is evaluated once per pooled element instead of once per input element.
__global__ void __launch_bounds__(256)
stem_bwd_reduce_kernel(__nv_bfloat16* __restrict__ dout, const uint8_t* __restrict__ argmax,
                       const __nv_bfloat16* __restrict__ y0, const float* __restrict__ coef, double* stats,
                       unsigned npool, unsigned IH, unsigned IW, unsigned OH, unsigned OW, unsigned mOH, unsigned mOW,
                       int swish) {
  constexpr unsigned C = 64;
  const unsigned g = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const F8 mean = ldf8(coef + g * 8), invstd = ldf8(coef + C + g * 8);
  const F8 sc = ldf8(coef + 2 * C + g * 8), sh = ldf8(coef + 3 * C + g * 8);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (unsigned pix = blockIdx.x * 32u + slot; pix < npool; pix += gridDim.x * 32u) {
    const unsigned t1 = fastdiv(pix, mOW), ow = pix - t1 * OW, n = fastdiv(t1, mOH), oh = t1 - n * OH;
    const unsigned o = pix * C + g * 8;
    const uint2 am = *reinterpret_cast<const uint2*>(argmax + o);
    const F8 d = ld8(dout + o);
    const unsigned base = ((n * IH + 2 * oh - 1) * IW + 2 * ow - 1) * C + g * 8;  // window origin (mod 2^32 is fine)
    float cv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned pos = ((k < 4 ? am.x : am.y) >> (8 * (k & 3))) & 0xff;
      const unsigned kh = (pos * 11u) >> 5, kw = pos - kh * 3u;  // pos / 3 for pos < 9
      cv[k] = __bfloat162float(y0[base + (kh * IW + kw) * C + k]);
    }
    F8 dzp;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float z = fmaf(cv[k], sc.v[k], sh.v[k]);
      dzp.v[k] = d.v[k] * (swish ? swish_grad_f(z) : gelu_grad_shared_exp(z));
      acc[k] += dzp.v[k];
      acc[8 + k] += dzp.v[k] * (cv[k] - mean.v[k]) * invstd.v[k];
    }
    st8(dout + o, dzp);
  }
  block_channel_reduce(acc, 8, C, stats);
}

// bf16x2 lane mask of channels (2j, 2j+1) from the byte-wise "selected" mask m (0xff per selected channel byte)
__device__ __forceinline__ uint32_t pair_mask(uint32_t m, int j) { return __byte_perm(m, 0, j ? 0x3322 : 0x1100); }

// APPLY: dc = scale * (dz - k1 - xhat * k2) with dz = sum of the dzp of the <= 4 windows that selected this input.
// One thread = 8 channels of a 2 x 2 QUAD of input pixels (2a, 2a+1) x (2b, 2b+1): the quad touches exactly the four
// pooling windows (a, b), (a, b+1), (a+1, b), (a+1, b+1), and which window position each (pixel, window) pair means is a
// compile-time constant -- pixel (2a, 2b) is the centre (4) of window (a, b); (2a, 2b+1) is position 5 of (a, b) and 3 of
// (a, b+1); (2a+1, 2b) is 7 of (a, b) and 1 of (a+1, b); (2a+1, 2b+1) is 8 / 6 / 2 / 0 of the four. Four window loads serve
// four pixels (the pixel-per-thread version issued sixteen), nine mask tests instead of sixteen.
__global__ void __launch_bounds__(256)
stem_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dzp, const uint8_t* __restrict__ argmax,
                      const __nv_bfloat16* __restrict__ y0, const float* __restrict__ coef,
                      const float* __restrict__ kcoef, __nv_bfloat16* __restrict__ dc, unsigned nquad, unsigned IH,
                      unsigned IW, unsigned OH, unsigned OW, unsigned QH, unsigned QW, unsigned mQH, unsigned mQW) {
  constexpr unsigned C = 64;
  const unsigned g = threadIdx.x & 7, slot = threadIdx.x >> 3;
  F8 sc, ca, cb;  // dc = dz*sc + (v*cb + ca): cb = -sc*k2*invstd, ca = -sc*k1 + sc*k2*mean*invstd
  {
    const F8 mean = ldf8(coef + g * 8), invstd = ldf8(coef + C + g * 8);
    const F8 k1 = ldf8(kcoef + g * 8), k2 = ldf8(kcoef + C + g * 8);
    sc = ldf8(coef + 2 * C + g * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cb.v[k] = -sc.v[k] * k2.v[k] * invstd.v[k];
      ca.v[k] = -sc.v[k] * k1.v[k] - cb.v[k] * mean.v[k];
    }
  }
  for (unsigned quad = blockIdx.x * 32u + slot; quad < nquad; quad += gridDim.x * 32u) {
    const unsigned t1 = fastdiv(quad, mQW), b = quad - t1 * QW, n = fastdiv(t1, mQH), a = t1 - n * QH;
    const unsigned ih0 = 2 * a, iw0 = 2 * b;
    const bool row1 = ih0 + 1 < IH, col1 = iw0 + 1 < IW;      // the quad's second pixel row / column exists
    const bool wr1 = a + 1 < OH, wc1 = b + 1 < OW;            // the windows below / to the right exist
    // windows: w[0] = (a, b), w[1] = (a, b+1), w[2] = (a+1, b), w[3] = (a+1, b+1); a missing one loads (a, b) again and is masked
    const unsigned o00 = ((n * OH + a) * OW + b) * C + g * 8;
    const unsigned ow_[4] = {o00, wc1 ? o00 + C : o00, wr1 ? o00 + OW * C : o00, (wr1 && wc1) ? o00 + (OW + 1) * C : o00};
    uint2 am[4];
    uint4 dd[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {  // all loads first
      am[w] = *reinterpret_cast<const uint2*>(argmax + ow_[w]);
      dd[w] = *reinterpret_cast<const uint4*>(dzp + ow_[w]);
    }
    const unsigned p00 = ((n * IH + ih0) * IW + iw0) * C + g * 8;
    const unsigned pix_[4] = {p00, p00 + C, p00 + IW * C, p00 + (IW + 1) * C};
    const bool pv[4] = {true, col1, row1, row1 && col1};
    F8 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = pv[q] ? ld8(y0 + pix_[q]) : F8{};
    float acc[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[q][k] = 0.f;
    // (pixel q, window w, position): the nine pairs of a quad
    auto add = [&](int q, int w, uint32_t pos, bool on) {
      const uint32_t pp = on ? pos * 0x01010101u : 0xffffffffu;  // a window that does not exist matches nothing
      const uint32_t mlo = __vcmpeq4(am[w].x, pp), mhi = __vcmpeq4(am[w].y, pp);
      const float2 x0 = unpack_bf16x2(dd[w].x & pair_mask(mlo, 0)), x1 = unpack_bf16x2(dd[w].y & pair_mask(mlo, 1));
      const float2 x2 = unpack_bf16x2(dd[w].z & pair_mask(mhi, 0)), x3 = unpack_bf16x2(dd[w].w & pair_mask(mhi, 1));
      acc[q][0] += x0.x, acc[q][1] += x0.y, acc[q][2] += x1.x, acc[q][3] += x1.y;
      acc[q][4] += x2.x, acc[q][5] += x2.y, acc[q][6] += x3.x, acc[q][7] += x3.y;
    };
    add(0, 0, 4u, true);
    add(1, 0, 5u, true), add(1, 1, 3u, wc1);
    add(2, 0, 7u, true), add(2, 2, 1u, wr1);
    add(3, 0, 8u, true), add(3, 1, 6u, wc1), add(3, 2, 2u, wr1), add(3, 3, 0u, wr1 && wc1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!pv[q]) continue;
      F8 out;
#pragma unroll
      for (int k = 0; k < 8; ++k) out.v[k] = fmaf(acc[q][k], sc.v[k], fmaf(v[q].v[k], cb.v[k], ca.v[k]));
      st8(dc + pix_[q], out);
    }
  }
}

// -------------------------------------------------------------------------------------------------
__global__ void meanpool_cls_kernel(const __nv_bfloat16* __restrict__ a, const float* __restrict__ cls,
                                    float* __restrict__ xs, int B, int T, int HW, int C, int ldx) {
  const int cg = C >> 3;
  const long long total = (long long)B * (T + 1) * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long row = i / cg;  // b*(T+1) + tt
    const int tt = (int)(row % (T + 1));
    const long long b = row / (T + 1);
    float* dst = xs + row * ldx + g * 8;
    if (tt == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[k] = cls[g * 8 + k];
      continue;
    }
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const __nv_bfloat16* src = a + ((b * T + (tt - 1)) * HW) * (long long)C + g * 8;
    for (int p = 0; p < HW; ++p) {
      const F8 v = ld8(src + (long long)p * C);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v.v[k];
    }
    const float inv = 1.0f / (float)HW;
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = acc[k] * inv;
  }
}

__global__ void meanpool_cls_bwd_kernel(const float* __restrict__ dx, __nv_bfloat16* __restrict__ dout,
                                        float* __restrict__ dcls, int B, int T, int HW, int C, int ldx) {
  const int cg = C >> 3;
  const long long total = (long long)B * (T + 1) * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long row = i / cg;
    const int tt = (int)(row % (T + 1));
    const long long b = row / (T + 1);
    const F8 d = ldf8(dx + row * ldx + g * 8);
    if (tt == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) atomicAdd(dcls + g * 8 + k, d.v[k]);
      continue;
    }
    F8 o;
    const float inv = 1.0f / (float)HW;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = d.v[k] * inv;
    __nv_bfloat16* dst = dout + ((b * T + (tt - 1)) * HW) * (long long)C + g * 8;
    for (int p = 0; p < HW; ++p) st8(dst + (long long)p * C, o);
  }
}

// -------------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wf,
                                        __nv_bfloat16* __restrict__ wd, int Cout, int Cin, int RS) {
  const long long total = (long long)Cout * Cin * RS;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int rs = (int)(i % RS);
    const int ci = (int)((i / RS) % Cin);
    const int co = (int)(i / ((long long)RS * Cin));
    const __nv_bfloat16 v = __float2bfloat16(w[i]);
    wf[(long long)co * RS * Cin + (long long)rs * Cin + ci] = v;
    if (wd) wd[(long long)ci * RS * Cout + (long long)rs * Cout + co] = v;
  }
}
__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ d, float* __restrict__ grad, int Cout, int Cin,
                                         int RS) {
  const long long total = (long long)Cout * Cin * RS;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int rs = (int)(i % RS);
    const int ci = (int)((i / RS) % Cin);
    const int co = (int)(i / ((long long)RS * Cin));
    grad[i] += d[((long long)rs * Cin + ci) * Cout + co];
  }
}
__global__ void pack_stem_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 64 * 320
  if (i >= 64 * 320) return;
  const int co = i / 320, k = i % 320;
  const int kt = k / 64, kh = (k % 64) / 8, kw = k % 8;
  float v = 0.f;
  if (kh < 7 && kw < 7) v = w[((co * 5 + kt) * 7 + kh) * 7 + kw];
  wp[i] = __float2bfloat16(v);
}
__global__ void unpack_stem_wgrad_kernel(const float* __restrict__ d, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over 64*245
  if (i >= 64 * 245) return;
  const int co = i / 245, r = i % 245;
  const int kt = r / 49, kh = (r % 49) / 7, kw = r % 7;
  grad[i] += d[(kt * 64 + kh * 8 + kw) * 64 + co];
}
__global__ void pack_linear_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wb,
                                          __nv_bfloat16* __restrict__ wt, int N, int K, int ldb, int ldt) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int n = n0 + j, k = k0 + threadIdx.x;
    float v = (n < N && k < K) ? w[(long long)n * K + k] : 0.f;
    tile[j][threadIdx.x] = v;
    if (n < N && k < K) wb[(long long)n * ldb + k] = __float2bfloat16(v);
  }
  __syncthreads();
  if (wt) {
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
      const int k = k0 + j, n = n0 + threadIdx.x;
      if (n < N && k < K) wt[(long long)k * ldt + n] = __float2bfloat16(tile[threadIdx.x][j]);
    }
  }
}
__global__ void __launch_bounds__(256) pack_all_kernel(const PackJob* __restrict__ jobs) {
  __shared__ float pack_smem[32 * (32 * 9 + 1)];  // conv tile (32 co x 32 ci x 9 taps) or linear tile [64][65]
  const PackJob jb = jobs[blockIdx.y];
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (jb.type == 0) {
    // conv [Cout][Cin][RS] -> fprop layout dst0 [Cout][RS][Cin] and dgrad layout dst1 [Cin][RS][Cout]: tiles of
    // 32 co x 32 ci x RS through shared memory, so the read (32*RS contiguous floats per co) and BOTH writes (32
    // contiguous bf16 per (row, tap)) are coalesced; element-wise scatter cost 2-byte writes to 22 M distinct sectors.
    const int Cout = jb.a, Cin = jb.b, RS = jb.c;
    if (RS <= 9 && Cout % 32 == 0 && Cin % 32 == 0) {
      float* ctile = pack_smem;
      const int pitch = 32 * RS + 1;
      const int tci = Cin / 32, ntiles = (Cout / 32) * tci;
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int co0 = (t / tci) * 32, ci0 = (t % tci) * 32;
        __syncthreads();
        for (int col = warp; col < 32; col += 8) {
          const float* srow = jb.src + ((long long)(co0 + col) * Cin + ci0) * RS;
          for (int e = lane; e < 32 * RS; e += 32) ctile[col * pitch + e] = srow[e];
        }
        __syncthreads();
        for (int pr = warp; pr < 32 * RS; pr += 8) {
          const int row = pr / RS, rs = pr - row * RS;
          // dst0: row = co_l, lane = ci_l
          jb.dst0[((long long)(co0 + row) * RS + rs) * Cin + ci0 + lane] = __float2bfloat16(ctile[row * pitch + lane * RS + rs]);
          // dst1: row = ci_l, lane = co_l
          if (jb.dst1)
            jb.dst1[((long long)(ci0 + row) * RS + rs) * Cout + co0 + lane] =
                __float2bfloat16(ctile[lane * pitch + row * RS + rs]);
        }
      }
    } else {
      const long long total = (long long)Cout * Cin * RS;
      for (long long i = i0; i < total; i += stride) {
        const int rs = (int)(i % RS);
        const int ci = (int)((i / RS) % Cin);
        const int co = (int)(i / ((long long)RS * Cin));
        const __nv_bfloat16 v = __float2bfloat16(jb.src[i]);
        jb.dst0[(long long)co * RS * Cin + (long long)rs * Cin + ci] = v;
        if (jb.dst1) jb.dst1[(long long)ci * RS * Cout + (long long)rs * Cout + co] = v;
      }
    }
  } else if (jb.type == 1 || (jb.type == 4 && (jb.a / 2) % 64 == 0)) {
    // linear [N,K]: 64x64 tiles through shared memory so that the plain copy AND the transposed copy are both written
    // with full rows (the encoder holds 38 M of the 63 M parameters). 16 loads per thread are in flight before the first
    // barrier: with 32x32 tiles (4 per thread) the pass was bound by two global round trips per 4 KB. GLU projections
    // (type 4, 25 M of them: rows [0,F) = values, [F,2F) = gates, gate rows remapped to start at Fp) take the same path
    // whenever a 64-row tile cannot straddle the value / gate boundary: all rows of a tile then share one row shift.
    float (*tile)[65] = reinterpret_cast<float (*)[65]>(pack_smem);
    const int N = jb.a, K = jb.b, ldb = jb.c, ldt = jb.d;
    const int F = N / 2, gate_shift = jb.type == 4 ? (F + 63) / 64 * 64 - F : 0;
    const int tk = (K + 63) / 64, tn = (N + 63) / 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int t = blockIdx.x; t < tk * tn; t += gridDim.x) {
      const int k0 = (t % tk) * 64, n0 = (t / tk) * 64;
      const int shift = (jb.type == 4 && n0 >= F) ? gate_shift : 0;  // output row = source row + shift
      __syncthreads();
      float v[8][2];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int n = n0 + ty + 8 * jj;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = k0 + tx + 32 * h;
          v[jj][h] = (n < N && k < K) ? jb.src[(long long)n * K + k] : 0.f;
        }
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = ty + 8 * jj, n = n0 + j;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k = k0 + tx + 32 * h;
          tile[j][tx + 32 * h] = v[jj][h];
          if (n < N && k < K) jb.dst0[(long long)(n + shift) * ldb + k] = __float2bfloat16(v[jj][h]);
        }
      }
      __syncthreads();
      if (jb.dst1) {
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = ty + 8 * jj, k = k0 + j;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int n = n0 + tx + 32 * h;
            if (n < N && k < K) jb.dst1[(long long)k * ldt + n + shift] = __float2bfloat16(tile[tx + 32 * h][j]);
          }
        }
      }
    }
  } else if (jb.type == 3 || jb.type == 4) {
    // fp32 vector -> zero-padded fp32 copy (type 3; dst0 reinterpreted as float*), or GLU linear [2F, K] -> bf16 rows
    // remapped to [2Fp, ldb] (value rows at 0, gate rows at Fp) + transposed [K, ldt] (type 4)
    const int N = jb.a, K = jb.b, ldb = jb.c, ldt = jb.d;
    if (jb.type == 3) {
      float* dst = reinterpret_cast<float*>(jb.dst0);
      const int glu = K;  // b = 1: [2F] -> [2Fp] with the gate half moved to Fp
      const int F = N / 2, Fp = (F + 63) / 64 * 64;
      for (long long i = i0; i < N; i += stride) dst[(glu && i >= F) ? i - F + Fp : i] = jb.src[i];
    } else {
      const int F = N / 2, Fp = (F + 63) / 64 * 64;
      for (long long i = i0; i < (long long)N * K; i += stride) {
        const int n = (int)(i / K), k = (int)(i % K);
        const int r = n >= F ? n - F + Fp : n;
        const __nv_bfloat16 v = __float2bfloat16(jb.src[i]);
        jb.dst0[(long long)r * ldb + k] = v;
        if (jb.dst1) jb.dst1[(long long)k * ldt + r] = v;
      }
    }
  } else {
    for (long long i = i0; i < 64 * 320; i += stride) {
      const int co = (int)(i / 320), k = (int)(i % 320);
      const int kt = k / 64, kh = (k % 64) / 8, kw = k % 8;
      float v = 0.f;
      if (kh < 7 && kw < 7) v = jb.src[((co * 5 + kt) * 7 + kh) * 7 + kw];
      jb.dst0[i] = __float2bfloat16(v);
    }
  }
}

// CutMix (LRW/video/src/augment.py:27-118) as one gather over the ORIGINAL batch: the host resolves the reference's
// sequential in-place frame swaps into source-clip tables (vsrc[i,t], asrc[i,a]); this kernel moves the frames / audio
// token rows and builds the mixed soft labels and word masks. One block per (clip, frame) for the video part.
__global__ void __launch_bounds__(256)
cutmix_video_kernel(const float4* __restrict__ vin, float4* __restrict__ vout, const int* __restrict__ vsrc, int T,
                    long long frame_vec4) {
  const long long bt = blockIdx.x;  // i * T + t
  const int t = (int)(bt % T);
  const long long src = (long long)vsrc[bt] * T + t;
  const float4* s = vin + src * frame_vec4;
  float4* d = vout + bt * frame_vec4;
  for (long long k = threadIdx.x; k < frame_vec4; k += blockDim.x) d[k] = s[k];
}
__global__ void cutmix_meta_kernel(const long long* __restrict__ ain, long long* __restrict__ aout,
                                   const int* __restrict__ asrc, int B, int Ta, int G, const long long* __restrict__ labels,
                                   const int* __restrict__ tgt, const float* __restrict__ rate,
                                   const unsigned char* __restrict__ mixed, float* __restrict__ soft, int num_labels,
                                   const float* __restrict__ wm_in, float* __restrict__ wm_out, int Tw) {
  const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = i0; i < (long long)B * Ta * G; i += stride) {
    const int g = (int)(i % G);
    const long long ba = i / G;
    const int a = (int)(ba % Ta);
    aout[i] = ain[((long long)asrc[ba] * Ta + a) * G + g];
  }
  for (long long i = i0; i < (long long)B * num_labels; i += stride) {
    const int b = (int)(i / num_labels), c = (int)(i % num_labels);
    const float own = labels[b] == c ? 1.f : 0.f;
    float v = own;
    if (mixed[b]) {  // (1.0 - mix_rate) * one_hot(org) + mix_rate * one_hot(tar), evaluated in fp32 like torch
      const float r = rate[b], q = (float)(1.0 - (double)r);
      v = q * own + r * (labels[tgt[b]] == c ? 1.f : 0.f);
    }
    soft[i] = v;
  }
  for (long long i = i0; i < (long long)B * Tw; i += stride) {
    const int b = (int)(i / Tw), t = (int)(i % Tw);
    float v = wm_in[i];
    if (mixed[b]) {
      const float r = rate[b], q = (float)(1.0 - (double)r);
      v = q * v + r * wm_in[(long long)tgt[b] * Tw + t];
    }
    wm_out[i] = v;
  }
}
// word-boundary column (lightning.py:145-150): x[b, 1+t, C] = word_mask[b, t], x[b, 0, C] = cls[C]; and d cls[C]
__global__ void wb_column_kernel(float* __restrict__ xs, const float* __restrict__ cls, const float* __restrict__ wm,
                                 int B, int T, int ldx, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (T + 1)) return;
  const int tt = i % (T + 1), b = i / (T + 1);
  xs[(long long)i * ldx + C] = tt == 0 ? cls[C] : wm[b * T + tt - 1];
}
__global__ void wb_column_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dcls, int B, int T, int ldx, int C) {
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += dx[(long long)b * (T + 1) * ldx + C];
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    dcls[C] += t;
  }
}
// grad[n, k] += tmp[row(n), k] for a padded weight-gradient scratch [Np, Kp]; glu: row(n) = n < F ? n : n - F + Fp
__global__ void unpack_linear_wgrad_kernel(const float* __restrict__ tmp, float* __restrict__ grad, int N, int K, int Kp,
                                           int glu) {
  const int F = N / 2, Fp = (F + 63) / 64 * 64;
  const long long total = (long long)N * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / K), k = (int)(i % K);
    const int r = (glu && n >= F) ? n - F + Fp : n;
    grad[i] += tmp[(long long)r * Kp + k];
  }
}
// db[col] += sum_r dy[r][col] (the bias gradient of a Linear). A thread owns 8 consecutive columns (one 16-byte load per
// row, four rows in flight); a block covers 64 columns x a strided slab of rows, reduces its 32 row lanes through shared
// memory and issues 64 atomics. ld % 8 == 0 and the row pitch covers N rounded up to 8 (bf16 GEMM operands do): the last
// group may read pad columns, which are never accumulated into db.
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dy, int ld, float* __restrict__ db,
                                                          int M, int N, const StepCtl ctl) {
  if (ctl_skipped(ctl)) return;
  const int cg = threadIdx.x & 7, lane_r = threadIdx.x >> 3;  // 8 column groups x 32 row lanes
  const int col = blockIdx.x * 64 + cg * 8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  if (col < N) {
    const int step = gridDim.y * 32;
    int r = blockIdx.y * 32 + lane_r;
    const __nv_bfloat16* p = dy + col;
    for (; r + 3 * step < M; r += 4 * step) {
      F8 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld8(p + (long long)(r + u * step) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[u].v[k];
    }
    for (; r < M; r += step) {
      const F8 v = ld8(p + (long long)r * ld);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v.v[k];
    }
  }
  __shared__ float s[32][65];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[lane_r][cg * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += s[i][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < N) atomicAdd(db + c, t);
  }
}
// dst = src when the sublayer's bit is set in the device-resident skip mask (residual stream passes through a dropped
// sublayer, lightning.py:95-105 layer_dropout), nothing otherwise
__global__ void copy_if_skipped_kernel(float4* __restrict__ dst, const float4* __restrict__ src, long long n4,
                                       const StepCtl ctl) {
  if (!ctl_skipped(ctl)) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
__global__ void set_step_ctl_kernel(unsigned* skip, unsigned long long* seed, unsigned skip_mask, unsigned long long s) {
  *skip = skip_mask, *seed = s;
}
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(x[i]);
}

__global__ void split_cast_last_kernel(const float* __restrict__ last, __nv_bfloat16* __restrict__ cls,
                                       __nv_bfloat16* __restrict__ frames, int B, int T, int D) {
  const long long total = (long long)B * (T + 1) * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long row = i / D;
    const int tt = (int)(row % (T + 1));
    const long long b = row / (T + 1);
    const __nv_bfloat16 v = __float2bfloat16(last[i]);
    if (tt == 0)
      cls[b * D + d] = v;
    else
      frames[(b * T + tt - 1) * D + d] = v;
  }
}

}  // namespace

#define LAUNCH_CHECK() \
  note_launch();       \
  SVSR_CHECK_CUDA(cudaGetLastError())

int stem_patch(const float* videos, __nv_bfloat16* patches, int B, int T, int H, int W, cudaStream_t s) {
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  const long long total = (long long)B * T * OH * OW * 8;
  stem_patch_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, s>>>(videos, patches, B, T, H, W, OH, OW);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bn_stats(const __nv_bfloat16* x, long long rows, int C, double* stats, cudaStream_t s) {
  SVSR_REQUIRE(C % 8 == 0 && 256 % (C / 8) == 0, "bn_stats: unsupported channel count %d", C);
  const int rpb = 256 / (C / 8);
  bn_stats_kernel<<<grid_for(rows, rpb * 8, 148 * 4), 256, 0, s>>>(x, rows, C, stats);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bn_finalize(const double* stats, long long rows, int C, const float* gamma, const float* beta, float eps,
                float momentum, float* running_mean, float* running_var, float* coef, int update_running,
                cudaStream_t s) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(stats, rows, C, gamma, beta, eps, momentum, running_mean,
                                                     running_var, coef, update_running);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bn_apply(const __nv_bfloat16* x, const float* coef, const __nv_bfloat16* res, const float* rcoef, int relu,
             __nv_bfloat16* out, long long rows, int C, cudaStream_t s) {
  SVSR_REQUIRE(C % 8 == 0 && C <= 2048, "bn_apply: unsupported channel count %d", C);
  bn_apply_kernel<<<grid_for(rows, (256 / (C / 8)) * 4, 148 * 2), 256, 0, s>>>(x, coef, res, rcoef, relu, out, rows, C);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bn_bwd_reduce(const __nv_bfloat16* dout, const __nv_bfloat16* relu_ref, const __nv_bfloat16* c, const float* coef,
                  long long rows, int C, double* stats, int self_mask, cudaStream_t s, const __nv_bfloat16* sw_res,
                  const float* sw_rcoef) {
  SVSR_REQUIRE(C % 8 == 0 && 256 % (C / 8) == 0, "bn_bwd_reduce: unsupported channel count %d", C);
  const int rpb = 256 / (C / 8);
  const unsigned grid = grid_for(rows, rpb * 8, 148 * 3);  // 3 CTAs per SM are resident (<= 85 registers)
  if (self_mask == 1)
    bn_bwd_reduce_kernel<1><<<grid, 256, 0, s>>>(dout, relu_ref, c, coef, rows, C, stats, self_mask, sw_res, sw_rcoef);
  else if (self_mask == 2)
    bn_bwd_reduce_kernel<2><<<grid, 256, 0, s>>>(dout, relu_ref, c, coef, rows, C, stats, self_mask, sw_res, sw_rcoef);
  else
    bn_bwd_reduce_kernel<0><<<grid, 256, 0, s>>>(dout, relu_ref, c, coef, rows, C, stats, self_mask, sw_res, sw_rcoef);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bn_bwd_finalize(const double* stats, long long rows, int C, float* dgamma, float* dbeta, float* kcoef,
                    cudaStream_t s) {
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, s>>>(stats, rows, C, dgamma, dbeta, kcoef);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bn_bwd_apply(const __nv_bfloat16* dout, const __nv_bfloat16* relu_ref, const __nv_bfloat16* c, const float* coef,
                 const float* kcoef, __nv_bfloat16* dc, __nv_bfloat16* gmask_out, long long rows, int C, int self_mask,
                 cudaStream_t s, const __nv_bfloat16* sw_res, const float* sw_rcoef, const double* fused_stats,
                 float* dgamma, float* dbeta) {
  SVSR_REQUIRE(C % 8 == 0 && C <= 2048, "bn_bwd_apply: unsupported channel count %d", C);
  SVSR_REQUIRE(fused_stats ? (dgamma && dbeta) : kcoef != nullptr, "bn_bwd_apply: kcoef, or the reduction sums with dgamma / dbeta");
  const unsigned grid = grid_for(rows, (256 / (C / 8)) * 4, 148 * 3);
  if (self_mask == 1)
    bn_bwd_apply_kernel<1><<<grid, 256, 0, s>>>(dout, relu_ref, c, coef, kcoef, dc, gmask_out, rows, C, sw_res, sw_rcoef,
                                                 fused_stats, dgamma, dbeta);
  else if (self_mask == 2)
    bn_bwd_apply_kernel<2><<<grid, 256, 0, s>>>(dout, relu_ref, c, coef, kcoef, dc, gmask_out, rows, C, sw_res, sw_rcoef,
                                                 fused_stats, dgamma, dbeta);
  else
    bn_bwd_apply_kernel<0><<<grid, 256, 0, s>>>(dout, relu_ref, c, coef, kcoef, dc, gmask_out, rows, C, sw_res, sw_rcoef,
                                                 fused_stats, dgamma, dbeta);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int stem_bn_gelu_pool(const __nv_bfloat16* y0, const float* coef, __nv_bfloat16* out, uint8_t* argmax, int N, int IH,
                      int IW, cudaStream_t s, int swish) {
  const int OH = (IH + 2 - 3) / 2 + 1, OW = (IW + 2 - 3) / 2 + 1;
  SVSR_REQUIRE((long long)N * OH * OW * 8 < (1LL << 31), "stem_bn_gelu_pool: %d frames exceed 32-bit indexing", N);
  stem_bn_gelu_pool_kernel<<<grid_for((long long)N * OH * OW * 8, 256 * 2), 256, 0, s>>>(y0, coef, out, argmax, N, IH,
                                                                                      IW, OH, OW, swish);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int stem_pool_gelu_bwd(const __nv_bfloat16* dout, const uint8_t* argmax, const __nv_bfloat16* y0, const float* coef,
                       __nv_bfloat16* dz, int N, int IH, int IW, cudaStream_t s) {
  const int OH = (IH + 2 - 3) / 2 + 1, OW = (IW + 2 - 3) / 2 + 1;
  stem_pool_gelu_bwd_kernel<<<grid_for((long long)N * IH * IW * 8, 256 * 2), 256, 0, s>>>(dout, argmax, y0, coef, dz,
                                                                                       N, IH, IW, OH, OW);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int stem_bwd_fused(__nv_bfloat16* dout, const uint8_t* argmax, const __nv_bfloat16* y0, const float* coef,
                   float* dgamma, float* dbeta, __nv_bfloat16* dc, double* stats_scratch, float* kcoef_scratch, int N,
                   int IH, int IW, cudaStream_t s, int swish) {
  const int OH = (IH + 2 - 3) / 2 + 1, OW = (IW + 2 - 3) / 2 + 1;
  const long long npix = (long long)N * IH * IW, npool = (long long)N * OH * OW;
  SVSR_REQUIRE(npix * 64 < (1LL << 31) && npix * (IH > IW ? IH : IW) < (1LL << 32),
               "stem_bwd_fused: %lld input pixels exceed 32-bit indexing", npix);
  auto magic = [](int d) { return (unsigned)((1ULL << 32) / (unsigned)d + 1ULL); };
  SVSR_CHECK_CUDA(cudaMemsetAsync(stats_scratch, 0, 2 * 64 * sizeof(double), s));
  stem_bwd_reduce_kernel<<<grid_for(npool, 32 * 4), 256, 0, s>>>(dout, argmax, y0, coef, stats_scratch, (unsigned)npool,
                                                                 (unsigned)IH, (unsigned)IW, (unsigned)OH, (unsigned)OW,
                                                                 magic(OH), magic(OW), swish);
  LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<1, 128, 0, s>>>(stats_scratch, npix, 64, dgamma, dbeta, kcoef_scratch);
  LAUNCH_CHECK();
  const int QH = (IH + 1) / 2, QW = (IW + 1) / 2;  // 2 x 2 quads of input pixels
  const long long nquad = (long long)N * QH * QW;
  stem_bwd_apply_kernel<<<grid_for(nquad, 32 * 4, 148 * 16), 256, 0, s>>>(dout, argmax, y0, coef, kcoef_scratch, dc,
                                                                          (unsigned)nquad, (unsigned)IH, (unsigned)IW,
                                                                          (unsigned)OH, (unsigned)OW, (unsigned)QH,
                                                                          (unsigned)QW, magic(QH), magic(QW));
  LAUNCH_CHECK();
  return SVSR_OK;
}
int meanpool_cls(const __nv_bfloat16* a, const float* cls, float* x_stream, int B, int T, int HW, int C,
                 cudaStream_t s, int ldx) {
  if (ldx <= 0) ldx = C;
  meanpool_cls_kernel<<<grid_for((long long)B * (T + 1) * (C / 8), 128), 128, 0, s>>>(a, cls, x_stream, B, T, HW, C,
                                                                                      ldx);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int meanpool_cls_bwd(const float* dx, __nv_bfloat16* dout, float* dcls, int B, int T, int HW, int C, cudaStream_t s,
                     int ldx) {
  if (ldx <= 0) ldx = C;
  meanpool_cls_bwd_kernel<<<grid_for((long long)B * (T + 1) * (C / 8), 128), 128, 0, s>>>(dx, dout, dcls, B, T, HW, C,
                                                                                          ldx);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int pack_conv_weight(const float* w, __nv_bfloat16* w_fprop, __nv_bfloat16* w_dgrad, int Cout, int Cin, int R, int S,
                     cudaStream_t s) {
  pack_conv_weight_kernel<<<grid_for((long long)Cout * Cin * R * S, 256), 256, 0, s>>>(w, w_fprop, w_dgrad, Cout, Cin,
                                                                                    R * S);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int unpack_conv_wgrad(const float* d, float* grad, int Cout, int Cin, int R, int S, cudaStream_t s) {
  unpack_conv_wgrad_kernel<<<grid_for((long long)Cout * Cin * R * S, 256), 256, 0, s>>>(d, grad, Cout, Cin, R * S);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int pack_stem_weight(const float* w, __nv_bfloat16* wp, cudaStream_t s) {
  pack_stem_weight_kernel<<<(64 * 320 + 255) / 256, 256, 0, s>>>(w, wp);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int unpack_stem_wgrad(const float* d, float* grad, cudaStream_t s) {
  unpack_stem_wgrad_kernel<<<(64 * 245 + 255) / 256, 256, 0, s>>>(d, grad);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int pack_linear_weight(const float* w, __nv_bfloat16* wb, __nv_bfloat16* wt, int N, int K, int ldb, int ldt,
                       cudaStream_t s) {
  dim3 grid((K + 31) / 32, (N + 31) / 32), block(32, 8);
  pack_linear_weight_kernel<<<grid, block, 0, s>>>(w, wb, wt, N, K, ldb, ldt);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int pack_all_weights(const PackJob* jobs_dev, int njobs, cudaStream_t s) {
  // 512 CTAs per job: the largest jobs (FF linears, 2048 tiles) then run 4 tiles per CTA instead of 32 in sequence --
  // with 64 the kernel lasted as long as its longest per-CTA chain (288 us for 466 MB); CTAs without a tile exit at once
  dim3 grid(512, (unsigned)njobs);
  pack_all_kernel<<<grid, 256, 0, s>>>(jobs_dev);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int cutmix_gather(const float* vin, float* vout, const int* vsrc, int B, int T, long long frame_elems,
                  const long long* ain, long long* aout, const int* asrc, int Ta, int G, const long long* labels,
                  const int* tgt, const float* rate, const unsigned char* mixed, float* soft, int num_labels,
                  const float* wm_in, float* wm_out, int Tw, cudaStream_t s) {
  SVSR_REQUIRE(frame_elems % 4 == 0, "cutmix: frame size %lld must be a multiple of 4 floats", frame_elems);
  cutmix_video_kernel<<<(unsigned)(B * T), 256, 0, s>>>(reinterpret_cast<const float4*>(vin),
                                                        reinterpret_cast<float4*>(vout), vsrc, T, frame_elems / 4);
  LAUNCH_CHECK();
  cutmix_meta_kernel<<<grid_for((long long)B * (Ta * G > num_labels ? Ta * G : num_labels), 256), 256, 0, s>>>(
      ain, aout, asrc, B, Ta, G, labels, tgt, rate, mixed, soft, num_labels, wm_in, wm_out, Tw);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int wb_column(float* xs, const float* cls, const float* wm, int B, int T, int ldx, int C, cudaStream_t s) {
  wb_column_kernel<<<(B * (T + 1) + 127) / 128, 128, 0, s>>>(xs, cls, wm, B, T, ldx, C);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int wb_column_bwd(const float* dx, float* dcls, int B, int T, int ldx, int C, cudaStream_t s) {
  wb_column_bwd_kernel<<<1, 256, 0, s>>>(dx, dcls, B, T, ldx, C);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int unpack_linear_wgrad(const float* tmp, float* grad, int N, int K, int Kp, int glu, cudaStream_t s) {
  unpack_linear_wgrad_kernel<<<grid_for((long long)N * K, 256), 256, 0, s>>>(tmp, grad, N, K, Kp, glu);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int colsum_bf16(const __nv_bfloat16* dy, int ld, float* db, int M, int N, cudaStream_t s, const StepCtl* ctl) {
  SVSR_REQUIRE(ld % 8 == 0 && ((N + 7) & ~7) <= ld, "colsum_bf16: pitch %d must be a multiple of 8 that covers N=%d rounded up to 8", ld, N);
  // about four blocks per SM; every block keeps at least 128 rows (4 rows in flight per thread) when M allows
  const int bx = (N + 63) / 64;
  int by = (148 * 4 + bx - 1) / bx;
  const int by_max = (M + 127) / 128;
  by = by < 1 ? 1 : (by > by_max ? by_max : by);
  dim3 grid((unsigned)bx, (unsigned)by);
  colsum_bf16_kernel<<<grid, 256, 0, s>>>(dy, ld, db, M, N, ctl ? *ctl : StepCtl());
  LAUNCH_CHECK();
  return SVSR_OK;
}
int copy_if_skipped(float* dst, const float* src, long long n, const StepCtl& ctl, cudaStream_t s) {
  SVSR_REQUIRE(n % 4 == 0, "copy_if_skipped: n must be a multiple of 4");
  copy_if_skipped_kernel<<<grid_for(n / 4, 256 * 2), 256, 0, s>>>(reinterpret_cast<float4*>(dst),
                                                                  reinterpret_cast<const float4*>(src), n / 4, ctl);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int set_step_ctl(unsigned* skip, unsigned long long* seed, unsigned skip_mask, unsigned long long seed_value,
                 cudaStream_t s) {
  set_step_ctl_kernel<<<1, 1, 0, s>>>(skip, seed, skip_mask, seed_value);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int cast_f32_to_bf16(const float* x, __nv_bfloat16* y, long long n, cudaStream_t s) {
  cast_f32_to_bf16_kernel<<<grid_for(n, 256 * 4), 256, 0, s>>>(x, y, n);
  LAUNCH_CHECK();
  return SVSR_OK;
}

int split_cast_last(const float* last, __nv_bfloat16* cls, __nv_bfloat16* frames, int B, int T, int D, cudaStream_t s) {
  split_cast_last_kernel<<<grid_for((long long)B * (T + 1) * D, 256 * 4), 256, 0, s>>>(last, cls, frames, B, T, D);
  LAUNCH_CHECK();
  return SVSR_OK;
}

}  // namespace svsr
