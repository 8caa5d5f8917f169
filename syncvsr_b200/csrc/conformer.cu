// Non-GEMM kernels of the LRS sentence-level path: see conformer.cuh for the operator contracts and the reference
// lines each one follows. The dense contractions around them (every Linear / pointwise Conv1d, forward, input- and
// weight-gradient) run on the tcgen05 implicit-GEMM kernels (igemm.cu / wgrad.cu).
#include "conformer.cuh"
#include "attention_rel_tc.cuh"

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
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
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

#define LAUNCH_CHECK() \
  note_launch();       \
  SVSR_CHECK_CUDA(cudaGetLastError())

// =================================================================================================
// LayerNorm: one warp per row, the row lives in registers (NV float4 per lane, D = 128 * NV)
// =================================================================================================
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ yb, float* __restrict__ yf, float* __restrict__ stats, int M,
                     float eps) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < M; r += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * D);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i] = xr[lane + 32 * i];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
    if (lane == 0 && stats) stats[2 * r] = mean, stats[2 * r + 1] = rstd;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      const float4 g = reinterpret_cast<const float4*>(gamma)[c4], bt = reinterpret_cast<const float4*>(beta)[c4];
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + bt.x, o.y = (v[i].y - mean) * rstd * g.y + bt.y;
      o.z = (v[i].z - mean) * rstd * g.z + bt.z, o.w = (v[i].w - mean) * rstd * g.w + bt.w;
      if (yf) reinterpret_cast<float4*>(yf + (size_t)r * D)[c4] = o;
      if (yb) {
        uint2 u;
        u.x = pack_bf16x2(o.x, o.y), u.y = pack_bf16x2(o.z, o.w);
        reinterpret_cast<uint2*>(yb + (size_t)r * D)[c4] = u;
      }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; dgamma += sum dy * xhat; dbeta += sum dy.
// Per-lane column partials of dgamma/dbeta stay in registers over the warp's rows, are reduced across the block's
// warps through shared memory and leave with one atomicAdd per column per block.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dyb, const float* dyf /* may alias dx */, const float* __restrict__ x,
                     const float* __restrict__ gamma, const float* __restrict__ stats, float* dx,
                     int accumulate, float* __restrict__ dgamma, float* __restrict__ dbeta, int M) {
  constexpr int D = NV * 128;
  __shared__ float red[8][D];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = make_float4(0.f, 0.f, 0.f, 0.f), ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = warp; r < M; r += nwarps) {
    const float mean = stats[2 * r], rstd = stats[2 * r + 1];
    float4 xh[NV], g[NV], prev[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      const float4 xv = reinterpret_cast<const float4*>(x + (size_t)r * D)[c4];
      // the value dx accumulates into is fetched with the other operands, ahead of the row reduction: the kernel is one
      // dependent chain per row (this warp is the only writer of the row, dyf may alias dx)
      prev[i] = accumulate ? reinterpret_cast<const float4*>(dx + (size_t)r * D)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 d;
      if (dyf) {
        d = reinterpret_cast<const float4*>(dyf + (size_t)r * D)[c4];
      } else {
        const uint2 u = reinterpret_cast<const uint2*>(dyb + (size_t)r * D)[c4];
        const float2 lo = unpack_bf16x2(u.x), hi = unpack_bf16x2(u.y);
        d = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
      const float4 gm = reinterpret_cast<const float4*>(gamma)[c4];
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      g[i] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += g[i].x + g[i].y + g[i].z + g[i].w;
      s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
      ag[i].x += d.x * xh[i].x, ag[i].y += d.y * xh[i].y, ag[i].z += d.z * xh[i].z, ag[i].w += d.w * xh[i].w;
      ab[i].x += d.x, ab[i].y += d.y, ab[i].z += d.z, ab[i].w += d.w;
    }
    const float c1 = warp_sum(s1) * (1.0f / D), c2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = lane + 32 * i;
      float4 o;
      o.x = rstd * (g[i].x - c1 - xh[i].x * c2), o.y = rstd * (g[i].y - c1 - xh[i].y * c2);
      o.z = rstd * (g[i].z - c1 - xh[i].z * c2), o.w = rstd * (g[i].w - c1 - xh[i].w * c2);
      float4* dst = reinterpret_cast<float4*>(dx + (size_t)r * D) + c4;
      o.x += prev[i].x, o.y += prev[i].y, o.z += prev[i].z, o.w += prev[i].w;
      *dst = o;
    }
  }
  // block reduction of the column partials (dgamma, then dbeta)
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) reinterpret_cast<float4*>(red[wib])[lane + 32 * i] = pass == 0 ? ag[i] : ab[i];
    __syncthreads();
    float* dst = pass == 0 ? dgamma : dbeta;
    for (int c = threadIdx.x; c < D; c += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][c];
      atomicAdd(dst + c, s);
    }
  }
}

// =================================================================================================
// GLU, depthwise conv, BatchNorm1d reductions
// =================================================================================================
__global__ void __launch_bounds__(256)
glu_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ u, long long M, int C) {
  const int cg = C >> 3;
  const long long total = M * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long r = i / cg;
    const F8 a = ld8(h + r * 2 * C + g * 8), b = ld8(h + r * 2 * C + C + g * 8);
    F8 o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = a.v[k] * sigmoid_f(b.v[k]);
    st8(u + r * C + g * 8, o);
  }
}
__global__ void __launch_bounds__(256)
glu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ du, __nv_bfloat16* __restrict__ dh,
               long long M, int C) {
  const int cg = C >> 3;
  const long long total = M * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long r = i / cg;
    const F8 a = ld8(h + r * 2 * C + g * 8), b = ld8(h + r * 2 * C + C + g * 8), d = ld8(du + r * C + g * 8);
    F8 da, db;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float sg = sigmoid_f(b.v[k]);
      da.v[k] = d.v[k] * sg;
      db.v[k] = d.v[k] * a.v[k] * sg * (1.0f - sg);
    }
    st8(dh + r * 2 * C + g * 8, da);
    st8(dh + r * 2 * C + C + g * 8, db);
  }
}

// y[b,t,c] = bias[c] + sum_k w[c, flip ? K-1-k : k] * x[b, t+k-pad, c]; one thread per (b, t, 8 channels)
__global__ void __launch_bounds__(256)
dwconv1d_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                __nv_bfloat16* __restrict__ y, int B, int T, int C, int K, int flip, int wt) {
  const int cg = C >> 3, pad = (K - 1) / 2;
  const long long total = (long long)B * T * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long bt = i / cg;
    const int t = (int)(bt % T);
    const long long b = bt / T;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[g * 8 + j] : 0.f;
    const __nv_bfloat16* xrow = x + (b * T) * (long long)C + g * 8;
    const float* wg = w + (g * 8) * K;
    for (int k = 0; k < K; ++k) {
      const int ts = t + k - pad;
      if (ts < 0 || ts >= T) continue;
      const F8 xv = ld8(xrow + (long long)ts * C);
      const int kk = flip ? K - 1 - k : k;
      if (wt) {  // w is the transposed copy [K, C]: 8 consecutive channels = two 16-byte loads, coalesced across lanes
        const F8 wv = ldf8(w + (long long)kk * C + g * 8);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wv.v[j], xv.v[j], acc[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(wg[j * K + kk], xv.v[j], acc[j]);
      }
    }
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = acc[j];
    st8(y + bt * C + g * 8, o);
  }
}

// dw[c,k] += sum_{b,t} dy[b,t,c] * x[b,t+k-pad,c]; dbias[c] += sum dy. Block: 32 tap slots x 8 channel groups
// (64 channels); grid.x = C/64, grid.y = row chunks of one clip (a chunk never crosses a clip boundary).
__global__ void __launch_bounds__(256)
dwconv1d_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw,
                      float* __restrict__ dbias, int B, int T, int C, int K, int chunk, int chunks_per_clip) {
  const int k = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 64 + g * 8;
  const int b = blockIdx.y / chunks_per_clip, t0 = (blockIdx.y % chunks_per_clip) * chunk;
  const int t1 = min(T, t0 + chunk), pad = (K - 1) / 2;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (k < K) {
    // four rows in flight (fixed trip count, predicated loads: eight independent 16-byte loads ahead of the FMAs)
    for (int t = t0; t < t1; t += 4) {
      uint4 dq[4], xq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tt = t + u, ts = tt + k - pad;
        const bool ok = tt < t1 && ts >= 0 && ts < T;
        dq[u] = ok ? *reinterpret_cast<const uint4*>(dy + ((long long)b * T + tt) * C + c0) : make_uint4(0u, 0u, 0u, 0u);
        xq[u] = ok ? *reinterpret_cast<const uint4*>(x + ((long long)b * T + ts) * C + c0) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 d0 = unpack_bf16x2(dq[u].x), d1 = unpack_bf16x2(dq[u].y), d2 = unpack_bf16x2(dq[u].z), d3 = unpack_bf16x2(dq[u].w);
        const float2 x0 = unpack_bf16x2(xq[u].x), x1 = unpack_bf16x2(xq[u].y), x2 = unpack_bf16x2(xq[u].z), x3 = unpack_bf16x2(xq[u].w);
        acc[0] = fmaf(d0.x, x0.x, acc[0]), acc[1] = fmaf(d0.y, x0.y, acc[1]);
        acc[2] = fmaf(d1.x, x1.x, acc[2]), acc[3] = fmaf(d1.y, x1.y, acc[3]);
        acc[4] = fmaf(d2.x, x2.x, acc[4]), acc[5] = fmaf(d2.y, x2.y, acc[5]);
        acc[6] = fmaf(d3.x, x3.x, acc[6]), acc[7] = fmaf(d3.y, x3.y, acc[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dw + (c0 + j) * K + k, acc[j]);
  } else if (k == 31 && dbias) {  // K <= 31: slot 31 is free and sums dy for the bias gradient
    for (int t = t0; t < t1; ++t) {
      const F8 d = ld8(dy + ((long long)b * T + t) * C + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += d.v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dbias + c0 + j, acc[j]);
  }
}

// Block: 8 channel groups (64 channels) x 32 row slots; grid (C/64, row chunks). See conformer.cuh for the modes.
__global__ void __launch_bounds__(256)
bn_col_reduce_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dout,
                     const float* __restrict__ coef, long long rows, int C, double* stats, int mode) {
  const int g = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + g * 8;
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = 0.f, a1[j] = 0.f;
  F8 mean, invstd, scl, shf;
  if (mode == 1) mean = ldf8(coef + c0), invstd = ldf8(coef + C + c0), scl = ldf8(coef + 2 * C + c0), shf = ldf8(coef + 3 * C + c0);
  for (long long r = (long long)blockIdx.y * 32 + slot; r < rows; r += (long long)gridDim.y * 32) {
    const F8 xv = ld8(x + r * C + c0);
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a0[j] += xv.v[j], a1[j] = fmaf(xv.v[j], xv.v[j], a1[j]);
    } else {
      const F8 d = ld8(dout + r * C + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gv = d.v[j] * swish_grad_f(xv.v[j] * scl.v[j] + shf.v[j]);
        a0[j] += gv;
        a1[j] = fmaf(gv, (xv.v[j] - mean.v[j]) * invstd.v[j], a1[j]);
      }
    }
  }
  __shared__ float red[2][32][65];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[0][slot][g * 8 + j] = a0[j], red[1][slot][g * 8 + j] = a1[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int q = threadIdx.x >> 6, c = threadIdx.x & 63;
    float s = 0.f;
#pragma unroll 8
    for (int sl = 0; sl < 32; ++sl) s += red[q][sl][c];
    atomicAdd(&stats[q * C + blockIdx.x * 64 + c], (double)s);
  }
}

__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // in [R, Cc] -> out [Cc, R]
  if (i < R * Cc) out[(i % Cc) * R + i / Cc] = in[i];
}
__global__ void add_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}
__global__ void cast_scale_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4, float alpha,
                                  float p, unsigned long long seed, const unsigned long long* seed_base) {
  seed = seed_plus(seed, seed_base);
  const float ks = p > 0.f ? alpha / (1.0f - p) : alpha;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    float v[4] = {a.x * ks, a.y * ks, a.z * ks, a.w * ks};
    if (p > 0.f) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (!dropout_keep(seed, (unsigned long long)(4 * i + k), p)) v[k] = 0.f;
    }
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]), u.y = pack_bf16x2(v[2], v[3]);
    reinterpret_cast<uint2*>(y)[i] = u;
  }
}
// y = dropout(x) (bf16 -> bf16), and dst(fp32) += dropout-mask * src(bf16): forward / backward of a stand-alone Dropout
__global__ void dropout_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n,
                                    float p, unsigned long long seed, const unsigned long long* seed_base) {
  seed = seed_plus(seed, seed_base);
  const float ks = 1.0f / (1.0f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(dropout_keep(seed, (unsigned long long)i, p) ? __bfloat162float(x[i]) * ks : 0.f);
}
__global__ void dropout_add_kernel(float* __restrict__ dst, const __nv_bfloat16* __restrict__ src, long long n, float p,
                                   unsigned long long seed, const unsigned long long* seed_base) {
  seed = seed_plus(seed, seed_base);
  const float ks = 1.0f / (1.0f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (dropout_keep(seed, (unsigned long long)i, p)) dst[i] += __bfloat162float(src[i]) * ks;
}
__global__ void dropout_mask_kernel(unsigned char* __restrict__ out, long long n, float p, unsigned long long seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = dropout_keep(seed, (unsigned long long)i, p) ? 1 : 0;
}
__global__ void lengths_kernel(const long long* __restrict__ in, int* __restrict__ out, int n, int maxv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    long long v = in[i];
    out[i] = (int)(v < 0 ? 0 : (v > maxv ? maxv : v));
  }
}

// feats[n, :] (bf16) = mean over HW of a[n, hw, :]  (AdaptiveAvgPool2d(1), resnet.py:126,175-176)
__global__ void meanpool_kernel(const __nv_bfloat16* __restrict__ a, __nv_bfloat16* __restrict__ out, long long N,
                                int HW, int C) {
  const int cg = C >> 3;
  const long long total = N * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long n = i / cg;
    F8 acc;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.v[k] = 0.f;
    for (int p = 0; p < HW; ++p) {
      const F8 v = ld8(a + (n * HW + p) * (long long)C + g * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc.v[k] += v.v[k];
    }
    const float inv = 1.0f / (float)HW;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.v[k] *= inv;
    st8(out + n * C + g * 8, acc);
  }
}
__global__ void meanpool_bwd_kernel(const __nv_bfloat16* __restrict__ df, __nv_bfloat16* __restrict__ dout, long long N,
                                    int HW, int C) {
  const int cg = C >> 3;
  const long long total = N * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    const long long n = i / cg;
    F8 d = ld8(df + n * C + g * 8);
    const float inv = 1.0f / (float)HW;
#pragma unroll
    for (int k = 0; k < 8; ++k) d.v[k] *= inv;
    for (int p = 0; p < HW; ++p) st8(dout + (n * HW + p) * (long long)C + g * 8, d);
  }
}

// =================================================================================================
// positional encodings, embedding
// =================================================================================================
__global__ void rel_pos_table_kernel(__nv_bfloat16* __restrict__ pe, int T, int D) {
  const int total = (2 * T - 1) * (D / 2);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / (D / 2), k = i % (D / 2);
    const float pos = (float)(T - 1 - r);
    const float div = expf((float)(2 * k) * -(logf(10000.0f) / (float)D));
    pe[(size_t)r * D + 2 * k] = __float2bfloat16(sinf(pos * div));
    pe[(size_t)r * D + 2 * k + 1] = __float2bfloat16(cosf(pos * div));
  }
}
__global__ void embed_posenc_kernel(const long long* __restrict__ tok, const float* __restrict__ emb,
                                    float* __restrict__ x, int rows, int L, int D, int V, float p,
                                    unsigned long long seed, const unsigned long long* seed_base) {
  seed = seed_plus(seed, seed_base);
  const long long total = (long long)rows * (D / 2);
  const float sc = sqrtf((float)D);
  const float ks = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % (D / 2));
    const long long r = i / (D / 2);
    const int l = (int)(r % L);
    long long t = tok[r];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    const float div = expf((float)(2 * k) * -(logf(10000.0f) / (float)D));
    const float2 e = reinterpret_cast<const float2*>(emb + t * D)[k];
    float2 o;
    o.x = e.x * sc + sinf((float)l * div);
    o.y = e.y * sc + cosf((float)l * div);
    if (p > 0.f) {
      o.x = dropout_keep(seed, (unsigned long long)(r * D + 2 * k), p) ? o.x * ks : 0.f;
      o.y = dropout_keep(seed, (unsigned long long)(r * D + 2 * k + 1), p) ? o.y * ks : 0.f;
    }
    reinterpret_cast<float2*>(x + r * D)[k] = o;
  }
}
__global__ void embed_bwd_kernel(const long long* __restrict__ tok, const float* __restrict__ dx,
                                 float* __restrict__ demb, int rows, int D, int V, float p, unsigned long long seed,
                                 const unsigned long long* seed_base) {
  seed = seed_plus(seed, seed_base);
  const long long total = (long long)rows * D;
  const float sc = sqrtf((float)D) * (p > 0.f ? 1.0f / (1.0f - p) : 1.0f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (p > 0.f && !dropout_keep(seed, (unsigned long long)i, p)) continue;
    const int d = (int)(i % D);
    const long long r = i / D;
    long long t = tok[r];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    atomicAdd(demb + t * D + d, dx[i] * sc);
  }
}

// =================================================================================================
// Multi-head attention core (d_k = 64) with optional relative-position term, key-length and causal masks.
// One CTA = AQT query rows of one (clip, head); K, V and the window of P rows those queries can reach stay in
// shared memory as bf16 with a 33-word row pitch (lanes iterate over keys: conflict-free); fp32 math.
// Register blocking: a warp owns 4 query rows and every lane 4 keys (16 accumulators), so one shared-memory word of K / P
// feeds 4 rows and one broadcast q pair feeds 4 keys: 8 LDS per 32 FMA instead of 5 per 8. The relative-position term
// is computed against the P window in the window's own index (rows shared by the 4 queries) and added into the score
// at its shifted position; P.V and dS.K / dS.P read the score rows as float4 broadcasts.
// =================================================================================================
constexpr int AQT = 32;   // query rows per CTA (8 warps x 4 rows)
constexpr int APITCH = 66;  // bf16 elements per shared-memory row (33 words)

struct AttnSmem {
  __nv_bfloat16 *Ks, *Vs, *Ps;
  float *S, *S2, *qu, *qv, *dO;
};
// row pitch (floats) of the score tiles: holds Tk keys or, in backward, the Tk + AQT - 1 window positions; multiple of 4
__host__ __device__ inline int attn_spitch(int Tk) { return (Tk + AQT - 1 + 3) & ~3; }
__host__ __device__ inline size_t attn_smem_bytes(int Tk, bool rel, bool bwd) {
  size_t b = (size_t)2 * Tk * APITCH * 2;
  if (rel) b += (size_t)(Tk + AQT - 1) * APITCH * 2;
  b = (b + 15) & ~size_t(15);
  b += (size_t)AQT * attn_spitch(Tk) * 4 * (bwd ? 2 : 1);
  b += (size_t)AQT * 64 * 4 * (bwd ? 3 : 2);
  return b + 16;
}
__device__ __forceinline__ AttnSmem attn_carve(uint8_t* base, int Tk, bool rel, bool bwd) {
  AttnSmem s;
  s.Ks = reinterpret_cast<__nv_bfloat16*>(base);
  s.Vs = s.Ks + (size_t)Tk * APITCH;
  s.Ps = s.Vs + (size_t)Tk * APITCH;
  size_t off = (size_t)2 * Tk * APITCH * 2 + (rel ? (size_t)(Tk + AQT - 1) * APITCH * 2 : 0);
  off = (off + 15) & ~size_t(15);
  s.S = reinterpret_cast<float*>(base + off);
  off += (size_t)AQT * attn_spitch(Tk) * 4;
  s.S2 = reinterpret_cast<float*>(base + off);
  if (bwd) off += (size_t)AQT * attn_spitch(Tk) * 4;
  s.qu = reinterpret_cast<float*>(base + off);
  s.qv = s.qu + AQT * 64;
  s.dO = s.qv + AQT * 64;
  return s;
}

// copy `rows` rows of 64 bf16 (global row r at src + r*ld, rows outside [0, nvalid) are zero) into pitch-66 smem
__device__ __forceinline__ void attn_load_rows(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int rows,
                                               long long first, long long nvalid) {
  for (int i = threadIdx.x; i < rows * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const long long gr = first + r;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (gr >= 0 && gr < nvalid) u = *reinterpret_cast<const uint4*>(src + gr * ld + c * 8);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + (size_t)r * APITCH + c * 8);
    d[0] = u.x, d[1] = u.y, d[2] = u.z, d[3] = u.w;
  }
}
// qd[ii][d] = q[i0+ii][d] (+ bias[d]) as fp32 (zero rows beyond Tq)
__device__ __forceinline__ void attn_load_q(float* qd, const __nv_bfloat16* src, long long ld, int i0, int Tq,
                                            const float* bias) {
  for (int i = threadIdx.x; i < AQT * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    F8 v;
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = 0.f;
    if (i0 + r < Tq) {
      v = ld8(src + (long long)(i0 + r) * ld + c * 8);
      if (bias) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v.v[k] += bias[c * 8 + k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) qd[r * 64 + c * 8 + k] = v.v[k];
  }
}

// acc[r][u] += q_{r} . M_{row(j0 + 32u + lane)} for the 4 consecutive fp32 query-side rows at qrows (64 floats each) and 4 x 32
// bf16 rows of M (row index clamped to [0, nrows): out-of-range results are discarded by the caller)
__device__ __forceinline__ void attn_dot4x4(const float* qrows, const __nv_bfloat16* M, int j0, int nrows,
                                            float (&acc)[4][4]) {
  const int lane = threadIdx.x & 31;
  const uint32_t* rows[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    int j = j0 + 32 * u + lane;
    j = j >= nrows ? nrows - 1 : j;
    rows[u] = reinterpret_cast<const uint32_t*>(M + (size_t)j * APITCH);
  }
#pragma unroll 4
  for (int w = 0; w < 32; ++w) {
    float2 qq[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) qq[r] = reinterpret_cast<const float2*>(qrows + r * 64)[w];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 kk = unpack_bf16x2(rows[u][w]);
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][u] = fmaf(qq[r].x, kk.x, fmaf(qq[r].y, kk.y, acc[r][u]));
    }
  }
}

// raw (unscaled, unmasked) scores of the warp's 4 query rows ii0..ii0+3 into S: (q+u).k_j + (q+v).p_{j-i+Tk-1}
__device__ __forceinline__ void attn_scores4(const AttnK& a, const AttnSmem& sm, int ii0, int sp) {
  const int lane = threadIdx.x & 31;
  for (int j0 = 0; j0 < a.Tk; j0 += 128) {
    float acc[4][4] = {};
    attn_dot4x4(sm.qu + ii0 * 64, sm.Ks, j0, a.Tk, acc);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u + lane;
      if (j < a.Tk) {
#pragma unroll
        for (int r = 0; r < 4; ++r) sm.S[(ii0 + r) * sp + j] = acc[r][u];
      }
    }
  }
  if (a.p) {
    __syncwarp();
    const int nP = a.Tk + AQT - 1;  // window row rl of query ii pairs with key j = rl - (AQT - 1 - ii)
    for (int r0 = 0; r0 < nP; r0 += 128) {
      float acc[4][4] = {};
      attn_dot4x4(sm.qv + ii0 * 64, sm.Ps, r0, nP, acc);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rl = r0 + 32 * u + lane;
        if (rl < nP) {
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int j = rl - (AQT - 1 - (ii0 + r));
            if (j >= 0 && j < a.Tk) sm.S[(ii0 + r) * sp + j] += acc[r][u];  // exactly one writer per (row, key)
          }
        }
      }
    }
  }
  __syncwarp();
}

// o[r] += sum_x W[(ii0 + r)][x] * M[x][2*lane .. 2*lane+1] over x in [0, n) (n rounded up to 4: the pad weights are 0);
// W rows (pitch sp, 16-byte aligned) are read as float4 broadcasts, M rows (bf16, pitch 33 words) one word per lane
__device__ __forceinline__ void attn_wsum4(const float* W, int sp, int ii0, const __nv_bfloat16* M, int n, int nrows,
                                           float2 (&o)[4]) {
  const int lane = threadIdx.x & 31;
  const uint32_t* Mw = reinterpret_cast<const uint32_t*>(M) + lane;
  for (int x = 0; x < n; x += 4) {
    float4 wv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) wv[r] = *reinterpret_cast<const float4*>(W + (ii0 + r) * sp + x);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int xx = x + t < nrows ? x + t : nrows - 1;
      const float2 mv = unpack_bf16x2(Mw[(size_t)xx * (APITCH / 2)]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float wgt = t == 0 ? wv[r].x : (t == 1 ? wv[r].y : (t == 2 ? wv[r].z : wv[r].w));
        o[r].x = fmaf(wgt, mv.x, o[r].x), o[r].y = fmaf(wgt, mv.y, o[r].y);
      }
    }
  }
}

__device__ __forceinline__ void attn_stage(const AttnK& a, const AttnSmem& sm, int b, int h, int i0) {
  attn_load_rows(sm.Ks, a.k + (long long)b * a.Tk * a.ldk + h * 64, a.ldk, a.Tk, 0, a.Tk);
  attn_load_rows(sm.Vs, a.v + (long long)b * a.Tk * a.ldv + h * 64, a.ldv, a.Tk, 0, a.Tk);
  if (a.p) {
    // window row rl holds P row rbase + rl, rbase = Tk - AQT - i0 (rows outside [0, 2Tk-1) are zero)
    attn_load_rows(sm.Ps, a.p + h * 64, a.ldp, a.Tk + AQT - 1, (long long)a.Tk - AQT - i0, 2LL * a.Tk - 1);
    attn_load_q(sm.qv, a.q + (long long)b * a.Tq * a.ldq + h * 64, a.ldq, i0, a.Tq, a.bv ? a.bv + h * 64 : nullptr);
  }
  attn_load_q(sm.qu, a.q + (long long)b * a.Tq * a.ldq + h * 64, a.ldq, i0, a.Tq, a.bu ? a.bu + h * 64 : nullptr);
}

__global__ void __launch_bounds__(256) attention_core_fwd_kernel(const AttnK a) {
  extern __shared__ __align__(16) uint8_t attn_smem_raw[];
  const AttnSmem sm = attn_carve(attn_smem_raw, a.Tk, a.p != nullptr, false);
  const int i0 = blockIdx.x * AQT, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sp = attn_spitch(a.Tk), ii0 = warp * 4;
  attn_stage(a, sm, b, h, i0);
  __syncthreads();
  if (i0 + ii0 >= a.Tq) return;  // no block-level barrier below
  const int klen = a.klen ? min(a.klen[b], a.Tk) : a.Tk;
  const int Tk4 = (a.Tk + 3) & ~3;
  attn_scores4(a, sm, ii0, sp);
  float inv[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ii0 + r;
    float* S = sm.S + (ii0 + r) * sp;
    float m = -INFINITY;
    for (int j = lane; j < a.Tk; j += 32) {
      const bool masked = j >= klen || (a.causal && j > i) || i >= a.Tq;
      const float sj = masked ? -INFINITY : S[j] * a.scale;
      S[j] = sj;
      m = fmaxf(m, sj);
    }
    m = warp_max(m);
    float sum = 0.f;
    if (m > -INFINITY) {
      const float ks = a.drop_p > 0.f ? 1.0f / (1.0f - a.drop_p) : 1.0f;
      const unsigned long long e0 = (((unsigned long long)b * a.H + h) * a.Tq + i) * a.Tk;
      for (int j = lane; j < a.Tk; j += 32) {
        const float e = __expf(S[j] - m);
        sum += e;  // the softmax normaliser is taken before dropout
        S[j] = (a.drop_p > 0.f && !dropout_keep(seed_plus(a.drop_seed, a.seed_base), e0 + j, a.drop_p)) ? 0.f : e * ks;
      }
      sum = warp_sum(sum);
    } else {  // every key masked: the reference's re-masked softmax row is all zero (attention.py:72-77)
      for (int j = lane; j < a.Tk; j += 32) S[j] = 0.f;
    }
    if (lane < Tk4 - a.Tk) S[a.Tk + lane] = 0.f;  // pad weights of the float4 reads
    inv[r] = sum > 0.f ? 1.0f / sum : 0.f;
    if (lane == 0 && a.lse && i < a.Tq) a.lse[((long long)b * a.H + h) * a.Tq + i] = sum > 0.f ? m + logf(sum) : 0.f;
  }
  __syncwarp();
  float2 o[4] = {};
  attn_wsum4(sm.S, sp, ii0, sm.Vs, Tk4, a.Tk, o);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ii0 + r;
    if (i < a.Tq)
      *reinterpret_cast<uint32_t*>(a.o + ((long long)b * a.Tq + i) * a.ldo + h * 64 + 2 * lane) =
          pack_bf16x2(o[r].x * inv[r], o[r].y * inv[r]);
  }
}

// Backward, kernel A: per query tile recompute p, ds = scale * p * (dp - delta); store p~ / ds for kernels B and C;
// dq = ds . K + ds . P_shift; per-CTA partial sums of dbias_u / dbias_v.
__global__ void __launch_bounds__(256) attention_core_bwd_q_kernel(const AttnK a) {
  extern __shared__ __align__(16) uint8_t attn_smem_raw[];
  const AttnSmem sm = attn_carve(attn_smem_raw, a.Tk, a.p != nullptr, true);
  const int i0 = blockIdx.x * AQT, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sp = attn_spitch(a.Tk), ii0 = warp * 4;
  attn_stage(a, sm, b, h, i0);
  attn_load_q(sm.dO, a.d_o + (long long)b * a.Tq * a.ldo + h * 64, a.ldo, i0, a.Tq, nullptr);
  __syncthreads();
  if (i0 + ii0 >= a.Tq) return;
  const int klen = a.klen ? min(a.klen[b], a.Tk) : a.Tk;
  const int Tk4 = (a.Tk + 3) & ~3, nP = a.Tk + AQT - 1, nP4 = (nP + 3) & ~3;
  attn_scores4(a, sm, ii0, sp);
  // dp[j] = dO_i . v_j for the 4 rows
  for (int j0 = 0; j0 < a.Tk; j0 += 128) {
    float acc[4][4] = {};
    attn_dot4x4(sm.dO + ii0 * 64, sm.Vs, j0, a.Tk, acc);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + 32 * u + lane;
      if (j < a.Tk) {
#pragma unroll
        for (int r = 0; r < 4; ++r) sm.S2[(ii0 + r) * sp + j] = acc[r][u];
      }
    }
  }
  __syncwarp();
  const float ks = a.drop_p > 0.f ? 1.0f / (1.0f - a.drop_p) : 1.0f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ii0 + r;
    float* S = sm.S + (ii0 + r) * sp;
    float* S2 = sm.S2 + (ii0 + r) * sp;
    const bool row_ok = i < a.Tq;
    const float lse = row_ok ? a.lse[((long long)b * a.H + h) * a.Tq + i] : 0.f;
    const unsigned long long e0 = (((unsigned long long)b * a.H + h) * a.Tq + i) * a.Tk;
    float delta = 0.f;
    for (int j = lane; j < a.Tk; j += 32) {
      const bool masked = j >= klen || (a.causal && j > i) || !row_ok;
      const float pj = masked ? 0.f : __expf(S[j] * a.scale - lse);
      // dropout mask on the probabilities: d p = mask * d p~ ; the key/value side uses p~ = mask * p
      const float mj = (a.drop_p > 0.f && !dropout_keep(seed_plus(a.drop_seed, a.seed_base), e0 + j, a.drop_p)) ? 0.f : ks;
      S[j] = pj;
      S2[j] *= mj;
      delta = fmaf(pj, S2[j], delta);
    }
    delta = warp_sum(delta);
    float* Pg = a.Pg + (((long long)b * a.H + h) * a.Tq + i) * a.Tk;
    float* DSg = a.DSg + (((long long)b * a.H + h) * a.Tq + i) * a.Tk;
    for (int j = lane; j < a.Tk; j += 32) {
      const float pj = S[j];
      const float ds = pj * (S2[j] - delta) * a.scale;
      const float mj = (a.drop_p > 0.f && !dropout_keep(seed_plus(a.drop_seed, a.seed_base), e0 + j, a.drop_p)) ? 0.f : ks;
      S2[j] = ds;
      if (row_ok) Pg[j] = pj * mj, DSg[j] = ds;
    }
    if (lane < Tk4 - a.Tk) S2[a.Tk + lane] = 0.f;
  }
  __syncwarp();
  float2 du[4] = {}, dv[4] = {};
  attn_wsum4(sm.S2, sp, ii0, sm.Ks, Tk4, a.Tk, du);
  if (a.p) {
    // ds re-indexed by window position: DSs[ii][rl] = ds[ii][rl - (AQT-1-ii)] (zero outside) into the S rows
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int shift = AQT - 1 - (ii0 + r);
      float* S = sm.S + (ii0 + r) * sp;
      const float* S2 = sm.S2 + (ii0 + r) * sp;
      for (int rl = lane; rl < nP4; rl += 32) {
        const int j = rl - shift;
        S[rl] = (j >= 0 && j < a.Tk) ? S2[j] : 0.f;
      }
    }
    __syncwarp();
    attn_wsum4(sm.S, sp, ii0, sm.Ps, nP4, nP, dv);
  }
  float2 su = make_float2(0.f, 0.f), sv = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ii0 + r;
    if (i < a.Tq) {
      *reinterpret_cast<uint32_t*>(a.dq + ((long long)b * a.Tq + i) * a.lddq + h * 64 + 2 * lane) =
          pack_bf16x2(du[r].x + dv[r].x, du[r].y + dv[r].y);
      su.x += du[r].x, su.y += du[r].y, sv.x += dv[r].x, sv.y += dv[r].y;
    }
  }
  if (a.dbu) atomicAdd(a.dbu + h * 64 + 2 * lane, su.x), atomicAdd(a.dbu + h * 64 + 2 * lane + 1, su.y);
  if (a.dbv) atomicAdd(a.dbv + h * 64 + 2 * lane, sv.x), atomicAdd(a.dbv + h * 64 + 2 * lane + 1, sv.y);
}

// Backward, kernel B: dk_j = sum_i ds[i][j] (q_i + u), dv_j = sum_i p[i][j] dO_i for a tile of 32 keys.
__global__ void __launch_bounds__(256) attention_core_bwd_kv_kernel(const AttnK a) {
  __shared__ float DSt[32][33], Pt[32][33], Qt[32][64], Ot[32][64];
  const int j0 = blockIdx.x * 32, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long bh = (long long)b * a.H + h;
  float2 dk[4], dv[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) dk[u] = make_float2(0.f, 0.f), dv[u] = make_float2(0.f, 0.f);
  for (int i0 = 0; i0 < a.Tq; i0 += 32) {
    __syncthreads();
    {
      const int r = threadIdx.x >> 3, c = (threadIdx.x & 7) * 4;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + r, j = j0 + c + u;
        const bool ok = i < a.Tq && j < a.Tk;
        DSt[r][c + u] = ok ? a.DSg[(bh * a.Tq + i) * a.Tk + j] : 0.f;
        Pt[r][c + u] = ok ? a.Pg[(bh * a.Tq + i) * a.Tk + j] : 0.f;
      }
      const int cc = threadIdx.x & 7;
      F8 qv, ov;
#pragma unroll
      for (int k = 0; k < 8; ++k) qv.v[k] = 0.f, ov.v[k] = 0.f;
      if (i0 + r < a.Tq) {
        qv = ld8(a.q + ((long long)b * a.Tq + i0 + r) * a.ldq + h * 64 + cc * 8);
        ov = ld8(a.d_o + ((long long)b * a.Tq + i0 + r) * a.ldo + h * 64 + cc * 8);
        if (a.bu) {
#pragma unroll
          for (int k = 0; k < 8; ++k) qv.v[k] += a.bu[h * 64 + cc * 8 + k];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) Qt[r][cc * 8 + k] = qv.v[k], Ot[r][cc * 8 + k] = ov.v[k];
    }
    __syncthreads();
#pragma unroll 4
    for (int ii = 0; ii < 32; ++ii) {
      const float2 q2 = reinterpret_cast<const float2*>(Qt[ii])[lane];
      const float2 o2 = reinterpret_cast<const float2*>(Ot[ii])[lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float ds = DSt[ii][warp * 4 + u], pp = Pt[ii][warp * 4 + u];
        dk[u].x = fmaf(ds, q2.x, dk[u].x), dk[u].y = fmaf(ds, q2.y, dk[u].y);
        dv[u].x = fmaf(pp, o2.x, dv[u].x), dv[u].y = fmaf(pp, o2.y, dv[u].y);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int j = j0 + warp * 4 + u;
    if (j < a.Tk) {
      *reinterpret_cast<uint32_t*>(a.dk + ((long long)b * a.Tk + j) * a.lddk + h * 64 + 2 * lane) =
          pack_bf16x2(dk[u].x, dk[u].y);
      *reinterpret_cast<uint32_t*>(a.dv + ((long long)b * a.Tk + j) * a.lddv + h * 64 + 2 * lane) =
          pack_bf16x2(dv[u].x, dv[u].y);
    }
  }
}

// Backward, kernel C: dP[r] += sum_i ds[i][r + i - (Tk-1)] (q_i + v) for a tile of 32 relative positions of one clip.
__global__ void __launch_bounds__(256) attention_core_bwd_pos_kernel(const AttnK a) {
  __shared__ float Dt[32][33], Qt[32][64];
  const int r0 = blockIdx.x * 32, h = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long bh = (long long)b * a.H + h;
  float2 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float2(0.f, 0.f);
  for (int i0 = 0; i0 < a.Tq; i0 += 32) {
    // keys reachable from this (relative position, query) tile: j = r + i - (Tk-1)
    if (r0 + 31 + i0 + 31 - (a.Tk - 1) < 0 || r0 + i0 - (a.Tk - 1) >= a.Tk) continue;  // uniform per CTA
    __syncthreads();
    {
      const int r = threadIdx.x >> 3, c = (threadIdx.x & 7) * 4;
      const int i = i0 + r;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = r0 + c + u + i - (a.Tk - 1);
        Dt[r][c + u] = (i < a.Tq && j >= 0 && j < a.Tk) ? a.DSg[(bh * a.Tq + i) * a.Tk + j] : 0.f;
      }
      const int cc = threadIdx.x & 7;
      F8 qv;
#pragma unroll
      for (int k = 0; k < 8; ++k) qv.v[k] = 0.f;
      if (i < a.Tq) {
        qv = ld8(a.q + ((long long)b * a.Tq + i) * a.ldq + h * 64 + cc * 8);
        if (a.bv) {
#pragma unroll
          for (int k = 0; k < 8; ++k) qv.v[k] += a.bv[h * 64 + cc * 8 + k];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) Qt[r][cc * 8 + k] = qv.v[k];
    }
    __syncthreads();
#pragma unroll 4
    for (int ii = 0; ii < 32; ++ii) {
      const float2 q2 = reinterpret_cast<const float2*>(Qt[ii])[lane];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float ds = Dt[ii][warp * 4 + u];
        acc[u].x = fmaf(ds, q2.x, acc[u].x), acc[u].y = fmaf(ds, q2.y, acc[u].y);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = r0 + warp * 4 + u;
    if (r < 2 * a.Tk - 1) {
      float* dst = a.dp + (long long)r * (a.H * 64) + h * 64 + 2 * lane;
      atomicAdd(dst, acc[u].x), atomicAdd(dst + 1, acc[u].y);
    }
  }
}


// =================================================================================================
// CTC (log-softmax + alpha-beta recursion + gradient w.r.t. the logits) and the label-smoothing KL loss
// =================================================================================================
__device__ __forceinline__ float lae(float a, float b) {  // log(exp(a) + exp(b)) with -inf operands
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + logf(expf(a - m) + expf(b - m));
}

__global__ void __launch_bounds__(256)
row_lse_kernel(const float* __restrict__ logits, int ld, int V, long long rows, float* __restrict__ lse) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* row = logits + r * ld;
    float m = -INFINITY;
    for (int j = lane; j < V; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float se = 0.f;
    for (int j = lane; j < V; j += 32) se += expf(row[j] - m);
    se = warp_sum(se);
    if (lane == 0) lse[r] = m + logf(se);
  }
}

// One CTA per sample, one thread per state of the blank-extended label sequence. alpha/beta [B, T, S] (log domain,
// both include the emission at their own frame, like torch's native ctc_loss). nll[b] = -log p(labels | x), +inf if
// no alignment exists (zero_infinity then scores it 0).
__global__ void ctc_alpha_beta_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ lse,
                                      const long long* __restrict__ labels, int Lmax, const int* __restrict__ in_len,
                                      int T, int S, float* __restrict__ alpha, float* __restrict__ beta,
                                      float* __restrict__ nll, double* acc, int slot) {
  extern __shared__ float ctc_sm[];  // [2][S]
  const int b = blockIdx.x, s = threadIdx.x;
  int Lb = 0;
  while (Lb < Lmax && labels[(long long)b * Lmax + Lb] >= 0) ++Lb;
  const int Sb = 2 * Lb + 1;
  int Tb = in_len[b];
  Tb = Tb > T ? T : Tb;
  const int ext = (s & 1) && s < Sb ? (int)labels[(long long)b * Lmax + (s >> 1)] : 0;
  const bool skip_ok = (s & 1) && s >= 3 && s < Sb && ext != (int)labels[(long long)b * Lmax + (s >> 1) - 1];
  const int ext_p2 = (s & 1) && s + 2 < Sb ? (int)labels[(long long)b * Lmax + (s >> 1) + 1] : -1;
  const bool skip_fw = (s & 1) && s + 2 < Sb && ext_p2 != ext;
  const float* lg = logits + (long long)b * T * ld;
  const float* ls = lse + (long long)b * T;
  float* al = alpha + (long long)b * T * S;
  float* be = beta + (long long)b * T * S;
  float* cur = ctc_sm;
  float* nxt = ctc_sm + S;
  if (Tb < 1) {
    if (s == 0) nll[b] = INFINITY;
    return;
  }
  // ---- alpha ----
  float a = -INFINITY;
  if (s < Sb && s < 2) a = lg[ext] - ls[0];
  if (s < S) cur[s] = a, al[s] = a;
  __syncthreads();
  for (int t = 1; t < Tb; ++t) {
    float v = -INFINITY;
    if (s < Sb) {
      v = cur[s];
      if (s >= 1) v = lae(v, cur[s - 1]);
      if (skip_ok) v = lae(v, cur[s - 2]);
      v += lg[(long long)t * ld + ext] - ls[t];
    }
    if (s < S) nxt[s] = v, al[(long long)t * S + s] = v;
    __syncthreads();
    float* tmp = cur;
    cur = nxt, nxt = tmp;
  }
  if (s == 0) {
    float l = cur[Sb - 1];
    if (Sb > 1) l = lae(l, cur[Sb - 2]);
    const float n = -l;
    nll[b] = n;
    if (n < INFINITY) atomicAdd(acc + slot, (double)n);
  }
  __syncthreads();
  // ---- beta ----
  float bv = -INFINITY;
  if (s < Sb && s >= Sb - 2) bv = lg[(long long)(Tb - 1) * ld + ext] - ls[Tb - 1];
  if (s < S) cur[s] = bv, be[(long long)(Tb - 1) * S + s] = bv;
  __syncthreads();
  for (int t = Tb - 2; t >= 0; --t) {
    float v = -INFINITY;
    if (s < Sb) {
      v = cur[s];
      if (s + 1 < Sb) v = lae(v, cur[s + 1]);
      if (skip_fw) v = lae(v, cur[s + 2]);
      v += lg[(long long)t * ld + ext] - ls[t];
    }
    if (s < S) nxt[s] = v, be[(long long)t * S + s] = v;
    __syncthreads();
    float* tmp = cur;
    cur = nxt, nxt = tmp;
  }
}

// One CTA per frame (b, t): dlogits = dscale * (softmax - occupancy); zero rows beyond the clip's length and for
// samples without a valid alignment (zero_infinity). Padding columns [V, ld) are zeroed.
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ logits, int ld, int V, const float* __restrict__ lse,
                const long long* __restrict__ labels, int Lmax, const int* __restrict__ in_len, int T, int S,
                const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ nll,
                __nv_bfloat16* __restrict__ dlogits, float dscale) {
  extern __shared__ float occ[];  // [V]
  const int t = blockIdx.x % T, b = blockIdx.x / T;
  const long long row = (long long)b * T + t;
  __nv_bfloat16* d = dlogits + row * ld;
  int Tb = in_len[b];
  Tb = Tb > T ? T : Tb;
  const float n = nll[b];
  if (t >= Tb || !(n < INFINITY)) {
    for (int v = threadIdx.x; v < ld; v += blockDim.x) d[v] = __float2bfloat16(0.f);
    return;
  }
  for (int v = threadIdx.x; v < V; v += blockDim.x) occ[v] = 0.f;
  __syncthreads();
  int Lb = 0;
  while (Lb < Lmax && labels[(long long)b * Lmax + Lb] >= 0) ++Lb;
  const int Sb = 2 * Lb + 1;
  const float* lg = logits + row * ld;
  const float l = lse[row];
  for (int s = threadIdx.x; s < Sb; s += blockDim.x) {
    const int ext = (s & 1) ? (int)labels[(long long)b * Lmax + (s >> 1)] : 0;
    const float ab = alpha[row * S + s] + beta[row * S + s];
    if (ab > -INFINITY) atomicAdd(&occ[ext], expf(ab + n - (lg[ext] - l)));
  }
  __syncthreads();
  for (int v = threadIdx.x; v < ld; v += blockDim.x)
    d[v] = __float2bfloat16(v < V ? dscale * (expf(lg[v] - l) - occ[v]) : 0.f);
}

// one warp per decoder position
__global__ void __launch_bounds__(256)
label_smoothing_kernel(const float* __restrict__ logits, int ld, int V, const long long* __restrict__ target, int rows,
                       float smoothing, __nv_bfloat16* __restrict__ dlogits, double* acc, int slot, float dscale) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const float conf = 1.0f - smoothing, low = smoothing / (float)(V - 1);
  for (int r = warp; r < rows; r += nwarps) {
    const long long tg = target[r];
    __nv_bfloat16* d = dlogits ? dlogits + (long long)r * ld : nullptr;
    if (tg < 0 || tg >= V) {  // ignore_id
      if (d)
        for (int v = lane; v < ld; v += 32) d[v] = __float2bfloat16(0.f);
      continue;
    }
    const float* row = logits + (long long)r * ld;
    float m = -INFINITY;
    int mi = 0x7fffffff;
    for (int v = lane; v < V; v += 32) {
      const float x = row[v];
      if (x > m) m = x, mi = v;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (om > m || (om == m && oi < mi)) m = om, mi = oi;
    }
    float se = 0.f, sl = 0.f;
    for (int v = lane; v < V; v += 32) {
      const float x = row[v];
      se += expf(x - m);
      sl += x;
    }
    se = warp_sum(se), sl = warp_sum(sl);
    const float lse = m + logf(se);
    if (lane == 0) {
      const float lpt = row[tg] - lse;
      const float sum_lp = sl - (float)V * lse;
      float kl = 0.f;
      if (conf > 0.f) kl += conf * (logf(conf) - lpt);
      if (low > 0.f) kl += low * ((float)(V - 1) * logf(low) - (sum_lp - lpt));
      atomicAdd(acc + slot, (double)kl);
      if (mi == (int)tg) atomicAdd(acc + slot + 1, 1.0);
      atomicAdd(acc + slot + 2, 1.0);
    }
    if (d)
      for (int v = lane; v < ld; v += 32) {
        float g = 0.f;
        if (v < V) g = dscale * (expf(row[v] - lse) - (v == (int)tg ? conf : low));
        d[v] = __float2bfloat16(g);
      }
  }
}

__global__ void lrs_finalize_kernel(const double* acc, float* out, int B, long long audio_rows, float mtlalpha,
                                    float audio_weight, int has_audio, const int* bad) {
  double la = has_audio ? acc[0] / (double)audio_rows : 0.0;
  if (has_audio && bad && bad[0]) la = (double)NAN;  // an audio token outside its vocabulary (see heads.cuh)
  const double lc = acc[1] / (double)B, lt = acc[2] / (double)B;
  out[0] = (float)((double)mtlalpha * lc + (1.0 - (double)mtlalpha) * lt + (has_audio ? la * (double)audio_weight : 0.0));
  out[1] = (float)lc;
  out[2] = (float)lt;
  out[3] = (float)la;
  out[4] = (float)(acc[4] > 0 ? acc[3] / acc[4] : 0.0);
}

}  // namespace

// =================================================================================================
// host launchers (part 1)
// =================================================================================================
#define LN_DISPATCH(NVV, KERNEL, ...)                                               \
  switch (NVV) {                                                                    \
    case 1: KERNEL<1><<<grid, 256, 0, s>>>(__VA_ARGS__); break;                     \
    case 2: KERNEL<2><<<grid, 256, 0, s>>>(__VA_ARGS__); break;                     \
    case 3: KERNEL<3><<<grid, 256, 0, s>>>(__VA_ARGS__); break;                     \
    case 4: KERNEL<4><<<grid, 256, 0, s>>>(__VA_ARGS__); break;                     \
    case 5: KERNEL<5><<<grid, 256, 0, s>>>(__VA_ARGS__); break;                     \
    case 6: KERNEL<6><<<grid, 256, 0, s>>>(__VA_ARGS__); break;                     \
    case 7: KERNEL<7><<<grid, 256, 0, s>>>(__VA_ARGS__); break;                     \
    default: KERNEL<8><<<grid, 256, 0, s>>>(__VA_ARGS__); break;                    \
  }

int layernorm_fwd(const float* x, const float* gamma, const float* beta, __nv_bfloat16* y_bf16, float* y_f32,
                  float* stats, int M, int D, float eps, cudaStream_t s) {
  SVSR_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, "layernorm: D=%d must be a multiple of 128 in [128,1024]", D);
  SVSR_REQUIRE(y_bf16 || y_f32, "layernorm: no output");
  const unsigned grid = grid_for(M, 8, 148 * 8);
  LN_DISPATCH(D / 128, layernorm_fwd_kernel, x, gamma, beta, y_bf16, y_f32, stats, M, eps);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int layernorm_bwd(const __nv_bfloat16* dy_bf16, const float* dy_f32, const float* x, const float* gamma,
                  const float* stats, float* dx, int accumulate, float* dgamma, float* dbeta, int M, int D,
                  cudaStream_t s) {
  SVSR_REQUIRE(D % 128 == 0 && D >= 128 && D <= 1024, "layernorm: D=%d must be a multiple of 128 in [128,1024]", D);
  SVSR_REQUIRE(dy_bf16 || dy_f32, "layernorm_bwd: no upstream gradient");
  // one row per warp: latency bound, so resident warps win over fewer column atomics (the same
  // measurement as rmsnorm_bwd in encoder.cu)
  const unsigned grid = grid_for(M, 8, 148 * 2);  // at D = 768 one CTA is resident per SM (173 registers): two waves
  LN_DISPATCH(D / 128, layernorm_bwd_kernel, dy_bf16, dy_f32, x, gamma, stats, dx, accumulate, dgamma, dbeta, M);
  LAUNCH_CHECK();
  return SVSR_OK;
}

int glu_fwd(const __nv_bfloat16* h, __nv_bfloat16* u, long long M, int C, cudaStream_t s) {
  SVSR_REQUIRE(C % 8 == 0, "glu: C=%d must be a multiple of 8", C);
  glu_fwd_kernel<<<grid_for(M * (C / 8), 256 * 2), 256, 0, s>>>(h, u, M, C);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int glu_bwd(const __nv_bfloat16* h, const __nv_bfloat16* du, __nv_bfloat16* dh, long long M, int C, cudaStream_t s) {
  SVSR_REQUIRE(C % 8 == 0, "glu: C=%d must be a multiple of 8", C);
  glu_bwd_kernel<<<grid_for(M * (C / 8), 256 * 2), 256, 0, s>>>(h, du, dh, M, C);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int transpose_f32(const float* in, float* out, int R, int Cc, cudaStream_t s) {
  transpose_f32_kernel<<<(R * Cc + 255) / 256, 256, 0, s>>>(in, out, R, Cc);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int dwconv1d_fwd(const __nv_bfloat16* x, const float* w, const float* bias, __nv_bfloat16* y, int B, int T, int C, int K,
                 int flip, cudaStream_t s, int w_transposed) {
  SVSR_REQUIRE(C % 8 == 0 && K % 2 == 1 && K <= 31, "dwconv1d: C=%d K=%d unsupported", C, K);
  dwconv1d_kernel<<<grid_for((long long)B * T * (C / 8), 256), 256, 0, s>>>(x, w, bias, y, B, T, C, K, flip,
                                                                           w_transposed);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int dwconv1d_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* dbias, int B, int T, int C, int K,
                   cudaStream_t s) {
  SVSR_REQUIRE(C % 64 == 0 && K % 2 == 1 && K <= 31, "dwconv1d_wgrad: C=%d K=%d unsupported", C, K);
  const int chunk = 32, cpc = (T + chunk - 1) / chunk;
  dim3 grid(C / 64, (unsigned)(B * cpc));
  dwconv1d_wgrad_kernel<<<grid, 256, 0, s>>>(x, dy, dw, dbias, B, T, C, K, chunk, cpc);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bn_col_reduce(const __nv_bfloat16* x, const __nv_bfloat16* dout, const float* coef, long long rows, int C,
                  double* stats, int mode, cudaStream_t s) {
  SVSR_REQUIRE(C % 64 == 0, "bn_col_reduce: C=%d must be a multiple of 64", C);
  SVSR_REQUIRE(mode == 0 || (dout && coef), "bn_col_reduce: mode 1 needs dout and coef");
  long long chunks = (rows + 127) / 128;
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  dim3 grid(C / 64, (unsigned)chunks);
  bn_col_reduce_kernel<<<grid, 256, 0, s>>>(x, dout, coef, rows, C, stats, mode);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int add_f32(float* dst, const float* src, long long n, cudaStream_t s) {
  SVSR_REQUIRE(n % 4 == 0, "add_f32: n must be a multiple of 4");
  add_f32_kernel<<<grid_for(n / 4, 256 * 2), 256, 0, s>>>(dst, src, n / 4);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int cast_scale_f32_bf16(const float* x, __nv_bfloat16* y, long long n, float alpha, cudaStream_t s, float drop_p,
                        unsigned long long drop_seed, const unsigned long long* seed_base) {
  SVSR_REQUIRE(n % 4 == 0, "cast_scale: n must be a multiple of 4");
  cast_scale_kernel<<<grid_for(n / 4, 256 * 2), 256, 0, s>>>(x, y, n / 4, alpha, drop_p, drop_seed, seed_base);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int dropout_bf16(const __nv_bfloat16* x, __nv_bfloat16* y, long long n, float p, unsigned long long seed,
                 cudaStream_t s, const unsigned long long* seed_base) {
  SVSR_REQUIRE(p > 0.f && p < 1.f, "dropout: p=%f out of (0,1)", p);
  dropout_bf16_kernel<<<grid_for(n, 256 * 4), 256, 0, s>>>(x, y, n, p, seed, seed_base);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int dropout_add_bf16_to_f32(float* dst, const __nv_bfloat16* src, long long n, float p, unsigned long long seed,
                            cudaStream_t s, const unsigned long long* seed_base) {
  SVSR_REQUIRE(p > 0.f && p < 1.f, "dropout: p=%f out of (0,1)", p);
  dropout_add_kernel<<<grid_for(n, 256 * 4), 256, 0, s>>>(dst, src, n, p, seed, seed_base);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int dropout_mask_u8(unsigned char* out, long long n, float p, unsigned long long seed, cudaStream_t s) {
  dropout_mask_kernel<<<grid_for(n, 256 * 4), 256, 0, s>>>(out, n, p, seed);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int lengths_i64_to_i32(const long long* in, int* out, int n, int maxv, cudaStream_t s) {
  lengths_kernel<<<(n + 127) / 128, 128, 0, s>>>(in, out, n, maxv);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int meanpool_bf16(const __nv_bfloat16* a, __nv_bfloat16* out, long long N, int HW, int C, cudaStream_t s) {
  meanpool_kernel<<<grid_for(N * (C / 8), 128), 128, 0, s>>>(a, out, N, HW, C);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int meanpool_bf16_bwd(const __nv_bfloat16* df, __nv_bfloat16* dout, long long N, int HW, int C, cudaStream_t s) {
  meanpool_bwd_kernel<<<grid_for(N * (C / 8), 128), 128, 0, s>>>(df, dout, N, HW, C);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int rel_pos_table(__nv_bfloat16* pe, int T, int D, cudaStream_t s) {
  SVSR_REQUIRE(D % 2 == 0 && T >= 1, "rel_pos_table: bad geometry");
  rel_pos_table_kernel<<<grid_for((long long)(2 * T - 1) * (D / 2), 256), 256, 0, s>>>(pe, T, D);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int embed_posenc_fwd(const long long* tok, const float* emb, float* x, int rows, int L, int D, int V, cudaStream_t s,
                     float drop_p, unsigned long long drop_seed, const unsigned long long* seed_base) {
  embed_posenc_kernel<<<grid_for((long long)rows * (D / 2), 256), 256, 0, s>>>(tok, emb, x, rows, L, D, V, drop_p,
                                                                              drop_seed, seed_base);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int embed_bwd(const long long* tok, const float* dx, float* demb, int rows, int D, int V, cudaStream_t s, float drop_p,
              unsigned long long drop_seed, const unsigned long long* seed_base) {
  embed_bwd_kernel<<<grid_for((long long)rows * D, 256), 256, 0, s>>>(tok, dx, demb, rows, D, V, drop_p, drop_seed,
                                                                      seed_base);
  LAUNCH_CHECK();
  return SVSR_OK;
}

// =================================================================================================
// host launchers (part 2): attention, CTC, label smoothing
// =================================================================================================
namespace {
int attn_fill(const AttnProblem& p, AttnK& k) {
  SVSR_REQUIRE(p.q && p.k && p.v && p.o, "attention: null operand");
  SVSR_REQUIRE(p.B > 0 && p.H > 0 && p.Tq > 0 && p.Tk > 0, "attention: bad geometry");
  SVSR_REQUIRE(p.ldq % 8 == 0 && p.ldk % 8 == 0 && p.ldv % 8 == 0 && p.ldo % 8 == 0 && p.ldp % 8 == 0,
               "attention: row pitches must be multiples of 8 elements");
  SVSR_REQUIRE(!p.p || p.Tq == p.Tk, "attention: the relative-position term needs Tq == Tk");
  k.q = p.q, k.k = p.k, k.v = p.v, k.p = p.p;
  k.ldq = p.ldq, k.ldk = p.ldk, k.ldv = p.ldv, k.ldp = p.ldp;
  k.bu = p.bias_u, k.bv = p.bias_v, k.klen = p.klen, k.causal = p.causal;
  k.B = p.B, k.H = p.H, k.Tq = p.Tq, k.Tk = p.Tk, k.scale = p.scale;
  k.o = p.o, k.ldo = p.ldo, k.lse = p.lse;
  k.d_o = nullptr, k.dq = k.dk = k.dv = nullptr, k.lddq = k.lddk = k.lddv = 0;
  k.dp = k.dbu = k.dbv = k.Pg = k.DSg = nullptr;
  SVSR_REQUIRE(p.drop_p >= 0.f && p.drop_p < 1.f, "attention: dropout %f out of [0,1)", p.drop_p);
  k.drop_p = p.drop_p, k.drop_seed = p.drop_seed, k.seed_base = p.seed_base;
  return SVSR_OK;
}
template <class K>
int attn_smem_attr(K kernel, size_t bytes, size_t* cur) {
  SVSR_REQUIRE(bytes <= 227 * 1024, "attention: Tk too long for the shared-memory resident kernel (%zu bytes)", bytes);
  if (bytes > *cur) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    *cur = bytes;
  }
  return SVSR_OK;
}
// tcgen05 kernels (attention_rel_tc.cu) unless SVSR_ATTN_TC=0
bool attn_use_tc() {  // read per call: the parity tests run both paths in one process
  const char* e = getenv("SVSR_ATTN_TC");
  return !(e && e[0] == '0');
}
}  // namespace

int attention_core_fwd(const AttnProblem& p, cudaStream_t s) {
  AttnK k;
  int rc = attn_fill(p, k);
  if (rc) return rc;
  if (attn_use_tc()) return attention_rel_tc_fwd(k, s);
  static size_t attr = 48 * 1024;
  const size_t smem = attn_smem_bytes(p.Tk, p.p != nullptr, false);
  rc = attn_smem_attr(attention_core_fwd_kernel, smem, &attr);
  if (rc) return rc;
  dim3 grid((p.Tq + AQT - 1) / AQT, p.H, p.B);
  attention_core_fwd_kernel<<<grid, 256, smem, s>>>(k);
  LAUNCH_CHECK();
  return SVSR_OK;
}
size_t attention_scratch_bytes(int B, int H, int Tq, int Tk) {
  const size_t cuda_core = (size_t)2 * B * H * Tq * Tk * sizeof(float), tc = attention_rel_tc_scratch_bytes(B, H, Tq, Tk);
  return cuda_core > tc ? cuda_core : tc;
}

int attention_core_bwd(const AttnProblem& p, const AttnGrads& g, cudaStream_t s) {
  AttnK k;
  int rc = attn_fill(p, k);
  if (rc) return rc;
  SVSR_REQUIRE(g.d_o && g.dq && g.dk && g.dv && g.scratch && p.lse, "attention_bwd: null operand");
  SVSR_REQUIRE(!p.p || g.dp, "attention_bwd: dp missing");
  SVSR_REQUIRE(g.lddq % 8 == 0 && g.lddk % 8 == 0 && g.lddv % 8 == 0, "attention_bwd: gradient pitches must be multiples of 8");
  k.d_o = g.d_o, k.dq = g.dq, k.dk = g.dk, k.dv = g.dv, k.lddq = g.lddq, k.lddk = g.lddk, k.lddv = g.lddv;
  k.dp = g.dp, k.dbu = g.dbias_u, k.dbv = g.dbias_v;
  if (attn_use_tc()) return attention_rel_tc_bwd(k, g.scratch, s);
  k.Pg = g.scratch, k.DSg = g.scratch + (size_t)p.B * p.H * p.Tq * p.Tk;
  static size_t attr = 48 * 1024;
  const size_t smem = attn_smem_bytes(p.Tk, p.p != nullptr, true);
  rc = attn_smem_attr(attention_core_bwd_q_kernel, smem, &attr);
  if (rc) return rc;
  dim3 gq((p.Tq + AQT - 1) / AQT, p.H, p.B);
  attention_core_bwd_q_kernel<<<gq, 256, smem, s>>>(k);
  LAUNCH_CHECK();
  dim3 gk((p.Tk + 31) / 32, p.H, p.B);
  attention_core_bwd_kv_kernel<<<gk, 256, 0, s>>>(k);
  LAUNCH_CHECK();
  if (p.p) {
    dim3 gp((2 * p.Tk - 1 + 31) / 32, p.H, p.B);
    attention_core_bwd_pos_kernel<<<gp, 256, 0, s>>>(k);
    LAUNCH_CHECK();
  }
  return SVSR_OK;
}

size_t ctc_scratch_bytes(int B, int T, int Lmax) {
  const size_t S = 2 * (size_t)Lmax + 1;
  return ((size_t)B * T + 2 * (size_t)B * T * S + (size_t)B + 64) * sizeof(float);
}
int ctc_loss_fwd_bwd(const float* logits, int ld, int V, const long long* labels, int Lmax, const int* in_len, int B,
                     int T, __nv_bfloat16* dlogits, double* acc, int slot, float dscale, float* scratch, cudaStream_t s) {
  SVSR_REQUIRE(logits && labels && in_len && acc && scratch, "ctc: null operand");
  const int S = 2 * Lmax + 1;
  SVSR_REQUIRE(Lmax >= 1 && S <= 1024, "ctc: label length %d unsupported (2L+1 must be <= 1024)", Lmax);
  SVSR_REQUIRE((size_t)V * 4 <= 160 * 1024, "ctc: vocabulary %d too large for the shared-memory occupancy table", V);
  float* lse = scratch;
  float* alpha = lse + (size_t)B * T;
  float* beta = alpha + (size_t)B * T * S;
  float* nll = beta + (size_t)B * T * S;
  const long long rows = (long long)B * T;
  row_lse_kernel<<<grid_for(rows, 8), 256, 0, s>>>(logits, ld, V, rows, lse);
  LAUNCH_CHECK();
  const int threads = (S + 31) / 32 * 32;
  ctc_alpha_beta_kernel<<<B, threads, 2 * S * sizeof(float), s>>>(logits, ld, lse, labels, Lmax, in_len, T, S, alpha, beta,
                                                                  nll, acc, slot);
  LAUNCH_CHECK();
  if (dlogits) {
    static size_t attr = 48 * 1024;
    const size_t smem = (size_t)V * 4;
    if (smem > attr) {
      SVSR_CHECK_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = smem;
    }
    ctc_grad_kernel<<<(unsigned)rows, 256, smem, s>>>(logits, ld, V, lse, labels, Lmax, in_len, T, S, alpha, beta, nll,
                                                     dlogits, dscale);
    LAUNCH_CHECK();
  }
  return SVSR_OK;
}
int label_smoothing_loss(const float* logits, int ld, int V, const long long* target, int rows, float smoothing,
                         __nv_bfloat16* dlogits, double* acc, int slot, float dscale, cudaStream_t s) {
  SVSR_REQUIRE(V >= 2, "label_smoothing_loss: V=%d", V);
  label_smoothing_kernel<<<grid_for(rows, 8), 256, 0, s>>>(logits, ld, V, target, rows, smoothing, dlogits, acc, slot,
                                                          dscale);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int lrs_finalize_metrics(const double* acc, float* out, int B, long long audio_rows, float mtlalpha, float audio_weight,
                         int has_audio, cudaStream_t s, const int* bad) {
  lrs_finalize_kernel<<<1, 1, 0, s>>>(acc, out, B, audio_rows, mtlalpha, audio_weight, has_audio, bad);
  LAUNCH_CHECK();
  return SVSR_OK;
}

}  // namespace svsr
