#include "encoder.cuh"
#include <stdlib.h>

namespace svsr {

namespace {

// erf via Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of what these kernels store):
// 2 MUFU + ~8 FMA instead of libdevice's branchy erff
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return copysignf(1.0f - p * t * __expf(-ax * ax), x);
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.0f + fast_erf(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y, v[4] = c.x, v[5] = c.y, v[6] = d.x, v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]), u.y = pack_bf16x2(v[2], v[3]), u.z = pack_bf16x2(v[4], v[5]), u.w = pack_bf16x2(v[6], v[7]);
  return u;
}

// ------------------------------------------------------------------------------------------------
// RMSNorm: one warp per row, D <= 1024 (D % 32 == 0); lanes own strided elements j = lane + 32*i
// ------------------------------------------------------------------------------------------------
template <int MAXI>
__global__ void __launch_bounds__(256)
rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g, __nv_bfloat16* __restrict__ y,
                   float* __restrict__ inv_out, int M, int D, float eps, int norm_dim, const StepCtl ctl) {
  if (ctl_skipped(ctl)) return;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int ni = D >> 5;
  const float rs = rsqrtf((float)norm_dim);
  for (int row = warp; row < M; row += nwarps) {
    const float* xr = x + (long long)row * D;
    float v[MAXI];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i)
      if (i < ni) {
        v[i] = xr[lane + 32 * i];
        ss += v[i] * v[i];
      }
    ss = warp_sum(ss);
    const float norm = sqrtf(ss) * rs;
    const float inv = 1.0f / fmaxf(norm, eps);
    if (lane == 0) inv_out[row] = inv;
#pragma unroll
    for (int i = 0; i < MAXI; ++i)
      if (i < ni) y[(long long)row * D + lane + 32 * i] = __float2bfloat16(v[i] * inv * g[lane + 32 * i]);
  }
}

template <int MAXI>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ g,
                   const float* __restrict__ inv_in, float* __restrict__ dx, __nv_bfloat16* __restrict__ dxb,
                   float* __restrict__ dg, int M, int D, float eps, int norm_dim, const StepCtl ctl) {
  if (ctl_skipped(ctl)) {  // dropped sublayer: the stream gradient passes through; only its bf16 copy moves on
    const long long n = (long long)M * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      dxb[i] = __float2bfloat16(dx[i]);
    return;
  }
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int ni = D >> 5;
  float dgacc[MAXI];
#pragma unroll
  for (int i = 0; i < MAXI; ++i) dgacc[i] = 0.f;
  for (int row = warp; row < M; row += nwarps) {
    const long long base = (long long)row * D;
    const float inv = inv_in[row];
    const bool clamped = inv >= 1.0f / eps;  // norm <= eps: the scale is a constant there
    float xv[MAXI], dv[MAXI];
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < MAXI; ++i)
      if (i < ni) {
        const int j = lane + 32 * i;
        xv[i] = x[base + j];
        dv[i] = __bfloat162float(dy[base + j]);
        t += dv[i] * g[j] * xv[i];
        dgacc[i] += dv[i] * xv[i] * inv;
      }
    t = warp_sum(t);
    const float coef = clamped ? 0.f : inv * inv * inv / (float)norm_dim * t;
#pragma unroll
    for (int i = 0; i < MAXI; ++i)
      if (i < ni) {
        const int j = lane + 32 * i;
        const float nv = dx[base + j] + g[j] * inv * dv[i] - coef * xv[i];
        dx[base + j] = nv;
        dxb[base + j] = __float2bfloat16(nv);
      }
  }
  // reduce dg over the block's warps, then one atomic per element per block
  __shared__ float sdg[8][32 * MAXI];
  const int w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < MAXI; ++i) sdg[w][lane + 32 * i] = dgacc[i];
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) s += sdg[ww][j];
    atomicAdd(dg + j, s);
  }
}

__global__ void rotary_table_kernel(float* tab, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 16) return;
  const int pos = i / 16, f = i % 16;
  const float inv_freq = 1.0f / powf(10000.0f, (float)(2 * f) / 32.0f);
  const float ang = (float)pos * inv_freq;
  tab[pos * 32 + f] = cosf(ang);
  tab[pos * 32 + 16 + f] = sinf(ang);
}

// ------------------------------------------------------------------------------------------------
// attention core, one CTA per (batch, head); n <= 64 tokens, head dim 64. Everything lives in smem as fp32.
// ------------------------------------------------------------------------------------------------
constexpr int AT_MAXN = 64;
constexpr int AT_D = 64;
constexpr int AT_LD = 65;  // padded row pitch (floats) against bank conflicts

// q, k, v (and dO in backward) of one (batch, head) in ONE pass: all global loads of a thread are in flight together and
// the CTA synchronises twice in total, instead of one dependent global round trip + two barriers per tensor.
__device__ __forceinline__ void load_qkv_rot(const __nv_bfloat16* __restrict__ base, int ld, int inner,
                                             const __nv_bfloat16* __restrict__ dout, const float* __restrict__ rot,
                                             float* sq, float* sk, float* sv, float* sdo, int n, bool rotary_v) {
  for (int i = threadIdx.x; i < n * 32; i += blockDim.x) {
    const int pos = i >> 5, d2 = (i & 31) * 2;
    const __nv_bfloat16* src = base + (long long)pos * ld + d2;
    const __nv_bfloat162 q = *reinterpret_cast<const __nv_bfloat162*>(src);
    const __nv_bfloat162 k = *reinterpret_cast<const __nv_bfloat162*>(src + inner);
    const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + 2 * inner);
    __nv_bfloat162 g = q;
    if (dout) g = *reinterpret_cast<const __nv_bfloat162*>(dout + (long long)pos * inner + d2);
    const int o = pos * AT_LD + d2;
    sq[o] = __low2float(q), sq[o + 1] = __high2float(q);
    sk[o] = __low2float(k), sk[o + 1] = __high2float(k);
    sv[o] = __low2float(v), sv[o + 1] = __high2float(v);
    if (dout) sdo[o] = __low2float(g), sdo[o + 1] = __high2float(g);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n * 16; i += blockDim.x) {
    const int pos = i >> 4, f = i & 15;
    const float c = rot[pos * 32 + f], s = rot[pos * 32 + 16 + f];
    const int o = pos * AT_LD + f;
    float a = sq[o], b = sq[o + 16];
    sq[o] = a * c - b * s, sq[o + 16] = b * c + a * s;
    a = sk[o], b = sk[o + 16];
    sk[o] = a * c - b * s, sk[o + 16] = b * c + a * s;
    if (rotary_v) {
      a = sv[o], b = sv[o + 16];
      sv[o] = a * c - b * s, sv[o + 16] = b * c + a * s;
    }
  }
  __syncthreads();
}

// out[r][c] = scale * sum_d A[r][d] * B[c][d]  (both [n x 64] in smem): register tile of 2 rows x 4 columns per thread,
// 8 FMAs per 6 shared loads, bank-conflict free with the 65-float pitch.
__device__ __forceinline__ void mm_abt_item(const float* sa, const float* sb, float* out, int n, float scale, int t) {
  const int ncq = (n + 3) >> 2;
  {
    const int rp = t / ncq, cq = t - rp * ncq;
    const int r0 = 2 * rp, r1 = min(r0 + 1, n - 1);
    int c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j] = min(4 * cq + j, n - 1);
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 8
    for (int d = 0; d < AT_D; ++d) {
      const float a0 = sa[r0 * AT_LD + d], a1 = sa[r1 * AT_LD + d];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float b = sb[c[j] * AT_LD + d];
        acc[0][j] += a0 * b;
        acc[1][j] += a1 * b;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * cq + j < n) {
        out[r0 * AT_LD + 4 * cq + j] = acc[0][j] * scale;
        if (r0 + 1 < n) out[(r0 + 1) * AT_LD + 4 * cq + j] = acc[1][j] * scale;
      }
  }
}

__device__ __forceinline__ void mm_abt(const float* sa, const float* sb, float* out, int n, float scale) {
  const int items = ((n + 1) >> 1) * ((n + 3) >> 2);
  for (int t = threadIdx.x; t < items; t += blockDim.x) mm_abt_item(sa, sb, out, n, scale, t);
}
// two independent products in one sweep (backward: S = Q K^T / 8 and dP = dO V^T), so that a 256-thread CTA runs both at once
__device__ __forceinline__ void mm_abt_pair(const float* a0, const float* b0, float* o0, float s0, const float* a1,
                                            const float* b1, float* o1, float s1, int n) {
  const int items = ((n + 1) >> 1) * ((n + 3) >> 2);
  const int split = (items + 31) & ~31;  // the second product starts on a warp boundary: no warp runs both loops
  for (int t = threadIdx.x; t < split + items; t += blockDim.x) {
    if (t < items)
      mm_abt_item(a0, b0, o0, n, s0, t);
    else if (t >= split)
      mm_abt_item(a1, b1, o1, n, s1, t - split);
  }
}

__device__ __forceinline__ void softmax_rows(float* sp, int n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < n; r += nw) {
    const float a = lane < n ? sp[r * AT_LD + lane] : -INFINITY;
    const float b = lane + 32 < n ? sp[r * AT_LD + lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(a, b));
    const float ea = lane < n ? expf(a - m) : 0.f, eb = lane + 32 < n ? expf(b - m) : 0.f;
    const float inv = 1.0f / warp_sum(ea + eb);
    if (lane < n) sp[r * AT_LD + lane] = ea * inv;
    if (lane + 32 < n) sp[r * AT_LD + lane + 32] = eb * inv;
  }
}

// acc[i][k] = sum_j L(r_i, j) * R[j][cq + 8k]; L(r, j) = Lm[r][j] or, transposed, Lm[j][r]. 2 rows x 8 strided columns
// per thread (16 FMAs per 10 shared loads); the column stride of 8 keeps a rotary pair (f, f+16) in one thread.
template <bool TRANS>
__device__ __forceinline__ void mm_tile_2x8(const float* Lm, const float* R, int n, int r0, int r1, int cq,
                                            float (&acc)[2][8]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
  for (int j = 0; j < n; ++j) {
    const float p0 = TRANS ? Lm[j * AT_LD + r0] : Lm[r0 * AT_LD + j];
    const float p1 = TRANS ? Lm[j * AT_LD + r1] : Lm[r1 * AT_LD + j];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float v = R[j * AT_LD + cq + 8 * k];
      acc[0][k] += p0 * v;
      acc[1][k] += p1 * v;
    }
  }
}

// p[i][j] *= keep / (1 - p_drop) (and, if `other` is given, the same mask applied to other[i][j]: dP in backward)
__device__ __forceinline__ void dropout_probs(float* sp, float* other, int n, unsigned long long base, float p,
                                              unsigned long long seed) {
  const float ks = 1.0f / (1.0f - p);
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
    const int i = t / n, j = t - i * n;
    const float m = dropout_keep(seed, base + (unsigned long long)t, p) ? ks : 0.f;
    sp[i * AT_LD + j] *= m;
    if (other) other[i * AT_LD + j] *= m;
  }
}

__global__ void __launch_bounds__(128)
attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rot,
                     __nv_bfloat16* __restrict__ o, int n, int heads, int rotary_v, float drop_p,
                     unsigned long long drop_seed, const StepCtl ctl) {
  if (ctl_skipped(ctl)) return;
  drop_seed = ctl_seed(ctl, drop_seed);
  extern __shared__ float sm[];
  float* sq = sm;
  float* sk = sq + n * AT_LD;
  float* sv = sk + n * AT_LD;
  float* sp = sv + n * AT_LD;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int inner = heads * AT_D, ld = 3 * inner;
  const __nv_bfloat16* base = qkv + (long long)b * n * ld + h * AT_D;
  load_qkv_rot(base, ld, inner, nullptr, rot, sq, sk, sv, nullptr, n, rotary_v != 0);
  mm_abt(sq, sk, sp, n, 0.125f);
  __syncthreads();
  softmax_rows(sp, n);
  __syncthreads();
  if (drop_p > 0.f) {  // Attention(dropout=attn_dropout): mask on the probabilities, index ((b*H+h)*n + i)*n + j
    dropout_probs(sp, nullptr, n, (unsigned long long)blockIdx.x * n * n, drop_p, drop_seed);
    __syncthreads();
  }
  const int nrp = (n + 1) >> 1;
  for (int t = threadIdx.x; t < nrp * 8; t += blockDim.x) {
    const int rp = t >> 3, cq = t & 7;
    const int r0 = 2 * rp, r1 = min(r0 + 1, n - 1);
    float acc[2][8];
    mm_tile_2x8<false>(sp, sv, n, r0, r1, cq, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (r0 + i >= n) continue;
      __nv_bfloat16* dst = o + ((long long)b * n + r0 + i) * inner + h * AT_D + cq;
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[8 * k] = __float2bfloat16(acc[i][k]);
    }
  }
}

__global__ void __launch_bounds__(256)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ rot,
                     const __nv_bfloat16* __restrict__ d_o, __nv_bfloat16* __restrict__ dqkv, int n, int heads,
                     int rotary_v, float drop_p, unsigned long long drop_seed, const StepCtl ctl) {
  if (ctl_skipped(ctl)) return;
  drop_seed = ctl_seed(ctl, drop_seed);
  extern __shared__ float sm[];
  float* sq = sm;
  float* sk = sq + n * AT_LD;
  float* sv = sk + n * AT_LD;
  float* sp = sv + n * AT_LD;
  float* sdo = sp + n * AT_LD;
  float* sds = sdo + n * AT_LD;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int inner = heads * AT_D, ld = 3 * inner;
  const __nv_bfloat16* base = qkv + (long long)b * n * ld + h * AT_D;
  load_qkv_rot(base, ld, inner, d_o + (long long)b * n * inner + h * AT_D, rot, sq, sk, sv, sdo, n, rotary_v != 0);
  mm_abt_pair(sq, sk, sp, 0.125f, sdo, sv, sds, 1.0f, n);  // S = Q' K'^T / 8 and dP = dO V'^T
  __syncthreads();
  softmax_rows(sp, n);
  __syncthreads();
  {  // dS = P o (dP - rowsum(dP o P))
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < n; r += nw) {
      const float pa = lane < n ? sp[r * AT_LD + lane] : 0.f, pb = lane + 32 < n ? sp[r * AT_LD + lane + 32] : 0.f;
      const float da = lane < n ? sds[r * AT_LD + lane] : 0.f, db = lane + 32 < n ? sds[r * AT_LD + lane + 32] : 0.f;
      float ma = 1.f, mb = 1.f;
      if (drop_p > 0.f) {  // d p = mask * d p~ ; dV below uses p~ = mask * p
        const float ks = 1.0f / (1.0f - drop_p);
        const unsigned long long base = ((unsigned long long)blockIdx.x * n + r) * n;
        ma = (lane < n && dropout_keep(drop_seed, base + lane, drop_p)) ? ks : 0.f;
        mb = (lane + 32 < n && dropout_keep(drop_seed, base + lane + 32, drop_p)) ? ks : 0.f;
      }
      const float dam = da * ma, dbm = db * mb;
      const float dot = warp_sum(pa * dam + pb * dbm);
      if (lane < n) sds[r * AT_LD + lane] = pa * (dam - dot), sp[r * AT_LD + lane] = pa * ma;
      if (lane + 32 < n) sds[r * AT_LD + lane + 32] = pb * (dbm - dot), sp[r * AT_LD + lane + 32] = pb * mb;
    }
  }
  __syncthreads();
  // dQ' = scale dS K', dK' = scale dS^T Q', dV' = P^T dO; rotate back (transpose of the rotation) and store
  __nv_bfloat16* obase = dqkv + (long long)b * n * ld + h * AT_D;
  const int nrp = (n + 1) >> 1;
  for (int t = threadIdx.x; t < nrp * 8 * 3; t += blockDim.x) {
    const int which = t / (nrp * 8);  // 0 = q, 1 = k, 2 = v
    const int rem = t - which * nrp * 8;
    const int rp = rem >> 3, cq = rem & 7;
    const int r0 = 2 * rp, r1 = min(r0 + 1, n - 1);
    float acc[2][8];
    if (which == 0)
      mm_tile_2x8<false>(sds, sk, n, r0, r1, cq, acc);
    else if (which == 1)
      mm_tile_2x8<true>(sds, sq, n, r0, r1, cq, acc);
    else
      mm_tile_2x8<true>(sp, sdo, n, r0, r1, cq, acc);
    const float sc = which < 2 ? 0.125f : 1.0f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + i;
      if (r >= n) continue;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = acc[i][k] * sc;
      if (which < 2 || rotary_v) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {  // columns f = cq + 8k (< 16) pair with f + 16 = cq + 8(k+2)
          const int f = cq + 8 * k;
          const float c = rot[r * 32 + f], s_ = rot[r * 32 + 16 + f];
          const float a = v[k], bb = v[k + 2];
          v[k] = a * c + bb * s_;
          v[k + 2] = -a * s_ + bb * c;
        }
      }
      __nv_bfloat16* dst = obase + which * inner + (long long)r * ld + cq;
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[8 * k] = __float2bfloat16(v[k]);
    }
  }
}

// u = dropout_p(value * gelu(gate))   (x-transformers FeedForward: GLU -> Dropout(ff_dropout) -> Linear)
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(const __nv_bfloat16* __restrict__ h, __nv_bfloat16* __restrict__ u, long long M, int F, float p,
                 unsigned long long seed, const StepCtl ctl) {
  if (ctl_skipped(ctl)) return;
  seed = ctl_seed(ctl, seed);
  const int fg = F >> 3;  // 8 hidden units (16 bytes) per thread
  const long long total = M * fg;
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / fg;
    const int c = (int)(i % fg) * 8;
    float v[8], gt[8], o[8];
    unpack8(*reinterpret_cast<const uint4*>(h + r * 2 * F + c), v);
    unpack8(*reinterpret_cast<const uint4*>(h + r * 2 * F + F + c), gt);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      o[k] = v[k] * gelu_f(gt[k]);
      if (p > 0.f) o[k] = dropout_keep(seed, (unsigned long long)(r * F + c + k), p) ? o[k] * keep_scale : 0.f;
    }
    *reinterpret_cast<uint4*>(u + r * F + c) = pack8(o);
  }
}
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ du,
                 __nv_bfloat16* __restrict__ dh, long long M, int F, float p, unsigned long long seed,
                 const StepCtl ctl) {
  if (ctl_skipped(ctl)) return;
  seed = ctl_seed(ctl, seed);
  const int fg = F >> 3;
  const long long total = M * fg;
  const float keep_scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / fg;
    const int c = (int)(i % fg) * 8;
    float v[8], gt[8], d[8], dv[8], dg[8];
    unpack8(*reinterpret_cast<const uint4*>(h + r * 2 * F + c), v);
    unpack8(*reinterpret_cast<const uint4*>(h + r * 2 * F + F + c), gt);
    unpack8(*reinterpret_cast<const uint4*>(du + r * F + c), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float dk = d[k];
      if (p > 0.f) dk = dropout_keep(seed, (unsigned long long)(r * F + c + k), p) ? dk * keep_scale : 0.f;
      dv[k] = dk * gelu_f(gt[k]);
      dg[k] = dk * v[k] * gelu_grad_f(gt[k]);
    }
    *reinterpret_cast<uint4*>(dh + r * 2 * F + c) = pack8(dv);
    *reinterpret_cast<uint4*>(dh + r * 2 * F + F + c) = pack8(dg);
  }
}

// ---- HuggingFace BERT encoder pieces (lightning.py:90-92,152-156; transformers BertEmbeddings / BertIntermediate) ----
// h = gelu(pre) (erf form, hidden_act "gelu"), dpre = dh * gelu'(pre); 8 elements per thread
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ pre, __nv_bfloat16* __restrict__ h, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8];
    unpack8(reinterpret_cast<const uint4*>(pre)[i], v);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = gelu_f(v[k]);
    reinterpret_cast<uint4*>(h)[i] = pack8(v);
  }
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ pre, const __nv_bfloat16* __restrict__ dh,
                __nv_bfloat16* __restrict__ dpre, long long n8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float v[8], d[8];
    unpack8(reinterpret_cast<const uint4*>(pre)[i], v);
    unpack8(reinterpret_cast<const uint4*>(dh)[i], d);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] *= gelu_grad_f(v[k]);
    reinterpret_cast<uint4*>(dpre)[i] = pack8(d);
  }
}
// E[m,:] = x[m,:] + pos[m % L, :] + tt[0,:]   (inputs_embeds + position_embeddings + token_type_embeddings(0))
__global__ void bert_embed_kernel(const float* __restrict__ x, const float* __restrict__ pos, const float* __restrict__ tt,
                                  float* __restrict__ E, long long M, int L, int D) {
  const long long total = M * (D / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (D / 4));
    const long long m = i / (D / 4);
    const float4 a = reinterpret_cast<const float4*>(x + m * D)[c];
    const float4 p = reinterpret_cast<const float4*>(pos + (m % L) * D)[c];
    const float4 t = reinterpret_cast<const float4*>(tt)[c];
    reinterpret_cast<float4*>(E + m * D)[c] = make_float4(a.x + p.x + t.x, a.y + p.y + t.y, a.z + p.z + t.z, a.w + p.w + t.w);
  }
}
// dpos[l,:] += sum_b dE[b*L + l, :]; dtt[0,:] += sum_m dE[m,:]
__global__ void bert_embed_bwd_kernel(const float* __restrict__ dE, float* __restrict__ dpos, float* __restrict__ dtt,
                                      int B, int L, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * D) return;
  const int l = i / D, d = i % D;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += dE[((long long)b * L + l) * D + d];
  dpos[i] += s;
  atomicAdd(dtt + d, s);
}
// Dropout on an fp32 tensor in place (+ optional bf16 copy): forward of nn.Dropout after the embedding LayerNorm, and
// (same call on the gradient) its backward
__global__ void dropout_f32_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ xb, long long n, float p,
                                   unsigned long long seed, const StepCtl ctl) {
  seed = ctl_seed(ctl, seed);
  const float ks = 1.0f / (1.0f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = dropout_keep(seed, (unsigned long long)i, p) ? x[i] * ks : 0.f;
    x[i] = v;
    if (xb) xb[i] = __float2bfloat16(v);
  }
}

}  // namespace

#define LAUNCH_CHECK() \
  note_launch();       \
  SVSR_CHECK_CUDA(cudaGetLastError())

int rmsnorm_fwd(const float* x, const float* g, __nv_bfloat16* y, float* inv, int M, int D, float eps, cudaStream_t s,
                int norm_dim, const StepCtl* ctl) {
  if (norm_dim <= 0) norm_dim = D;
  SVSR_REQUIRE(D % 32 == 0 && D <= 1024, "rmsnorm: D=%d unsupported", D);
  const int blocks = (M + 7) / 8 < 148 * 4 ? (M + 7) / 8 : 148 * 4;
  if (D <= 512)
    rmsnorm_fwd_kernel<16><<<blocks, 256, 0, s>>>(x, g, y, inv, M, D, eps, norm_dim, ctl ? *ctl : StepCtl());
  else
    rmsnorm_fwd_kernel<32><<<blocks, 256, 0, s>>>(x, g, y, inv, M, D, eps, norm_dim, ctl ? *ctl : StepCtl());
  LAUNCH_CHECK();
  return SVSR_OK;
}
// D % 128 == 0 variant (the dim-512 stream): every lane owns NCH chunks of 4 consecutive columns, so a row is NCH
// 16-byte loads per operand instead of 4 * NCH scalar ones, and the old dx values are fetched together with x / dy --
// before the row reduction, not after it. The kernel is latency bound (one row per warp, ncu: 12 % warps active): what
// it costs is the length of that dependent chain.
template <int NCH>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_vec_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ g,
                       const float* __restrict__ inv_in, float* __restrict__ dx, __nv_bfloat16* __restrict__ dxb,
                       float* __restrict__ dg, int M, int D, float eps, int norm_dim, const StepCtl ctl) {
  if (ctl_skipped(ctl)) {  // dropped sublayer: the stream gradient passes through; only its bf16 copy moves on
    const long long n = (long long)M * D;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
      dxb[i] = __float2bfloat16(dx[i]);
    return;
  }
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 gv[NCH], dgacc[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    gv[i] = *reinterpret_cast<const float4*>(g + 128 * i + 4 * lane);
    dgacc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = warp; row < M; row += nwarps) {
    const long long base = (long long)row * D + 4 * lane;
    float4 xv[NCH], dv[NCH], ov[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      xv[i] = *reinterpret_cast<const float4*>(x + base + 128 * i);
      const uint2 d2 = *reinterpret_cast<const uint2*>(dy + base + 128 * i);
      const float2 lo = unpack_bf16x2(d2.x), hi = unpack_bf16x2(d2.y);
      dv[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
      ov[i] = *reinterpret_cast<const float4*>(dx + base + 128 * i);
    }
    const float inv = inv_in[row];
    const bool clamped = inv >= 1.0f / eps;  // norm <= eps: the scale is a constant there
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      t += dv[i].x * gv[i].x * xv[i].x + dv[i].y * gv[i].y * xv[i].y + dv[i].z * gv[i].z * xv[i].z +
           dv[i].w * gv[i].w * xv[i].w;
      dgacc[i].x += dv[i].x * xv[i].x * inv, dgacc[i].y += dv[i].y * xv[i].y * inv;
      dgacc[i].z += dv[i].z * xv[i].z * inv, dgacc[i].w += dv[i].w * xv[i].w * inv;
    }
    t = warp_sum(t);
    const float coef = clamped ? 0.f : inv * inv * inv / (float)norm_dim * t;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float4 nv;
      nv.x = ov[i].x + gv[i].x * inv * dv[i].x - coef * xv[i].x;
      nv.y = ov[i].y + gv[i].y * inv * dv[i].y - coef * xv[i].y;
      nv.z = ov[i].z + gv[i].z * inv * dv[i].z - coef * xv[i].z;
      nv.w = ov[i].w + gv[i].w * inv * dv[i].w - coef * xv[i].w;
      *reinterpret_cast<float4*>(dx + base + 128 * i) = nv;
      uint2 pk;
      pk.x = pack_bf16x2(nv.x, nv.y), pk.y = pack_bf16x2(nv.z, nv.w);
      *reinterpret_cast<uint2*>(dxb + base + 128 * i) = pk;
    }
  }
  // reduce dg over the block's warps, then one atomic per element per block
  __shared__ float4 sdg[8][32 * NCH];
  const int w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NCH; ++i) sdg[w][32 * i + lane] = dgacc[i];  // float4 slot 32 i + lane = columns 128 i + 4 lane ..
  __syncthreads();
  const float* sflat = reinterpret_cast<const float*>(&sdg[0][0]);
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float sum = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) sum += sflat[ww * 128 * NCH + j];
    atomicAdd(dg + j, sum);
  }
}

int rmsnorm_bwd(const __nv_bfloat16* dy, const float* x, const float* g, const float* inv, float* dx,
                __nv_bfloat16* dx_bf16, float* dg, int M, int D, float eps, cudaStream_t s, int norm_dim,
                const StepCtl* ctl) {
  if (norm_dim <= 0) norm_dim = D;
  SVSR_REQUIRE(D % 32 == 0 && D <= 1024, "rmsnorm: D=%d unsupported", D);
  // one row per warp: the kernel is latency bound (ncu: 12 % warps active, every pipe < 5 %), so more resident warps
  // win over fewer column atomics (measured: 240 CTAs 19 us, 120 CTAs 32 us at M = 1920)
  const int blocks = (M + 7) / 8 < 148 * 2 ? (M + 7) / 8 : 148 * 2;
  const bool al16 = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx_bf16)) & 7) == 0 &&
                    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  if (D == 512 && al16)
    rmsnorm_bwd_vec_kernel<4><<<blocks, 256, 0, s>>>(dy, x, g, inv, dx, dx_bf16, dg, M, D, eps, norm_dim, ctl ? *ctl : StepCtl());
  else if (D <= 512)
    rmsnorm_bwd_kernel<16><<<blocks, 256, 0, s>>>(dy, x, g, inv, dx, dx_bf16, dg, M, D, eps, norm_dim, ctl ? *ctl : StepCtl());
  else
    rmsnorm_bwd_kernel<32><<<blocks, 256, 0, s>>>(dy, x, g, inv, dx, dx_bf16, dg, M, D, eps, norm_dim, ctl ? *ctl : StepCtl());
  LAUNCH_CHECK();
  return SVSR_OK;
}
int rotary_table(float* tab, int n, cudaStream_t s) {
  rotary_table_kernel<<<(n * 16 + 127) / 128, 128, 0, s>>>(tab, n);
  LAUNCH_CHECK();
  return SVSR_OK;
}
// SVSR_ATTN_TC=0 keeps the CUDA-core fp32 kernels below (A/B); the default is the tcgen05 pair of attention_tc.cu.
static bool attn_use_tc() {
  const char* e = getenv("SVSR_ATTN_TC");
  return !(e && e[0] == '0');
}
int attention_fwd(const __nv_bfloat16* qkv, const float* rot, __nv_bfloat16* o, int B, int n, int heads,
                  int rotary_v, cudaStream_t s, float drop_p, unsigned long long drop_seed, const StepCtl* ctl) {
  SVSR_REQUIRE(n >= 1 && n <= AT_MAXN, "attention: n=%d must be in [1,%d]", n, AT_MAXN);
  if (attn_use_tc()) return attention_tc_fwd(qkv, rot, o, B, n, heads, rotary_v, s, drop_p, drop_seed, ctl);
  const int smem = 4 * n * AT_LD * sizeof(float);
  const int smem_max = 4 * AT_MAXN * AT_LD * sizeof(float);
  static bool done = false;
  if (!done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    done = true;
  }
  attention_fwd_kernel<<<B * heads, 128, smem, s>>>(qkv, rot, o, n, heads, rotary_v, drop_p, drop_seed,
                                                    ctl ? *ctl : StepCtl());
  LAUNCH_CHECK();
  return SVSR_OK;
}
int attention_bwd(const __nv_bfloat16* qkv, const float* rot, const __nv_bfloat16* d_o, __nv_bfloat16* dqkv, int B,
                  int n, int heads, int rotary_v, cudaStream_t s, float drop_p, unsigned long long drop_seed,
                  const StepCtl* ctl) {
  SVSR_REQUIRE(n >= 1 && n <= AT_MAXN, "attention: n=%d must be in [1,%d]", n, AT_MAXN);
  if (attn_use_tc()) return attention_tc_bwd(qkv, rot, d_o, dqkv, B, n, heads, rotary_v, s, drop_p, drop_seed, ctl);
  const int smem = 6 * n * AT_LD * sizeof(float);
  const int smem_max = 6 * AT_MAXN * AT_LD * sizeof(float);
  static bool done = false;
  if (!done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    done = true;
  }
  // 256 threads: S and dP (2 x 120 register tiles at n = 30) in one round, the three gradient products (360) in two
  attention_bwd_kernel<<<B * heads, 256, smem, s>>>(qkv, rot, d_o, dqkv, n, heads, rotary_v, drop_p, drop_seed,
                                                    ctl ? *ctl : StepCtl());
  LAUNCH_CHECK();
  return SVSR_OK;
}
int geglu_fwd(const __nv_bfloat16* h, __nv_bfloat16* u, int M, int F, float p, unsigned long long seed, cudaStream_t s,
              const StepCtl* ctl) {
  SVSR_REQUIRE(F % 8 == 0, "geglu: F=%d must be a multiple of 8", F);
  const long long total = (long long)M * (F / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  geglu_fwd_kernel<<<blocks, 256, 0, s>>>(h, u, M, F, p, seed, ctl ? *ctl : StepCtl());
  LAUNCH_CHECK();
  return SVSR_OK;
}
int geglu_bwd(const __nv_bfloat16* h, const __nv_bfloat16* du, __nv_bfloat16* dh, int M, int F, float p,
              unsigned long long seed, cudaStream_t s, const StepCtl* ctl) {
  SVSR_REQUIRE(F % 8 == 0, "geglu: F=%d must be a multiple of 8", F);
  const long long total = (long long)M * (F / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  geglu_bwd_kernel<<<blocks, 256, 0, s>>>(h, du, dh, M, F, p, seed, ctl ? *ctl : StepCtl());
  LAUNCH_CHECK();
  return SVSR_OK;
}

int gelu_fwd(const __nv_bfloat16* pre, __nv_bfloat16* h, long long n, cudaStream_t s) {
  SVSR_REQUIRE(n % 8 == 0, "gelu: n must be a multiple of 8");
  const long long n8 = n / 8;
  gelu_fwd_kernel<<<(unsigned)((n8 + 255) / 256 < 148 * 8 ? (n8 + 255) / 256 : 148 * 8), 256, 0, s>>>(pre, h, n8);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int gelu_bwd(const __nv_bfloat16* pre, const __nv_bfloat16* dh, __nv_bfloat16* dpre, long long n, cudaStream_t s) {
  SVSR_REQUIRE(n % 8 == 0, "gelu: n must be a multiple of 8");
  const long long n8 = n / 8;
  gelu_bwd_kernel<<<(unsigned)((n8 + 255) / 256 < 148 * 8 ? (n8 + 255) / 256 : 148 * 8), 256, 0, s>>>(pre, dh, dpre, n8);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bert_embed_fwd(const float* x, const float* pos, const float* tt, float* E, long long M, int L, int D, cudaStream_t s) {
  SVSR_REQUIRE(D % 4 == 0, "bert_embed: D must be a multiple of 4");
  const long long total = M * (D / 4);
  bert_embed_kernel<<<(unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8), 256, 0, s>>>(x, pos, tt, E, M, L, D);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int bert_embed_bwd(const float* dE, float* dpos, float* dtt, int B, int L, int D, cudaStream_t s) {
  bert_embed_bwd_kernel<<<(L * D + 255) / 256, 256, 0, s>>>(dE, dpos, dtt, B, L, D);
  LAUNCH_CHECK();
  return SVSR_OK;
}
int dropout_f32_inplace(float* x, __nv_bfloat16* xb, long long n, float p, unsigned long long seed, cudaStream_t s,
                        const StepCtl* ctl) {
  SVSR_REQUIRE(p > 0.f && p < 1.f, "dropout: p=%f out of (0,1)", p);
  dropout_f32_kernel<<<(unsigned)((n + 1023) / 1024 < 148 * 8 ? (n + 1023) / 1024 : 148 * 8), 256, 0, s>>>(x, xb, n, p, seed, ctl ? *ctl : StepCtl());
  LAUNCH_CHECK();
  return SVSR_OK;
}

}  // namespace svsr
