#include "precise.cuh"

namespace svsr {
namespace {

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ void split2(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(x);
  lo = __float2bfloat16(x - __bfloat162float(hi));
}
#define GRID_STRIDE(i, total) \
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (total); i += (long long)gridDim.x * blockDim.x)

// x fp32 [rows, C] at pitch ldx -> [rows, 3 Cp] = hi | lo | hi, columns C..Cp of every third are zero
__global__ void split3_kernel(const float* __restrict__ x, int ldx, __nv_bfloat16* __restrict__ out, long long rows, int C,
                              int Cp) {
  GRID_STRIDE(i, rows * Cp) {
    const long long r = i / Cp;
    const int c = (int)(i % Cp);
    __nv_bfloat16 hi, lo;
    split2(c < C ? x[r * ldx + c] : 0.f, hi, lo);
    __nv_bfloat16* o = out + r * 3 * Cp;
    o[c] = hi, o[Cp + c] = lo, o[2 * Cp + c] = hi;
  }
}
__global__ void pack_conv_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin,
                                       int RS) {
  GRID_STRIDE(i, (long long)Cout * Cin * RS) {
    const int rs = (int)(i % RS);
    const int ci = (int)((i / RS) % Cin);
    const int co = (int)(i / ((long long)RS * Cin));
    __nv_bfloat16 hi, lo;
    split2(w[i], hi, lo);
    __nv_bfloat16* o = out + ((long long)co * RS + rs) * 3 * Cin;
    o[ci] = hi, o[Cin + ci] = hi, o[2 * Cin + ci] = lo;
  }
}
__global__ void pack_linear_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int N, int K,
                                         int Kp) {
  GRID_STRIDE(i, (long long)N * Kp) {
    const long long n = i / Kp;
    const int k = (int)(i % Kp);
    __nv_bfloat16 hi, lo;
    split2(k < K ? w[n * K + k] : 0.f, hi, lo);
    __nv_bfloat16* o = out + n * 3 * Kp;
    o[k] = hi, o[Kp + k] = hi, o[2 * Kp + k] = lo;
  }
}
__global__ void pack_stem_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  GRID_STRIDE(i, 64 * 5 * 64) {
    const int co = (int)(i / 320), k = (int)(i % 320);
    const int kt = k / 64, slot = k % 64, kh = slot / 8, kw = slot % 8;
    float v = 0.f;
    if (kh < 7 && kw < 7) v = w[((co * 5 + kt) * 7 + kh) * 7 + kw];
    __nv_bfloat16 hi, lo;
    split2(v, hi, lo);
    __nv_bfloat16* o = out + (long long)co * 5 * 192 + kt * 192;
    o[slot] = hi, o[64 + slot] = hi, o[128 + slot] = lo;
  }
}
__global__ void stem_patch_f32_kernel(const float* __restrict__ x, float* __restrict__ P, int B, int T, int H, int W,
                                      int OH, int OW) {
  GRID_STRIDE(i, (long long)B * T * OH * OW * 64) {
    const int slot = (int)(i & 63), kh = slot >> 3, kw = slot & 7;
    long long pix = i >> 6;
    const int ow = (int)(pix % OW);
    long long t1 = pix / OW;
    const int oh = (int)(t1 % OH);
    const long long bt = t1 / OH;
    const int ih = 2 * oh + kh - 3, iw = 2 * ow + kw - 3;
    float v = 0.f;
    if (kh < 7 && kw < 7 && ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(bt * H + ih) * (long long)W + iw];
    P[i] = v;
  }
}
__global__ void bn_apply_f32_kernel(const float* __restrict__ x, const float* __restrict__ coef,
                                    const float* __restrict__ res, const float* __restrict__ rcoef, int relu,
                                    float* __restrict__ out, long long rows, int C) {
  GRID_STRIDE(i, rows * C) {
    const int c = (int)(i % C);
    float v = x[i] * coef[2 * C + c] + coef[3 * C + c];
    if (res) v += rcoef ? res[i] * rcoef[2 * C + c] + rcoef[3 * C + c] : res[i];
    out[i] = relu ? fmaxf(v, 0.f) : v;
  }
}
__global__ void stem_pool_f32_kernel(const float* __restrict__ y0, const float* __restrict__ coef,
                                     float* __restrict__ out, int N, int IH, int IW, int OH, int OW) {
  GRID_STRIDE(i, (long long)N * OH * OW * 64) {
    const int c = (int)(i & 63);
    long long pix = i >> 6;
    const int ow = (int)(pix % OW);
    long long t1 = pix / OW;
    const int oh = (int)(t1 % OH);
    const long long n = t1 / OH;
    float best = -INFINITY;
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = 2 * oh + kh - 1;
      if (ih < 0 || ih >= IH) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = 2 * ow + kw - 1;
        if (iw < 0 || iw >= IW) continue;
        best = fmaxf(best, gelu_f(y0[((n * IH + ih) * IW + iw) * 64 + c] * coef[128 + c] + coef[192 + c]));
      }
    }
    out[i] = best;
  }
}
__global__ void meanpool_cls_f32_kernel(const float* __restrict__ a, const float* __restrict__ cls,
                                        float* __restrict__ xs, int B, int T, int HW, int C, int ld) {
  GRID_STRIDE(i, (long long)B * (T + 1) * C) {
    const int c = (int)(i % C);
    const long long row = i / C;
    const int tt = (int)(row % (T + 1));
    const long long b = row / (T + 1);
    if (tt == 0) {
      xs[row * ld + c] = cls[c];
      continue;
    }
    const float* src = a + ((b * T + (tt - 1)) * HW) * (long long)C + c;
    float acc = 0.f;
    for (int p = 0; p < HW; ++p) acc += src[(long long)p * C];
    xs[row * ld + c] = acc / (float)HW;
  }
}
__global__ void rmsnorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ y,
                                   int M, int D, int ld, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < M; row += nwarps) {
    const float* xr = x + (long long)row * ld;
    float ss = 0.f;
    for (int j = lane; j < D; j += 32) ss += xr[j] * xr[j];
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss) * rsqrtf((float)D), eps);
    for (int j = lane; j < D; j += 32) y[(long long)row * ld + j] = xr[j] * inv * g[j];
  }
}
// one CTA per (batch, head); plain fp32, n <= 64, head dim 64
__global__ void __launch_bounds__(128)
attention_f32_kernel(const float* __restrict__ qkv, const float* __restrict__ rot, float* __restrict__ o, int n,
                     int heads, int rotary_v) {
  extern __shared__ float sm[];
  float* sq = sm;
  float* sk = sq + n * 65;
  float* sv = sk + n * 65;
  float* sp = sv + n * 65;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int inner = heads * 64, ld = 3 * inner;
  for (int which = 0; which < 3; ++which) {
    float* dst = which == 0 ? sq : (which == 1 ? sk : sv);
    const float* src = qkv + (long long)b * n * ld + which * inner + h * 64;
    for (int i = threadIdx.x; i < n * 64; i += blockDim.x) dst[(i >> 6) * 65 + (i & 63)] = src[(long long)(i >> 6) * ld + (i & 63)];
    __syncthreads();
    if (rot && (which < 2 || rotary_v)) {
      for (int i = threadIdx.x; i < n * 16; i += blockDim.x) {
        const int pos = i >> 4, f = i & 15;
        const float c = rot[pos * 32 + f], s_ = rot[pos * 32 + 16 + f];
        const float a = dst[pos * 65 + f], bb = dst[pos * 65 + 16 + f];
        dst[pos * 65 + f] = a * c - bb * s_;
        dst[pos * 65 + 16 + f] = bb * c + a * s_;
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
    const int r = i / n, c = i - r * n;
    float acc = 0.f;
    for (int d = 0; d < 64; ++d) acc += sq[r * 65 + d] * sk[c * 65 + d];
    sp[r * 65 + c] = acc * 0.125f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < n; r += 4) {
    const float a = lane < n ? sp[r * 65 + lane] : -INFINITY, bq = lane + 32 < n ? sp[r * 65 + lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(a, bq));
    const float ea = lane < n ? expf(a - m) : 0.f, eb = lane + 32 < n ? expf(bq - m) : 0.f;
    const float inv = 1.0f / warp_sum(ea + eb);
    if (lane < n) sp[r * 65 + lane] = ea * inv;
    if (lane + 32 < n) sp[r * 65 + lane + 32] = eb * inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n * 64; i += blockDim.x) {
    const int r = i >> 6, d = i & 63;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc += sp[r * 65 + j] * sv[j * 65 + d];
    o[((long long)b * n + r) * inner + h * 64 + d] = acc;
  }
}
// torch.nn.LayerNorm over the last dimension (biased variance), one warp per row
__global__ void layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                                     float* __restrict__ y, int M, int D, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < M; row += nwarps) {
    const float* xr = x + (long long)row * D;
    float sum = 0.f;
    for (int j = lane; j < D; j += 32) sum += xr[j];
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.f;
    for (int j = lane; j < D; j += 32) sq += (xr[j] - mean) * (xr[j] - mean);
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
    for (int j = lane; j < D; j += 32) y[(long long)row * D + j] = (xr[j] - mean) * rstd * g[j] + b[j];
  }
}
__global__ void gelu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  GRID_STRIDE(i, n) y[i] = gelu_f(x[i]);
}
__global__ void geglu_f32_kernel(const float* __restrict__ h, float* __restrict__ u, long long M, int F) {
  GRID_STRIDE(i, M * F) {
    const long long r = i / F;
    const int c = (int)(i % F);
    u[i] = h[r * 2 * F + c] * gelu_f(h[r * 2 * F + F + c]);
  }
}
__global__ void split_last_f32_kernel(const float* __restrict__ last, int ld, float* __restrict__ cls,
                                      float* __restrict__ frames, int B, int T, int D) {
  GRID_STRIDE(i, (long long)B * (T + 1) * D) {
    const int d = (int)(i % D);
    const long long row = i / D;
    const int tt = (int)(row % (T + 1));
    const long long b = row / (T + 1);
    const float v = last[row * ld + d];
    if (tt == 0)
      cls[b * D + d] = v;
    else
      frames[(b * T + tt - 1) * D + d] = v;
  }
}

inline unsigned nblk(long long total) {
  long long b = (total + 255) / 256;
  return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}
#define LAUNCHED()  \
  note_launch();    \
  SVSR_CHECK_CUDA(cudaGetLastError()); \
  return SVSR_OK

}  // namespace

int split3_f32(const float* x, int ldx, __nv_bfloat16* out, long long rows, int C, int Cp, cudaStream_t s) {
  SVSR_REQUIRE(ldx >= C && Cp >= C && Cp % 8 == 0, "split3_f32: pitch %d / padded width %d do not cover C=%d", ldx, Cp, C);
  split3_kernel<<<nblk(rows * Cp), 256, 0, s>>>(x, ldx, out, rows, C, Cp);
  LAUNCHED();
}
int pack_conv_weight_split(const float* w, __nv_bfloat16* out, int Cout, int Cin, int RS, cudaStream_t s) {
  pack_conv_split_kernel<<<nblk((long long)Cout * Cin * RS), 256, 0, s>>>(w, out, Cout, Cin, RS);
  LAUNCHED();
}
int pack_linear_weight_split(const float* w, __nv_bfloat16* out, int N, int K, int Kp, cudaStream_t s) {
  SVSR_REQUIRE(Kp >= K && Kp % 8 == 0, "pack_linear_weight_split: padded width %d does not cover K=%d", Kp, K);
  pack_linear_split_kernel<<<nblk((long long)N * Kp), 256, 0, s>>>(w, out, N, K, Kp);
  LAUNCHED();
}
int pack_stem_weight_split(const float* w, __nv_bfloat16* out, cudaStream_t s) {
  pack_stem_split_kernel<<<nblk(64 * 320), 256, 0, s>>>(w, out);
  LAUNCHED();
}
int stem_patch_f32(const float* videos, float* patches, int B, int T, int H, int W, cudaStream_t s) {
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  stem_patch_f32_kernel<<<nblk((long long)B * T * OH * OW * 64), 256, 0, s>>>(videos, patches, B, T, H, W, OH, OW);
  LAUNCHED();
}
int bn_apply_f32(const float* x, const float* coef, const float* res, const float* rcoef, int relu, float* out,
                 long long rows, int C, cudaStream_t s) {
  bn_apply_f32_kernel<<<nblk(rows * C), 256, 0, s>>>(x, coef, res, rcoef, relu, out, rows, C);
  LAUNCHED();
}
int stem_bn_gelu_pool_f32(const float* y0, const float* coef, float* out, int N, int IH, int IW, cudaStream_t s) {
  const int OH = (IH + 2 - 3) / 2 + 1, OW = (IW + 2 - 3) / 2 + 1;
  stem_pool_f32_kernel<<<nblk((long long)N * OH * OW * 64), 256, 0, s>>>(y0, coef, out, N, IH, IW, OH, OW);
  LAUNCHED();
}
int meanpool_cls_f32(const float* a, const float* cls, float* x_stream, int B, int T, int HW, int C, int ld,
                     cudaStream_t s) {
  meanpool_cls_f32_kernel<<<nblk((long long)B * (T + 1) * C), 256, 0, s>>>(a, cls, x_stream, B, T, HW, C, ld);
  LAUNCHED();
}
int rmsnorm_fwd_f32(const float* x, const float* g, float* y, int M, int D, int ld, float eps, cudaStream_t s) {
  rmsnorm_f32_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, g, y, M, D, ld, eps);
  LAUNCHED();
}
int attention_fwd_f32(const float* qkv, const float* rot, float* o, int B, int n, int heads, int rotary_v,
                      cudaStream_t s) {
  SVSR_REQUIRE(n >= 1 && n <= 64, "attention_f32: n=%d out of range", n);
  const int smem = 4 * n * 65 * sizeof(float);
  static bool done = false;
  if (!done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(attention_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         4 * 64 * 65 * (int)sizeof(float)));
    done = true;
  }
  attention_f32_kernel<<<B * heads, 128, smem, s>>>(qkv, rot, o, n, heads, rotary_v);
  LAUNCHED();
}
int layernorm_fwd_f32(const float* x, const float* g, const float* b, float* y, int M, int D, float eps, cudaStream_t s) {
  layernorm_f32_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, g, b, y, M, D, eps);
  LAUNCHED();
}
int gelu_fwd_f32(const float* x, float* y, long long n, cudaStream_t s) {
  gelu_f32_kernel<<<nblk(n), 256, 0, s>>>(x, y, n);
  LAUNCHED();
}
int geglu_fwd_f32(const float* h, float* u, int M, int F, cudaStream_t s) {
  geglu_f32_kernel<<<nblk((long long)M * F), 256, 0, s>>>(h, u, M, F);
  LAUNCHED();
}
int split_last_f32(const float* last, int ld, float* cls, float* frames, int B, int T, int D, cudaStream_t s) {
  split_last_f32_kernel<<<nblk((long long)B * (T + 1) * D), 256, 0, s>>>(last, ld, cls, frames, B, T, D);
  LAUNCHED();
}

}  // namespace svsr
