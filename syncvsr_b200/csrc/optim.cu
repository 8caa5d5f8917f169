// Fused optimizer step over the flat fp32 arenas: global-norm gradient clipping (Trainer gradient_clip_val,
// LRW/video/src/train.py:32) + AdamW with decoupled weight decay on the ndim >= 2 parameters only
// (lightning.py:216-221). The arena is laid out [decayed | non-decayed] so the decay flag is a single boundary.
#include "common.cuh"

namespace svsr {
namespace {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* acc) {
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) s += g[i] * g[i];
  s = warp_sum(s);
  __shared__ float sp[8];
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += sp[i];
    atomicAdd(acc, t);
  }
}

// scratch[0] = sum of squares (fp64); scratch[1] (as float at byte 8) = clip coefficient; scratch float[3] = norm
__global__ void clip_coef_kernel(double* scratch, float max_norm, float grad_div) {
  const double norm = sqrt(scratch[0]) / (double)grad_div;
  float* f = reinterpret_cast<float*>(scratch + 1);
  float coef = 1.0f / grad_div;
  if (max_norm > 0.f) {
    const double c = (double)max_norm / (norm + 1e-6);
    if (c < 1.0) coef *= (float)c;
  }
  f[0] = coef;
  f[1] = (float)norm;
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, const double* __restrict__ scratch, float lr, float b1, float b2, float eps, float wd,
             float bc1, float bc2_sqrt) {
  const float coef = reinterpret_cast<const float*>(scratch + 1)[0];
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x;
    const float* gg = &gv.x;
    float* mm = &mv.x;
    float* vq = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gg[k] * coef;
      mm[k] = b1 * mm[k] + (1.f - b1) * gr;
      vq[k] = b2 * vq[k] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(vq[k]) / bc2_sqrt + eps;
      pp[k] = pp[k] * (1.f - lr * wd) - (lr / bc1) * (mm[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

// Segmented variant: the arena is cut into ranges [begin[i], begin[i+1]) that carry their own Adam step count.
// A range whose parameters received no gradient this step (a sublayer dropped by x-transformers' layer_dropout:
// `p.grad is None` in the reference, so torch.optim.AdamW leaves the parameter, its moments and its step count
// untouched -- lightning.py:216-221) has bc1 = 0 and is skipped entirely: no decay, no moments.
struct AdamSegs {
  long long begin[SVSR_ADAMW_MAX_SEGMENTS + 1];  // element offsets, multiples of 4, ascending; begin[n] = end
  float bc1[SVSR_ADAMW_MAX_SEGMENTS];            // 1 - beta1^step of the range; 0 = skip the range this step
  float bc2s[SVSR_ADAMW_MAX_SEGMENTS];           // sqrt(1 - beta2^step)
  int n;
};

__global__ void __launch_bounds__(256)
adamw_seg_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 long long first, long long n, long long n_decay, const double* __restrict__ scratch, float lr, float b1,
                 float b2, float eps, float wd, const __grid_constant__ AdamSegs segs) {
  const float coef = reinterpret_cast<const float*>(scratch + 1)[0];
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = first + (i << 2);
    int lo = 0, hi = segs.n - 1;  // last range whose begin <= e
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (segs.begin[mid] <= e) lo = mid; else hi = mid - 1;
    }
    const float bc1 = segs.bc1[lo], bc2_sqrt = segs.bc2s[lo];
    if (bc1 == 0.f) continue;
    const float decay = e < n_decay ? 1.f - lr * wd : 1.f;
    float4 pv = reinterpret_cast<float4*>(p + first)[i];
    const float4 gv = reinterpret_cast<const float4*>(g + first)[i];
    float4 mv = reinterpret_cast<float4*>(m + first)[i], vv = reinterpret_cast<float4*>(v + first)[i];
    float* pp = &pv.x;
    const float* gg = &gv.x;
    float* mm = &mv.x;
    float* vq = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gg[k] * coef;
      mm[k] = b1 * mm[k] + (1.f - b1) * gr;
      vq[k] = b2 * vq[k] + (1.f - b2) * gr * gr;
      const float denom = sqrtf(vq[k]) / bc2_sqrt + eps;
      pp[k] = pp[k] * decay - (lr / bc1) * (mm[k] / denom);
    }
    reinterpret_cast<float4*>(p + first)[i] = pv;
    reinterpret_cast<float4*>(m + first)[i] = mv;
    reinterpret_cast<float4*>(v + first)[i] = vv;
  }
}

}  // namespace
}  // namespace svsr

using namespace svsr;

extern "C" {

// params/grads/exp_avg/exp_avg_sq: fp32 [n_total], n_decay leading elements get weight decay (both multiples of 4).
// grad_div divides the gradients first (world size after a SUM all-reduce). scratch: >= 32 bytes of device memory.
// After the call scratch holds: fp64 sum of squares, fp32 clip coefficient, fp32 gradient norm.
int svsr_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_decay,
                    int64_t n_total, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    float max_norm, float grad_div, void* scratch, void* stream) {
  SVSR_REQUIRE(params && grads && exp_avg && exp_avg_sq && scratch, "adamw: null pointer");
  SVSR_REQUIRE(n_decay % 4 == 0 && n_total % 4 == 0 && n_decay <= n_total && step >= 1, "adamw: bad sizes/step");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* sc = static_cast<double*>(scratch);
  SVSR_CHECK_CUDA(cudaMemsetAsync(sc, 0, 32, s));
  sumsq_kernel<<<148 * 4, 256, 0, s>>>(grads, n_total, sc);
  clip_coef_kernel<<<1, 1, 0, s>>>(sc, max_norm, grad_div);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  if (n_decay > 0)
    adamw_kernel<<<148 * 8, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, n_decay, sc, lr, beta1, beta2, eps,
                                         weight_decay, bc1, bc2s);
  if (n_total > n_decay)
    adamw_kernel<<<148 * 2, 256, 0, s>>>(params + n_decay, grads + n_decay, exp_avg + n_decay, exp_avg_sq + n_decay,
                                         n_total - n_decay, sc, lr, beta1, beta2, eps, 0.f, bc1, bc2s);
  for (int i = 0; i < 3 + (n_total > n_decay ? 1 : 0); ++i) note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

// Per-range step counts (see AdamSegs). seg_begin: host array of nseg+1 ascending element offsets (multiples of 4,
// seg_begin[0] = 0, seg_begin[nseg] = n_total); seg_step: host array of nseg Adam step counts, 0 = the range got no
// gradient this step and is left untouched. One launch over the whole arena.
int svsr_adamw_step_segmented(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_decay,
                              int64_t n_total, float lr, float beta1, float beta2, float eps, float weight_decay,
                              const int64_t* seg_begin, const int32_t* seg_step, int nseg, float max_norm,
                              float grad_div, void* scratch, void* stream) {
  SVSR_REQUIRE(params && grads && exp_avg && exp_avg_sq && scratch && seg_begin && seg_step, "adamw: null pointer");
  SVSR_REQUIRE(nseg >= 1 && nseg <= SVSR_ADAMW_MAX_SEGMENTS, "adamw: %d ranges (max %d)", nseg, SVSR_ADAMW_MAX_SEGMENTS);
  SVSR_REQUIRE(n_decay % 4 == 0 && n_total % 4 == 0 && n_decay <= n_total, "adamw: bad sizes");
  SVSR_REQUIRE(seg_begin[0] == 0 && seg_begin[nseg] == n_total, "adamw: ranges must cover [0, n_total)");
  AdamSegs segs;
  segs.n = nseg;
  for (int i = 0; i < nseg; ++i) {
    SVSR_REQUIRE(seg_begin[i] % 4 == 0 && seg_begin[i] < seg_begin[i + 1] && seg_step[i] >= 0,
                 "adamw: range %d is not ascending / 4-aligned / has a negative step", i);
    segs.begin[i] = seg_begin[i];
    segs.bc1[i] = seg_step[i] > 0 ? 1.f - powf(beta1, (float)seg_step[i]) : 0.f;
    segs.bc2s[i] = seg_step[i] > 0 ? sqrtf(1.f - powf(beta2, (float)seg_step[i])) : 1.f;
  }
  segs.begin[nseg] = n_total;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* sc = static_cast<double*>(scratch);
  SVSR_CHECK_CUDA(cudaMemsetAsync(sc, 0, 32, s));
  sumsq_kernel<<<148 * 4, 256, 0, s>>>(grads, n_total, sc);
  clip_coef_kernel<<<1, 1, 0, s>>>(sc, max_norm, grad_div);
  adamw_seg_kernel<<<148 * 8, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, 0, n_total, n_decay, sc, lr, beta1, beta2,
                                           eps, weight_decay, segs);
  for (int i = 0; i < 3; ++i) note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // extern "C"
