#include "heads.cuh"

namespace svsr {
namespace {

// one warp per (row, channel-group) = one softmax over V classes
__global__ void __launch_bounds__(256)
audio_ce_kernel(const float* __restrict__ logits, int ld, const long long* __restrict__ tokens, long long tok_stride_b,
                int T, int A, int G, int V, long long nrows, __nv_bfloat16* __restrict__ dlogits, double* acc,
                int* bad_token, float dscale) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int AG = A * G;
  float local = 0.f;
  for (long long r = warp; r < nrows; r += nwarps) {
    const int c = (int)(r % AG);
    const long long bt = r / AG;
    const int t = (int)(bt % T);
    const long long b = bt / T;
    const int a = c / G, g = c - a * G;
    const long long tgt = tokens[b * tok_stride_b + (long long)(t * A + a) * G + g];
    if (tgt < 0 || tgt >= V) {  // F.cross_entropy would raise: flag it (the step's metrics become NaN) and leave a
      if (lane == 0) *bad_token = 1;  // zero gradient row instead of last step's stale one
      if (dlogits) {
        __nv_bfloat16* drow = dlogits + bt * ld + (long long)c * V;
        for (int j = lane; j < V; j += 32) drow[j] = __float2bfloat16(0.f);
      }
      continue;
    }
    const float* row = logits + bt * ld + (long long)c * V;
    float m = -INFINITY;
    for (int j = lane; j < V; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float se = 0.f;
    for (int j = lane; j < V; j += 32) se += expf(row[j] - m);
    se = warp_sum(se);
    const float lse = m + logf(se);
    if (lane == 0) local += lse - row[tgt];
    if (dlogits) {
      __nv_bfloat16* drow = dlogits + bt * ld + (long long)c * V;
      for (int j = lane; j < V; j += 32) {
        const float p = expf(row[j] - lse);
        drow[j] = __float2bfloat16((p - (j == (int)tgt ? 1.f : 0.f)) * dscale);
      }
    }
  }
  // lanes 0 of each warp hold partial sums
  __shared__ float sp[8];
  if (lane == 0) sp[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sp[i];
    atomicAdd(acc, s);
  }
}

__global__ void __launch_bounds__(256)
category_ce_kernel(const float* __restrict__ logits, int ld, const long long* __restrict__ labels,
                   const float* __restrict__ soft, int B, int C, float eps, __nv_bfloat16* __restrict__ dlogits,
                   int ldd, double* acc, float dscale, int* bad_label) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < B; r += nwarps) {
    const float* row = logits + (long long)r * ld;
    if (!soft && (labels[r] < 0 || labels[r] >= C)) {  // out-of-range class index: flag, zero gradient row, no reads
      if (lane == 0 && bad_label) *bad_label = 1;
      if (dlogits)
        for (int j = lane; j < ldd; j += 32) dlogits[(long long)r * ldd + j] = __float2bfloat16(0.f);
      continue;
    }
    float m = -INFINITY;
    for (int j = lane; j < C; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float se = 0.f;
    for (int j = lane; j < C; j += 32) se += expf(row[j] - m);
    se = warp_sum(se);
    const float lse = m + logf(se);
    // hard target index (argmax of the soft labels when soft, first maximum like torch.argmax)
    int hard;
    if (soft) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int j = lane; j < C; j += 32) {
        const float v = soft[(long long)r * C + j];
        if (v > bv) bv = v, bi = j;
      }
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
      }
      hard = bi;
    } else {
      hard = (int)labels[r];
    }
    // loss = -sum_j q_j log p_j with q = (1-eps)*target + eps/C
    float lsum = 0.f;
    int greater = 0;
    const float thard = row[hard];
    for (int j = lane; j < C; j += 32) {
      const float lp = row[j] - lse;
      const float tq = soft ? soft[(long long)r * C + j] : (j == hard ? 1.f : 0.f);
      const float qj = (1.f - eps) * tq + eps / (float)C;
      lsum -= qj * lp;
      greater += (row[j] > thard) ? 1 : 0;
      if (dlogits) dlogits[(long long)r * ldd + j] = __float2bfloat16((expf(lp) - qj) * dscale);
    }
    if (dlogits)
      for (int j = C + lane; j < ldd; j += 32) dlogits[(long long)r * ldd + j] = __float2bfloat16(0.f);
    lsum = warp_sum(lsum);
    greater = __reduce_add_sync(0xffffffffu, greater);
    if (lane == 0) {
      atomicAdd(acc + 1, (double)lsum);
      if (greater == 0) atomicAdd(acc + 2, 1.0);
      if (greater < 5) atomicAdd(acc + 3, 1.0);
    }
  }
}

__global__ void finalize_metrics_kernel(const double* acc, float* out, float lambda_audio, int B, long long audio_rows,
                                        const int* bad) {
  double la = acc[0] / (double)audio_rows, lc = acc[1] / (double)B;
  // an audio token / class label outside its vocabulary (F.cross_entropy raises a device assert in the reference):
  // the native step cannot raise from a kernel, so its losses read NaN -- loud in any log and any `isfinite` guard
  if (bad && bad[0]) la = lc = (double)NAN;
  out[0] = (float)(lc + la * (double)lambda_audio);
  out[1] = (float)lc;
  out[2] = (float)la;
  out[3] = (float)(acc[2] / (double)B);
  out[4] = (float)(acc[3] / (double)B);
}

// Second half of the fused audio head's forward (igemm.cuh, IgemmCe mode 1): one thread per (frame, c) softmax merges the
// (max, sum exp) partials its V/64 column slots left behind into the log-sum-exp, adds lse - logit[target] to the loss.
__global__ void __launch_bounds__(256)
ce_finalize_kernel(const float2* __restrict__ part, const float* __restrict__ xt, const long long* __restrict__ tokens,
                   long long tok_stride_b, int T, int A, int G, int V, long long nrows, float* __restrict__ lse,
                   double* acc) {
  const int AG = A * G, spc = V >> 6, nslots = AG * spc;
  float local = 0.f;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nrows; r += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(r % AG);
    const long long bt = r / AG;
    const float2* pp = part + bt * nslots + (long long)c * spc;
    float m = -INFINITY;
    for (int i = 0; i < spc; ++i) m = fmaxf(m, pp[i].x);
    float se = 0.f;
    for (int i = 0; i < spc; ++i) se += pp[i].y * exp2f((pp[i].x - m) * 1.4426950408889634f);
    const float l = m + logf(se);
    lse[r] = l;
    const int t = (int)(bt % T);
    const long long b = bt / T;
    const int a = c / G, g = c - a * G;
    const long long tgt = tokens[b * tok_stride_b + (long long)(t * A + a) * G + g];
    if (tgt >= 0 && tgt < V) local += l - xt[r];
  }
  local = warp_sum(local);
  __shared__ float sp[8];
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < 8; ++i) s += sp[i];
    atomicAdd(acc, s);
  }
}

__global__ void scale_bf16_kernel(__nv_bfloat16* x, long long n, const float* __restrict__ scale) {
  const float sc = *scale;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __float2bfloat16(__bfloat162float(x[i]) * sc);
}

}  // namespace

int scale_bf16_by_device_scalar(__nv_bfloat16* x, long long n, const float* scale, cudaStream_t s) {
  long long blocks = (n + 1023) / 1024;
  if (blocks > 148 * 8) blocks = 148 * 8;
  scale_bf16_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, n, scale);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

int audio_ce(const float* logits, int ld, const long long* tokens, long long tok_stride_b, int B, int T, int A, int G,
             int V, __nv_bfloat16* dlogits, double* acc, int* bad_token, float dscale, cudaStream_t s) {
  const long long nrows = (long long)B * T * A * G;
  long long blocks = (nrows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  audio_ce_kernel<<<(unsigned)blocks, 256, 0, s>>>(logits, ld, tokens, tok_stride_b, T, A, G, V, nrows, dlogits, acc,
                                                   bad_token, dscale);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}
int ce_finalize(const float2* part, const float* xt, const long long* tokens, long long tok_stride_b, int B, int T, int A,
                int G, int V, float* lse, double* acc, cudaStream_t s) {
  const long long nrows = (long long)B * T * A * G;
  long long blocks = (nrows + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ce_finalize_kernel<<<(unsigned)blocks, 256, 0, s>>>(part, xt, tokens, tok_stride_b, T, A, G, V, nrows, lse, acc);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}
int category_ce(const float* logits, int ld, const long long* labels, const float* soft_labels, int B, int C, float eps,
                __nv_bfloat16* dlogits, int ldd, double* acc, float dscale, cudaStream_t s, int* bad_label) {
  SVSR_REQUIRE((labels != nullptr) != (soft_labels != nullptr), "category_ce: exactly one of labels/soft_labels");
  const int blocks = (B + 7) / 8;
  category_ce_kernel<<<blocks, 256, 0, s>>>(logits, ld, labels, soft_labels, B, C, eps, dlogits, ldd, acc, dscale,
                                            bad_label);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}
int finalize_metrics(const double* acc, float* out, float lambda_audio, int B, long long audio_rows, cudaStream_t s,
                     const int* bad) {
  finalize_metrics_kernel<<<1, 1, 0, s>>>(acc, out, lambda_audio, B, audio_rows, bad);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // namespace svsr
