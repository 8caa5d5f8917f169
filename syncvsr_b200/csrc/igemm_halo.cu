// 3x3 / stride-1 / 64 -> 64 channel convolution (resnet.layer1: fprop and input gradient) with ONE activation load per
// tile instead of one per filter tap.
//
// The generic implicit GEMM (igemm.cu) fetches the A tile once per tap: 9 x the activation bytes cross L2 -> SM, which is
// what bounds the 64-channel layers (ncu: tensor pipe 22 %, L2 -> SM ~ 1 GB per launch). Here a tile is RT image rows
// of ONE image laid out on the zero-padded "halo grid" of pitch P = W + 2: a single 4-D TMA box (64 ch, P, RT + 2, 1)
// starting at (w, h) = (-1, h0 - 1) lands the (RT + 2) x P halo pixels densely in shared memory (out-of-bounds = the
// convolution's zero padding). Accumulator row m is halo position (h0 + m / P, m % P - 1); for tap (dh, dw) its input is
// the shared-memory row m + (1 + dh) * P + dw, i.e. the SAME tile addressed through a UMMA descriptor whose start
// address is shifted by a whole number of 128-byte rows (valid for SWIZZLE_128B K-major operands with base_offset 0:
// probed in debug_probe.cu / profiles/r1_gpu_check_wgrad_dgrad_rowshift.txt). Rows at the two halo columns (and rows
// m >= RT * P) compute garbage and are never stored. The 9 x [64 x 64] weight tiles (72 KB) stay resident in shared
// memory for the lifetime of the persistent CTA. L2 -> SM traffic per tile: 21.5 KB instead of 147 KB + 72 KB.
#include "igemm.cuh"
#include "tmap.h"
#include <stdlib.h>

namespace svsr {

namespace {

constexpr int HALO_STAGES = 4;
constexpr int HALO_STAGE_BYTES = 24576;  // 1024 B lead (row -1) + up to 184 halo rows of 128 B
constexpr int HALO_B_BYTES = 9 * 8192;

struct HaloParams {
  int N, H, W, P, RT, tiles_per_img, total_tiles;
  int tap_row[9];    // (1 + dh) * P + dw of tap t (row shift inside the halo tile, >= -1)
  int tap_kbase[9];  // column of the weight matrix where tap t's 64-wide block starts
  int a_bytes;       // bytes of one halo box
  const void* resid;  // bf16 [N, H, W, 64] or null
  double* bn_stats;   // fp64 [2][64] (+=) or null
  const void* relu_mask;  // bf16 [N, H, W, 64] or null: zero the result where mask <= 0
  IgemmBnBwd bnb;         // fused BatchNorm-backward statistics of ONE BatchNorm (igemm.cuh), n = 0: off
};

// Transposed warp reduction (see igemm.cu): afterwards lane l holds in v[0] the sum over the 32 lanes of value l.
__device__ __forceinline__ void warp_transpose_reduce32h(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? v[i] : v[i + off];
      const float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

struct HaloSmem {
  static constexpr int B_OFFSET = 0;
  static constexpr int A_OFFSET = HALO_B_BYTES;
  static constexpr int STAGING_OFFSET = A_OFFSET + HALO_STAGES * HALO_STAGE_BYTES;  // 2 x 16 KB
  static constexpr int BAR_OFFSET = STAGING_OFFSET + 2 * 16384;
  static constexpr int STATS_OFFSET = BAR_OFFSET + 256;  // fp32 [8 warps][2][64]
  static constexpr int TOTAL = STATS_OFFSET + 8 * 128 * 4 + 1024;
  static_assert(TOTAL <= 232448, "exceeds 227 KB of shared memory");
};

template <bool BNB>  // BNB: fused BatchNorm-backward statistics (+ activation masks) in the epilogue
__global__ void __launch_bounds__(320, 1)
conv3x3_c64_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmC, const HaloParams p) {
  using L = HaloSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem + L::B_OFFSET;
  uint8_t* sA = smem + L::A_OFFSET;
  uint8_t* s_stage = smem + L::STAGING_OFFSET;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + HALO_STAGES;
  uint64_t* tmem_full_bar = empty_bar + HALO_STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]
  uint64_t* b_bar = tmem_empty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_bar + 1);
  float* s_stats = reinterpret_cast<float*>(smem + L::STATS_OFFSET);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 128;  // two 64-column accumulators

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < HALO_STAGES; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&tmem_full_bar[a], 1), mbar_init(&tmem_empty_bar[a], 4);
    mbar_init(b_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) s_stats[i] = 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // weights: resident for the whole kernel
      mbar_expect_tx(b_bar, HALO_B_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d(sB + t * 8192, &tmB, b_bar, p.tap_kbase[t], 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = tile / p.tiles_per_img, h0 = (tile - n * p.tiles_per_img) * p.RT;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)p.a_bytes);
        tma_load_4d(sA + stage * HALO_STAGE_BYTES + 1024, &tmA, &full_bar[stage], 0, -1, h0 - 1, n);
        if (++stage == HALO_STAGES) stage = 0, phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      mbar_wait(b_bar, 0);
      tcgen05_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int j = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++j) {
        const int acc = j & 1;
        mbar_wait(&tmem_empty_bar[acc], ((uint32_t)(j >> 1) & 1) ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
        const uint32_t a0 = smem_u32(sA + stage * HALO_STAGE_BYTES + 1024);
        const uint32_t b0 = smem_u32(sB);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t a_desc = umma_smem_desc_sw128(a0 + (uint32_t)(p.tap_row[t] * 128), 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(b0 + (uint32_t)(t * 8192), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (t | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full_bar[acc]);
        if (++stage == HALO_STAGES) stage = 0, phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> registers (+ residual) -> dense staging tile -> TMA store ----------------
    // Two warpgroups alternate tiles (group g owns accumulator g and staging buffer g), so one tile's TMEM drain,
    // residual add, BN partial sums and store overlap the next tile's; the residual rows are fetched before the
    // accumulator is waited for.
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;       // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;  // accumulator row = halo position of the tile
    const int hr = r / p.P, c = r - hr * p.P;
    const bool col_ok = hr < p.RT && c >= 1 && c <= p.W;
    const int srow = hr * p.W + (c - 1);  // dense row of the staged [RT][W][64] tile
    const bool leader = threadIdx.x == 64 + 128 * g;
    const int bar_id = 1 + g;
    uint8_t* stg = s_stage + g * 16384;
    float st_sum[2] = {0.f, 0.f}, st_sq[2] = {0.f, 0.f};
    float bg[2] = {0.f, 0.f}, bx[2] = {0.f, 0.f};  // fused BatchNorm-backward sums of column ch*32 + lane
    int j = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++j) {
      if ((j & 1) != g) continue;
      const int n = tile / p.tiles_per_img, h0 = (tile - n * p.tiles_per_img) * p.RT;
      const bool valid = col_ok && (h0 + hr) < p.H;
      uint4 rv[8];
      if (p.resid && valid && !BNB) {
        const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) +
                                                         (((long long)n * p.H + h0 + hr) * p.W + (c - 1)) * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = __ldg(rp + i);
      }
      const long long pix_off = (((long long)n * p.H + h0 + hr) * p.W + (c - 1)) * 64;
      mbar_wait(&tmem_full_bar[g], (uint32_t)((j >> 1) & 1));
      tcgen05_fence_after();
      if (leader) tma_store_wait_read<0>();  // this group's previous store has read the staging buffer
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64 + ch * 32), v);
        uint4 mv[4], cq[4];  // activation-mask reference and BatchNorm input of this row's 32 channels
        if (BNB && p.relu_mask && valid) {
          const uint4* mp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.relu_mask) + pix_off + ch * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) mv[i] = __ldg(mp + i);
        }
        if (BNB && valid) {
          const uint4* cp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bnb.c[0]) + pix_off + ch * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) cq[i] = __ldg(cp + i);
          if (p.resid) {  // (the fused-statistics path fetches the residual chunk by chunk: fewer live registers)
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.resid) + pix_off + ch * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) rv[i] = __ldg(rp + i);
          }
        }
        tmem_ld_wait();
        if constexpr (BNB) {  // all 32 lanes take part in the transposed reduction: rows outside the image contribute zeros
          float f[32], cv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = valid ? __uint_as_float(v[i]) : 0.f;
          if (p.resid && valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 t = rv[i];
              const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), cc = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
              f[8 * i] += a.x, f[8 * i + 1] += a.y, f[8 * i + 2] += b.x, f[8 * i + 3] += b.y;
              f[8 * i + 4] += cc.x, f[8 * i + 5] += cc.y, f[8 * i + 6] += d.x, f[8 * i + 7] += d.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (valid) {
              const float2 a = unpack_bf16x2(cq[i].x), b = unpack_bf16x2(cq[i].y), cc = unpack_bf16x2(cq[i].z), d = unpack_bf16x2(cq[i].w);
              cv[8 * i] = a.x, cv[8 * i + 1] = a.y, cv[8 * i + 2] = b.x, cv[8 * i + 3] = b.y;
              cv[8 * i + 4] = cc.x, cv[8 * i + 5] = cc.y, cv[8 * i + 6] = d.x, cv[8 * i + 7] = d.y;
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) cv[8 * i + k] = 0.f;
            }
          }
          if (p.relu_mask && valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 a = unpack_bf16x2(mv[i].x), b = unpack_bf16x2(mv[i].y), cc = unpack_bf16x2(mv[i].z), d = unpack_bf16x2(mv[i].w);
              f[8 * i] = a.x > 0.f ? f[8 * i] : 0.f, f[8 * i + 1] = a.y > 0.f ? f[8 * i + 1] : 0.f;
              f[8 * i + 2] = b.x > 0.f ? f[8 * i + 2] : 0.f, f[8 * i + 3] = b.y > 0.f ? f[8 * i + 3] : 0.f;
              f[8 * i + 4] = cc.x > 0.f ? f[8 * i + 4] : 0.f, f[8 * i + 5] = cc.y > 0.f ? f[8 * i + 5] : 0.f;
              f[8 * i + 6] = d.x > 0.f ? f[8 * i + 6] : 0.f, f[8 * i + 7] = d.y > 0.f ? f[8 * i + 7] : 0.f;
            }
          }
          const float* cf = p.bnb.coef[0];
          if (p.bnb.self_mask) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(cf + 128 + ch * 32) + i);
              const float4 sh = __ldg(reinterpret_cast<const float4*>(cf + 192 + ch * 32) + i);
              f[4 * i] = fmaf(cv[4 * i], sc.x, sh.x) > 0.f ? f[4 * i] : 0.f;
              f[4 * i + 1] = fmaf(cv[4 * i + 1], sc.y, sh.y) > 0.f ? f[4 * i + 1] : 0.f;
              f[4 * i + 2] = fmaf(cv[4 * i + 2], sc.z, sh.z) > 0.f ? f[4 * i + 2] : 0.f;
              f[4 * i + 3] = fmaf(cv[4 * i + 3], sc.w, sh.w) > 0.f ? f[4 * i + 3] : 0.f;
            }
          }
          if (valid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 t;
              t.x = pack_bf16x2(f[8 * i], f[8 * i + 1]), t.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
              t.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]), t.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
              const int chunk = (ch * 4 + i) ^ (srow & 7);
              *reinterpret_cast<uint4*>(stg + srow * 128 + chunk * 16) = t;
            }
          }
          // sums of the STORED (bf16-rounded) gradient: sum g and sum g * (c - mean), invstd applied at the flush
          float a[32], b[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 mu = __ldg(reinterpret_cast<const float4*>(cf + ch * 32) + i);
#pragma unroll
            for (int k = 0; k < 4; ++k) a[4 * i + k] = __bfloat162float(__float2bfloat16(f[4 * i + k]));
            b[4 * i] = a[4 * i] * (cv[4 * i] - mu.x), b[4 * i + 1] = a[4 * i + 1] * (cv[4 * i + 1] - mu.y);
            b[4 * i + 2] = a[4 * i + 2] * (cv[4 * i + 2] - mu.z), b[4 * i + 3] = a[4 * i + 3] * (cv[4 * i + 3] - mu.w);
          }
          warp_transpose_reduce32h(a, lane);
          warp_transpose_reduce32h(b, lane);
          bg[ch] += a[0], bx[ch] += b[0];
        }
        if (!BNB && valid) {
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (p.resid) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint4 t = rv[ch * 4 + i];
              const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), cc = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
              f[8 * i] += a.x, f[8 * i + 1] += a.y, f[8 * i + 2] += b.x, f[8 * i + 3] += b.y;
              f[8 * i + 4] += cc.x, f[8 * i + 5] += cc.y, f[8 * i + 6] += d.x, f[8 * i + 7] += d.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t;
            t.x = pack_bf16x2(f[8 * i], f[8 * i + 1]), t.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
            t.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]), t.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
            const int chunk = (ch * 4 + i) ^ (srow & 7);  // 16-byte chunk position after the 128B swizzle
            *reinterpret_cast<uint4*>(stg + srow * 128 + chunk * 16) = t;
          }
        }
      }
      // the accumulator has been read: hand it back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[g]);
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (leader) {
        tma_store_4d(&tmC, stg, 0, 0, h0, n);
        tma_store_commit();
      }
      if (p.bn_stats) {
        // statistics of the staged (bf16-rounded) rows: lane l of warp q owns channels (2l, 2l+1) over rows [32q, 32q+32)
        const int nrows = min(p.RT, p.H - h0) * p.W;
        const int rbeg = q * 32, rend = min(rbeg + 32, nrows);
        const uint8_t* colbase = stg + (lane & 3) * 4;
        const int cpos = lane >> 2;
        for (int rr = rbeg; rr < rend; ++rr) {
          const float2 v2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(colbase + rr * 128 + ((cpos ^ (rr & 7)) << 4)));
          st_sum[0] += v2.x, st_sum[1] += v2.y;
          st_sq[0] = fmaf(v2.x, v2.x, st_sq[0]), st_sq[1] = fmaf(v2.y, v2.y, st_sq[1]);
        }
      }
    }
    if (leader) tma_store_wait_all();
    if constexpr (BNB) {
      float* sl = s_stats + (warp - 2) * 128;
      sl[lane] = bg[0], sl[32 + lane] = bg[1], sl[64 + lane] = bx[0], sl[96 + lane] = bx[1];
      asm volatile("bar.sync 3, 256;" ::: "memory");
      const int t = threadIdx.x - 64;  // t < 64: sum g of channel t; 64 <= t < 128: sum g * xhat of channel t - 64
      if (t < 128) {
        double sacc = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sacc += (double)s_stats[w * 128 + t];
        if (t >= 64) sacc *= (double)__ldg(p.bnb.coef[0] + 64 + (t - 64));  // x invstd
        atomicAdd(p.bnb.stats[0] + t, sacc);
      }
    }
    if (p.bn_stats) {
      float* sl = s_stats + (warp - 2) * 128 + 2 * lane;
      sl[0] = st_sum[0], sl[1] = st_sum[1], sl[64] = st_sq[0], sl[65] = st_sq[1];
      asm volatile("bar.sync 3, 256;" ::: "memory");
      const int t = threadIdx.x - 64;  // t < 64 -> sum of channel t, 64 <= t < 128 -> sum of squares of channel t - 64
      if (t < 128) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += (double)s_stats[w * 128 + t];
        atomicAdd(p.bn_stats + t, s);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

// y[N,H,W,64] = sum_t x[n, h + dh_t, w + dw_t, :] . Wm[:, kbase_t .. kbase_t+64)^T (+ resid); Wm bf16 [64, 576]
int conv3x3_c64_halo(const void* x, const void* wm, int w_pitch, void* y, const void* resid, int N, int H, int W,
                     const int* tap_dh, const int* tap_dw, const int* tap_kbase, double* bn_stats, cudaStream_t stream,
                     const void* relu_mask = nullptr, const IgemmBnBwd* bnb = nullptr) {
  HaloParams p{};
  p.relu_mask = relu_mask;
  if (bnb) p.bnb = *bnb;
  SVSR_REQUIRE(p.bnb.n <= 1 && !(p.bnb.n && bn_stats), "halo conv: at most one fused BatchNorm backward, not with forward stats");
  SVSR_REQUIRE(!p.bnb.n || (p.bnb.c[0] && p.bnb.coef[0] && p.bnb.stats[0]), "halo conv: bnb buffers missing");
  p.N = N, p.H = H, p.W = W, p.P = W + 2;
  p.RT = 128 / p.P;
  SVSR_REQUIRE(p.P <= 27 && p.RT >= 1, "halo conv: image width %d unsupported", W);
  SVSR_REQUIRE(1024 + (128 + 2 * p.P + 2) * 128 <= HALO_STAGE_BYTES + 128, "halo conv: tile does not fit its stage");
  p.tiles_per_img = (H + p.RT - 1) / p.RT;
  p.total_tiles = N * p.tiles_per_img;
  for (int t = 0; t < 9; ++t) {
    SVSR_REQUIRE(tap_dh[t] >= -1 && tap_dh[t] <= 1 && tap_dw[t] >= -1 && tap_dw[t] <= 1, "halo conv: tap out of range");
    p.tap_row[t] = (1 + tap_dh[t]) * p.P + tap_dw[t];
    p.tap_kbase[t] = tap_kbase[t];
  }
  p.a_bytes = (p.RT + 2) * p.P * 128;
  p.resid = resid, p.bn_stats = bn_stats;
  CUtensorMap tmA, tmB, tmC;
  {
    uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    uint32_t box[4] = {64, (uint32_t)p.P, (uint32_t)(p.RT + 2), 1};
    int rc = make_tmap_bf16(&tmA, x, 4, dims, strides, box, nullptr, true);
    if (rc) return rc;
    uint32_t boxc[4] = {64, (uint32_t)W, (uint32_t)p.RT, 1};
    rc = make_tmap_bf16(&tmC, y, 4, dims, strides, boxc, nullptr, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)w_pitch, 64};
    uint64_t strides[1] = {(uint64_t)w_pitch * 2};
    uint32_t box[2] = {64, 64};
    int rc = make_tmap_bf16(&tmB, wm, 2, dims, strides, box, nullptr, true);
    if (rc) return rc;
  }
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_c64_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         HaloSmem::TOTAL));
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_c64_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         HaloSmem::TOTAL));
    attr_done = true;
  }
  const int grid = p.total_tiles < 148 ? p.total_tiles : 148;
  prof_begin(PROF_IGEMM, 2.0 * N * H * W * 64.0 * 576.0, stream);
  if (p.bnb.n || p.relu_mask)
    conv3x3_c64_halo_kernel<true><<<grid, 320, HaloSmem::TOTAL, stream>>>(tmA, tmB, tmC, p);
  else
    conv3x3_c64_halo_kernel<false><<<grid, 320, HaloSmem::TOTAL, stream>>>(tmA, tmB, tmC, p);
  note_launch();
  prof_end(stream);
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}


// True when `p` is exactly the shape this kernel covers; igemm_launch() routes such problems here.
bool igemm_halo_matches(const IgemmProblem& p) {
  if (!(p.ntaps == 9 && p.cin == 64 && p.b_rows == 64 && p.stride == 1 && p.a_C == 64 && p.a_coff == 0)) return false;
  if (p.out_fp32 || p.ldc != 64 || p.c_off != 0 || p.o_sh != 1 || p.o_sw != 1 || p.o_oh != 0 || p.o_ow != 0) return false;
  if (p.OH != p.a_H || p.OW != p.a_W || p.o_H != p.OH || p.o_W != p.OW || p.o_N != p.a_N) return false;
  if (p.bias || p.alpha != 1.0f || p.relu || p.drop_p > 0.f || (p.resid && p.resid_fp32) || p.ce.mode || p.bnb.n > 1) return false;
  if (p.relu_mask && !p.bnb.n) return false;  // the masked epilogue only exists with the fused statistics
  if (p.a_W + 2 > 27 || p.a_W < 2) return false;
  for (int t = 0; t < 9; ++t)
    if (p.tap_dh[t] < -1 || p.tap_dh[t] > 1 || p.tap_dw[t] < -1 || p.tap_dw[t] > 1) return false;
  const char* e = getenv("SVSR_HALO_CONV");
  return !(e && e[0] == '0');
}

int igemm_halo_launch(const IgemmProblem& p, cudaStream_t stream) {
  return conv3x3_c64_halo(p.a, p.b, p.b_cols, p.out, p.resid, p.a_N, p.a_H, p.a_W, p.tap_dh, p.tap_dw, p.tap_kbase,
                          p.bn_stats, stream, p.relu_mask, &p.bnb);
}

}  // namespace svsr
