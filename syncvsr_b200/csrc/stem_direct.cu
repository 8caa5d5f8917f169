// The 3-D conv stem without a patch tensor: Conv3d(1, 64, (5,7,7), (1,2,2), (2,3,3)) (lightning.py:49-50,
// conv3d_extractor.py:31-33) forward and weight gradient as 5-tap temporal implicit GEMMs (igemm_stem.cu /
// wgrad_stem.cu) whose A operand -- the 7x7/s2 window of every output pixel, 8x8 = 64 slots per frame -- is BUILT IN
// SHARED MEMORY by producer warps from a bf16 copy of the video instead of being read back from a materialised
// [clips, frames, pixels, 64] tensor (460 MB at the bench geometry, written once and read twice per step).
//
// Tile = 8 frames x 16 pixels of one clip; its operand is the 12 x 16 halo rows (frames t0-2 .. t0+9) of 128 B:
// row (slot f, pixel px) chunk kh (16 B) = the video row ih = 2 oh + kh - 3 at columns 2 ow - 3 .. 2 ow + 3 (+ one zero).
// The video is addressed as 32-bit pixel pairs P[j] = (x[2j], x[2j+1]): the seven values are the upper half of
// P[ow-2], P[ow-1], P[ow] and P[ow+1], stitched with three byte-permutes; 16 lanes = 16 consecutive pixels, so every
// load is one contiguous 64-byte segment. The chunk lands at the SWIZZLE_128B position a TMA box load of the patch
// tensor would have used (chunk kh ^ (row & 7)), so the MMA warp, the epilogue, the tap-as-descriptor-offset trick and
// the accumulation order are the halo kernels' own and the results are bit-identical to the patch-tensor path.
#include "igemm.cuh"
#include "wgrad.cuh"
#include "stem_direct.cuh"
#include "tmap.h"
#include <stdlib.h>

#define SD_RC(expr)           \
  do {                        \
    const int rc_ = (expr);   \
    if (rc_) return rc_;      \
  } while (0)

namespace svsr {

namespace {

constexpr int SD_STAGES = 4;
constexpr int SD_FT = 8, SD_PW = 16;
constexpr int SD_SLOTS = SD_FT + 4;
constexpr int SD_A_BYTES = SD_SLOTS * SD_PW * 128;  // 24 KB halo tile
constexpr int SD_B_BYTES = 5 * 8192;                // resident weights (forward)
constexpr int SD_DZ_BYTES = SD_FT * SD_PW * 128;    // 16 KB gradient tile (weight gradient)
constexpr int SD_FWD_PRODUCERS = 256;  // forward: eight producer warps
constexpr int SD_WG_PRODUCERS = 128;   // weight gradient: the four drain warps produce during the main loop

struct StemDirectParams {
  const uint32_t* vid;  // bf16 video as pixel pairs [N * T * H][W / 2]
  int N, T, H, W2, OW;
  int wtiles, ttiles, total_tiles;
  double* bn_stats;  // forward: fp64 [2][64] (+=) or null
  float* out;        // weight gradient: fp32 [320][ldo] (+=)
  int ldo;
};

// one producer thread's share (pixel ptid & 15, every (NPROD/16)-th (slot, kh) pair) of the halo tile at sA. All of the
// thread's loads are issued before the first store: the producers run at the latency of ONE round trip to L2 per tile.
template <int NPROD>
__device__ __forceinline__ void build_patch_tile(uint8_t* sA, const StemDirectParams& p, int n, int t0, int w0, int ptid) {
  constexpr int STEP = NPROD / 16, ITERS = (SD_SLOTS * 7 + STEP - 1) / STEP;
  const int px = ptid & 15;
  const int pix = w0 + px;
  const int oh = pix / p.OW, ow = pix - oh * p.OW;
  const int sw = px & 7;
  uint8_t* rowbase = sA + px * 128;
  const bool in_m2 = ow >= 2, in_m1 = ow >= 1, in_p1 = ow + 1 < p.W2;
  const uint32_t* clip = p.vid + (long long)n * p.T * p.H * p.W2 + ow;
  uint32_t a[ITERS], b[ITERS], d[ITERS], e[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int rest = (ptid >> 4) + it * STEP;
    const int slot = rest / 7, kh = rest - slot * 7;
    const int t = t0 - 2 + slot, ih = 2 * oh + kh - 3;
    const bool ok = rest < SD_SLOTS * 7 && t >= 0 && t < p.T && ih >= 0 && ih < p.H;
    const uint32_t* src = clip + (long long)(t * p.H + ih) * p.W2;  // dereferenced only when ok
    a[it] = (ok && in_m2) ? __ldg(src - 2) : 0u;
    b[it] = (ok && in_m1) ? __ldg(src - 1) : 0u;
    d[it] = ok ? __ldg(src) : 0u;
    e[it] = (ok && in_p1) ? __ldg(src + 1) : 0u;
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int rest = (ptid >> 4) + it * STEP;
    const int slot = rest / 7, kh = rest - slot * 7;
    uint4 c;
    c.x = __byte_perm(a[it], b[it], 0x5432);  // x[2ow-3], x[2ow-2]
    c.y = __byte_perm(b[it], d[it], 0x5432);  // x[2ow-1], x[2ow]
    c.z = __byte_perm(d[it], e[it], 0x5432);  // x[2ow+1], x[2ow+2]
    c.w = e[it] >> 16;                        // x[2ow+3], 0
    if (rest < SD_SLOTS * 7) *reinterpret_cast<uint4*>(rowbase + slot * (SD_PW * 128) + ((kh ^ sw) << 4)) = c;
  }
}

// chunk kh = 7 of every row is zero for every tile and no producer ever writes it: clear the operand stages once
__device__ __forceinline__ void clear_stages(uint8_t* base, int stage_stride, int nstages) {
  for (int s = 0; s < nstages; ++s)
    for (int i = threadIdx.x; i < SD_A_BYTES / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(base + s * stage_stride)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
}

// ---------------------------------------------------------------------------------------------------------------
// forward: y0[n, t, pixel, :] = sum_kt window(n, t + kt - 2, pixel) . Wm[:, kt*64 .. kt*64+64)^T  (+ BN statistics)
// warps: 0 = weight load, 1 = MMA issue, 2..9 = epilogue (two warpgroups alternating tiles), 10..17 = producers
// ---------------------------------------------------------------------------------------------------------------
struct FwdSmem {
  static constexpr int B_OFFSET = 0;
  static constexpr int A_OFFSET = SD_B_BYTES;
  static constexpr int STAGING_OFFSET = A_OFFSET + SD_STAGES * SD_A_BYTES;  // 2 x 16 KB
  static constexpr int BAR_OFFSET = STAGING_OFFSET + 2 * 16384;
  static constexpr int STATS_OFFSET = BAR_OFFSET + 256;  // fp32 [8 warps][2][64]
  static constexpr int TOTAL = STATS_OFFSET + 8 * 128 * 4 + 1024;
  static_assert(A_OFFSET % 1024 == 0 && SD_A_BYTES % 1024 == 0 && STAGING_OFFSET % 1024 == 0, "swizzle atoms");
  static_assert(TOTAL <= 232448, "exceeds 227 KB of shared memory");
};

__global__ void __launch_bounds__(320 + SD_FWD_PRODUCERS, 1)
conv_stem_direct_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                        const StemDirectParams p) {
  using L = FwdSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem + L::B_OFFSET;
  uint8_t* sA = smem + L::A_OFFSET;
  uint8_t* s_stage = smem + L::STAGING_OFFSET;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + SD_STAGES;
  uint64_t* tmem_full_bar = empty_bar + SD_STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]
  uint64_t* b_bar = tmem_empty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_bar + 1);
  float* s_stats = reinterpret_cast<float*>(smem + L::STATS_OFFSET);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 128;  // two 64-column accumulators
  const int tiles_per_clip = p.ttiles * p.wtiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < SD_STAGES; ++s) mbar_init(&full_bar[s], SD_FWD_PRODUCERS), mbar_init(&empty_bar[s], 1);
    for (int a = 0; a < 2; ++a) mbar_init(&tmem_full_bar[a], 1), mbar_init(&tmem_empty_bar[a], 4);
    mbar_init(b_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) s_stats[i] = 0.f;
  clear_stages(sA, SD_A_BYTES, SD_STAGES);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {  // weights: resident for the whole kernel
      mbar_expect_tx(b_bar, SD_B_BYTES);
      for (int t = 0; t < 5; ++t) tma_load_2d(sB + t * 8192, &tmB, b_bar, t * 64, 0);
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
        const uint32_t a0 = smem_u32(sA + stage * SD_A_BYTES);
        const uint32_t b0 = smem_u32(sB);
#pragma unroll
        for (int t = 0; t < 5; ++t) {  // tap t reads halo rows 16 t .. 16 t + 127: a 2 KB shift of the descriptor start
          const uint64_t a_desc = umma_smem_desc_sw128(a0 + (uint32_t)(t * SD_PW * 128), 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(b0 + (uint32_t)(t * 8192), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (t | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full_bar[acc]);
        if (++stage == SD_STAGES) stage = 0, phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 10) {
    // ---------------- producers: build the halo tile of the patch rows in shared memory ----------------
    const int ptid = threadIdx.x - 320;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_clip, rem = tile - n * tiles_per_clip;
      const int tt = rem / p.wtiles, wt = rem - tt * p.wtiles;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      build_patch_tile<SD_FWD_PRODUCERS>(sA + stage * SD_A_BYTES, p, n, tt * SD_FT, wt * SD_PW, ptid);
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(&full_bar[stage]);
      if (++stage == SD_STAGES) stage = 0, phase ^= 1;
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> staging tile -> TMA store (igemm_stem.cu's) ----------------
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;       // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;  // accumulator row = (frame r / 16, pixel r % 16) of the tile = staged row
    const int ft = r >> 4, px = r & (SD_PW - 1);
    const bool leader = threadIdx.x == 64 + 128 * g;
    const int bar_id = 1 + g;
    uint8_t* stg = s_stage + g * 16384;
    float st_sum[2] = {0.f, 0.f}, st_sq[2] = {0.f, 0.f};
    const int W = p.wtiles * SD_PW;
    int j = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++j) {
      if ((j & 1) != g) continue;
      const int n = tile / tiles_per_clip, rem = tile - n * tiles_per_clip;
      const int tt = rem / p.wtiles, wt = rem - tt * p.wtiles;
      const int t0 = tt * SD_FT, w0 = wt * SD_PW;
      const bool valid = (t0 + ft) < p.T && (w0 + px) < W;
      mbar_wait(&tmem_full_bar[g], (uint32_t)((j >> 1) & 1));
      tcgen05_fence_after();
      if (leader) tma_store_wait_read<0>();  // this group's previous store has read the staging buffer
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 64 + ch * 32), v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 t;
            t.x = pack_bf16x2(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1]));
            t.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
            t.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
            t.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
            const int chunk = (ch * 4 + i) ^ (r & 7);
            *reinterpret_cast<uint4*>(stg + r * 128 + chunk * 16) = t;
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[g]);
      fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (leader) {
        tma_store_4d(&tmC, stg, 0, w0, t0, n);  // frames >= T are clipped by the tensor map
        tma_store_commit();
      }
      if (p.bn_stats) {
        const int nrows = min(SD_FT, p.T - t0) * SD_PW;
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

// ---------------------------------------------------------------------------------------------------------------
// weight gradient: dW[kt*64 + slot, co] += sum over (n, t, pixel) window(n, t + kt - 2, pixel)[slot] * dz[n, t, pixel, co]
// warps: 0 = gradient-tile TMA, 1 = MMA issue, 2..5 = producers during the main loop, then the TMEM drain
// ---------------------------------------------------------------------------------------------------------------
constexpr int WD_STAGE_BYTES = SD_A_BYTES + SD_DZ_BYTES;
constexpr int WD_PAIRS = 3;

struct WgradSmem {
  static constexpr int BAR_OFFSET = SD_STAGES * WD_STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
  static_assert(WD_STAGE_BYTES % 1024 == 0, "swizzle atoms");
  static_assert(TOTAL <= 232448, "exceeds 227 KB of shared memory");
};

__global__ void __launch_bounds__(192, 1)
wgrad_stem_direct_kernel(const __grid_constant__ CUtensorMap tmDZ, const StemDirectParams p) {
  using L = WgradSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + SD_STAGES;
  uint64_t* tmem_full_bar = empty_bar + SD_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 256;  // 3 x 64 accumulator columns
  const int tiles_per_clip = p.ttiles * p.wtiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDZ);
    // a stage is full when the 128 producers have arrived and the gradient box's bytes have landed
    for (int s = 0; s < SD_STAGES; ++s) mbar_init(&full_bar[s], SD_WG_PRODUCERS + 1), mbar_init(&empty_bar[s], 1);
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  clear_stages(smem, WD_STAGE_BYTES, SD_STAGES);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int my_tiles = blockIdx.x < p.total_tiles ? (p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_clip, rem = tile - n * tiles_per_clip;
        const int tt = rem / p.wtiles, wt = rem - tt * p.wtiles;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], (uint32_t)SD_DZ_BYTES);  // full-size box: frames >= T are zero-filled
        tma_load_4d(smem + stage * WD_STAGE_BYTES + SD_A_BYTES, &tmDZ, &full_bar[stage], 0, wt * SD_PW, tt * SD_FT, n);
        if (++stage == SD_STAGES) stage = 0, phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);  // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int tt = (tile % tiles_per_clip) / p.wtiles;
        const int ksteps = min(SD_FT, p.T - tt * SD_FT);  // one k-step = one frame = 16 contraction rows
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t a0 = smem_u32(smem + stage * WD_STAGE_BYTES);
        const uint32_t b0 = a0 + SD_A_BYTES;
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t b_desc = umma_smem_desc_sw128(b0 + ks * 2048, 16384, 1024);
#pragma unroll
          for (int j = 0; j < WD_PAIRS; ++j) {
            // taps (2j, 2j+1) as one M = 128 operand: the second tap's rows start one frame (2 KB) later; the fifth
            // tap is alone (its upper 64 accumulator rows are never read)
            const uint64_t a_desc =
                umma_smem_desc_sw128(a0 + (uint32_t)(2 * j * SD_PW * 128) + ks * 2048, j < 2 ? SD_PW * 128 : 0, 1024);
            umma_bf16(tmem_base + (uint32_t)(j * 64), a_desc, b_desc, idesc, !(first && ks == 0));
          }
        }
        first = false;
        umma_commit(&empty_bar[stage]);
        if (++stage == SD_STAGES) stage = 0, phase ^= 1;
      }
      umma_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    {  // producers
      const int ptid = threadIdx.x - 64;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_clip, rem = tile - n * tiles_per_clip;
        const int tt = rem / p.wtiles, wt = rem - tt * p.wtiles;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        build_patch_tile<SD_WG_PRODUCERS>(smem + stage * WD_STAGE_BYTES, p, n, tt * SD_FT, wt * SD_PW, ptid);
        fence_proxy_async_smem();
        mbar_arrive(&full_bar[stage]);
        if (++stage == SD_STAGES) stage = 0, phase ^= 1;
      }
    }
    if (my_tiles > 0) {  // drain: accumulator row r of pair j = weight row (tap 2j + r / 64, slot r % 64)
      const int q = warp & 3;
      const int r = q * 32 + lane;
      mbar_wait(tmem_full_bar, 0);
      tcgen05_fence_after();
#pragma unroll 1
      for (int j = 0; j < WD_PAIRS; ++j) {
        const int tap = (j == 2 && r >= 64) ? -1 : 2 * j + (r >> 6);
        float* orow = p.out + (long long)((tap < 0 ? 0 : tap) * 64 + (r & 63)) * p.ldo;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 64 + ch * 32), v);
          tmem_ld_wait();
          if (tap < 0) continue;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + ch * 32 + 4 * i),
                         "f"(__uint_as_float(v[4 * i])), "f"(__uint_as_float(v[4 * i + 1])),
                         "f"(__uint_as_float(v[4 * i + 2])), "f"(__uint_as_float(v[4 * i + 3]))
                         : "memory");
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

int fill_params(StemDirectParams& p, const void* video_bf16, int B, int T, int H, int W) {
  SVSR_REQUIRE(stem_direct_supported(H, W), "stem_direct: %dx%d frames are not covered (even width, output pixels a multiple of 16)", H, W);
  SVSR_REQUIRE((reinterpret_cast<uintptr_t>(video_bf16) & 3) == 0, "stem_direct: the bf16 video must be 4-byte aligned");
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  p.vid = static_cast<const uint32_t*>(video_bf16);
  p.N = B, p.T = T, p.H = H, p.W2 = W / 2, p.OW = OW;
  p.wtiles = OH * OW / SD_PW;
  p.ttiles = (T + SD_FT - 1) / SD_FT;
  p.total_tiles = B * p.ttiles * p.wtiles;
  return SVSR_OK;
}

}  // namespace

bool stem_direct_supported(int H, int W) {
  if (H < 1 || W < 2 || (W & 1)) return false;
  const int OH = (H + 6 - 7) / 2 + 1, OW = (W + 6 - 7) / 2 + 1;
  return OW == W / 2 && (OH * OW) % SD_PW == 0;
}

int stem_direct_fwd(const void* video_bf16, const __nv_bfloat16* w_packed, __nv_bfloat16* y0, double* bn_stats, int B,
                    int T, int H, int W, double algo_flops, cudaStream_t stream) {
  StemDirectParams p{};
  SD_RC(fill_params(p, video_bf16, B, T, H, W));
  p.bn_stats = bn_stats;
  const uint64_t npix = (uint64_t)p.wtiles * SD_PW;
  CUtensorMap tmB, tmC;
  {
    uint64_t dims[4] = {64, npix, (uint64_t)T, (uint64_t)B};
    uint64_t strides[3] = {128, npix * 128, (uint64_t)T * npix * 128};
    uint32_t boxc[4] = {64, SD_PW, SD_FT, 1};
    SD_RC(make_tmap_bf16(&tmC, y0, 4, dims, strides, boxc, nullptr, true));
  }
  {
    uint64_t dims[2] = {320, 64};
    uint64_t strides[1] = {320 * 2};
    uint32_t box[2] = {64, 64};
    SD_RC(make_tmap_bf16(&tmB, w_packed, 2, dims, strides, box, nullptr, true));
  }
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(conv_stem_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    attr_done = true;
  }
  const int grid = p.total_tiles < 148 ? p.total_tiles : 148;
  prof_begin(PROF_IGEMM, algo_flops > 0 ? algo_flops : 2.0 * B * T * (double)npix * 64.0 * 320.0, stream);
  conv_stem_direct_kernel<<<grid, 320 + SD_FWD_PRODUCERS, FwdSmem::TOTAL, stream>>>(tmB, tmC, p);
  note_launch();
  prof_end(stream);
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

int stem_direct_wgrad(const void* video_bf16, const __nv_bfloat16* dz, float* out, int ldo, int B, int T, int H, int W,
                      double algo_flops, cudaStream_t stream) {
  StemDirectParams p{};
  SD_RC(fill_params(p, video_bf16, B, T, H, W));
  SVSR_REQUIRE(ldo >= 64 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
               "stem_direct_wgrad: the gradient rows must be 16-byte aligned (ldo=%d)", ldo);
  p.out = out, p.ldo = ldo;
  const uint64_t npix = (uint64_t)p.wtiles * SD_PW;
  CUtensorMap tmDZ;
  uint64_t dims[4] = {64, npix, (uint64_t)T, (uint64_t)B};
  uint64_t strides[3] = {128, npix * 128, (uint64_t)T * npix * 128};
  uint32_t boxz[4] = {64, SD_PW, SD_FT, 1};
  SD_RC(make_tmap_bf16(&tmDZ, dz, 4, dims, strides, boxz, nullptr, true));
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_stem_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem::TOTAL));
    attr_done = true;
  }
  const int grid = p.total_tiles < 148 ? p.total_tiles : 148;
  prof_begin(PROF_WGRAD, algo_flops > 0 ? algo_flops : 2.0 * B * T * (double)npix * 320.0 * 64.0, stream);
  wgrad_stem_direct_kernel<<<grid, 192, WgradSmem::TOTAL, stream>>>(tmDZ, p);
  note_launch();
  prof_end(stream);
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // namespace svsr
