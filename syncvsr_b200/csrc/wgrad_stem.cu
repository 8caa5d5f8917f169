// Weight gradient of the stem's temporal 5-tap / 64 -> 64 column contraction (backward-weight of
// Conv3d(1, 64, (5,7,7), ...), lightning.py:50, over the [clips, frames, pixels, 64] patch rows) with ONE load of the
// patch tile per pixel tile instead of one per temporal tap.
//
// The generic kernel (wgrad.cu) fetches, per 128-pixel tile, the patch box once for each of the 5 taps plus the gradient
// box once per tap pair; the launch sits alone at the very end of the backward pass (its gradient operand is the last
// tensor backward produces), so its 360 us are exposed step time. Here a pixel tile is FT = 8 frames x PW = 16 pixel
// rows of one clip (as in igemm_stem.cu):
//   * x:  one TMA box (64 ch, 16 px, 8 + 4 frames, 1 clip) at frame t0 - 2 -> 192 halo rows of 128 B (frames outside
//     [0, T) zero-filled = the temporal padding);
//   * dz: one TMA box (64 ch, 16 px, 8 frames, 1 clip) at frame t0 -> 128 contraction rows (frames >= T zero-filled),
//     so contraction index k = (frame, pixel) of the OUTPUT pairs dz[k] with x[k + 16 kt] for tap kt;
//   * both are MN-major UMMA operands; a tap is a 2 KB offset of the A descriptor's start address, and a PAIR of taps
//     (kt, kt + 1) forms one M = 128 operand whose two 64-column halves are LBO = 2 KB apart inside the same tile.
// D (fp32, TMEM): 3 tap pairs x [128 x 64] (the sixth half is unused), accumulated over all the tiles a persistent CTA
// owns and reduced into the gradient with red.add once at the end. L2 -> SM traffic per tile: 40 KB instead of 128 KB.
#include "wgrad.cuh"
#include "tmap.h"
#include <stdlib.h>

namespace svsr {

namespace {

constexpr int WS_STAGES = 4;
constexpr int WS_FT = 8, WS_PW = 16;
constexpr int WS_A_BYTES = (WS_FT + 4) * WS_PW * 128;  // 24 KB halo tile
constexpr int WS_B_BYTES = WS_FT * WS_PW * 128;        // 16 KB = 128 contraction rows
constexpr int WS_STAGE_BYTES = WS_A_BYTES + WS_B_BYTES;
constexpr int WS_PAIRS = 3;

struct WgradStemParams {
  int N, T, W, wtiles, ttiles, total_tiles;
  int pair_off[WS_PAIRS];  // byte offset of the first tap's rows of pair j inside the halo tile
  int pair_lbo[WS_PAIRS];  // byte distance to the second tap's rows (0 for the unpaired fifth tap)
  int pair_tap[WS_PAIRS][2];
  float* out;
  int ldo;
};

struct WgradStemSmem {
  static constexpr int BAR_OFFSET = WS_STAGES * WS_STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
  static_assert(WS_STAGE_BYTES % 1024 == 0 && WS_A_BYTES % 1024 == 0, "swizzle atoms");
  static_assert(TOTAL <= 232448, "exceeds 227 KB of shared memory");
};

__global__ void __launch_bounds__(192, 1)
wgrad_t5_c64_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDZ,
                         const WgradStemParams p) {
  using L = WgradStemSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + WS_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WS_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 256;  // 3 x 64 accumulator columns
  const int tiles_per_clip = p.ttiles * p.wtiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDZ);
    for (int s = 0; s < WS_STAGES; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr_smem, TMEM_COLS);
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
        uint8_t* sA = smem + stage * WS_STAGE_BYTES;
        // both boxes are always full size (out-of-range frames are zero-filled), so every row of the stage is rewritten
        mbar_expect_tx(&full_bar[stage], (uint32_t)(WS_A_BYTES + WS_B_BYTES));
        tma_load_4d(sA, &tmX, &full_bar[stage], 0, wt * WS_PW, tt * WS_FT - 2, n);
        tma_load_4d(sA + WS_A_BYTES, &tmDZ, &full_bar[stage], 0, wt * WS_PW, tt * WS_FT, n);
        if (++stage == WS_STAGES) stage = 0, phase ^= 1;
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
        // one k-step = 16 contraction rows = one frame of the tile; frames >= T hold zero gradients: skip them
        const int ksteps = min(WS_FT, p.T - tt * WS_FT);
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t a0 = smem_u32(smem + stage * WS_STAGE_BYTES);
        const uint32_t b0 = a0 + WS_A_BYTES;
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          // one MMA consumes 16 contraction rows = two 8-row swizzle atoms (SBO = 1024 B apart)
          const uint64_t b_desc = umma_smem_desc_sw128(b0 + ks * 2048, 16384, 1024);
#pragma unroll
          for (int j = 0; j < WS_PAIRS; ++j) {
            const uint64_t a_desc =
                umma_smem_desc_sw128(a0 + (uint32_t)p.pair_off[j] + ks * 2048, (uint32_t)p.pair_lbo[j], 1024);
            umma_bf16(tmem_base + (uint32_t)(j * 64), a_desc, b_desc, idesc, !(first && ks == 0));
          }
        }
        first = false;
        umma_commit(&empty_bar[stage]);
        if (++stage == WS_STAGES) stage = 0, phase ^= 1;
      }
      umma_commit(tmem_full_bar);
    }
    __syncwarp();
  } else if (my_tiles > 0) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
#pragma unroll 1
    for (int j = 0; j < WS_PAIRS; ++j) {
      const int tap = p.pair_tap[j][r >> 6];
      float* orow = p.out + (long long)(tap * 64 + (r & 63)) * p.ldo;
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

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

bool wgrad_stem_matches(const WgradProblem& p) {
  if (!(p.ntaps == 5 && p.a_cin == 64 && p.n_cols == 64 && p.a_stride == 1 && p.a_C == 64 && p.b_C == 64)) return false;
  if (p.a_coff != 0 || p.b_coff != 0 || p.k_H != p.a_H || p.k_W != p.a_W || p.k_N != p.a_N) return false;
  if (p.a_W % WS_PW != 0 || p.a_W < WS_PW || p.ldo % 4 != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return false;
  if (p.m_valid > 0 && p.m_valid != 320) return false;
  for (int t = 0; t < 5; ++t) {
    if (p.tap_dw[t] != 0 || p.tap_dh[t] < -2 || p.tap_dh[t] > 2) return false;
    for (int u = 0; u < t; ++u)
      if (p.tap_dh[u] == p.tap_dh[t]) return false;
  }
  const char* e = getenv("SVSR_STEM_HALO");
  return !(e && e[0] == '0');
}

int wgrad_stem_launch(const WgradProblem& p, cudaStream_t stream) {
  WgradStemParams kp{};
  kp.N = p.a_N, kp.T = p.a_H, kp.W = p.a_W;
  kp.wtiles = kp.W / WS_PW;
  kp.ttiles = (kp.T + WS_FT - 1) / WS_FT;
  kp.total_tiles = kp.N * kp.ttiles * kp.wtiles;
  // taps sorted by frame shift, paired (0,1) (2,3) (4,-): the second tap of a pair must not precede the first
  int order[5], shift[5];
  for (int t = 0; t < 5; ++t) order[t] = t, shift[t] = (2 + p.tap_dh[t]) * WS_PW;  // in 128-byte rows
  for (int i = 0; i < 5; ++i)
    for (int j = i + 1; j < 5; ++j)
      if (shift[order[j]] < shift[order[i]]) { const int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
  for (int j = 0; j < WS_PAIRS; ++j) {
    const int ta = order[2 * j], tb = j < 2 ? order[2 * j + 1] : -1;
    kp.pair_off[j] = shift[ta] * 128;
    kp.pair_lbo[j] = tb >= 0 ? (shift[tb] - shift[ta]) * 128 : 0;
    kp.pair_tap[j][0] = ta, kp.pair_tap[j][1] = tb;
  }
  kp.out = p.out, kp.ldo = p.ldo;
  CUtensorMap tmX, tmDZ;
  uint64_t dims[4] = {64, (uint64_t)kp.W, (uint64_t)kp.T, (uint64_t)kp.N};
  uint64_t strides[3] = {128, (uint64_t)kp.W * 128, (uint64_t)kp.T * kp.W * 128};
  uint32_t boxx[4] = {64, WS_PW, WS_FT + 4, 1};
  uint32_t boxz[4] = {64, WS_PW, WS_FT, 1};
  int rc = make_tmap_bf16(&tmX, p.a, 4, dims, strides, boxx, nullptr, true);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmDZ, p.b, 4, dims, strides, boxz, nullptr, true);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_t5_c64_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         WgradStemSmem::TOTAL));
    attr_done = true;
  }
  const int grid = kp.total_tiles < 148 ? kp.total_tiles : 148;
  const double flops = p.algo_flops > 0 ? p.algo_flops : 2.0 * kp.N * kp.T * (double)kp.W * 320.0 * 64.0;
  prof_begin(PROF_WGRAD, flops, stream);
  wgrad_t5_c64_halo_kernel<<<grid, 192, WgradStemSmem::TOTAL, stream>>>(tmX, tmDZ, kp);
  note_launch();
  prof_end(stream);
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // namespace svsr
