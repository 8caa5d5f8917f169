// Weight gradient of the 3x3 / stride-1 / 64 -> 64 channel convolutions (resnet.layer1) with ONE load of the
// activation tile per pixel tile instead of one per filter-tap pair.
//
// The generic kernel (wgrad.cu) fetches, per pixel tile, the activation box once for each of the 9 taps plus the
// gradient box once for each of the 5 tap pairs: 240 KB through L2 -> SM for 31 KB of distinct data, and the five
// 64-channel weight gradients ran at 220 us per launch (L2-bound). Here a pixel tile is RT image rows of one image on the
// zero-padded halo grid of pitch P = W + 2 (as in igemm_halo.cu):
//   * x: one TMA box (64 ch, P, RT + 2, 1) at (w, h) = (-1, h0 - 1)  -> (RT + 2) * P halo pixels, 128 B each;
//   * dy: one TMA box (64 ch, P, RT, 1) at (w, h) = (-1, h0)        -> RT * P rows whose halo columns are zero-filled,
//     so contraction index k (= halo position of the OUTPUT pixel) pairs dy[k] with x[k + (1+dh) * P + dw] for tap (dh, dw);
//   * both are MN-major UMMA operands (the pixel is the strided dimension); a tap is a row offset of the A descriptor's
//     start address, and a PAIR of taps forms one M = 128 operand whose two 64-channel halves are
//     LBO = (shift_b - shift_a) * 128 bytes apart inside the same tile.
// D (fp32, TMEM): 5 tap pairs x [128 x 64], accumulated over all the tiles a persistent CTA owns and reduced into the
// gradient with red.add once at the end. L2 -> SM traffic per tile: 36.5 KB instead of 240 KB.
#include "wgrad.cuh"
#include "tmap.h"
#include <stdlib.h>

namespace svsr {

namespace {

constexpr int WH_STAGES = 4;
constexpr int WH_A_BYTES = 25600;  // 1024 B lead (row -1) + up to 192 halo rows
constexpr int WH_B_BYTES = 16384;  // 128 contraction rows
constexpr int WH_STAGE_BYTES = WH_A_BYTES + WH_B_BYTES;

struct WgradHaloParams {
  int N, H, W, P, RT, tiles_per_img, total_tiles;
  int pair_row[5];  // row shift of the first tap of pair j
  int pair_lbo[5];  // byte distance to the second tap's rows (0 for the unpaired ninth tap)
  int pair_tap[5][2];
  int a_bytes, b_bytes;
  float* out;
  int ldo;
};

struct WgradHaloSmem {
  static constexpr int BAR_OFFSET = WH_STAGES * WH_STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
  static_assert(TOTAL <= 232448, "exceeds 227 KB of shared memory");
};

__global__ void __launch_bounds__(192, 1)
wgrad3x3_c64_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                         const WgradHaloParams p) {
  using L = WgradHaloSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + WH_STAGES;
  uint64_t* tmem_full_bar = empty_bar + WH_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;  // 5 x 64 accumulator columns

  // Rows the TMA boxes never write (beyond the halo tile / beyond RT * P gradient rows) must be zero: they meet zero
  // gradients or feed nothing, but a stale NaN pattern times zero would still poison the sum.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < L::BAR_OFFSET / 16; i += blockDim.x) z[i] = zero;
  }
  fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < WH_STAGES; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
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
        const int n = tile / p.tiles_per_img, h0 = (tile - n * p.tiles_per_img) * p.RT;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * WH_STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], (uint32_t)(p.a_bytes + p.b_bytes));
        tma_load_4d(sA + 1024, &tmX, &full_bar[stage], 0, -1, h0 - 1, n);
        tma_load_4d(sA + WH_A_BYTES, &tmDY, &full_bar[stage], 0, -1, h0, n);
        if (++stage == WH_STAGES) stage = 0, phase ^= 1;
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
        const int h0 = (tile % p.tiles_per_img) * p.RT;
        // gradient rows beyond the image's last row are zero-filled: skip their k-steps (the ragged last tile of a
        // 22-row image has 2 of 5 rows)
        const int ksteps = (min(p.RT, p.H - h0) * p.P + 15) >> 4;
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t a0 = smem_u32(smem + stage * WH_STAGE_BYTES + 1024);
        const uint32_t b0 = smem_u32(smem + stage * WH_STAGE_BYTES + WH_A_BYTES);
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          // one MMA consumes 16 contraction rows = two 8-row swizzle atoms (SBO = 1024 B apart)
          const uint64_t b_desc = umma_smem_desc_sw128(b0 + ks * 2048, 16384, 1024);
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const uint64_t a_desc =
                umma_smem_desc_sw128(a0 + (uint32_t)(p.pair_row[j] * 128) + ks * 2048, (uint32_t)p.pair_lbo[j], 1024);
            umma_bf16(tmem_base + (uint32_t)(j * 64), a_desc, b_desc, idesc, !(first && ks == 0));
          }
        }
        first = false;
        umma_commit(&empty_bar[stage]);
        if (++stage == WH_STAGES) stage = 0, phase ^= 1;
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
    for (int j = 0; j < 5; ++j) {
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

bool wgrad_halo_matches(const WgradProblem& p) {
  if (!(p.ntaps == 9 && p.a_cin == 64 && p.n_cols == 64 && p.a_stride == 1 && p.a_C == 64 && p.b_C == 64)) return false;
  if (p.a_coff != 0 || p.b_coff != 0 || p.k_H != p.a_H || p.k_W != p.a_W || p.k_N != p.a_N) return false;
  if (p.a_W + 2 > 27 || p.a_W < 2 || p.ldo % 4 != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15) != 0) return false;
  if (p.m_valid > 0 && p.m_valid != 576) return false;
  for (int t = 0; t < 9; ++t)
    if (p.tap_dh[t] < -1 || p.tap_dh[t] > 1 || p.tap_dw[t] < -1 || p.tap_dw[t] > 1) return false;
  const char* e = getenv("SVSR_HALO_CONV");
  return !(e && e[0] == '0');
}

int wgrad_halo_launch(const WgradProblem& p, cudaStream_t stream) {
  WgradHaloParams kp{};
  kp.N = p.a_N, kp.H = p.a_H, kp.W = p.a_W, kp.P = p.a_W + 2;
  kp.RT = 128 / kp.P;
  kp.tiles_per_img = (kp.H + kp.RT - 1) / kp.RT;
  kp.total_tiles = kp.N * kp.tiles_per_img;
  SVSR_REQUIRE(1024 + (128 + 2 * kp.P + 2) * 128 <= WH_A_BYTES, "wgrad halo: tile does not fit its stage");
  // taps sorted by row shift, paired (0,1) (2,3) (4,5) (6,7) (8,-): the second tap of a pair must not precede the first
  int order[9], shift[9];
  for (int t = 0; t < 9; ++t) order[t] = t, shift[t] = (1 + p.tap_dh[t]) * kp.P + p.tap_dw[t];
  for (int i = 0; i < 9; ++i)
    for (int j = i + 1; j < 9; ++j)
      if (shift[order[j]] < shift[order[i]]) { const int tmp = order[i]; order[i] = order[j]; order[j] = tmp; }
  for (int j = 0; j < 5; ++j) {
    const int ta = order[2 * j], tb = j < 4 ? order[2 * j + 1] : -1;
    kp.pair_row[j] = shift[ta];
    kp.pair_lbo[j] = tb >= 0 ? (shift[tb] - shift[ta]) * 128 : 0;
    kp.pair_tap[j][0] = ta, kp.pair_tap[j][1] = tb;
  }
  kp.a_bytes = (kp.RT + 2) * kp.P * 128, kp.b_bytes = kp.RT * kp.P * 128;
  kp.out = p.out, kp.ldo = p.ldo;
  CUtensorMap tmX, tmDY;
  uint64_t dims[4] = {64, (uint64_t)kp.W, (uint64_t)kp.H, (uint64_t)kp.N};
  uint64_t strides[3] = {128, (uint64_t)kp.W * 128, (uint64_t)kp.H * kp.W * 128};
  uint32_t boxx[4] = {64, (uint32_t)kp.P, (uint32_t)(kp.RT + 2), 1};
  uint32_t boxy[4] = {64, (uint32_t)kp.P, (uint32_t)kp.RT, 1};
  int rc = make_tmap_bf16(&tmX, p.a, 4, dims, strides, boxx, nullptr, true);
  if (rc) return rc;
  rc = make_tmap_bf16(&tmDY, p.b, 4, dims, strides, boxy, nullptr, true);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad3x3_c64_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         WgradHaloSmem::TOTAL));
    attr_done = true;
  }
  const int grid = kp.total_tiles < 148 ? kp.total_tiles : 148;
  const double flops = p.algo_flops > 0 ? p.algo_flops : 2.0 * kp.N * kp.H * kp.W * 576.0 * 64.0;
  prof_begin(PROF_WGRAD, flops, stream);
  wgrad3x3_c64_halo_kernel<<<grid, 192, WgradHaloSmem::TOTAL, stream>>>(tmX, tmDY, kp);
  note_launch();
  prof_end(stream);
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // namespace svsr
