// tcgen05 weight-gradient kernel (see wgrad.cuh). Same warp roles as igemm.cu.
// smem stage = A: two 64-channel blocks [128 pixel rows x 128 B] + B: BN/64 such blocks, all SWIZZLE_128B, MN-major.
#include "wgrad.cuh"
#include <stdlib.h>
#include "igemm.cuh"  // igemm_choose_box
#include "tmap.h"

namespace svsr {

struct WgradKParams {
  int tiles_h, tiles_w, ktiles;
  int bn, bh, bw, box_rows;
  int a_stride, a_coff, a_cblocks;  // a_cblocks = a_cin / 64
  int ngroups;                      // ntaps * a_cblocks: number of 64-row groups of D
  int mb_per_cta;
  int tap_dh[WGRAD_MAX_TAPS];
  int tap_dw[WGRAD_MAX_TAPS];
  int b_coff, n_cols;
  float* out;
  int ldo;
  int m_valid;
  int vec_ok;  // output rows 16-byte aligned: red.global.add.v4.f32 usable
  StepCtl ctl;
};

// CG = 2 (CTA pair, cta_group::2): the pair's MMA covers two 128-row blocks of D (one per CTA) against ONE gradient tile whose
// 64-channel blocks are split between the two CTAs.
template <int BN, int STAGES, int CG = 1>
struct WgradSmem {
  static constexpr int BLK = 128 * 128;  // one 64-channel block: 128 pixel rows x 128 B
  static constexpr int A_BYTES = 2 * BLK;
  static constexpr int B_BYTES = (BN / 64 / CG) * BLK;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;
};

template <int BN, int STAGES, int CG = 1>
__global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradKParams p) {
  if (ctl_skipped(p.ctl)) return;
  using L = WgradSmem<BN, STAGES, CG>;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CG = 1: this CTA owns blocks mb0 .. mb0 + mb_cnt - 1. CG = 2: the pair owns block PAIRS, MMA i of the pair covers blocks
  // mb0 + 2 i (leader's accumulator rows) and mb0 + 2 i + 1 (peer's); block_of(i) is this CTA's block
  const int mb0 = (int)(blockIdx.x / CG) * p.mb_per_cta * CG;
  const int nb = blockIdx.y;
  const int split = blockIdx.z, nsplit = gridDim.z;
  const int num_mblocks = (p.ngroups + 1) / 2;
  const int mb_cnt = min(p.mb_per_cta, (num_mblocks - mb0 + CG - 1) / CG);
  auto block_of = [&](int i) { return mb0 + CG * i + rank; };
  auto nsub_of = [&](int mb) { return max(0, min(2, p.ngroups - 2 * mb)); };
  const int my_ktiles = (p.ktiles - split + nsplit - 1) / nsplit;  // kt = split, split+nsplit, ...
  constexpr uint32_t TMEM_COLS = 512;

  // Rows >= box_rows of every block are never written by TMA: zero them once so they contribute 0 to the sums.
  if (p.box_rows < 128) {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < L::BAR_OFFSET / 16; i += blockDim.x) z[i] = zero;
  }
  fence_proxy_async_smem();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2sm(tmem_ptr_smem, TMEM_COLS);
    else tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t box_bytes = (uint32_t)p.box_rows * 128u;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_ktiles; ++i) {
        const int kt = split + i * nsplit;
        const int tw = kt % p.tiles_w;
        const int th = (kt / p.tiles_w) % p.tiles_h;
        const int tn = kt / (p.tiles_w * p.tiles_h);
        const int n0 = tn * p.bn, h0 = th * p.bh, w0 = tw * p.bw;
        for (int mbi = 0; mbi < mb_cnt; ++mbi) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * L::STAGE_BYTES;
          uint8_t* sB = sA + L::A_BYTES;
          const int mb = block_of(mbi);
          const int g0 = 2 * mb;
          const int nsub = nsub_of(mb);
          // pair: the leader's barrier counts the bytes of BOTH CTAs' loads (an odd tail block of the peer loads no A)
          if (rank == 0)
            mbar_expect_tx(&full_bar[stage],
                           box_bytes * (uint32_t)(nsub + (CG == 2 ? nsub_of(mb + 1) : 0) + BN / 64));
          for (int sub = 0; sub < nsub; ++sub) {
            const int g = g0 + sub;
            const int tap = g / p.a_cblocks;
            const int cb = g - tap * p.a_cblocks;
            if (CG == 2)
              tma_load_4d_2sm(sA + sub * L::BLK, &tmA, &full_bar[stage], p.a_coff + cb * 64,
                              w0 * p.a_stride + p.tap_dw[tap], h0 * p.a_stride + p.tap_dh[tap], n0);
            else
              tma_load_4d(sA + sub * L::BLK, &tmA, &full_bar[stage], p.a_coff + cb * 64,
                          w0 * p.a_stride + p.tap_dw[tap], h0 * p.a_stride + p.tap_dh[tap], n0);
          }
#pragma unroll
          for (int blk = 0; blk < BN / 64 / CG; ++blk) {
            if (CG == 2)
              tma_load_4d_2sm(sB + blk * L::BLK, &tmB, &full_bar[stage],
                              p.b_coff + nb * BN + (rank * (BN / 128) + blk) * 64, w0, h0, n0);
            else
              tma_load_4d(sB + blk * L::BLK, &tmB, &full_bar[stage], p.b_coff + nb * BN + blk * 64, w0, h0, n0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {  // pair: the leader issues the MMAs for both CTAs' accumulator rows
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CG, BN, 1, 1);  // both operands MN-major
      const int ksteps = (p.box_rows + 15) / 16;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_ktiles; ++i) {
        for (int mbi = 0; mbi < mb_cnt; ++mbi) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t b_addr = a_addr + L::A_BYTES;
          for (int ks = 0; ks < ksteps; ++ks) {
            // one MMA consumes 16 pixel rows = two 8-row swizzle atoms (SBO = 1024 B apart);
            // 64-channel blocks are LBO = 16 KB apart.
            const uint64_t a_desc = umma_smem_desc_sw128(a_addr + ks * 2048, L::BLK, 1024);
            const uint64_t b_desc = umma_smem_desc_sw128(b_addr + ks * 2048, L::BLK, 1024);
            if (CG == 2) umma_bf16_2sm(tmem_base + (uint32_t)(mbi * BN), a_desc, b_desc, idesc, (i | ks) != 0);
            else umma_bf16(tmem_base + (uint32_t)(mbi * BN), a_desc, b_desc, idesc, (i | ks) != 0);
          }
          if (CG == 2) umma_commit_2sm(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (CG == 2) umma_commit_2sm(tmem_full_bar);
      else umma_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    for (int mbi = 0; mbi < mb_cnt; ++mbi) {
      const int g = 2 * block_of(mbi) + (r >> 6);
      const bool row_valid = (g < p.ngroups) && (my_ktiles > 0) && (g * 64 + (r & 63) < p.m_valid);
      float* orow = p.out + (long long)(g * 64 + (r & 63)) * p.ldo;
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mbi * BN + ch * 32), v);
        tmem_ld_wait();
        const int col0 = nb * BN + ch * 32;
        if (!row_valid) continue;
        if (col0 + 32 <= p.n_cols && p.vec_ok) {
          // vector reduction: 8 x red.global.add.v4.f32 instead of 32 scalar atomics
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow + col0 + 4 * j),
                         "f"(__uint_as_float(v[4 * j])), "f"(__uint_as_float(v[4 * j + 1])),
                         "f"(__uint_as_float(v[4 * j + 2])), "f"(__uint_as_float(v[4 * j + 3]))
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.n_cols) atomicAdd(orow + col0 + j, __uint_as_float(v[j]));
        }
      }
    }
  }

  tcgen05_fence_before();
  if (CG == 2) {
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, int STAGES>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradKParams& kp, dim3 grid,
                       cudaStream_t stream) {
  using L = WgradSmem<BN, STAGES, 2>;
  auto kernel = wgrad_kernel<BN, STAGES, 2>;
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = dim3(192), cfg.dynamicSmemBytes = L::TOTAL, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  SVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, kp));
  note_launch();
  return SVSR_OK;
}

template <int BN, int STAGES>
static int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const WgradKParams& kp, dim3 grid,
                    cudaStream_t stream) {
  using L = WgradSmem<BN, STAGES>;
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL));
    attr_done = true;
  }
  wgrad_kernel<BN, STAGES><<<grid, 192, L::TOTAL, stream>>>(tmA, tmB, kp);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

int wgrad_launch(const WgradProblem& p, cudaStream_t stream) {
  SVSR_REQUIRE(p.a && p.b && p.out, "wgrad: null operand");
  SVSR_REQUIRE(p.a_cin > 0 && p.a_cin % 64 == 0, "wgrad: a_cin=%d must be a positive multiple of 64", p.a_cin);
  SVSR_REQUIRE(p.a_C % 8 == 0 && p.a_coff % 8 == 0 && p.b_C % 8 == 0 && p.b_coff % 8 == 0,
               "wgrad: channel pitches/offsets must be multiples of 8");
  SVSR_REQUIRE(p.ntaps >= 1 && p.ntaps <= WGRAD_MAX_TAPS, "wgrad: ntaps=%d out of range", p.ntaps);
  SVSR_REQUIRE(p.a_stride == 1 || p.a_stride == 2, "wgrad: stride must be 1 or 2");
  SVSR_REQUIRE(p.n_cols > 0 && p.ldo >= p.n_cols, "wgrad: bad output geometry");
  if (wgrad_halo_matches(p)) return wgrad_halo_launch(p, stream);
  if (wgrad_stem_matches(p)) return wgrad_stem_launch(p, stream);

  WgradKParams kp{};
  igemm_choose_box(p.k_N, p.k_H, p.k_W, &kp.bn, &kp.bh, &kp.bw);
  kp.tiles_h = (p.k_H + kp.bh - 1) / kp.bh;
  kp.tiles_w = (p.k_W + kp.bw - 1) / kp.bw;
  const int tiles_n = (p.k_N + kp.bn - 1) / kp.bn;
  kp.ktiles = tiles_n * kp.tiles_h * kp.tiles_w;
  kp.box_rows = kp.bn * kp.bh * kp.bw;
  kp.a_stride = p.a_stride, kp.a_coff = p.a_coff, kp.a_cblocks = p.a_cin / 64;
  kp.ngroups = p.ntaps * kp.a_cblocks;
  for (int t = 0; t < p.ntaps; ++t) kp.tap_dh[t] = p.tap_dh[t], kp.tap_dw[t] = p.tap_dw[t];
  kp.b_coff = p.b_coff, kp.n_cols = p.n_cols;
  kp.out = p.out, kp.ldo = p.ldo;
  kp.m_valid = p.m_valid > 0 ? p.m_valid : p.ntaps * p.a_cin;
  kp.vec_ok = (p.ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
  kp.ctl = p.ctl;

  const int BN = p.n_cols <= 64 ? 64 : (p.n_cols <= 128 ? 128 : 256);
  const int num_mblocks = (kp.ngroups + 1) / 2;
  // CTA pairs (cta_group::2; SVSR_IGEMM_2CTA=0 turns them off): M = 256 MMAs over two 128-row blocks of D, the gradient tile's
  // 64-channel blocks split between the two CTAs
  static const bool pair_env = [] {
    const char* e = getenv("SVSR_IGEMM_2CTA");
    return !(e && e[0] == '0');
  }();
  const bool pair = pair_env && BN >= 128 && num_mblocks >= 2;
  const int cg = pair ? 2 : 1;
  kp.mb_per_cta = 512 / BN;  // accumulators [128 x BN] per CTA
  const int units = (num_mblocks + cg - 1) / cg;  // blocks (or block pairs) to distribute
  if (kp.mb_per_cta > units) kp.mb_per_cta = units;
  const int gx = cg * ((units + kp.mb_per_cta - 1) / kp.mb_per_cta);
  const int gy = (p.n_cols + BN - 1) / BN;
  // split-K factor from a small cost model (microseconds): one CTA per SM is resident (192 KB of smem), so the launch
  // runs in ceil(ctas / 148) waves of  t_fixed + mb_per_cta * (ktiles_per_cta * t_stage + t_epilogue);  the fp32 red.add
  // traffic of all splits is added at L2 atomic throughput. The previous rule (ceil(296 / (gx*gy)) splits) produced
  // 297-, 300-, 306- and 324-CTA grids: a third wave with a handful of CTAs, i.e. +50 % time on the layer2-4 convs.
  const double t_stage = BN == 256 ? 0.93 : (BN == 128 ? 0.53 : 0.33);  // 8 MMAs: (smem operand read + math) cycles
  const double t_epi = 0.5 + 2.0 * BN / 256.0, t_fixed = 2.5;
  int nsplit = 1;
  double best = 1e30;
  for (int n = 1; n <= kp.ktiles && n <= 1024; ++n) {
    const int ctas = gx * gy * n, waves = (ctas + 147) / 148, kt = (kp.ktiles + n - 1) / n;
    const double t_cta = t_fixed + kp.mb_per_cta * (kt * t_stage + t_epi);
    const double red_bytes = (double)ctas * kp.mb_per_cta * 128.0 * BN * 4.0;
    const double cost = waves * t_cta + 0.5 * red_bytes / 3.0e6;
    if (cost < best * 0.98) best = cost, nsplit = n;  // ties go to fewer splits (less atomic traffic)
  }
  if (const char* e = getenv("SVSR_WGRAD_NSPLIT")) {  // tuning override (tools/wgrad_bench.py)
    const int n = atoi(e);
    if (n >= 1) nsplit = n < kp.ktiles ? n : kp.ktiles;
  }
  if (const char* e = getenv("SVSR_WGRAD_SPLIT_LEGACY")) {
    if (e[0] == '1') {
      nsplit = (2 * 148 + gx * gy - 1) / (gx * gy);
      if (nsplit > (kp.ktiles + 3) / 4) nsplit = (kp.ktiles + 3) / 4;
      if (nsplit < 1) nsplit = 1;
    }
  }

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p.a_C, (uint64_t)p.a_W, (uint64_t)p.a_H, (uint64_t)p.a_N};
    uint64_t strides[3] = {(uint64_t)p.a_C * 2, (uint64_t)p.a_W * p.a_C * 2, (uint64_t)p.a_H * p.a_W * p.a_C * 2};
    uint32_t box[4] = {64, (uint32_t)((kp.bw - 1) * p.a_stride + 1), (uint32_t)((kp.bh - 1) * p.a_stride + 1),
                       (uint32_t)kp.bn};
    uint32_t es[4] = {1, (uint32_t)p.a_stride, (uint32_t)p.a_stride, 1};
    int rc = make_tmap_bf16(&tmA, p.a, 4, dims, strides, box, es, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)p.b_C, (uint64_t)p.k_W, (uint64_t)p.k_H, (uint64_t)p.k_N};
    uint64_t strides[3] = {(uint64_t)p.b_C * 2, (uint64_t)p.k_W * p.b_C * 2, (uint64_t)p.k_H * p.k_W * p.b_C * 2};
    uint32_t box[4] = {64, (uint32_t)kp.bw, (uint32_t)kp.bh, (uint32_t)kp.bn};
    int rc = make_tmap_bf16(&tmB, p.b, 4, dims, strides, box, nullptr, true);
    if (rc) return rc;
  }
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)nsplit);
  const double flops = p.algo_flops > 0
                           ? p.algo_flops
                           : 2.0 * p.k_N * p.k_H * p.k_W * (double)p.ntaps * p.a_cin * (double)p.n_cols;
  prof_begin(PROF_WGRAD, flops, stream);
  int rc;
  if (pair) {
    rc = BN == 128 ? launch_pair<128, 4>(tmA, tmB, kp, grid, stream)   // 4 x (32 + 16) KB
                   : launch_pair<256, 3>(tmA, tmB, kp, grid, stream);  // 3 x (32 + 32) KB
    prof_end(stream);
    return rc;
  }
  switch (BN) {
    case 64: rc = launch_t<64, 3>(tmA, tmB, kp, grid, stream); break;
    case 128: rc = launch_t<128, 3>(tmA, tmB, kp, grid, stream); break;
    default: rc = launch_t<256, 2>(tmA, tmB, kp, grid, stream); break;
  }
  prof_end(stream);
  return rc;
}

}  // namespace svsr
