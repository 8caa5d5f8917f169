// tcgen05 implicit-GEMM kernel (see igemm.cuh). Warp roles (320 threads):
//   warp 0    : TMA producer  (one elected lane)
//   warp 1    : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warps 2-5 : epilogue group 0, warps 6-9 : epilogue group 1. Each warp reads the TMEM lane quarter warp_id % 4
//               (32 accumulator rows); the two groups take the two COLUMN HALVES of the same tile, so a launch that
//               owns one tile per CTA (the encoder's small GEMMs, the K = 2048 / 4096 input gradients) drains it in
//               half the time, and a multi-tile CTA drains tile j in half the MMA time of tile j+1.
#include "igemm.cuh"
#include "tmap.h"

namespace svsr {

struct IgemmKParams {
  int tiles_h, tiles_w;
  int m_tiles, n_tiles;
  int bn, bh, bw;
  int o_N, OH, OW;
  int stride;
  int ntaps, cblocks;
  int a_coff;
  int a_box_bytes;
  int tap_dh[IGEMM_MAX_TAPS];
  int tap_dw[IGEMM_MAX_TAPS];
  int tap_kbase[IGEMM_MAX_TAPS];
  void* out;
  const float* bias;
  const void* resid;
  int out_fp32, resid_fp32;
  int ldc, c_off;
  int o_H, o_W, o_sh, o_sw, o_oh, o_ow;
  int n_cols;
  float alpha;
  float bias_scale;
  int relu;
  const void* relu_mask;
  double* bn_stats;
  float drop_p;
  unsigned long long drop_seed;
  int tma_store;  // bf16 output goes smem-staged through a TMA tensor store (full-line writes, hardware clipping)
  IgemmCe ce;     // fused cross-entropy epilogue (mode 0 = off)
  IgemmBnBwd bnb;  // fused BatchNorm-backward statistics (n = 0: off)
  StepCtl ctl;
};

// A pipeline stage holds KPS consecutive 64-wide k-blocks (A sub-tile + B sub-tile each): one mbarrier round trip
// (~200 cycles of issue-side latency) is then amortised over 4*KPS MMAs, which matters when N is small.
// CG = 2 (CTA pair, cta_group::2): a CTA stages its own 128 rows of A and HALF of the B tile.
template <int BN, int STAGES, int KPS, int CG = 1>
struct IgemmSmem {
  static constexpr int A_BYTES = 128 * 128;  // 128 pixel rows x 64 bf16 (one 128B swizzle row each)
  static constexpr int B_BYTES = BN / CG * 128;
  static constexpr int SUB_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = KPS * SUB_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  // epilogue staging tiles [128 rows x 128 B] (1024-aligned): a double buffer per epilogue group (BN = 64: the two
  // groups fill the two halves of ONE 64-column tile, so one double buffer)
  static constexpr int N_STAGING = BN >= 128 ? 4 : 2;
  static constexpr int STAGING_OFFSET = BAR_OFFSET;
  static constexpr int BAR_OFFSET2 = STAGING_OFFSET + N_STAGING * 16384;
  static constexpr int STATS_OFFSET = BAR_OFFSET2 + 256;  // fp32 [4 lane quarters][2][512] per-CTA BatchNorm partial sums
  static constexpr int VALID_OFFSET = STATS_OFFSET + 4 * 4096;  // 128 row-validity bytes of the current tile
  static constexpr int TOTAL = VALID_OFFSET + 128 + 1024;  // + alignment slack
  static_assert(TOTAL <= 232448, "exceeds 227 KB of shared memory");
  static_assert((2 * STAGES + 4) * 8 + 8 <= 256, "barrier block too small");
};

// Transposed warp reduction: every lane holds 32 values (one accumulator row); afterwards lane l holds in v[0] the
// sum over the 32 lanes of value l. 31 shuffles (halving exchange) instead of 32 x 5.
__device__ __forceinline__ void warp_transpose_reduce32(float (&v)[32], int lane) {
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

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr int IGEMM_THREADS = 320;

template <int BN, int STAGES, int KPS, bool BNB, int CG = 1>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ IgemmKParams p) {
  if (ctl_skipped(p.ctl)) return;  // sublayer dropped this step (device-resident layer_dropout mask)
  // CG = 2: the two CTAs of a (2,1,1) cluster work on M-tiles (2 t, 2 t + 1) of the same column tile as ONE M = 256 MMA:
  // each stages its own A box and half of the B tile, the leader (rank 0) issues the MMAs and multicasts the commits,
  // both drain their own 128 accumulator rows with the unchanged epilogue.
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int cta = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int ncta = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // Persistent: CTA c processes tiles c, c + gridDim.x, ... The TMA warp runs ahead across tile boundaries, the MMA
  // warp alternates between two TMEM accumulators, and the epilogue of tile j overlaps the MMAs of tile j+1.
  using L = IgemmSmem<BN, STAGES, KPS, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET2);
  uint8_t* s_stage = smem + L::STAGING_OFFSET;
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_stats = reinterpret_cast<float*>(smem + L::STATS_OFFSET);  // [4][2][512], one slice per TMEM lane quarter
  uint8_t* s_valid = smem + L::VALID_OFFSET;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.ntaps * p.cblocks;
  const int m_tiles = (p.m_tiles + CG - 1) / CG;  // (pairs of) M-tiles per column tile
  const int total_tiles = m_tiles * p.n_tiles;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 8 * CG);  // one arrive per epilogue warp (of both CTAs of a pair, on the leader's)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) tmem_alloc_2sm(tmem_ptr_smem, TMEM_COLS);
    else tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  }
  if (p.bn_stats || BNB)
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_stats[i] = 0.f;
  tcgen05_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them remotely
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cta; tile < total_tiles; tile += ncta) {
        const int nt = tile / m_tiles, mt = CG * (tile - nt * m_tiles) + rank;  // (an odd tail tile: boxes out of range = zeros)
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        const int n0 = tn * p.bn, oh0 = th * p.bh, ow0 = tw * p.bw;
        for (int kb0 = 0; kb0 < num_kb; kb0 += KPS) {
          const int nk = min(KPS, num_kb - kb0);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          // pair: the leader's barrier counts the bytes of BOTH CTAs' loads
          if (rank == 0) mbar_expect_tx(&full_bar[stage], (uint32_t)(nk * CG * (p.a_box_bytes + L::B_BYTES)));
          for (int u = 0; u < nk; ++u) {
            const int kb = kb0 + u;
            const int tap = kb / p.cblocks;
            const int cc = kb - tap * p.cblocks;
            uint8_t* sA = smem + stage * L::STAGE_BYTES + u * L::SUB_BYTES;
            uint8_t* sB = sA + L::A_BYTES;
            if (CG == 2) {
              tma_load_4d_2sm(sA, &tmA, &full_bar[stage], p.a_coff + cc * 64, ow0 * p.stride + p.tap_dw[tap],
                              oh0 * p.stride + p.tap_dh[tap], n0);
              tma_load_2d_2sm(sB, &tmB, &full_bar[stage], p.tap_kbase[tap] + cc * 64, nt * BN + rank * (BN / 2));
            } else {
              tma_load_4d(sA, &tmA, &full_bar[stage], p.a_coff + cc * 64, ow0 * p.stride + p.tap_dw[tap],
                          oh0 * p.stride + p.tap_dh[tap], n0);
              tma_load_2d(sB, &tmB, &full_bar[stage], p.tap_kbase[tap] + cc * 64, nt * BN);
            }
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
    if (lane == 0 && rank == 0) {  // pair: only the leader issues MMAs (for both CTAs' accumulators)
      constexpr uint32_t idesc = umma_idesc_bf16(128 * CG, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int j = 0;  // local tile counter
      for (int tile = cta; tile < total_tiles; tile += ncta, ++j) {
        const int acc = j & 1;
        const uint32_t use = (uint32_t)(j >> 1);
        mbar_wait(&tmem_empty_bar[acc], (use & 1) ^ 1);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb0 = 0; kb0 < num_kb; kb0 += KPS) {
          const int nk = min(KPS, num_kb - kb0);
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          for (int u = 0; u < nk; ++u) {
            const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES + u * L::SUB_BYTES);
            const uint32_t b_addr = a_addr + L::A_BYTES;
            const uint64_t a_desc = umma_smem_desc_sw128(a_addr, 16, 1024);
            const uint64_t b_desc = umma_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128B swizzle row: +2 in (addr >> 4) units
              if (CG == 2) umma_bf16_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb0 | u | k) != 0);
              else umma_bf16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb0 | u | k) != 0);
            }
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above retire
          if (CG == 2) umma_commit_2sm(&empty_bar[stage]);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (CG == 2) umma_commit_2sm(&tmem_full_bar[acc]);
        else umma_commit(&tmem_full_bar[acc]);
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    constexpr bool SHARED_TILE = (BN == 64);          // both groups fill one 64-column staging tile
    constexpr int CPG = SHARED_TILE ? 1 : BN / 64;    // 32-column chunks per epilogue group and tile
    constexpr int NPAIR = SHARED_TILE ? 1 : BN / 128;  // 64-column staging tiles per group and tile
    const int grp = (warp - 2) >> 2;                  // epilogue group = column half
    const int ewg = (warp - 2) & 3;                   // warp index inside the group (row block of the staged tile)
    const int q = warp & 3;                           // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;                      // accumulator row == pixel slot in the box
    const int bar_id = SHARED_TILE ? 1 : 1 + grp;
    const int bar_n = SHARED_TILE ? 256 : 128;
    const bool leader = SHARED_TILE ? (threadIdx.x == 64) : (threadIdx.x == 64 + 128 * grp);  // issues the TMA stores
    const int hw = p.bh * p.bw;
    const int dn = r / hw;
    const int rem = r - dn * hw;
    const int dh = rem / p.bw;
    const int dw = rem - dh * p.bw;
    int j = 0;
    int gcount = 0;
    float st_sum[NPAIR][2], st_sq[NPAIR][2];  // per-thread BatchNorm partial sums (smem-staged path)
#pragma unroll
    for (int g2 = 0; g2 < NPAIR; ++g2) st_sum[g2][0] = st_sum[g2][1] = st_sq[g2][0] = st_sq[g2][1] = 0.f;
    int stat_nt = -1;
    const bool do_staged_stats = p.bn_stats && p.tma_store && (!SHARED_TILE || grp == 0);
    float ce_scale = 1.f;
    if (p.ce.mode == 2) ce_scale = p.ce.dscale * (p.ce.grad_scale ? __ldg(p.ce.grad_scale) : 1.f);
    for (int tile = cta; tile < total_tiles; tile += ncta, ++j) {
      const int nt = tile / m_tiles, mt = CG * (tile - nt * m_tiles) + rank;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int n0 = tn * p.bn, oh0 = th * p.bh, ow0 = tw * p.bw;
      const int n = n0 + dn, oh = oh0 + dh, ow = ow0 + dw;
      const bool row_valid = (r < p.bn * hw) && (n < p.o_N) && (oh < p.OH) && (ow < p.OW);
      const long long pix =
          ((long long)n * p.o_H + (long long)oh * p.o_sh + p.o_oh) * p.o_W + (long long)ow * p.o_sw + p.o_ow;
      const long long row_off = pix * p.ldc + p.c_off;
      const int acc = j & 1;
      const bool tile_all_valid = (n0 + p.bn <= p.o_N) && (oh0 + p.bh <= p.OH) && (ow0 + p.bw <= p.OW);
      if (p.bn_stats && p.tma_store && !tile_all_valid) s_valid[r] = row_valid ? 1 : 0;  // read after the group barriers
      // fused cross-entropy: this row is frame (b, t) of the clip batch
      int ce_b = 0, ce_t = 0;
      float ce_m = -INFINITY, ce_s = 0.f;
      if (p.ce.mode) ce_b = n / p.ce.T, ce_t = n - ce_b * p.ce.T;
      mbar_wait(&tmem_full_bar[acc], (uint32_t)((j >> 1) & 1));
      tcgen05_fence_after();

#pragma unroll 1
      for (int cc = 0; cc < CPG; ++cc) {
        const int ch = grp * CPG + cc;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + ch * 32), v);
        tmem_ld_wait();
        const int col0 = nt * BN + ch * 32;
        if (p.bn_stats && !p.tma_store) {  // register path (fp32 outputs): transposed warp reduction of the chunk
          float a[32], b[32];
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            a[jj] = row_valid ? __uint_as_float(v[jj]) : 0.f;
            b[jj] = a[jj] * a[jj];
          }
          warp_transpose_reduce32(a, lane);
          warp_transpose_reduce32(b, lane);
          if (col0 + lane < p.n_cols) {  // slice per lane quarter; the two groups own disjoint columns
            s_stats[q * 1024 + col0 + lane] += a[0];
            s_stats[q * 1024 + 512 + col0 + lane] += b[0];
          }
        }
        if (!p.tma_store && (!row_valid || col0 >= p.n_cols)) continue;
        float f[32];
  #pragma unroll
        for (int jj = 0; jj < 32; ++jj) f[jj] = __uint_as_float(v[jj]) * p.alpha;
        if (p.bias) {
  #pragma unroll
          for (int jj = 0; jj < 32; ++jj)
            if (col0 + jj < p.n_cols) f[jj] = fmaf(__ldg(p.bias + col0 + jj), p.bias_scale, f[jj]);
        }
        const bool full_chunk = (col0 + 32 <= p.n_cols);
        if (p.ce.mode) {
          // ---- fused projection + reshape + log-softmax + NLL (lightning.py:168-171, e2e_asr_transformer.py:198-201) ----
          // columns [c*V, (c+1)*V) of this row are the V logits of target tokens[b, t*A + a, g], c = a*G + g (int64,
          // bit-exact index arithmetic); V % 64 == 0, so a 64-column slot never straddles two softmaxes.
          const int c = col0 / p.ce.V;
          const int a = c / p.ce.G, g = c - a * p.ce.G;
          long long tgt = -1;
          if (row_valid && col0 < p.n_cols)
            tgt = p.ce.tokens[(long long)ce_b * p.ce.tok_stride_b + (long long)(ce_t * p.ce.A + a) * p.ce.G + g];
          const bool bad = tgt < 0 || tgt >= p.ce.V;
          const int tj = bad ? -1 : (int)tgt - (col0 - c * p.ce.V);  // position of the target inside this chunk (if any)
          if (p.ce.mode == 1) {
            if (row_valid && col0 < p.n_cols) {
              if (bad) *p.ce.bad_token = 1;
              float m = f[0];
#pragma unroll
              for (int jj = 1; jj < 32; ++jj) m = fmaxf(m, f[jj]);
              float se = 0.f, xt = 0.f;
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) {
                se += exp2f((f[jj] - m) * 1.4426950408889634f);
                if (jj == tj) xt = f[jj];
              }
              const long long rc = (long long)n * p.ce.AG + c;
              if (tj >= 0 && tj < 32) p.ce.xt[rc] = xt;  // exactly one chunk of the row holds the target logit
              // merge the two 32-column chunks of a 64-column slot in registers, one (max, sum) partial per slot
              if ((ch & 1) == 0) {
                ce_m = m, ce_s = se;
              } else {
                const float mm = fmaxf(ce_m, m);
                const float ss = ce_s * exp2f((ce_m - mm) * 1.4426950408889634f) + se * exp2f((m - mm) * 1.4426950408889634f);
                p.ce.part[(long long)n * (p.n_cols >> 6) + (col0 >> 6)] = make_float2(mm, ss);
              }
            }
            continue;  // forward: nothing is stored but the partials -- the logits never leave the SM
          }
          // backward: d logits = (softmax - onehot) * dscale from the recomputed tile and the saved log-sum-exp
          float lse = 0.f;
          if (row_valid && col0 < p.n_cols) lse = p.ce.lse[(long long)n * p.ce.AG + c];
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const float pr = exp2f((f[jj] - lse) * 1.4426950408889634f);
            f[jj] = bad ? 0.f : (pr - (jj == tj ? 1.f : 0.f)) * ce_scale;
          }
        }
        if (p.drop_p > 0.f && row_valid) {
          const float ks = 1.0f / (1.0f - p.drop_p);
          const unsigned long long e0 = (unsigned long long)(row_off + col0);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) f[jj] = dropout_keep(ctl_seed(p.ctl, p.drop_seed), e0 + jj, p.drop_p) ? f[jj] * ks : 0.f;
        }
        if (p.resid && row_valid) {
          if (p.resid_fp32) {
            const float* rp = reinterpret_cast<const float*>(p.resid) + row_off + col0;
            if (full_chunk) {
  #pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                float4 t = reinterpret_cast<const float4*>(rp)[jj];
                f[4 * jj] += t.x, f[4 * jj + 1] += t.y, f[4 * jj + 2] += t.z, f[4 * jj + 3] += t.w;
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (col0 + jj < p.n_cols) f[jj] += rp[jj];
            }
          } else {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.resid) + row_off + col0;
            if (full_chunk) {
  #pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                uint4 t = reinterpret_cast<const uint4*>(rp)[jj];
                float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
                f[8 * jj] += a.x, f[8 * jj + 1] += a.y, f[8 * jj + 2] += b.x, f[8 * jj + 3] += b.y;
                f[8 * jj + 4] += c.x, f[8 * jj + 5] += c.y, f[8 * jj + 6] += d.x, f[8 * jj + 7] += d.y;
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj)
                if (col0 + jj < p.n_cols) f[jj] += __bfloat162float(rp[jj]);
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) f[jj] = fmaxf(f[jj], 0.f);
        }
        if (p.relu_mask && row_valid) {
          const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(p.relu_mask) + row_off + col0;
          if (full_chunk) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const uint4 t = reinterpret_cast<const uint4*>(mp)[jj];
              const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
              f[8 * jj] = a.x > 0.f ? f[8 * jj] : 0.f, f[8 * jj + 1] = a.y > 0.f ? f[8 * jj + 1] : 0.f;
              f[8 * jj + 2] = b.x > 0.f ? f[8 * jj + 2] : 0.f, f[8 * jj + 3] = b.y > 0.f ? f[8 * jj + 3] : 0.f;
              f[8 * jj + 4] = c.x > 0.f ? f[8 * jj + 4] : 0.f, f[8 * jj + 5] = c.y > 0.f ? f[8 * jj + 5] : 0.f;
              f[8 * jj + 6] = d.x > 0.f ? f[8 * jj + 6] : 0.f, f[8 * jj + 7] = d.y > 0.f ? f[8 * jj + 7] : 0.f;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (col0 + jj < p.n_cols) f[jj] = __bfloat162float(mp[jj]) > 0.f ? f[jj] : 0.f;
          }
        }
        if constexpr (BNB) {
          // ---- fused BatchNorm-backward statistics (IgemmBnBwd): mask, then per-channel sums over the tile's 128 rows ----
          const bool live = row_valid && full_chunk;
          const int sstride = p.bnb.n == 2 ? 256 : 512;
          float gq[32];
#pragma unroll 1
          for (int bi = 0; bi < p.bnb.n; ++bi) {
            float cv[32];
            if (live) {
              const uint4* cp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bnb.c[bi]) + row_off + col0);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const uint4 t = __ldg(cp + jj);
                const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
                cv[8 * jj] = a.x, cv[8 * jj + 1] = a.y, cv[8 * jj + 2] = b.x, cv[8 * jj + 3] = b.y;
                cv[8 * jj + 4] = c.x, cv[8 * jj + 5] = c.y, cv[8 * jj + 6] = d.x, cv[8 * jj + 7] = d.y;
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) cv[jj] = 0.f;
            }
            const float* cf = p.bnb.coef[bi];
            if (bi == 0) {
              if (p.bnb.self_mask) {  // the ReLU that follows BN 0: its mask is the sign of c*scale + shift
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const float4 sc = __ldg(reinterpret_cast<const float4*>(cf + 2 * p.n_cols + col0) + jj);
                  const float4 sh = __ldg(reinterpret_cast<const float4*>(cf + 3 * p.n_cols + col0) + jj);
                  f[4 * jj] = fmaf(cv[4 * jj], sc.x, sh.x) > 0.f ? f[4 * jj] : 0.f;
                  f[4 * jj + 1] = fmaf(cv[4 * jj + 1], sc.y, sh.y) > 0.f ? f[4 * jj + 1] : 0.f;
                  f[4 * jj + 2] = fmaf(cv[4 * jj + 2], sc.z, sh.z) > 0.f ? f[4 * jj + 2] : 0.f;
                  f[4 * jj + 3] = fmaf(cv[4 * jj + 3], sc.w, sh.w) > 0.f ? f[4 * jj + 3] : 0.f;
                }
              }
              // the statistics are those of the STORED (bf16-rounded) gradient, exactly what bn_bwd_apply reads back
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) gq[jj] = live ? __bfloat162float(__float2bfloat16(f[jj])) : 0.f;
              float a[32];
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) a[jj] = gq[jj];
              warp_transpose_reduce32(a, lane);
              if (col0 + lane < p.n_cols) s_stats[q * 1024 + col0 + lane] += a[0];
            }
            float b[32];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const float4 mu = __ldg(reinterpret_cast<const float4*>(cf + col0) + jj);
              b[4 * jj] = gq[4 * jj] * (cv[4 * jj] - mu.x), b[4 * jj + 1] = gq[4 * jj + 1] * (cv[4 * jj + 1] - mu.y);
              b[4 * jj + 2] = gq[4 * jj + 2] * (cv[4 * jj + 2] - mu.z), b[4 * jj + 3] = gq[4 * jj + 3] * (cv[4 * jj + 3] - mu.w);
            }
            warp_transpose_reduce32(b, lane);
            if (col0 + lane < p.n_cols) s_stats[q * 1024 + (1 + bi) * sstride + col0 + lane] += b[0];
          }
        }
        if (p.out_fp32) {
          float* op = reinterpret_cast<float*>(p.out) + row_off + col0;
          if (full_chunk) {
  #pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              reinterpret_cast<float4*>(op)[jj] = make_float4(f[4 * jj], f[4 * jj + 1], f[4 * jj + 2], f[4 * jj + 3]);
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (col0 + jj < p.n_cols) op[jj] = f[jj];
          }
        } else if (p.tma_store) {
          // stage 64 output channels (two 32-column chunks) per pixel row in a SWIZZLE_128B tile, then one TMA store.
          // BN >= 128: the group owns whole 64-column tiles (its own double buffer and 128-thread barrier);
          // BN == 64: group g writes half g of the CTA's single tile (256-thread barrier, thread 64 stores).
          const int half = SHARED_TILE ? grp : (cc & 1);
          const bool first_half = SHARED_TILE || (cc & 1) == 0;
          const bool last_half = SHARED_TILE || (cc & 1) == 1;
          const int pair = SHARED_TILE ? 0 : (cc >> 1);
          const int bufi = SHARED_TILE ? (gcount & 1) : (grp * 2 + ((gcount + pair) & 1));
          uint8_t* stg = s_stage + bufi * 16384;
          if (first_half) {
            if (leader) tma_store_wait_read<1>();  // the store that last used this buffer has read it
            named_bar_sync(bar_id, bar_n);
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            uint4 t;
            t.x = pack_bf16x2(f[8 * jj], f[8 * jj + 1]);
            t.y = pack_bf16x2(f[8 * jj + 2], f[8 * jj + 3]);
            t.z = pack_bf16x2(f[8 * jj + 4], f[8 * jj + 5]);
            t.w = pack_bf16x2(f[8 * jj + 6], f[8 * jj + 7]);
            const int chunk = (half * 4 + jj) ^ (r & 7);  // 16-byte chunk position after the 128B swizzle
            *reinterpret_cast<uint4*>(stg + r * 128 + chunk * 16) = t;
          }
          if (last_half) {
            fence_proxy_async_smem();
            named_bar_sync(bar_id, bar_n);
            const int cgrp = SHARED_TILE ? 0 : (grp * NPAIR + pair);  // 64-column group of the tile
            if (leader) {
              tma_store_4d(&tmC, stg, p.c_off + nt * BN + cgrp * 64, ow0 * p.o_sw + p.o_ow, oh0 * p.o_sh + p.o_oh, n0);
              tma_store_commit();
            }
            if (do_staged_stats) {
              // fused BatchNorm statistics from the staged (bf16-rounded = exactly what BN will normalise) tile: lane l of
              // group warp w owns the column pair (2l, 2l+1) over rows [32w, 32w+32): 32 conflict-free LDS.32 per
              // 64-column group; partial sums stay in registers across all tiles of this CTA.
              const int rows_in_box = p.bn * hw;
              const int rbeg = ewg * 32, rend = min(rbeg + 32, rows_in_box);
              float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
              const uint8_t* colbase = stg + (lane & 3) * 4;
              const int cpos = lane >> 2;
              if (tile_all_valid) {
#pragma unroll 8
                for (int rr = rbeg; rr < rend; ++rr) {
                  const float2 v2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(colbase + rr * 128 + ((cpos ^ (rr & 7)) << 4)));
                  s1a += v2.x, s1b += v2.y;
                  s2a = fmaf(v2.x, v2.x, s2a), s2b = fmaf(v2.y, v2.y, s2b);
                }
              } else {
                for (int rr = rbeg; rr < rend; ++rr) {
                  if (!s_valid[rr]) continue;
                  const float2 v2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(colbase + rr * 128 + ((cpos ^ (rr & 7)) << 4)));
                  s1a += v2.x, s1b += v2.y;
                  s2a = fmaf(v2.x, v2.x, s2a), s2b = fmaf(v2.y, v2.y, s2b);
                }
              }
              if (nt != stat_nt) {  // switched to another column tile: spill the register partials first
                if (stat_nt >= 0) {
#pragma unroll
                  for (int g2 = 0; g2 < NPAIR; ++g2) {
                    float* sl = s_stats + ewg * 1024 + stat_nt * BN + (SHARED_TILE ? 0 : (grp * NPAIR + g2) * 64) + 2 * lane;
                    sl[0] += st_sum[g2][0], sl[1] += st_sum[g2][1];
                    sl[512] += st_sq[g2][0], sl[513] += st_sq[g2][1];
                    st_sum[g2][0] = st_sum[g2][1] = st_sq[g2][0] = st_sq[g2][1] = 0.f;
                  }
                }
                stat_nt = nt;
              }
#pragma unroll
              for (int g2 = 0; g2 < NPAIR; ++g2)
                if (g2 == pair) st_sum[g2][0] += s1a, st_sum[g2][1] += s1b, st_sq[g2][0] += s2a, st_sq[g2][1] += s2b;
            }
          }
        } else {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + col0;
          if (full_chunk) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              uint4 t;
              t.x = pack_bf16x2(f[8 * jj], f[8 * jj + 1]);
              t.y = pack_bf16x2(f[8 * jj + 2], f[8 * jj + 3]);
              t.z = pack_bf16x2(f[8 * jj + 4], f[8 * jj + 5]);
              t.w = pack_bf16x2(f[8 * jj + 6], f[8 * jj + 7]);
              reinterpret_cast<uint4*>(op)[jj] = t;
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (col0 + jj < p.n_cols) op[jj] = __float2bfloat16(f[jj]);
          }
        }
      }
      gcount += NPAIR;
      // all tcgen05.ld of this warp have completed (wait::ld above): hand the accumulator back to the MMA warp
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_leader(&tmem_empty_bar[acc]);
        else mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
    if (p.tma_store && leader) tma_store_wait_all();  // smem must outlive the last bulk store
    if (do_staged_stats && stat_nt >= 0) {
#pragma unroll
      for (int g2 = 0; g2 < NPAIR; ++g2) {
        float* sl = s_stats + ewg * 1024 + stat_nt * BN + (SHARED_TILE ? 0 : (grp * NPAIR + g2) * 64) + 2 * lane;
        sl[0] += st_sum[g2][0], sl[1] += st_sum[g2][1];
        sl[512] += st_sq[g2][0], sl[513] += st_sq[g2][1];
      }
    }
    if constexpr (BNB) {  // fused BatchNorm-backward sums: sum g is shared, sum g*(c - mean) gets its invstd here
      named_bar_sync(3, 256);
      const int sstride = p.bnb.n == 2 ? 256 : 512;
      for (int cidx = threadIdx.x - 64; cidx < p.n_cols; cidx += 256) {
        const double sg = (double)s_stats[cidx] + (double)s_stats[1024 + cidx] + (double)s_stats[2048 + cidx] +
                          (double)s_stats[3072 + cidx];
        for (int bi = 0; bi < p.bnb.n; ++bi) {
          const int o = (1 + bi) * sstride + cidx;
          const double sx = (double)s_stats[o] + (double)s_stats[1024 + o] + (double)s_stats[2048 + o] +
                            (double)s_stats[3072 + o];
          atomicAdd(p.bnb.stats[bi] + cidx, sg);
          atomicAdd(p.bnb.stats[bi] + p.n_cols + cidx, sx * (double)__ldg(p.bnb.coef[bi] + p.n_cols + cidx));
        }
      }
    }
    if (p.bn_stats) {  // flush this CTA's partial sums once (fp64 across CTAs)
      named_bar_sync(3, 256);  // the eight epilogue warps only
      const int t = threadIdx.x - 64;
      for (int cidx = t; cidx < p.n_cols; cidx += 256) {
        const double su = (double)s_stats[cidx] + (double)s_stats[1024 + cidx] + (double)s_stats[2048 + cidx] +
                          (double)s_stats[3072 + cidx];
        const double sq2 = (double)s_stats[512 + cidx] + (double)s_stats[1536 + cidx] + (double)s_stats[2560 + cidx] +
                           (double)s_stats[3584 + cidx];
        atomicAdd(p.bn_stats + cidx, su);
        atomicAdd(p.bn_stats + p.n_cols + cidx, sq2);
      }
    }
  }

  tcgen05_fence_before();
  if (CG == 2) {
    cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the peer's MMAs / remote arrives can still touch it
    if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
void igemm_choose_box(int o_N, int OH, int OW, int* pbn, int* pbh, int* pbw) {
  long long best_tiles = -1;
  int best[3] = {1, 1, 1};
  for (int bw = 1; bw <= OW && bw <= 128; ++bw) {
    for (int bh = 1; bh <= OH && bh * bw <= 128; ++bh) {
      int bn = 1;
      if (bh == OH && bw == OW) bn = 128 / (OH * OW);
      if (bn > o_N) bn = o_N;
      if (bn > 128) bn = 128;
      if (bn < 1) bn = 1;
      long long tiles = (long long)((o_N + bn - 1) / bn) * ((OH + bh - 1) / bh) * ((OW + bw - 1) / bw);
      if (best_tiles < 0 || tiles < best_tiles) {
        best_tiles = tiles;
        best[0] = bn, best[1] = bh, best[2] = bw;
      }
    }
  }
  *pbn = best[0], *pbh = best[1], *pbw = best[2];
}

template <int BN, int STAGES, int KPS, bool BNB>
static int launch_tb(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const IgemmKParams& kp,
                     dim3 grid, cudaStream_t stream) {
  using L = IgemmSmem<BN, STAGES, KPS>;
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(igemm_kernel<BN, STAGES, KPS, BNB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL));
    attr_done = true;
  }
  igemm_kernel<BN, STAGES, KPS, BNB><<<grid, IGEMM_THREADS, L::TOTAL, stream>>>(tmA, tmB, tmC, kp);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}
// CTA pairs (cta_group::2): a (2,1,1) cluster per pair of M-tiles, `pairs` clusters
template <int BN, int STAGES, int KPS>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const IgemmKParams& kp,
                       int pairs, cudaStream_t stream) {
  using L = IgemmSmem<BN, STAGES, KPS, 2>;
  auto kernel = igemm_kernel<BN, STAGES, KPS, false, 2>;
  static bool attr_done = false;
  if (!attr_done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs)), cfg.blockDim = dim3(IGEMM_THREADS), cfg.dynamicSmemBytes = L::TOTAL, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  SVSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, tmC, kp));
  note_launch();
  return SVSR_OK;
}
// (the fused BatchNorm-backward statistics are a separate instantiation: the common launches carry none of their registers)
template <int BN, int STAGES, int KPS>
static int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const IgemmKParams& kp,
                    dim3 grid, cudaStream_t stream) {
  return kp.bnb.n ? launch_tb<BN, STAGES, KPS, true>(tmA, tmB, tmC, kp, grid, stream)
                  : launch_tb<BN, STAGES, KPS, false>(tmA, tmB, tmC, kp, grid, stream);
}

int igemm_launch(const IgemmProblem& p, cudaStream_t stream) {
  SVSR_REQUIRE(p.a && p.b && (p.out || p.ce.mode == 1), "igemm: null operand");
  if (p.ce.mode) {
    SVSR_REQUIRE(p.ce.mode == 1 || p.ce.mode == 2, "igemm: ce.mode %d unknown", p.ce.mode);
    SVSR_REQUIRE(p.ntaps == 1 && p.OH == 1 && p.OW == 1 && p.o_H == 1 && p.o_W == 1 && p.stride == 1,
                 "igemm: the fused cross-entropy epilogue needs a dense GEMM");
    SVSR_REQUIRE(p.ce.V > 0 && p.ce.V % 64 == 0 && p.ce.AG == p.ce.A * p.ce.G && p.b_rows == p.ce.AG * p.ce.V &&
                     p.b_rows >= 128 &&
                     p.ce.T > 0 && p.o_N % p.ce.T == 0,
                 "igemm: fused cross-entropy geometry (V=%d must be a multiple of 64, N=%d = A*G*V, rows %d = B*T)", p.ce.V,
                 p.b_rows, p.o_N);
    SVSR_REQUIRE(p.ce.tokens && p.ce.bad_token && (p.ce.mode == 1 ? (p.ce.part && p.ce.xt) : (p.ce.lse && !p.out_fp32)),
                 "igemm: fused cross-entropy buffers missing");
  }
  SVSR_REQUIRE(p.cin > 0 && p.cin % 64 == 0, "igemm: cin=%d must be a positive multiple of 64", p.cin);
  SVSR_REQUIRE(p.a_C % 8 == 0 && p.a_coff % 8 == 0, "igemm: A channel pitch/offset must be multiples of 8");
  SVSR_REQUIRE(p.b_cols % 8 == 0, "igemm: B pitch %d must be a multiple of 8", p.b_cols);
  SVSR_REQUIRE(p.ntaps >= 1 && p.ntaps <= IGEMM_MAX_TAPS, "igemm: ntaps=%d out of range", p.ntaps);
  SVSR_REQUIRE(p.ldc % 8 == 0 && p.c_off % 8 == 0, "igemm: output pitch/offset must be multiples of 8");
  SVSR_REQUIRE(p.stride == 1 || p.stride == 2, "igemm: stride must be 1 or 2");
  if (igemm_halo_matches(p)) return igemm_halo_launch(p, stream);
  if (igemm_stem_matches(p)) return igemm_stem_launch(p, stream);

  IgemmKParams kp{};
  igemm_choose_box(p.o_N, p.OH, p.OW, &kp.bn, &kp.bh, &kp.bw);
  kp.tiles_h = (p.OH + kp.bh - 1) / kp.bh;
  kp.tiles_w = (p.OW + kp.bw - 1) / kp.bw;
  const int tiles_n = (p.o_N + kp.bn - 1) / kp.bn;
  kp.o_N = p.o_N, kp.OH = p.OH, kp.OW = p.OW;
  kp.stride = p.stride;
  kp.ntaps = p.ntaps;
  kp.cblocks = p.cin / 64;
  kp.a_coff = p.a_coff;
  kp.a_box_bytes = kp.bn * kp.bh * kp.bw * 128;
  for (int t = 0; t < p.ntaps; ++t) {
    kp.tap_dh[t] = p.tap_dh[t];
    kp.tap_dw[t] = p.tap_dw[t];
    kp.tap_kbase[t] = p.tap_kbase[t];
  }
  kp.out = p.out, kp.bias = p.bias, kp.resid = p.resid;
  kp.out_fp32 = p.out_fp32, kp.resid_fp32 = p.resid_fp32;
  kp.ldc = p.ldc, kp.c_off = p.c_off;
  kp.o_H = p.o_H, kp.o_W = p.o_W, kp.o_sh = p.o_sh, kp.o_sw = p.o_sw, kp.o_oh = p.o_oh, kp.o_ow = p.o_ow;
  kp.n_cols = p.b_rows;
  kp.alpha = p.alpha;
  kp.bias_scale = p.bias_scale, kp.relu = p.relu, kp.relu_mask = p.relu_mask;
  kp.bn_stats = p.bn_stats;
  kp.drop_p = p.drop_p, kp.drop_seed = p.drop_seed;
  SVSR_REQUIRE(!p.bn_stats || p.b_rows <= 512, "igemm: fused BN statistics support at most 512 output channels");

  // A: [a_N, a_H, a_W, a_C] NHWC, box = (64 ch, bw, bh, bn) pixels, traversal stride for strided convs.
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)p.a_C, (uint64_t)p.a_W, (uint64_t)p.a_H, (uint64_t)p.a_N};
    uint64_t strides[3] = {(uint64_t)p.a_C * 2, (uint64_t)p.a_W * p.a_C * 2, (uint64_t)p.a_H * p.a_W * p.a_C * 2};
    uint32_t box[4] = {64, (uint32_t)((kp.bw - 1) * p.stride + 1), (uint32_t)((kp.bh - 1) * p.stride + 1),
                       (uint32_t)kp.bn};
    uint32_t es[4] = {1, (uint32_t)p.stride, (uint32_t)p.stride, 1};
    int rc = make_tmap_bf16(&tmA, p.a, 4, dims, strides, box, es, true);
    if (rc) return rc;
  }
  // N tile: the widest that still gives ~one wave of CTAs (small-M GEMMs of the encoder are latency bound)
  const long long m_tiles = (long long)tiles_n * kp.tiles_h * kp.tiles_w;
  int BN = p.b_rows <= 64 ? 64 : (p.b_rows <= 128 ? 128 : 256);
  if (BN == 256 && m_tiles * ((p.b_rows + 255) / 256) < 100) BN = 128;
  if (BN == 128 && m_tiles * ((p.b_rows + 127) / 128) < 100 && !p.ce.mode) BN = 64;
  if (const char* force = getenv("SVSR_IGEMM_BN")) {  // measurement only (tools/gemm_bench.py): override the N tile
    const int f = atoi(force);
    if ((f == 64 || f == 128 || f == 256) && !(p.ce.mode && f == 64)) BN = f;
  }
  // CTA pairs (cta_group::2, M = 256 MMAs; SVSR_IGEMM_2CTA=0 turns them off): each CTA of a pair stages half of the B tile
  static const bool pair_env = [] {
    const char* e = getenv("SVSR_IGEMM_2CTA");
    return !(e && e[0] == '0');
  }();
  const bool pair = pair_env && BN >= 128 && !p.bnb.n && m_tiles >= 2;  // (the fused BN-backward variant keeps single CTAs)
  {
    uint64_t dims[2] = {(uint64_t)p.b_cols, (uint64_t)p.b_rows};
    uint64_t strides[1] = {(uint64_t)p.b_cols * 2};
    uint32_t box[2] = {64, (uint32_t)(pair ? BN / 2 : BN)};
    int rc = make_tmap_bf16(&tmB, p.b, 2, dims, strides, box, nullptr, true);
    if (rc) return rc;
  }
  kp.m_tiles = (int)m_tiles;
  kp.n_tiles = (p.b_rows + BN - 1) / BN;
  const long long total_tiles = m_tiles * kp.n_tiles;
  dim3 grid((unsigned)(total_tiles < 148 ? total_tiles : 148));
  // bf16 outputs whose channel count is a multiple of 64 leave through TMA tensor stores (the map mirrors the output
  // geometry: pitch ldc, pixel strides o_s*, so hardware clips pixels outside the tensor)
  CUtensorMap tmC = tmA;
  kp.tma_store = (!p.out_fp32 && p.b_rows % 64 == 0 && p.ce.mode != 1) ? 1 : 0;
  kp.ce = p.ce;
  kp.ctl = p.ctl;
  kp.bnb = p.bnb;
  if (p.bnb.n) {
    SVSR_REQUIRE(p.bnb.n == 1 || p.bnb.n == 2, "igemm: bnb.n = %d", p.bnb.n);
    SVSR_REQUIRE(kp.tma_store && !p.bn_stats && p.c_off == 0 && p.b_rows == p.ldc && p.b_rows <= (p.bnb.n == 2 ? 256 : 512),
                 "igemm: fused BatchNorm-backward statistics need a bf16 output of %d..%d channels (multiple of 64) that "
                 "fills the pixel", 64, p.bnb.n == 2 ? 256 : 512);
    for (int i = 0; i < p.bnb.n; ++i)
      SVSR_REQUIRE(p.bnb.c[i] && p.bnb.coef[i] && p.bnb.stats[i], "igemm: bnb buffers of BatchNorm %d missing", i);
  }
  if (kp.tma_store) {
    uint64_t dims[4] = {(uint64_t)p.ldc, (uint64_t)p.o_W, (uint64_t)p.o_H, (uint64_t)p.o_N};
    uint64_t strides[3] = {(uint64_t)p.ldc * 2, (uint64_t)p.o_W * p.ldc * 2, (uint64_t)p.o_H * p.o_W * p.ldc * 2};
    uint32_t box[4] = {64, (uint32_t)((kp.bw - 1) * p.o_sw + 1), (uint32_t)((kp.bh - 1) * p.o_sh + 1), (uint32_t)kp.bn};
    uint32_t es[4] = {1, (uint32_t)p.o_sw, (uint32_t)p.o_sh, 1};
    int rc2 = make_tmap_bf16(&tmC, p.out, 4, dims, strides, box, es, true);
    if (rc2) return rc2;
  }
  const double flops = p.algo_flops > 0 ? p.algo_flops
                                        : 2.0 * p.o_N * p.OH * p.OW * (double)p.b_rows * p.ntaps * p.cin;
  prof_begin(PROF_IGEMM, flops, stream);
  int rc;
  if (pair) {
    const long long pair_tiles = (m_tiles + 1) / 2 * kp.n_tiles;
    const int pairs = (int)(pair_tiles < 74 ? pair_tiles : 74);
    rc = BN == 128 ? launch_pair<128, 3, 2>(tmA, tmB, tmC, kp, pairs, stream)   // 3 x 2 x (16 + 8) KB
                   : launch_pair<256, 4, 1>(tmA, tmB, tmC, kp, pairs, stream);  // 4 x (16 + 16) KB
    prof_end(stream);
    return rc;
  }
  switch (BN) {
    case 64: rc = launch_t<64, 3, 2>(tmA, tmB, tmC, kp, grid, stream); break;    // 3 x 48 KB
    case 128: rc = launch_t<128, 2, 2>(tmA, tmB, tmC, kp, grid, stream); break;  // 2 x 64 KB
    default: rc = launch_t<256, 3, 1>(tmA, tmB, tmC, kp, grid, stream); break;   // 3 x 48 KB
  }
  prof_end(stream);
  return rc;
}

}  // namespace svsr
