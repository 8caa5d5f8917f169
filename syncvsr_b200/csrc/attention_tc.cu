// x-transformers attention core on the tensor cores (reference: x_transformers.Encoder sublayer "a" called at
// LRW/video/src/lightning.py:95-105,158 -- SURVEY.md Appendix A: rotary on the first 32 dims of q, k (and v), scale 64^-1/2,
// fp32 softmax, attn_dropout on the probabilities): n <= 64 tokens, 64-dim heads.
//
// The sequence is tiny (n = 30 at T = 29), so ONE 128-row UMMA tile carries PPT = 128 / BS (batch, head) pairs, each
// padded to BS = 32 (or 64) rows: S = Q' K'^T is a single [128 x 128] tcgen05 product whose diagonal BS x BS blocks are
// the pairs' score matrices (the off-diagonal blocks are never read), P is written back block-diagonally (zeros
// elsewhere) and O = P V' is a second [128 x 64] product. Thread r of the 128-thread CTA owns tile row r = TMEM lane r:
// it loads its q / k / v row (16-byte loads), applies the rotation in registers, stores the bf16 row into the
// 128-byte-swizzled operand tile, and later reads its own score row from TMEM -- softmax needs no shuffles.
// The same shared-memory bytes serve as K-major operands (rows = M/N index, 64 K elements per 128-byte row) and as
// MN-major operands (rows = K index), so the backward's five products
//     S = Q'K'^T   dP = dO V'^T   dV' = P^T dO   dQ' = dS K'   dK' = dS^T Q'
// need no transposed copies: 4 + 4 + 8 + 8 + 8 MMAs on tiles that are written once.
#include "encoder.cuh"
#include "attention_tc.cuh"
#include "tmap.h"

namespace svsr {
namespace {
using namespace attn_tc;

// softmax of this thread's score row (BS columns of its pair's block, columns >= n masked); p[] = probabilities
template <int BS>
__device__ __forceinline__ void softmax_row(const float (&s)[BS], int n, float (&p)[BS]) {
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < BS; ++j)
    if (j < n) m = fmaxf(m, s[j]);
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < BS; ++j) {
    p[j] = j < n ? exp2f((s[j] - m) * (0.125f * LOG2E)) : 0.f;  // scale 64^-1/2 folded in: max commutes with it
    sum += p[j];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int j = 0; j < BS; ++j) p[j] *= inv;
}

// bf16 row of a block-diagonal [128 x 128] two-block tile: this row's BS columns start at column blk * BS
template <int BS>
__device__ __forceinline__ void store_block_row(uint8_t* tile2, int r, int blk, const float (&x)[BS]) {
#pragma unroll
  for (int c = 0; c < BS / 8; ++c) {
    const int cg = blk * (BS / 8) + c;  // 16-byte chunk index inside the 256-byte logical row
    *reinterpret_cast<uint4*>(sw_chunk(tile2 + (cg >> 3) * TILE_BYTES, r, cg & 7)) = pack8(&x[8 * c]);
  }
}

struct AttnTcParams {
  const __nv_bfloat16* qkv;
  const float* rot;
  const __nv_bfloat16* d_o;
  __nv_bfloat16* o;     // forward output / unused in backward
  __nv_bfloat16* dqkv;  // backward output
  int n, heads, pairs, rotary_v;
  float drop_p;
  unsigned long long drop_seed;
  StepCtl ctl;
};

template <int BS>
__global__ void __launch_bounds__(128, 1) attention_tc_fwd_kernel(const AttnTcParams p) {
  if (ctl_skipped(p.ctl)) return;
  const unsigned long long drop_seed = ctl_seed(p.ctl, p.drop_seed);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sP = sV + TILE_BYTES;  // two blocks
  uint64_t* bar = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  constexpr int PPT = 128 / BS;
  const int r = threadIdx.x, warp = r >> 5;
  const int blk = r / BS, t = r - blk * BS;
  const int pair = blockIdx.x * PPT + blk;
  const bool valid = t < p.n && pair < p.pairs;
  const int b = pair / p.heads, h = pair - b * p.heads;
  const int inner = p.heads * TC_D, ld = 3 * inner;

  if (r == 0) {
    mbar_init(&bar[0], 1), mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  {  // P is block diagonal: everything outside the pairs' own blocks stays zero
    uint4* z = reinterpret_cast<uint4*>(sP);
    for (int i = r; i < 2 * TILE_BYTES / 16; i += 128) z[i] = make_uint4(0, 0, 0, 0);
  }
  float cs[32] = {};
  if (valid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.rot + t * 32) + i);
      cs[4 * i] = v.x, cs[4 * i + 1] = v.y, cs[4 * i + 2] = v.z, cs[4 * i + 3] = v.w;
    }
  }
  const __nv_bfloat16* src = p.qkv + ((long long)b * p.n + t) * ld + h * TC_D;
  load_rot_store(src, valid, true, cs, sQ, r);
  load_rot_store(src + inner, valid, true, cs, sK, r);
  load_rot_store(src + 2 * inner, valid, p.rotary_v != 0, cs, sV, r);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (r == 0) {
    mma_k64(tmem, smem_u32(sQ), smem_u32(sK), umma_idesc_bf16(128, 128, 0, 0));  // S -> columns [0, 128)
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  tcgen05_fence_after();
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  {
    float s[BS], pr[BS];
    tmem_row<BS>(lane_base + (uint32_t)(blk * BS), s);
    softmax_row<BS>(s, p.n, pr);
    if (p.drop_p > 0.f) {  // Attention(dropout=attn_dropout): element index ((b*H + h)*n + i)*n + j
      const float ks = 1.0f / (1.0f - p.drop_p);
      const unsigned long long base = ((unsigned long long)pair * p.n + t) * p.n;
#pragma unroll
      for (int j = 0; j < BS; ++j)
        pr[j] = (j < p.n && dropout_keep(drop_seed, base + j, p.drop_p)) ? pr[j] * ks : 0.f;
    }
    if (!valid) {
#pragma unroll
      for (int j = 0; j < BS; ++j) pr[j] = 0.f;
    }
    store_block_row<BS>(sP, r, blk, pr);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (r == 0) {
    mma_k128(tmem + 128, smem_u32(sP), smem_u32(sV), false);  // O = P V' -> columns [128, 192)
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tcgen05_fence_after();
  {
    float ov[64];
    tmem_row<64>(lane_base + 128u, ov);
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(p.o + ((long long)b * p.n + t) * inner + h * TC_D);
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[c] = pack8(ov + 8 * c);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// rotate a gradient row back (transpose of the rotation) and store it as bf16
__device__ __forceinline__ void unrot_store(float (&x)[64], bool rotate, const float* cs, __nv_bfloat16* dst) {
  if (rotate) {
#pragma unroll
    for (int f = 0; f < 16; ++f) {
      const float a = x[f], b = x[f + 16];
      x[f] = a * cs[f] + b * cs[16 + f];
      x[f + 16] = b * cs[f] - a * cs[16 + f];
    }
  }
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int c = 0; c < 8; ++c) d4[c] = pack8(x + 8 * c);
}

template <int BS>
__global__ void __launch_bounds__(128, 1) attention_tc_bwd_kernel(const AttnTcParams p) {
  if (ctl_skipped(p.ctl)) return;
  const unsigned long long drop_seed = ctl_seed(p.ctl, p.drop_seed);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sO = sV + TILE_BYTES;   // dO
  uint8_t* sP = sO + TILE_BYTES;   // two blocks
  uint8_t* sS = sP + 2 * TILE_BYTES;  // dS, two blocks
  uint64_t* bar = reinterpret_cast<uint64_t*>(sS + 2 * TILE_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  constexpr int PPT = 128 / BS;
  const int r = threadIdx.x, warp = r >> 5;
  const int blk = r / BS, t = r - blk * BS;
  const int pair = blockIdx.x * PPT + blk;
  const bool valid = t < p.n && pair < p.pairs;
  const int b = pair / p.heads, h = pair - b * p.heads;
  const int inner = p.heads * TC_D, ld = 3 * inner;

  if (r == 0) {
    mbar_init(&bar[0], 1), mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  {
    uint4* z = reinterpret_cast<uint4*>(sP);  // P and dS are contiguous
    for (int i = r; i < 4 * TILE_BYTES / 16; i += 128) z[i] = make_uint4(0, 0, 0, 0);
  }
  float cs[32] = {};
  if (valid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.rot + t * 32) + i);
      cs[4 * i] = v.x, cs[4 * i + 1] = v.y, cs[4 * i + 2] = v.z, cs[4 * i + 3] = v.w;
    }
  }
  const long long row = (long long)b * p.n + t;
  const __nv_bfloat16* src = p.qkv + row * ld + h * TC_D;
  load_rot_store(src, valid, true, cs, sQ, r);
  load_rot_store(src + inner, valid, true, cs, sK, r);
  load_rot_store(src + 2 * inner, valid, p.rotary_v != 0, cs, sV, r);
  load_rot_store(p.d_o + row * inner + h * TC_D, valid, false, cs, sO, r);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (r == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    mma_k64(tmem, smem_u32(sQ), smem_u32(sK), idesc);        // S  -> columns [0, 128)
    mma_k64(tmem + 128, smem_u32(sO), smem_u32(sV), idesc);  // dP -> columns [128, 256)
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  tcgen05_fence_after();
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  {
    float s[BS], pr[BS], dp[BS];
    tmem_row<BS>(lane_base + (uint32_t)(blk * BS), s);
    softmax_row<BS>(s, p.n, pr);
    tmem_row<BS>(lane_base + 128u + (uint32_t)(blk * BS), dp);
    // d p = mask * d p~ ; dS = P o (dP - rowsum(dP o P)) * scale ; dV' uses p~ = mask * p
    float dot = 0.f;
    if (p.drop_p > 0.f) {
      const float ks = 1.0f / (1.0f - p.drop_p);
      const unsigned long long base = ((unsigned long long)pair * p.n + t) * p.n;
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        const float m = (j < p.n && dropout_keep(drop_seed, base + j, p.drop_p)) ? ks : 0.f;
        dp[j] *= m;
        dot += pr[j] * dp[j];
        s[j] = pr[j] * m;  // p~
      }
    } else {
#pragma unroll
      for (int j = 0; j < BS; ++j) dot += pr[j] * dp[j], s[j] = pr[j];
    }
#pragma unroll
    for (int j = 0; j < BS; ++j) {
      dp[j] = valid ? pr[j] * (dp[j] - dot) * 0.125f : 0.f;  // dS (the 64^-1/2 of the scores folded in)
      if (!valid) s[j] = 0.f;
    }
    store_block_row<BS>(sP, r, blk, s);
    store_block_row<BS>(sS, r, blk, dp);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (r == 0) {
    mma_k128(tmem + 256, smem_u32(sP), smem_u32(sO), true);   // dV' = P~^T dO -> [256, 320)
    mma_k128(tmem + 320, smem_u32(sS), smem_u32(sK), false);  // dQ' = dS K'   -> [320, 384)
    mma_k128(tmem + 384, smem_u32(sS), smem_u32(sQ), true);   // dK' = dS^T Q' -> [384, 448)
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tcgen05_fence_after();
  {
    __nv_bfloat16* dst = p.dqkv + row * ld + h * TC_D;
    float g[64];
    tmem_row<64>(lane_base + 320u, g);
    if (valid) unrot_store(g, true, cs, dst);
    tmem_row<64>(lane_base + 384u, g);
    if (valid) unrot_store(g, true, cs, dst + inner);
    tmem_row<64>(lane_base + 256u, g);
    if (valid) unrot_store(g, p.rotary_v != 0, cs, dst + 2 * inner);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused QKV projection + rotary + softmax + PV (forward of the x-transformers attention sublayer, n <= 32): north_star's
// "fused QKV-project + softmax + AV kernel fed by TMA with tcgen05". One CTA = one head x four clips (the 128 rows of the UMMA
// tile are 4 clips x 32 token slots): a 4-stage TMA pipeline streams the normalised activations [128 x 64] (one 3-D box:
// 64 channels x 32 tokens x 4 clips, token slots >= n and clips >= B zero-filled by the hardware) and the head's q | k | v
// weight rows [192 x 64] per 64-wide k-block; ONE N = 192 MMA per 16 channels accumulates [Q | K | V] in TMEM. Each thread
// then reads its row, stores the projections as bf16 for the backward pass, rotates q / k (/ v) in registers and writes
// the operand tiles over the drained pipeline stages; from there on it is the kernel above.
constexpr int QKV_STAGES = 4;
constexpr int QKV_A_BYTES = TILE_BYTES;          // [128 rows x 64 channels]
constexpr int QKV_B_BYTES = 3 * 64 * 128;        // [192 weight rows x 64 channels]
constexpr int QKV_STAGE_BYTES = QKV_A_BYTES + QKV_B_BYTES;
constexpr uint32_t QKV_TMEM = 256;               // [Q | K | V] accumulator columns (S at 0, O at 128 as above)

struct AttnQkvParams {
  const float* rot;
  __nv_bfloat16* qkv;  // [B*n, 3*inner] projections (read by the backward kernel)
  __nv_bfloat16* o;    // [B*n, inner]
  int n, heads, B, rotary_v, nkb;
  float drop_p;
  unsigned long long drop_seed;
  StepCtl ctl;
};

__global__ void __launch_bounds__(128, 1)
attention_qkv_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                            const AttnQkvParams p) {
  if (ctl_skipped(p.ctl)) return;
  const unsigned long long drop_seed = ctl_seed(p.ctl, p.drop_seed);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;  // the operand tiles of the attention part reuse the drained pipeline stages
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sP = sV + TILE_BYTES;  // two blocks
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + QKV_STAGES * QKV_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + QKV_STAGES;
  uint64_t* bar = empty_bar + QKV_STAGES;  // [0] projections done, [1] S done, [2] O done
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 3);
  constexpr int BS = 32;
  const int r = threadIdx.x, warp = r >> 5;
  const int blk = r / BS, t = r - blk * BS;
  const int h = blockIdx.x, b = blockIdx.y * 4 + blk;
  const bool valid = t < p.n && b < p.B;
  const int inner = p.heads * TC_D, ld = 3 * inner;

  if (r == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < QKV_STAGES; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
    mbar_init(&bar[0], 1), mbar_init(&bar[1], 1), mbar_init(&bar[2], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  float cs[32] = {};
  if (valid) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.rot + t * 32) + i);
      cs[4 * i] = v.x, cs[4 * i + 1] = v.y, cs[4 * i + 2] = v.z, cs[4 * i + 3] = v.w;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (r == 0) {  // TMA producer + MMA issuer of the projection
    auto issue = [&](int kb) {
      const int s = kb % QKV_STAGES;
      uint8_t* sA = smem + s * QKV_STAGE_BYTES;
      mbar_expect_tx(&full_bar[s], (uint32_t)QKV_STAGE_BYTES);
      tma_load_4d(sA, &tmX, &full_bar[s], kb * 64, 0, (int)blockIdx.y * 4, 0);
#pragma unroll
      for (int j = 0; j < 3; ++j)
        tma_load_2d(sA + QKV_A_BYTES + j * 8192, &tmW, &full_bar[s], kb * 64, j * inner + h * TC_D);
    };
    for (int kb = 0; kb < p.nkb && kb < QKV_STAGES; ++kb) issue(kb);
    constexpr uint32_t idesc = umma_idesc_bf16(128, 192, 0, 0);
    for (int kb = 0; kb < p.nkb; ++kb) {
      const int s = kb % QKV_STAGES;
      const uint32_t ph = (uint32_t)((kb / QKV_STAGES) & 1);
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * QKV_STAGE_BYTES);
      const uint64_t a_desc = umma_smem_desc_sw128(a_addr, 16, 1024);
      const uint64_t b_desc = umma_smem_desc_sw128(a_addr + QKV_A_BYTES, 16, 1024);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem + QKV_TMEM, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
      umma_commit(&empty_bar[s]);
      if (kb + QKV_STAGES < p.nkb) {
        mbar_wait(&empty_bar[s], ph);
        issue(kb + QKV_STAGES);
      }
    }
    umma_commit(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  tcgen05_fence_after();
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  {  // P is block diagonal: everything outside the pairs' own blocks stays zero (the stages are drained: all MMAs retired)
    uint4* z = reinterpret_cast<uint4*>(sP);
    for (int i = r; i < 2 * TILE_BYTES / 16; i += 128) z[i] = make_uint4(0, 0, 0, 0);
  }
  {
    __nv_bfloat16* dst = p.qkv + ((long long)b * p.n + t) * ld + h * TC_D;
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {  // q, k, v
      float x[64];
      tmem_row<64>(lane_base + QKV_TMEM + (uint32_t)(j * 64), x);
      if (valid) {
        uint4* d4 = reinterpret_cast<uint4*>(dst + j * inner);
#pragma unroll
        for (int c = 0; c < 8; ++c) d4[c] = pack8(x + 8 * c);
      }
      if (valid && (j < 2 || p.rotary_v != 0)) {
        // the backward kernel rotates the bf16 projections it reads back: rotate the same rounded values here
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = pack8(x + 8 * c);
          unpack8(u, x + 8 * c);
        }
#pragma unroll
        for (int f = 0; f < 16; ++f) {
          const float a = x[f], bb = x[f + 16];
          x[f] = a * cs[f] - bb * cs[16 + f];
          x[f + 16] = bb * cs[f] + a * cs[16 + f];
        }
      }
      uint8_t* tile = j == 0 ? sQ : (j == 1 ? sK : sV);
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sw_chunk(tile, r, c)) = pack8(x + 8 * c);
    }
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (r == 0) {
    mma_k64(tmem, smem_u32(sQ), smem_u32(sK), umma_idesc_bf16(128, 128, 0, 0));  // S -> columns [0, 128)
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tcgen05_fence_after();
  const int pair = b * p.heads + h;
  {
    float s[BS], pr[BS];
    tmem_row<BS>(lane_base + (uint32_t)(blk * BS), s);
    softmax_row<BS>(s, p.n, pr);
    if (p.drop_p > 0.f) {  // Attention(dropout=attn_dropout): element index ((b*H + h)*n + i)*n + j
      const float ks = 1.0f / (1.0f - p.drop_p);
      const unsigned long long base = ((unsigned long long)pair * p.n + t) * p.n;
#pragma unroll
      for (int j = 0; j < BS; ++j)
        pr[j] = (j < p.n && dropout_keep(drop_seed, base + j, p.drop_p)) ? pr[j] * ks : 0.f;
    }
    if (!valid) {
#pragma unroll
      for (int j = 0; j < BS; ++j) pr[j] = 0.f;
    }
    store_block_row<BS>(sP, r, blk, pr);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (r == 0) {
    mma_k128(tmem + 128, smem_u32(sP), smem_u32(sV), false);  // O = P V' -> columns [128, 192)
    umma_commit(&bar[2]);
  }
  mbar_wait(&bar[2], 0);
  tcgen05_fence_after();
  {
    float ov[64];
    tmem_row<64>(lane_base + 128u, ov);
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(p.o + ((long long)b * p.n + t) * inner + h * TC_D);
#pragma unroll
      for (int c = 0; c < 8; ++c) dst[c] = pack8(ov + 8 * c);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int BS>
int launch_tc(const AttnTcParams& p, bool bwd, cudaStream_t s) {
  constexpr int PPT = 128 / BS;
  const int grid = (p.pairs + PPT - 1) / PPT;
  const int smem = (bwd ? 8 : 5) * TILE_BYTES + 64 + 1024;
  static bool done[2] = {false, false};
  if (!done[bwd]) {
    if (bwd)
      SVSR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_bwd_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else
      SVSR_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_fwd_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done[bwd] = true;
  }
  if (bwd)
    attention_tc_bwd_kernel<BS><<<grid, 128, smem, s>>>(p);
  else
    attention_tc_fwd_kernel<BS><<<grid, 128, smem, s>>>(p);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // namespace

int attention_tc_fwd(const __nv_bfloat16* qkv, const float* rot, __nv_bfloat16* o, int B, int n, int heads, int rotary_v,
                     cudaStream_t s, float drop_p, unsigned long long drop_seed, const StepCtl* ctl) {
  SVSR_REQUIRE(n >= 1 && n <= 64, "attention: n=%d must be in [1,64]", n);
  AttnTcParams p{qkv, rot, nullptr, o, nullptr, n, heads, B * heads, rotary_v, drop_p, drop_seed, ctl ? *ctl : StepCtl()};
  return n <= 32 ? launch_tc<32>(p, false, s) : launch_tc<64>(p, false, s);
}
int attention_tc_bwd(const __nv_bfloat16* qkv, const float* rot, const __nv_bfloat16* d_o, __nv_bfloat16* dqkv, int B, int n,
                     int heads, int rotary_v, cudaStream_t s, float drop_p, unsigned long long drop_seed,
                     const StepCtl* ctl) {
  SVSR_REQUIRE(n >= 1 && n <= 64, "attention: n=%d must be in [1,64]", n);
  AttnTcParams p{qkv, rot, d_o, nullptr, dqkv, n, heads, B * heads, rotary_v, drop_p, drop_seed, ctl ? *ctl : StepCtl()};
  return n <= 32 ? launch_tc<32>(p, true, s) : launch_tc<64>(p, true, s);
}

// Fused projection + attention forward (see attention_qkv_tc_fwd_kernel). xn [B*n, ldx] bf16 normalised activations, w [3*inner,
// Kp] bf16 (q | k | v weight rows, K-major, Kp % 64 == 0 columns contracted), qkv [B*n, 3*inner] receives the projections.
int attention_qkv_tc_fwd(const __nv_bfloat16* xn, int ldx, const __nv_bfloat16* w, int Kp, const float* rot,
                         __nv_bfloat16* qkv, __nv_bfloat16* o, int B, int n, int heads, int rotary_v, cudaStream_t s,
                         float drop_p, unsigned long long drop_seed, const StepCtl* ctl) {
  SVSR_REQUIRE(n >= 1 && n <= 32, "fused qkv attention: n=%d must be in [1,32]", n);
  SVSR_REQUIRE(Kp % 64 == 0 && Kp <= ldx && ldx % 8 == 0, "fused qkv attention: Kp=%d ldx=%d", Kp, ldx);
  const int inner = heads * TC_D;
  CUtensorMap tmX, tmW;
  {
    uint64_t dims[4] = {(uint64_t)Kp, (uint64_t)n, (uint64_t)B, 1};
    uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)n * ldx * 2, (uint64_t)B * n * ldx * 2};
    uint32_t box[4] = {64, 32, 4, 1};
    int rc = make_tmap_bf16(&tmX, xn, 4, dims, strides, box, nullptr, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Kp, (uint64_t)(3 * inner)};
    uint64_t strides[1] = {(uint64_t)Kp * 2};
    uint32_t box[2] = {64, 64};
    int rc = make_tmap_bf16(&tmW, w, 2, dims, strides, box, nullptr, true);
    if (rc) return rc;
  }
  AttnQkvParams p{rot, qkv, o, n, heads, B, rotary_v, Kp / 64, drop_p, drop_seed, ctl ? *ctl : StepCtl()};
  constexpr int smem = QKV_STAGES * QKV_STAGE_BYTES + 128 + 1024;
  static bool done = false;
  if (!done) {
    SVSR_CHECK_CUDA(cudaFuncSetAttribute(attention_qkv_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  dim3 grid((unsigned)heads, (unsigned)((B + 3) / 4));
  attention_qkv_tc_fwd_kernel<<<grid, 128, smem, s>>>(tmX, tmW, p);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // namespace svsr
