// d_k = 64 attention core of the LRS path on the tensor cores (tcgen05.mma, accumulators in TMEM).
// Reference: espnet/nets/pytorch_backend/transformer/attention.py:192-278 (RelPositionMultiHeadedAttention: scores =
// ((q + u) k^T + rel_shift((q + v) p^T)) / sqrt(d_k)), :38-118 (MultiHeadedAttention: decoder self / source attention).
//
// Tiles: 128 query rows (one per thread = one TMEM lane) x 64 keys. Per tile pair
//     S  = (Q + u) K^T                    [128 x 64]    tcgen05, K-major operands
//     W  = (Q + v) Pwin^T                 [128 x 192]   Pwin = the 191 rows of p this tile pair can reach
//     BD[ii][jj] = W[ii][jj - ii + 127]                  rel_shift = a per-row skew: every thread copies ITS row of W
//                                                        from TMEM into its own fp32 staging row at the shifted
//                                                        position (bank-conflict-free: pitch 68, lanes step 69 words)
// so neither the [B,h,T,2T-1] tensor nor its shifted copy ever exists. Forward: online softmax over the key tiles in
// registers, O_tile = P~ V on the tensor core, rescaled per tile. Backward, three kernels:
//   query side : recompute p from the saved log-sum-exp, dS = p o (mask o dP - delta) / sqrt(d_k), delta = dO . O;
//                dQ = dS K + dW Pwin (dW = dS skewed back into window coordinates, a bf16 operand tile whose band
//                positions never move, so its zeros are written once); P~, dS and dW go to a bf16 scratch (dW rows are
//                stored at window coordinates, origin chosen so that every 16-byte chunk of the tile is a 16-byte
//                chunk of the scratch row; the chunk two key tiles share travels in a register)
//   key side   : dV = P~^T dO, dK = dS^T (Q + u): the scratch tiles are read as MN-major operands (no transposes)
//   p side     : dP[r] = sum_b sum_i dW_b[i][r] (q_i + v): the same MN-major product over the dW scratch, accumulated
//                over a chunk of the batch in TMEM before the fp32 atomics
// 256 threads per CTA: two threads share a query row (32 keys / 32 output columns each; TMEM lane quarters repeat every
// four warps); key / value / window tiles arrive by cp.async (16-byte, zero-filled outside the clip).
#include "attention_rel_tc.cuh"
#include "attention_tc.cuh"

namespace svsr {
namespace {
using namespace attn_tc;

constexpr int NT = 256;                   // threads per CTA: thread (ii = tid & 127, hf = tid >> 7) owns half a query row
constexpr int QT = 128;                   // query rows per tile (= TMEM lanes)
constexpr int KT = 64;                    // keys per tile
constexpr int WR = 192;                   // window rows per tile pair: rl = jj - ii + 127 in [0, 191)
constexpr int SP = 68;                    // fp32 words per staging row (64 + max / sum exchange slots)
constexpr int KTILE_BYTES = KT * 128;     // [64 x 64] bf16
constexpr int WTILE_BYTES = WR * 128;     // [192 x 64] bf16
constexpr int STAGE_BYTES = QT * SP * 4;  // 34816 = 34 KB
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~uintptr_t(1023));
}
__host__ __device__ inline int round_up64(int x) { return (x + 63) & ~63; }
// dW scratch rows: element (i, r) of the window-coordinate score gradient lives at column r + dw_off(Tk); with this origin
// the 16-byte chunks of every (query tile, key tile) operand tile are 16-byte chunks of the row
__host__ __device__ inline int dw_off(int Tk) { return (8 - (Tk & 7)) & 7; }
__host__ __device__ inline int dw_pitch(int Tk) { return (Tk + round_up64(Tk) + 24 + 7) & ~7; }
// keys >= the returned bound are masked for EVERY row of the query tile starting at i0
__device__ __forceinline__ int active_keys(const AttnK& a, int b, int i0) {
  int jend = a.klen ? min(max(__ldg(a.klen + b), 0), a.Tk) : a.Tk;
  if (a.causal) jend = min(jend, i0 + QT);
  return jend;
}

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 16 : 0)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// rows [0, nrows) of a K-major SWIZZLE_128B tile [nrows x 64 bf16] <- global rows g = g0 + r of `src` (pitch ld
// elements; the head's column offset is already applied), zeros where g is outside [lo, hi); `bias` (64 fp32) is added
// before the rounding. 8 threads per row, one 16-byte chunk each (coalesced 128-byte rows, conflict-free stores).
__device__ __forceinline__ void load_tile(uint8_t* tile, int nrows, const __nv_bfloat16* __restrict__ src, long long ld,
                                          long long g0, long long lo, long long hi, const float* __restrict__ bias) {
  for (int idx = threadIdx.x; idx < nrows * 8; idx += NT) {
    const int r = idx >> 3, c = idx & 7;
    const long long g = g0 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (g >= lo && g < hi) {
      v = __ldg(reinterpret_cast<const uint4*>(src + g * ld) + c);
      if (bias) {
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += __ldg(bias + c * 8 + k);
        v = pack8(f);
      }
    }
    *reinterpret_cast<uint4*>(sw_chunk(tile, r, c)) = v;
  }
}
// (Q + u) and (Q + v) operand tiles from ONE read of the query rows
__device__ __forceinline__ void load_q_tiles(uint8_t* tile_u, uint8_t* tile_v, const __nv_bfloat16* __restrict__ src, long long ld,
                                             int i0, int Tq, const float* __restrict__ bu, const float* __restrict__ bv) {
  for (int idx = threadIdx.x; idx < QT * 8; idx += NT) {
    const int r = idx >> 3, c = idx & 7;
    uint4 vu = make_uint4(0u, 0u, 0u, 0u), vv = vu;
    if (i0 + r < Tq) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (long long)(i0 + r) * ld) + c);
      float f[8], g[8];
      unpack8(v, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = f[k] + (bv ? __ldg(bv + c * 8 + k) : 0.f), f[k] += bu ? __ldg(bu + c * 8 + k) : 0.f;
      vu = pack8(f), vv = pack8(g);
    }
    *reinterpret_cast<uint4*>(sw_chunk(tile_u, r, c)) = vu;
    *reinterpret_cast<uint4*>(sw_chunk(tile_v, r, c)) = vv;
  }
}
// the same without a bias, as asynchronous 16-byte copies (all of a thread's chunks in flight at once)
__device__ __forceinline__ void load_tile_async(uint8_t* tile, int nrows, const __nv_bfloat16* __restrict__ src, long long ld,
                                                long long g0, long long lo, long long hi) {
  for (int idx = threadIdx.x; idx < nrows * 8; idx += NT) {
    const int r = idx >> 3, c = idx & 7;
    const long long g = g0 + r;
    const bool ok = g >= lo && g < hi;
    cp_async16(sw_chunk(tile, r, c), ok ? src + g * ld + c * 8 : src, ok);
  }
}

// D[128 x 64] (+)= X[128 x 64 nblk (K)] . Y[64 nblk (K) x 64]: X = nblk K-major blocks of [128 x 64] (block kb = K columns
// [64 kb, 64 kb + 64)), Y MN-major (rows = K index, 64 N columns, rows contiguous over the blocks)
__device__ __forceinline__ void mma_ab(uint32_t d_tmem, uint32_t x_addr, int nblk, uint32_t y_addr, bool accumulate) {
  constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
  for (int ks = 0; ks < 4 * nblk; ++ks) {
    const uint64_t a = umma_smem_desc_sw128(x_addr + (ks >> 2) * TILE_BYTES + (ks & 3) * 32, 16, 1024);
    const uint64_t b = umma_smem_desc_sw128(y_addr + ks * 2048, TILE_BYTES, 1024);
    umma_bf16(d_tmem, a, b, idesc, (accumulate || ks != 0) ? 1u : 0u);
  }
}

// rel_shift for the 32 keys [32 hf, 32 hf + 32) of row ii = 32 wq + lane: srow[jj] = W[ii][jj + 127 - ii]. The thread only
// needs the two 32-column chunks hf - wq + 3 and hf - wq + 4 of its row of W (TMEM columns at w_lane_base).
__device__ __forceinline__ void skew_bd(uint32_t w_lane_base, int wq, int hf, int ii, float* srow) {
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int cw = hf - wq + 3 + t;
    float v[32];
    tmem_row<32>(w_lane_base + (uint32_t)(cw * 32), v);
    const int jb = cw * 32 - (QT - 1) + ii - hf * 32;  // (key - 32 hf) of element 0
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const int jr = jb + e;
      if ((unsigned)jr < 32u) srow[hf * 32 + jr] = v[e];
    }
  }
}

__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const float* x) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int c = 0; c < 4; ++c) d4[c] = pack8(x + 8 * c);
}

#define ATTN_SYNC_FOR_MMA()   \
  fence_proxy_async_smem();   \
  tcgen05_fence_before();     \
  __syncthreads();            \
  tcgen05_fence_after()

// ================================================================================================= forward
template <bool REL, bool DROP>
__global__ void __launch_bounds__(NT, 2) attn_rel_fwd_kernel(const AttnK a) {
  extern __shared__ uint8_t attn_rel_smem[];
  uint8_t* smem = align1024(attn_rel_smem);
  uint8_t* sQu = smem;
  uint8_t* sQv = sQu + TILE_BYTES;
  uint8_t* sK = sQv + TILE_BYTES;
  uint8_t* sV = sK + KTILE_BYTES;
  uint8_t* sPw = sV + KTILE_BYTES;  // window rows of p; once W is done its first 16 KB carry the P~ operand tile
  float* stage = reinterpret_cast<float*>(sPw + WTILE_BYTES);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stage) + STAGE_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, wq = warp & 3, hf = warp >> 2, ii = tid & (QT - 1);
  const int i0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int i = i0 + ii;
  const bool row_ok = i < a.Tq;
  const bool warp_rows = i0 + wq * 32 < a.Tq;
  if (tid == 0) {
    mbar_init(&bar[0], 1), mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  const __nv_bfloat16* qsrc = a.q + (long long)b * a.Tq * a.ldq + h * 64;
  if (REL)
    load_q_tiles(sQu, sQv, qsrc, a.ldq, i0, a.Tq, a.bu ? a.bu + h * 64 : nullptr, a.bv ? a.bv + h * 64 : nullptr);
  else
    load_tile(sQu, QT, qsrc, a.ldq, i0, 0, a.Tq, a.bu ? a.bu + h * 64 : nullptr);
  const int klen = a.klen ? min(max(__ldg(a.klen + b), 0), a.Tk) : a.Tk;
  const int nk = (active_keys(a, b, i0) + KT - 1) / KT;
  const float c2 = a.scale * LOG2E;
  const float ks = DROP ? 1.0f / (1.0f - a.drop_p) : 1.0f;
  const unsigned long long dseed = DROP ? seed_plus(a.drop_seed, a.seed_base) : 0ULL;
  const unsigned long long e0 = (((unsigned long long)b * a.H + h) * a.Tq + i) * a.Tk;
  float m_run = -INFINITY, l_run = 0.f;  // l_run: this thread's 32 keys only
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.f;
  uint32_t ph = 0;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
  float* srow = stage + ii * SP;

  for (int kt = 0; kt < nk; ++kt) {
    const int j0 = kt * KT;
    load_tile_async(sK, KT, a.k + (long long)b * a.Tk * a.ldk + h * 64, a.ldk, j0, 0, a.Tk);
    load_tile_async(sV, KT, a.v + (long long)b * a.Tk * a.ldv + h * 64, a.ldv, j0, 0, a.Tk);
    if (REL) load_tile_async(sPw, WR, a.p + h * 64, a.ldp, (long long)j0 - i0 - (QT - 1) + a.Tk - 1, 0, 2LL * a.Tk - 1);
    cp_async_wait_all();
    ATTN_SYNC_FOR_MMA();
    if (tid == 0) {
      mma_k64(tmem, smem_u32(sQu), smem_u32(sK), umma_idesc_bf16(128, KT, 0, 0));                  // S -> [0, 64)
      if (REL) mma_k64(tmem + 64, smem_u32(sQv), smem_u32(sPw), umma_idesc_bf16(128, WR, 0, 0));  // W -> [64, 256)
      umma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], ph);
    tcgen05_fence_after();
    float t[32];
    float mt = -INFINITY;
    if (!warp_rows) {  // (uniform per warp) a tail tile: none of this warp's rows exists
#pragma unroll
      for (int e = 0; e < 32; ++e) t[e] = -INFINITY;
    } else {
      if (REL) skew_bd(lane_base + 64u, wq, hf, ii, srow);
      float s[32];
      tmem_row<32>(lane_base + (uint32_t)(hf * 32), s);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        float4 bd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (REL) bd = reinterpret_cast<const float4*>(srow)[hf * 8 + q4];
        const float add[4] = {bd.x, bd.y, bd.z, bd.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = q4 * 4 + u, j = j0 + hf * 32 + e;
          const bool masked = j >= klen || (a.causal && j > i) || !row_ok;
          const float tv = masked ? -INFINITY : (s[e] + add[u]) * c2;
          t[e] = tv;
          mt = fmaxf(mt, tv);
        }
      }
    }
    srow[64 + hf] = mt;  // the row maximum is over both halves
    __syncthreads();
    const float m_new = fmaxf(m_run, fmaxf(srow[64], srow[65]));
    const float alpha = m_new == -INFINITY ? 1.0f : exp2f(m_run - m_new);
    float sum = 0.f;
    if (warp_rows) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        float pv = t[e] == -INFINITY ? 0.f : exp2f(t[e] - m_new);
        sum += pv;  // the softmax normaliser is taken before dropout
        if (DROP) pv = dropout_keep(dseed, e0 + (unsigned long long)(j0 + hf * 32 + e), a.drop_p) ? pv * ks : 0.f;
        t[e] = pv;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) t[e] = 0.f;
    }
    l_run = l_run * alpha + sum;
    m_run = m_new;
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(sw_chunk(sPw, ii, hf * 4 + c)) = pack8(t + 8 * c);
    ATTN_SYNC_FOR_MMA();
    if (tid == 0) {
      mma_ab(tmem, smem_u32(sPw), 1, smem_u32(sV), false);  // O_tile = P~ V -> [0, 64) (S is consumed)
      umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], ph);
    tcgen05_fence_after();
    ph ^= 1u;
    if (warp_rows) {
      float s[32];
      tmem_row<32>(lane_base + (uint32_t)(hf * 32), s);
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], alpha, s[e]);
    }
    tcgen05_fence_before();
  }
  srow[66 + hf] = l_run;
  __syncthreads();
  const float l_tot = srow[66] + srow[67];
  if (row_ok) {
    const float inv = l_tot > 0.f ? 1.0f / l_tot : 0.f;  // every key masked: the reference's re-masked row is all zero
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] *= inv;
    store_row32(a.o + ((long long)b * a.Tq + i) * a.ldo + h * 64 + hf * 32, o);
    if (a.lse && hf == 0) a.lse[((long long)b * a.H + h) * a.Tq + i] = l_tot > 0.f ? fmaf(m_run, LN2, logf(l_tot)) : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ================================================================================================= backward, query side
template <bool REL, bool DROP>
__global__ void __launch_bounds__(NT, 1)
attn_rel_bwd_q_kernel(const AttnK a, __nv_bfloat16* __restrict__ Pg, __nv_bfloat16* __restrict__ DSg,
                      __nv_bfloat16* __restrict__ DWg, int Tkp) {
  extern __shared__ uint8_t attn_rel_smem[];
  uint8_t* smem = align1024(attn_rel_smem);
  uint8_t* sQu = smem;
  uint8_t* sQv = sQu + TILE_BYTES;
  uint8_t* sdO = sQv + TILE_BYTES;
  uint8_t* sK = sdO + TILE_BYTES;
  uint8_t* sV = sK + KTILE_BYTES;
  uint8_t* sPw = sV + KTILE_BYTES;
  float* stage = reinterpret_cast<float*>(sPw + WTILE_BYTES);
  uint8_t* sdS = reinterpret_cast<uint8_t*>(stage) + STAGE_BYTES;  // [128 x 64 keys], one K-major block
  uint8_t* sdW = sdS + TILE_BYTES;                                 // [128 x 192 window positions], three blocks
  uint64_t* bar = reinterpret_cast<uint64_t*>(sdW + 3 * TILE_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  constexpr uint32_t C_S = 0, C_DP = 64, C_W = 128, C_DQA = 320, C_DQB = 384;

  const int tid = threadIdx.x, warp = tid >> 5, wq = warp & 3, hf = warp >> 2, ii = tid & (QT - 1);
  const int i0 = blockIdx.x * QT, h = blockIdx.y, b = blockIdx.z;
  const int i = i0 + ii;
  const bool row_ok = i < a.Tq;
  const bool warp_rows = i0 + wq * 32 < a.Tq;  // a tail tile: warps without a single valid row only keep the barriers
  const long long bh = (long long)b * a.H + h;
  if (tid == 0) {
    mbar_init(&bar[0], 1), mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  const __nv_bfloat16* qsrc = a.q + (long long)b * a.Tq * a.ldq + h * 64;
  const __nv_bfloat16* dosrc = a.d_o + (long long)b * a.Tq * a.ldo + h * 64;
  load_tile_async(sdO, QT, dosrc, a.ldo, i0, 0, a.Tq);
  if (REL)
    load_q_tiles(sQu, sQv, qsrc, a.ldq, i0, a.Tq, a.bu ? a.bu + h * 64 : nullptr, a.bv ? a.bv + h * 64 : nullptr);
  else
    load_tile(sQu, QT, qsrc, a.ldq, i0, 0, a.Tq, a.bu ? a.bu + h * 64 : nullptr);
  {  // dS rows of warps without valid rows stay zero; dW: the band a row's two threads write never moves, the rest stays zero
    uint4* z = reinterpret_cast<uint4*>(sdS);
    for (int x = tid; x < (REL ? 4 : 1) * TILE_BYTES / 16; x += NT) z[x] = make_uint4(0u, 0u, 0u, 0u);
  }
  float delta = 0.f, lse2 = 0.f;  // delta = sum_j p~ dP~ = dO . O ; lse in log2 units
  if (row_ok) {
    const uint4* po = reinterpret_cast<const uint4*>(a.o + ((long long)b * a.Tq + i) * a.ldo + h * 64);
    const uint4* pd = reinterpret_cast<const uint4*>(dosrc + (long long)i * a.ldo);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float x[8], y[8];
      unpack8(__ldg(po + c), x), unpack8(__ldg(pd + c), y);
#pragma unroll
      for (int k = 0; k < 8; ++k) delta = fmaf(x[k], y[k], delta);
    }
    lse2 = a.lse[bh * a.Tq + i] * LOG2E;
  }
  const int klen = a.klen ? min(max(__ldg(a.klen + b), 0), a.Tk) : a.Tk;
  const int jend = active_keys(a, b, i0);
  const float c2 = a.scale * LOG2E;
  const float ks = DROP ? 1.0f / (1.0f - a.drop_p) : 1.0f;
  const unsigned long long dseed = DROP ? seed_plus(a.drop_seed, a.seed_base) : 0ULL;
  const unsigned long long e0 = ((unsigned long long)bh * a.Tq + i) * a.Tk;
  __nv_bfloat16* prow = Pg + (bh * a.Tq + i) * (long long)Tkp + hf * 32;
  __nv_bfloat16* dsrow = DSg + (bh * a.Tq + i) * (long long)Tkp + hf * 32;
  // dW scratch row: the chunk holding window position rl = 8 wc of key tile j0 starts at column j0 + dwc0 + 8 wc
  const int c0 = (QT - 1 - ii) >> 3;  // first window chunk this row's band touches
  const int dwc0 = a.Tk - QT - i0 + dw_off(a.Tk);
  __nv_bfloat16* dwrow = REL ? DWg + (bh * a.Tq + i) * (long long)dw_pitch(a.Tk) + dwc0 + 8 * c0 : nullptr;
  uint4 carry = make_uint4(0u, 0u, 0u, 0u);  // (hf == 1) the chunk this tile's band shares with the next tile's
  uint32_t ph = 0;
  bool any = false;
  cp_async_wait_all();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
  float* srow = stage + ii * SP;

  for (int j0 = 0; j0 < Tkp; j0 += KT) {
    if (j0 >= jend) {  // every (row, key) of this tile pair is masked: P~ = dS = 0 (uniform per CTA)
      if (row_ok) {
        uint4* z0 = reinterpret_cast<uint4*>(prow + j0);
        uint4* z1 = reinterpret_cast<uint4*>(dsrow + j0);
#pragma unroll
        for (int c = 0; c < 4; ++c) z0[c] = make_uint4(0u, 0u, 0u, 0u), z1[c] = make_uint4(0u, 0u, 0u, 0u);
      }
      continue;
    }
    load_tile_async(sK, KT, a.k + (long long)b * a.Tk * a.ldk + h * 64, a.ldk, j0, 0, a.Tk);
    load_tile_async(sV, KT, a.v + (long long)b * a.Tk * a.ldv + h * 64, a.ldv, j0, 0, a.Tk);
    if (REL) load_tile_async(sPw, WR, a.p + h * 64, a.ldp, (long long)j0 - i0 - (QT - 1) + a.Tk - 1, 0, 2LL * a.Tk - 1);
    cp_async_wait_all();
    ATTN_SYNC_FOR_MMA();
    if (tid == 0) {
      constexpr uint32_t id64 = umma_idesc_bf16(128, KT, 0, 0);
      mma_k64(tmem + C_S, smem_u32(sQu), smem_u32(sK), id64);   // S
      mma_k64(tmem + C_DP, smem_u32(sdO), smem_u32(sV), id64);  // dP~ = dO V^T
      if (REL) mma_k64(tmem + C_W, smem_u32(sQv), smem_u32(sPw), umma_idesc_bf16(128, WR, 0, 0));
      umma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], ph);
    tcgen05_fence_after();
    if (warp_rows) {
      if (REL) skew_bd(lane_base + C_W, wq, hf, ii, srow);
      float s[32], dp[32];
      tmem_row<32>(lane_base + C_S + (uint32_t)(hf * 32), s);
      tmem_row<32>(lane_base + C_DP + (uint32_t)(hf * 32), dp);
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        float4 bd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (REL) bd = reinterpret_cast<const float4*>(srow)[hf * 8 + q4];
        const float add[4] = {bd.x, bd.y, bd.z, bd.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = q4 * 4 + u, j = j0 + hf * 32 + e;
          const bool masked = j >= klen || (a.causal && j > i) || !row_ok;
          const float pj = masked ? 0.f : exp2f((s[e] + add[u]) * c2 - lse2);
          // dropout on the probabilities: d p = mask o d p~ ; the key/value side uses p~ = mask o p
          const float mj = (DROP && !dropout_keep(dseed, e0 + (unsigned long long)j, a.drop_p)) ? 0.f : ks;
          s[e] = pj * mj;
          dp[e] = pj * (dp[e] * mj - delta) * a.scale;
        }
      }
      if (row_ok) {
        uint4* gp = reinterpret_cast<uint4*>(prow + j0);
        uint4* gs = reinterpret_cast<uint4*>(dsrow + j0);
#pragma unroll
        for (int c = 0; c < 4; ++c) gp[c] = pack8(s + 8 * c), gs[c] = pack8(dp + 8 * c);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(sw_chunk(sdS, ii, hf * 4 + c)) = pack8(dp + 8 * c);
      if (REL) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int rl = hf * 32 + e + (QT - 1) - ii;
          uint8_t* dst = sdW + (rl >> 6) * TILE_BYTES + ii * 128 + ((((rl & 63) >> 3) ^ (ii & 7)) << 4) + (rl & 7) * 2;
          *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(dp[e]);
        }
      }
    }
    ATTN_SYNC_FOR_MMA();
    if (tid == 0) {
      mma_ab(tmem + C_DQA, smem_u32(sdS), 1, smem_u32(sK), any);             // dQ (content term) += dS K
      if (REL) mma_ab(tmem + C_DQB, smem_u32(sdW), 3, smem_u32(sPw), any);  // dQ (position term) += dW Pwin
      umma_commit(&bar[1]);
    }
    if (REL && row_ok) {  // while the MMAs run: this row of dW -> scratch (window chunks c0 .. c0 + 8 of the tile)
      __nv_bfloat16* dst = dwrow + j0;
      auto wchunk = [&](int k) {
        const int wc = c0 + k;
        return *reinterpret_cast<const uint4*>(sw_chunk(sdW + (wc >> 3) * TILE_BYTES, ii, wc & 7));
      };
      if (hf == 0) {
#pragma unroll
        for (int k = 1; k <= 4; ++k) *reinterpret_cast<uint4*>(dst + 8 * k) = wchunk(k);
      } else {
        uint4 f = wchunk(0);
        f.x |= carry.x, f.y |= carry.y, f.z |= carry.z, f.w |= carry.w;
        *reinterpret_cast<uint4*>(dst) = f;
#pragma unroll
        for (int k = 5; k <= 7; ++k) *reinterpret_cast<uint4*>(dst + 8 * k) = wchunk(k);
        carry = wchunk(8);
      }
    }
    mbar_wait(&bar[1], ph);
    tcgen05_fence_after();
    ph ^= 1u;
    any = true;
    tcgen05_fence_before();
  }
  if (REL && row_ok && hf == 1 && jend > 0)  // the last band's tail chunk
    *reinterpret_cast<uint4*>(dwrow + round_up64(jend)) = carry;

  float ga[32], gb[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) ga[d] = 0.f, gb[d] = 0.f;
  if (any) {
    tcgen05_fence_after();
    tmem_row<32>(lane_base + C_DQA + (uint32_t)(hf * 32), ga);
    if (REL) tmem_row<32>(lane_base + C_DQB + (uint32_t)(hf * 32), gb);
  }
  if (row_ok) {
    float g[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) g[d] = ga[d] + gb[d];
    store_row32(a.dq + ((long long)b * a.Tq + i) * a.lddq + h * 64 + hf * 32, g);
  }
  if (REL && (a.dbu || a.dbv)) {  // dbias_u / dbias_v += column sums of the two dQ terms (rows >= Tq are zero)
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      float* dst = which == 0 ? a.dbu : a.dbv;
      __syncthreads();
      if (dst) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          reinterpret_cast<float4*>(srow)[hf * 8 + c] =
              which == 0 ? make_float4(ga[4 * c], ga[4 * c + 1], ga[4 * c + 2], ga[4 * c + 3])
                         : make_float4(gb[4 * c], gb[4 * c + 1], gb[4 * c + 2], gb[4 * c + 3]);
      }
      __syncthreads();
      if (dst && tid < 64) {
        float acc = 0.f;
        for (int r = 0; r < QT; ++r) acc += stage[r * SP + tid];
        atomicAdd(dst + h * 64 + tid, acc);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ================================================================================================= backward, key/value side
// One CTA per 128 keys of one (clip, head): dV = P~^T dO, dK = dS^T (Q + u), summed over the query tiles in TMEM.
__global__ void __launch_bounds__(NT) attn_rel_bwd_kv_kernel(const AttnK a, const __nv_bfloat16* __restrict__ Pg,
                                                             const __nv_bfloat16* __restrict__ DSg, int Tkp) {
  extern __shared__ uint8_t attn_rel_smem[];
  uint8_t* smem = align1024(attn_rel_smem);
  uint8_t* sP = smem;                    // two blocks: [128 queries x 128 keys]
  uint8_t* sS = sP + 2 * TILE_BYTES;     // dS, same shape
  uint8_t* sdO = sS + 2 * TILE_BYTES;    // [128 queries x 64]
  uint8_t* sQu = sdO + TILE_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sQu + TILE_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, wq = warp & 3, hf = warp >> 2, jj = tid & 127;
  const int jk0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const long long bh = (long long)b * a.H + h;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 128);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
  const int klen = a.klen ? min(max(__ldg(a.klen + b), 0), a.Tk) : a.Tk;
  uint32_t ph = 0;
  bool any = false;
  const int nq = jk0 < klen ? (a.Tq + QT - 1) / QT : 0;  // a key tile past the clip's length: gradients are exactly zero
  for (int qt = 0; qt < nq; ++qt) {
    const int i0 = qt * QT;
    if (a.causal && i0 + QT - 1 < jk0) continue;  // every query of the tile precedes every key
    for (int idx = tid; idx < QT * 16; idx += NT) {
      const int r = idx >> 4, c16 = idx & 15;
      const int i = i0 + r, col = jk0 + c16 * 8;
      const bool ok = i < a.Tq && col < Tkp;
      const long long off = ok ? (bh * a.Tq + i) * (long long)Tkp + col : 0;
      cp_async16(sw_chunk(sP + (c16 >> 3) * TILE_BYTES, r, c16 & 7), Pg + off, ok);
      cp_async16(sw_chunk(sS + (c16 >> 3) * TILE_BYTES, r, c16 & 7), DSg + off, ok);
    }
    load_tile_async(sdO, QT, a.d_o + (long long)b * a.Tq * a.ldo + h * 64, a.ldo, i0, 0, a.Tq);
    load_tile(sQu, QT, a.q + (long long)b * a.Tq * a.ldq + h * 64, a.ldq, i0, 0, a.Tq, a.bu ? a.bu + h * 64 : nullptr);
    cp_async_wait_all();
    ATTN_SYNC_FOR_MMA();
    if (tid == 0) {
      mma_k128(tmem, smem_u32(sP), smem_u32(sdO), true, any);        // dV += P~^T dO
      mma_k128(tmem + 64, smem_u32(sS), smem_u32(sQu), true, any);   // dK += dS^T (Q + u)
      umma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], ph);
    tcgen05_fence_after();
    ph ^= 1u;
    any = true;
    tcgen05_fence_before();
  }
  const int j = jk0 + jj;
  float g[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) g[d] = 0.f;
  if (any) {
    tcgen05_fence_after();
    tmem_row<32>(lane_base + (uint32_t)(hf * 32), g);
  }
  if (j < a.Tk) store_row32(a.dv + ((long long)b * a.Tk + j) * a.lddv + h * 64 + hf * 32, g);
  if (any) tmem_row<32>(lane_base + 64u + (uint32_t)(hf * 32), g);
  if (j < a.Tk) store_row32(a.dk + ((long long)b * a.Tk + j) * a.lddk + h * 64 + hf * 32, g);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// ================================================================================================= backward, p side
// One CTA per 128 columns u = r + dw_off of the dW scratch (r = relative position), one head and a chunk of the batch:
// dP[r] += sum_b sum_i dW_b[i][r] (q_i + v) = (dW tile)^T (Q + v), the tile read MN-major straight from the scratch rows.
__global__ void __launch_bounds__(NT) attn_rel_bwd_pos_kernel(const AttnK a, const __nv_bfloat16* __restrict__ DWg, int bper) {
  extern __shared__ uint8_t attn_rel_smem[];
  uint8_t* smem = align1024(attn_rel_smem);
  uint8_t* sW = smem;                  // two blocks: [128 queries x 128 window columns]
  uint8_t* sQv = sW + 2 * TILE_BYTES;  // [128 queries x 64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sQv + TILE_BYTES);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, wq = warp & 3, hf = warp >> 2;
  const int u0 = blockIdx.x * 128, h = blockIdx.y;
  const int b_begin = blockIdx.z * bper, b_end = min(a.B, b_begin + bper);
  const int off = dw_off(a.Tk), RW = dw_pitch(a.Tk), T1 = a.Tk - 1;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
  uint32_t ph = 0;
  bool any = false;
  for (int b = b_begin; b < b_end; ++b) {
    const long long bh = (long long)b * a.H + h;
    for (int i0 = 0; i0 < a.Tq; i0 += QT) {
      // row i of the scratch holds its band at columns [lo_i, lo_i + wlen), lo_i = floor8(T-1-i+off); nothing else was written
      const int jend = active_keys(a, b, i0);
      if (jend <= 0) continue;
      const int wlen = round_up64(jend) + 8;
      const int ilast = min(i0 + QT - 1, a.Tq - 1);
      if (((T1 - ilast + off) & ~7) >= u0 + 128 || ((T1 - i0 + off) & ~7) + wlen <= u0) continue;  // uniform per CTA
      for (int idx = tid; idx < QT * 16; idx += NT) {
        const int r = idx >> 4, c16 = idx & 15;
        const int i = i0 + r, u = u0 + c16 * 8;
        const int lo = (T1 - i + off) & ~7;
        const bool ok = i < a.Tq && u >= lo && u < lo + wlen;
        cp_async16(sw_chunk(sW + (c16 >> 3) * TILE_BYTES, r, c16 & 7), DWg + (ok ? (bh * a.Tq + i) * (long long)RW + u : 0), ok);
      }
      load_tile(sQv, QT, a.q + (long long)b * a.Tq * a.ldq + h * 64, a.ldq, i0, 0, a.Tq, a.bv ? a.bv + h * 64 : nullptr);
      cp_async_wait_all();
      ATTN_SYNC_FOR_MMA();
      if (tid == 0) {
        mma_k128(tmem, smem_u32(sW), smem_u32(sQv), true, any);  // dP tile += dW^T (Q + v)
        umma_commit(&bar[0]);
      }
      mbar_wait(&bar[0], ph);
      tcgen05_fence_after();
      ph ^= 1u;
      any = true;
      tcgen05_fence_before();
    }
  }
  if (any) {
    float g[32];
    tcgen05_fence_after();
    tmem_row<32>(lane_base + (uint32_t)(hf * 32), g);
    float* stg = reinterpret_cast<float*>(smem);  // [128][SP] fp32 over the (finished) operand tiles
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c)
      reinterpret_cast<float4*>(stg + (tid & 127) * SP)[hf * 8 + c] = make_float4(g[4 * c], g[4 * c + 1], g[4 * c + 2], g[4 * c + 3]);
    __syncthreads();
    const int nrel = 2 * a.Tk - 1;
    for (int idx = tid; idx < 128 * 64; idx += NT) {
      const int rr = idx >> 6, d = idx & 63;
      const int r = u0 + rr - off;
      if (r >= 0 && r < nrel) atomicAdd(a.dp + (long long)r * (a.H * 64) + h * 64 + d, stg[rr * SP + d]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

constexpr int SMEM_FWD = 2 * TILE_BYTES + 2 * KTILE_BYTES + WTILE_BYTES + STAGE_BYTES + 64 + 1024;
constexpr int SMEM_BQ = 3 * TILE_BYTES + 2 * KTILE_BYTES + WTILE_BYTES + STAGE_BYTES + 4 * TILE_BYTES + 64 + 1024;
constexpr int SMEM_BKV = 6 * TILE_BYTES + 64 + 1024;
constexpr int SMEM_BP = 3 * TILE_BYTES + 64 + 1024;
static_assert(SMEM_BQ <= 227 * 1024, "query-side backward tile set exceeds the shared memory of one CTA");
static_assert(3 * TILE_BYTES >= STAGE_BYTES, "p-side output staging must fit over its operand tiles");

template <class K>
int set_smem(K kernel, int bytes) {
  SVSR_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return SVSR_OK;
}
int init_attrs() {
  static bool done = false;
  if (done) return SVSR_OK;
  int rc;
  if ((rc = set_smem(attn_rel_fwd_kernel<true, true>, SMEM_FWD))) return rc;
  if ((rc = set_smem(attn_rel_fwd_kernel<true, false>, SMEM_FWD))) return rc;
  if ((rc = set_smem(attn_rel_fwd_kernel<false, true>, SMEM_FWD))) return rc;
  if ((rc = set_smem(attn_rel_fwd_kernel<false, false>, SMEM_FWD))) return rc;
  if ((rc = set_smem(attn_rel_bwd_q_kernel<true, true>, SMEM_BQ))) return rc;
  if ((rc = set_smem(attn_rel_bwd_q_kernel<true, false>, SMEM_BQ))) return rc;
  if ((rc = set_smem(attn_rel_bwd_q_kernel<false, true>, SMEM_BQ))) return rc;
  if ((rc = set_smem(attn_rel_bwd_q_kernel<false, false>, SMEM_BQ))) return rc;
  if ((rc = set_smem(attn_rel_bwd_kv_kernel, SMEM_BKV))) return rc;
  if ((rc = set_smem(attn_rel_bwd_pos_kernel, SMEM_BP))) return rc;
  done = true;
  return SVSR_OK;
}

}  // namespace

size_t attention_rel_tc_scratch_bytes(int B, int H, int Tq, int Tk) {
  // P~ and dS: [B,H,Tq,roundup64(Tk)] each; dW (relative-position attention only, Tq == Tk): [B,H,Tq,dw_pitch(Tk)]
  return ((size_t)2 * B * H * Tq * round_up64(Tk) + (size_t)B * H * Tq * dw_pitch(Tk)) * sizeof(__nv_bfloat16);
}

int attention_rel_tc_fwd(const AttnK& k, cudaStream_t s) {
  int rc = init_attrs();
  if (rc) return rc;
  dim3 grid((k.Tq + QT - 1) / QT, k.H, k.B);
  const bool drop = k.drop_p > 0.f;
  if (k.p) {
    if (drop) attn_rel_fwd_kernel<true, true><<<grid, NT, SMEM_FWD, s>>>(k);
    else attn_rel_fwd_kernel<true, false><<<grid, NT, SMEM_FWD, s>>>(k);
  } else {
    if (drop) attn_rel_fwd_kernel<false, true><<<grid, NT, SMEM_FWD, s>>>(k);
    else attn_rel_fwd_kernel<false, false><<<grid, NT, SMEM_FWD, s>>>(k);
  }
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

int attention_rel_tc_bwd(const AttnK& k, void* scratch, cudaStream_t s) {
  int rc = init_attrs();
  if (rc) return rc;
  const int Tkp = round_up64(k.Tk);
  __nv_bfloat16* Pg = static_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* DSg = Pg + (size_t)k.B * k.H * k.Tq * Tkp;
  __nv_bfloat16* DWg = DSg + (size_t)k.B * k.H * k.Tq * Tkp;
  dim3 gq((k.Tq + QT - 1) / QT, k.H, k.B);
  const bool drop = k.drop_p > 0.f;
  if (k.p) {
    if (drop) attn_rel_bwd_q_kernel<true, true><<<gq, NT, SMEM_BQ, s>>>(k, Pg, DSg, DWg, Tkp);
    else attn_rel_bwd_q_kernel<true, false><<<gq, NT, SMEM_BQ, s>>>(k, Pg, DSg, DWg, Tkp);
  } else {
    if (drop) attn_rel_bwd_q_kernel<false, true><<<gq, NT, SMEM_BQ, s>>>(k, Pg, DSg, DWg, Tkp);
    else attn_rel_bwd_q_kernel<false, false><<<gq, NT, SMEM_BQ, s>>>(k, Pg, DSg, DWg, Tkp);
  }
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  dim3 gk((k.Tk + 127) / 128, k.H, k.B);
  attn_rel_bwd_kv_kernel<<<gk, NT, SMEM_BKV, s>>>(k, Pg, DSg, Tkp);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  if (k.p) {
    const int nu = (2 * k.Tk - 1 + dw_off(k.Tk) + 127) / 128;
    int chunks = 148 / (nu * k.H);  // about one CTA per SM; each sums its clips in TMEM before the atomics
    chunks = chunks < 1 ? 1 : (chunks > k.B ? k.B : chunks);
    const int bper = (k.B + chunks - 1) / chunks;
    dim3 gp(nu, k.H, (k.B + bper - 1) / bper);
    attn_rel_bwd_pos_kernel<<<gp, NT, SMEM_BP, s>>>(k, DWg, bper);
    note_launch();
    SVSR_CHECK_CUDA(cudaGetLastError());
  }
  return SVSR_OK;
}

}  // namespace svsr
