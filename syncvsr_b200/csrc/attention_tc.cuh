// Shared pieces of the tcgen05 attention kernels (attention_tc.cu: x-transformers core, n <= 64; attention_rel_tc.cu:
// Conformer rel-pos / decoder attention): 128-byte-swizzled [128 x 64] bf16 operand tiles written from registers, UMMA
// issue helpers for K-major and MN-major reads of the same tile, TMEM row reads.
#pragma once
#include "common.cuh"

namespace svsr {
namespace attn_tc {

constexpr int TC_D = 64;
constexpr int TILE_BYTES = 128 * 128;  // [128 rows x 64 bf16], SWIZZLE_128B
constexpr float LOG2E = 1.4426950408889634f;

// 16-byte chunk `c` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint8_t* sw_chunk(uint8_t* tile, int r, int c) { return tile + r * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ void unpack8(const uint4 u, float* f) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y, f[4] = c.x, f[5] = c.y, f[6] = d.x, f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]), u.y = pack_bf16x2(f[2], f[3]), u.z = pack_bf16x2(f[4], f[5]), u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// Loads one 64-wide bf16 row (or zeros), rotates the pairs (f, f + 16), f < 16, by the row's angles when `rotate`, and
// stores it as bf16 into row r of a swizzled tile. cs = this row's [16 cos | 16 sin].
__device__ __forceinline__ void load_rot_store(const __nv_bfloat16* __restrict__ src, bool valid, bool rotate,
                                               const float* cs, uint8_t* tile, int r) {
  float x[64];
  if (valid) {
    const uint4* p = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int c = 0; c < 8; ++c) unpack8(__ldg(p + c), x + 8 * c);
    if (rotate) {
#pragma unroll
      for (int f = 0; f < 16; ++f) {
        const float a = x[f], b = x[f + 16];
        x[f] = a * cs[f] - b * cs[16 + f];
        x[f + 16] = b * cs[f] + a * cs[16 + f];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 64; ++i) x[i] = 0.f;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(sw_chunk(tile, r, c)) = pack8(x + 8 * c);
}

// K-major operand whose K extent is one 128-byte row (64 elements): 4 MMAs of K = 16
__device__ __forceinline__ void mma_k64(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc) {
  const uint64_t a = umma_smem_desc_sw128(a_addr, 16, 1024), b = umma_smem_desc_sw128(b_addr, 16, 1024);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a + (uint64_t)(2 * k), b + (uint64_t)(2 * k), idesc, k != 0);
}
// D[128 x 64] = X[128 x 128 (K)] . Y[128 (K) x 64]: X = two-block tile (block kb = K columns [64 kb, 64 kb + 64)) read
// K-major, or -- a_mn -- the same tile read MN-major (D = X^T . Y, M = the tile's columns); Y MN-major (rows = K index)
// `accumulate`: add to what the accumulator already holds (a sum over several key / window blocks)
__device__ __forceinline__ void mma_k128(uint32_t d_tmem, uint32_t x_addr, uint32_t y_addr, bool a_mn,
                                         bool accumulate = false) {
  const uint32_t idesc = umma_idesc_bf16(128, 64, a_mn ? 1 : 0, 1);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint64_t a = a_mn ? umma_smem_desc_sw128(x_addr + ks * 2048, TILE_BYTES, 1024)
                            : umma_smem_desc_sw128(x_addr + (ks >> 2) * TILE_BYTES + (ks & 3) * 32, 16, 1024);
    const uint64_t b = umma_smem_desc_sw128(y_addr + ks * 2048, TILE_BYTES, 1024);
    umma_bf16(d_tmem, a, b, idesc, accumulate || ks != 0);
  }
}

template <int BS>
__device__ __forceinline__ void tmem_row(uint32_t taddr, float (&s)[BS]) {
#pragma unroll
  for (int c = 0; c < BS / 32; ++c) {
    uint32_t v[32];
    tmem_ld_32x32(taddr + (uint32_t)(c * 32), v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) s[c * 32 + j] = __uint_as_float(v[j]);
  }
}

}  // namespace attn_tc
}  // namespace svsr
