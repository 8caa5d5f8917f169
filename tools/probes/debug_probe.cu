// Hardware probe (developer tool, NOT part of libsvsr.so: tools/gpu_check.py builds it on demand into
// tools/probes/_build/libprobe.so and calls svsr_debug_rowshift from there): does a UMMA shared-memory descriptor
// whose start address is offset by a whole number of 128-byte rows (not a multiple of the 1024-byte swizzle atom)
// still address a SWIZZLE_128B tile correctly? This decides whether one halo tile in shared memory can serve all
// filter taps of a convolution (tap shift == row offset) instead of one TMA load per tap.
#include "../../syncvsr_b200/csrc/common.cuh"
#include "../../syncvsr_b200/csrc/tmap.h"

namespace svsr {

// mode bit0: 0 = K-major probe  (D[128,64] = A[shift:shift+128, 0:64] . B[64,64]^T)
//            1 = MN-major probe (D[128,64] = sum_k A[shift+k, 0:128]^T B[shift+k, 0:64], k < 128)
// mode bit1: put ((start_addr >> 7) & 7) into the descriptor's base_offset field
__global__ void __launch_bounds__(128, 1)
rowshift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out,
                      int shift, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // layout: A0 [256 rows x 128B] | A1 [256 x 128B] | B [256 x 128B] | barriers
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * 32768;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 3 * 32768);
  uint64_t* done = bar + 1;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool mn = mode & 1;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tptr, 64);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tptr;

  if (threadIdx.x == 0) {
    if (!mn) {
      mbar_expect_tx(bar, 32768 + 8192);
      tma_load_2d(sA, &tmA, bar, 0, 0);            // rows 0..127
      tma_load_2d(sA + 16384, &tmA, bar, 0, 128);  // rows 128..255
      tma_load_2d(sB, &tmB, bar, 0, 0);            // 64 rows
    } else {
      mbar_expect_tx(bar, 3 * 32768);
      for (int h = 0; h < 2; ++h) {
        tma_load_2d(sA + h * 16384, &tmA, bar, 0, h * 128);           // channels 0..63
        tma_load_2d(sA + 32768 + h * 16384, &tmA, bar, 64, h * 128);  // channels 64..127
        tma_load_2d(sB + h * 16384, &tmB, bar, 0, h * 128);
      }
    }
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    const uint32_t a_addr = smem_u32(sA) + shift * 128;
    const uint32_t b_addr = smem_u32(sB) + (mn ? shift * 128 : 0);
    uint64_t bo_a = (mode & 2) ? ((uint64_t)((a_addr >> 7) & 7) << 49) : 0;
    uint64_t bo_b = (mode & 2) ? ((uint64_t)((b_addr >> 7) & 7) << 49) : 0;
    if (!mn) {
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, (umma_smem_desc_sw128(a_addr, 16, 1024) | bo_a) + 2 * k,
                  (umma_smem_desc_sw128(b_addr, 16, 1024) | bo_b) + 2 * k, idesc, k != 0);
    } else {
      const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      for (int ks = 0; ks < 8; ++ks)
        umma_bf16(tmem, umma_smem_desc_sw128(a_addr + ks * 2048, 32768, 1024) | bo_a,
                  umma_smem_desc_sw128(b_addr + ks * 2048, 32768, 1024) | bo_b, idesc, ks != 0);
    }
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0);
  tcgen05_fence_after();
  const int r = warp * 32 + lane;
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[r * 64 + ch * 32 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int debug_rowshift(const void* a, const void* b, float* out, int shift, int mode, cudaStream_t stream) {
  // K-major: a is [256, 64] bf16, b is [64, 64]. MN-major: a is [256, 128], b is [256, 64].
  const bool mn = mode & 1;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {(uint64_t)(mn ? 128 : 64), 256};
    uint64_t strides[1] = {(uint64_t)(mn ? 256 : 128)};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&tmA, a, 2, dims, strides, box, nullptr, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {64, (uint64_t)(mn ? 256 : 64)};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, (uint32_t)(mn ? 128 : 64)};
    int rc = make_tmap_bf16(&tmB, b, 2, dims, strides, box, nullptr, true);
    if (rc) return rc;
  }
  const int smem_bytes = 3 * 32768 + 256 + 1024;
  SVSR_CHECK_CUDA(cudaFuncSetAttribute(rowshift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  rowshift_probe_kernel<<<1, 128, smem_bytes, stream>>>(tmA, tmB, out, shift, mode);
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // namespace svsr

extern "C" int svsr_debug_rowshift(const void* a, const void* b, float* out, int shift, int mode, void* stream) {
  return svsr::debug_rowshift(a, b, out, shift, mode, static_cast<cudaStream_t>(stream));
}
