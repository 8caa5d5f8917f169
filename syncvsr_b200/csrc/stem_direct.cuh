// The 3-D conv stem (lightning.py:49-50) without a materialised patch tensor: forward and weight gradient build the
// 7x7/s2 window rows in shared memory from a bf16 copy of the video (stem_direct.cu).
#pragma once
#include "common.cuh"

namespace svsr {

// frames H x W the direct kernels cover: even width, (H/2 * W/2) output pixels a multiple of the 16-pixel tile
bool stem_direct_supported(int H, int W);
// y0 bf16 [B, T, OH*OW, 64] = Conv3d(1, 64, (5,7,7), (1,2,2), (2,3,3))(video); video_bf16: [B, T, H, W] bf16;
// w_packed: bf16 [64, 320], column kt*64 + kh*8 + kw (kh, kw < 7, else zero); bn_stats: fp64 [2][64] (+=) or null
int stem_direct_fwd(const void* video_bf16, const __nv_bfloat16* w_packed, __nv_bfloat16* y0, double* bn_stats, int B,
                    int T, int H, int W, double algo_flops, cudaStream_t stream);
// out fp32 [320, ldo] (row kt*64 + kh*8 + kw, column = output channel) += window^T . dz; dz bf16 [B, T, OH*OW, 64]
int stem_direct_wgrad(const void* video_bf16, const __nv_bfloat16* dz, float* out, int ldo, int B, int T, int H, int W,
                      double algo_flops, cudaStream_t stream);

}  // namespace svsr
