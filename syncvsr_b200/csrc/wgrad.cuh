// Weight-gradient GEMM on tcgen05: D[m, n] = sum over pixels p of A[p (+tap), m] * B[p, n]
// Both operands are NHWC bf16 tensors whose contraction index (the pixel) is the strided dimension, i.e. they
// are "MN-major" UMMA operands: a 4-D TMA box of (64 channels x pixels) lands in shared memory exactly in the
// canonical MN-major SWIZZLE_128B layout, so no transpose of activations or gradients is ever materialised.
// The pixel range is split across CTAs (split-K); partial tiles are reduced into the fp32 output with red.add.
#pragma once
#include "common.cuh"

namespace svsr {

constexpr int WGRAD_MAX_TAPS = 16;

struct WgradProblem {
  // A side (rows of D): tensor [a_N, a_H, a_W, a_C], channels a_coff .. a_coff+a_cin, shifted by taps.
  const void* a = nullptr;
  int a_N = 0, a_H = 1, a_W = 1, a_C = 0, a_coff = 0, a_cin = 0;
  int a_stride = 1;  // A pixel coordinate = K-grid coordinate * a_stride + tap offset
  int ntaps = 1;
  int tap_dh[WGRAD_MAX_TAPS] = {0};
  int tap_dw[WGRAD_MAX_TAPS] = {0};
  // B side (columns of D): tensor [k_N, k_H, k_W, b_C], channels b_coff .. b_coff+n_cols; also defines the K grid.
  const void* b = nullptr;
  int b_C = 0, b_coff = 0, n_cols = 0;
  int k_N = 0, k_H = 1, k_W = 1;
  // D: fp32 [ntaps * a_cin, n_cols] with pitch ldo, row index = tap * a_cin + channel. Accumulated into (+=).
  float* out = nullptr;
  int ldo = 0;
  int m_valid = 0;  // rows of D actually written (0 = all ntaps * a_cin)
  double algo_flops = 0;  // algorithmic FLOPs for the profiler (0 = 2*pixels*ntaps*a_cin*n_cols)
  StepCtl ctl;            // device-resident skip predicate (generic kernel only)
};

int wgrad_launch(const WgradProblem& p, cudaStream_t stream);

// wgrad_halo.cu: single-load halo-tile kernel for 3x3 / stride-1 / 64 -> 64 channel problems (resnet.layer1);
// wgrad_launch() dispatches to it when wgrad_halo_matches() (SVSR_HALO_CONV=0 keeps the generic kernel).
bool wgrad_halo_matches(const WgradProblem& p);
int wgrad_halo_launch(const WgradProblem& p, cudaStream_t stream);

// wgrad_stem.cu: temporal-halo kernel for the stem's 5-tap / 64 -> 64 column weight gradient over [clips, frames,
// pixels, 64] rows; wgrad_launch() dispatches to it when wgrad_stem_matches() (SVSR_STEM_HALO=0 keeps the generic kernel).
bool wgrad_stem_matches(const WgradProblem& p);
int wgrad_stem_launch(const WgradProblem& p, cudaStream_t stream);

}  // namespace svsr
