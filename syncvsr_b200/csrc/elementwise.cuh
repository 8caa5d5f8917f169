// HBM-bound kernels of the LRW hot path: BatchNorm statistics / apply / backward, the stem's patch gather and
// BN+GELU+max-pool, spatial mean pooling, weight packing. NHWC bf16 activations, 8 channels (16 B) per thread,
// fp32 math, fp64 cross-block accumulation for the batch statistics.
#pragma once
#include "common.cuh"

namespace svsr {

// stem: videos fp32 [B,1,T,H,W] -> patches bf16 [B,T,OH*OW,64]; slot kh*8+kw = x[2oh+kh-3, 2ow+kw-3] (7x7, zero padded)
int stem_patch(const float* videos, __nv_bfloat16* patches, int B, int T, int H, int W, cudaStream_t s);

// per-channel sum / sum of squares of x[rows, C] accumulated (+=) into fp64 stats[0..C) and stats[C..2C)
int bn_stats(const __nv_bfloat16* x, long long rows, int C, double* stats, cudaStream_t s);
// finalize: mean/invstd/scale/shift (fp32 [4][C] in `coef`) + running stats update (momentum, unbiased variance)
// update_running: 1 = train (batch stats, update buffers), 0 = batch stats only, -1 = eval (use running stats)
int bn_finalize(const double* stats, long long rows, int C, const float* gamma, const float* beta, float eps,
                float momentum, float* running_mean, float* running_var, float* coef, int update_running,
                cudaStream_t s);
// out = act(x*scale+shift (+ res*rscale+rshift | + res)); coef/rcoef are the [4][C] blocks of bn_finalize.
// relu: 0 = identity, 1 = ReLU, 2 = Swish (LRS frontend)
int bn_apply(const __nv_bfloat16* x, const float* coef, const __nv_bfloat16* res, const float* rcoef, int relu,
             __nv_bfloat16* out, long long rows, int C, cudaStream_t s);
// self_mask = 1: the ReLU that follows this BN is recomputed from c (mask = c*scale+shift > 0), relu_ref unused.
// self_mask = 2: Swish follows: g = dout * swish'(c*scale+shift [+ sw_res*rscale+rshift | + sw_res]) (sw_rcoef optional).
// backward reductions: g = dout * (ref > 0 if ref) ; stats[0..C) += sum g ; stats[C..2C) += sum g * xhat
int bn_bwd_reduce(const __nv_bfloat16* dout, const __nv_bfloat16* relu_ref, const __nv_bfloat16* c, const float* coef,
                  long long rows, int C, double* stats, int self_mask, cudaStream_t s,
                  const __nv_bfloat16* sw_res = nullptr, const float* sw_rcoef = nullptr);
// dgamma += sum g*xhat ; dbeta += sum g ; kcoef[0..C) = sum g / rows ; kcoef[C..2C) = sum g*xhat / rows
int bn_bwd_finalize(const double* stats, long long rows, int C, float* dgamma, float* dbeta, float* kcoef,
                    cudaStream_t s);
// dc = scale * (g - k1 - xhat*k2); optionally also writes g (the relu-masked upstream gradient) to gmask_out
int bn_bwd_apply(const __nv_bfloat16* dout, const __nv_bfloat16* relu_ref, const __nv_bfloat16* c, const float* coef,
                 const float* kcoef, __nv_bfloat16* dc, __nv_bfloat16* gmask_out, long long rows, int C, int self_mask,
                 cudaStream_t s, const __nv_bfloat16* sw_res = nullptr, const float* sw_rcoef = nullptr,
                 // bn_bwd_finalize folded into this launch: the fp64 sums of bn_bwd_reduce ([2][C]) instead of kcoef; d gamma /
                 // d beta (+=) are then written here
                 const double* fused_stats = nullptr, float* dgamma = nullptr, float* dbeta = nullptr);

// stem epilogue: y0 [N,IH,IW,64] -> max_pool3x3s2p1(gelu(bn(y0))) [N,OH,OW,64] + argmax slot (uint8)
// swish = 1: Swish instead of GELU (LRS frontend3D, conv3d_extractor.py:31-38)
int stem_bn_gelu_pool(const __nv_bfloat16* y0, const float* coef, __nv_bfloat16* out, uint8_t* argmax, int N, int IH,
                      int IW, cudaStream_t s, int swish = 0);
// dz[N,IH,IW,64] = (scatter of dout through argmax) * gelu'(bn(y0))
int stem_pool_gelu_bwd(const __nv_bfloat16* dout, const uint8_t* argmax, const __nv_bfloat16* y0, const float* coef,
                       __nv_bfloat16* dz, int N, int IH, int IW, cudaStream_t s);

// Fused stem backward: the three passes above (pool scatter * GELU', BN reduce, BN apply) as two, dz never stored.
// dgamma/dbeta += ; dc = d loss / d y0 (bf16). dout is OVERWRITTEN (dout * gelu'(z_selected), the routed gradient). stats_scratch: fp64 [128], kcoef_scratch: fp32 [128].
int stem_bwd_fused(__nv_bfloat16* dout, const uint8_t* argmax, const __nv_bfloat16* y0, const float* coef,
                   float* dgamma, float* dbeta, __nv_bfloat16* dc, double* stats_scratch, float* kcoef_scratch, int N,
                   int IH, int IW, cudaStream_t s, int swish = 0);

// x_stream[b, t+1, :] = mean over HW of a[b*T+t, :, :]; x_stream[b, 0, :] = cls  (fp32 [B,T+1,C])
int meanpool_cls(const __nv_bfloat16* a, const float* cls, float* x_stream, int B, int T, int HW, int C,
                 cudaStream_t s, int ldx = 0);  // ldx: row pitch of x_stream (0 = C)
// dout[b*T+t, hw, :] = dx[b, t+1, :] / HW (bf16); dcls += sum_b dx[b,0,:]
int meanpool_cls_bwd(const float* dx, __nv_bfloat16* dout, float* dcls, int B, int T, int HW, int C, cudaStream_t s,
                     int ldx = 0);
// CutMix gather (augment.py:27-118): vout[i,t] = vin[vsrc[i,t], t]; aout[i,a] = ain[asrc[i,a], a]; soft labels and word
// masks mixed with (1 - rate, rate) for the clips flagged in `mixed` (tgt = the partner clip)
int cutmix_gather(const float* vin, float* vout, const int* vsrc, int B, int T, long long frame_elems,
                  const long long* ain, long long* aout, const int* asrc, int Ta, int G, const long long* labels,
                  const int* tgt, const float* rate, const unsigned char* mixed, float* soft, int num_labels,
                  const float* wm_in, float* wm_out, int Tw, cudaStream_t s);
// word-boundary channel (lightning.py:145-150): column C of the stream = word_mask[b,t] (frames) / cls[C] (CLS row)
int wb_column(float* xs, const float* cls, const float* wm, int B, int T, int ldx, int C, cudaStream_t s);
int wb_column_bwd(const float* dx, float* dcls, int B, int T, int ldx, int C, cudaStream_t s);  // dcls[C] += sum_b
// grad[N,K] += scratch[Np,Kp] (zero-padded weight-gradient scratch; glu = 1: gate rows live at Fp = ceil64(N/2))
int unpack_linear_wgrad(const float* tmp, float* grad, int N, int K, int Kp, int glu, cudaStream_t s);

// ---- weight packing (fp32 master -> bf16 operand layouts) and gradient unpacking ----
int pack_conv_weight(const float* w, __nv_bfloat16* w_fprop, __nv_bfloat16* w_dgrad, int Cout, int Cin, int R, int S,
                     cudaStream_t s);
int unpack_conv_wgrad(const float* d, float* grad, int Cout, int Cin, int R, int S, cudaStream_t s);
int pack_stem_weight(const float* w, __nv_bfloat16* wp, cudaStream_t s);  // [64,1,5,7,7] -> [64, 5*64]
int unpack_stem_wgrad(const float* d, float* grad, cudaStream_t s);        // [5*64, 64] -> [64,1,5,7,7] (+=)
int pack_linear_weight(const float* w, __nv_bfloat16* wb, __nv_bfloat16* wt, int N, int K, int ldb, int ldt,
                       cudaStream_t s);
// All weight packs of a step in ONE launch: a device table of jobs, blockIdx.y selects the job.
struct PackJob {
  const float* src;
  __nv_bfloat16* dst0;  // conv: fprop layout; linear: [N, ldb]; stem: [64, 320]
  __nv_bfloat16* dst1;  // conv: dgrad layout; linear: transposed [K, ldt] (may be null)
  int type;             // 0 = conv [Cout,Cin,R,S], 1 = linear [N,K], 2 = stem, 3 = fp32 vector pad copy (a = N,
                        // b = glu remap), 4 = GLU linear [2F,K] -> rows remapped to [2*ceil64(F), ldb]
  int a, b, c, d;       // conv: Cout, Cin, RS, -; linear: N, K, ldb, ldt
};
int pack_all_weights(const PackJob* jobs_dev, int njobs, cudaStream_t s);
// db[N] += column sums of dy[M, N] (bf16, pitch ld)
int colsum_bf16(const __nv_bfloat16* dy, int ld, float* db, int M, int N, cudaStream_t s, const StepCtl* ctl = nullptr);
// device-resident step control (common.cuh StepCtl): dst = src iff the sublayer's bit is set; write the control words
int copy_if_skipped(float* dst, const float* src, long long n, const StepCtl& ctl, cudaStream_t s);
int set_step_ctl(unsigned* skip, unsigned long long* seed, unsigned skip_mask, unsigned long long seed_value,
                 cudaStream_t s);
int cast_f32_to_bf16(const float* x, __nv_bfloat16* y, long long n, cudaStream_t s);
// last [B, T+1, D] fp32 -> cls rows [B, D] and frame rows [B*T, D], both bf16
int split_cast_last(const float* last, __nv_bfloat16* cls, __nv_bfloat16* frames, int B, int T, int D, cudaStream_t s);

}  // namespace svsr
