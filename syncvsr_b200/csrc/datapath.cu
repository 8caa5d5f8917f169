// Data path on the device: what the reference does per sample on CPU DataLoader workers
// (LRW/video/src/data.py:32-68 Dataset.__getitem__, data.py:156-171 transform pipelines) becomes, per BATCH,
//   1. jpeg_decode_gray: baseline grayscale JPEG -> u8 frames (data.py:41 `TurboJPEG.decode(img, TJPF_GRAY)`):
//      one thread per frame walks the Huffman-coded scan (frames are independent; 1856 frames per B=64 batch keep
//      the machine busy), one warp-parallel pass does dequantisation + libjpeg's integer "islow" inverse DCT, so
//      the pixels are bit-identical to libjpeg-turbo's default decoder;
//   2. video_transform: u8 [B,T,H,W] -> f32 [B,1,T,OH,OW] in one pass: x/255 -> horizontal flip -> crop + antialiased
//      bilinear resize (RandomResizedCrop / Resize / CenterCrop) -> TimeMask (fill with the clip mean) -> Normalize.
// All random decisions are drawn on the host in the reference's order (syncvsr_b200/data.py) and passed as tables.
// Host -> device traffic per clip drops from 4 bytes per pixel (f32 tensors from the workers) to the JPEG bytes.
#include "../../include/svsr.h"
#include "common.cuh"
#include <string.h>

namespace svsr {
namespace {

// ------------------------------------------------------------------------------------------------ transform ------
// ATen's antialiased bilinear weights (UpSampleKernel.cpp, _compute_indices_weights_aa with the triangle filter):
// scale = in/out, support = max(scale, 1), taps [xmin, xmin + xsize) with weights tri((j + xmin - center + 0.5) / max(scale,1))
// normalised to sum 1. For scale <= 1 this is ordinary align_corners=False bilinear interpolation.
struct AaTaps {
  int lo, n;
  float inv_total, center, invscale;
};

__device__ __forceinline__ AaTaps aa_taps(int i, int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.f ? scale : 1.f;
  AaTaps t;
  t.invscale = scale >= 1.f ? 1.f / scale : 1.f;
  t.center = scale * ((float)i + 0.5f);
  t.lo = max((int)(t.center - support + 0.5f), 0);
  t.n = min((int)(t.center + support + 0.5f), in_size) - t.lo;
  float total = 0.f;
  for (int j = 0; j < t.n; ++j) {
    const float x = fabsf(((float)(j + t.lo) - t.center + 0.5f) * t.invscale);
    total += x < 1.f ? 1.f - x : 0.f;
  }
  t.inv_total = total != 0.f ? 1.f / total : 0.f;
  return t;
}

__device__ __forceinline__ float aa_weight(const AaTaps& t, int j) {
  const float x = fabsf(((float)(j + t.lo) - t.center + 0.5f) * t.invscale);
  return (x < 1.f ? 1.f - x : 0.f) * t.inv_total;
}

// grid (ceil(T*OH*OW / 256), B); xf[b] = {flip, top, left, crop_h, crop_w, mask_t0, mask_t1, 0}
__global__ void __launch_bounds__(256)
video_transform_kernel(const uint8_t* __restrict__ frames, const int* __restrict__ xf, float* __restrict__ out,
                       double* __restrict__ clip_sum, int T, int H, int W, int OH, int OW, float mean, float stdv) {
  const int b = blockIdx.y;
  const int* x8 = xf + b * 8;
  const int flip = x8[0], top = x8[1], left = x8[2], ch = x8[3], cw = x8[4];
  const long long per_clip = (long long)T * OH * OW;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
  if (idx < per_clip) {
    const int ox = (int)(idx % OW), oy = (int)((idx / OW) % OH), t = (int)(idx / ((long long)OW * OH));
    const uint8_t* src = frames + ((long long)b * T + t) * H * W;
    const AaTaps ty = aa_taps(oy, ch, OH), tx = aa_taps(ox, cw, OW);
    float acc = 0.f;
    for (int jy = 0; jy < ty.n; ++jy) {
      const uint8_t* row = src + (long long)(top + ty.lo + jy) * W;
      float h = 0.f;  // horizontal pass first, like the separable ATen kernel
      for (int jx = 0; jx < tx.n; ++jx) {
        const int xc = left + tx.lo + jx;  // column of the (flipped) image
        const float p = __fdiv_rn((float)row[flip ? W - 1 - xc : xc], 255.0f);
        h += aa_weight(tx, jx) * p;
      }
      acc += aa_weight(ty, jy) * h;
    }
    v = acc;
    out[(long long)b * per_clip + idx] = __fdiv_rn(v - mean, stdv);
  }
  if (clip_sum) {  // sum of the un-normalised clip for TimeMask's `cloned.mean()` (augment.py:141)
    float s = warp_sum(v);
    __shared__ float ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < 8; ++i) tot += ws[i];
      atomicAdd(clip_sum + b, (double)tot);
    }
  }
}

// frames [mask_t0, mask_t1) of clip b <- Normalize(mean of the clip)
__global__ void __launch_bounds__(256)
video_timemask_kernel(const int* __restrict__ xf, float* __restrict__ out, const double* __restrict__ clip_sum, int T,
                      int OH, int OW, float mean, float stdv) {
  const int b = blockIdx.y;
  const int m0 = xf[b * 8 + 5], m1 = xf[b * 8 + 6];
  const long long frame = (long long)OH * OW, n = (long long)(m1 - m0) * frame;
  if (n <= 0) return;
  const float fill = __fdiv_rn((float)(clip_sum[b] / (double)((long long)T * frame)) - mean, stdv);
  float* dst = out + ((long long)b * T + m0) * frame;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = fill;
}


// ------------------------------------------------------------------------------------------------ JPEG -----------
// Frame descriptor (int32 x SVSR_JPEG_DESC_INTS), written by svsr_jpeg_parse on the host:
//   [0] scan offset in the blob  [1] scan bytes  [2] width  [3] height  [4] restart interval (MCUs, 0 = none)
//   [5] components in the scan (1 or 3)  [6 + 5c ..] per component: h, v, quant table, DC table, AC table (pool indices)
// Only component 0 (luminance) is reconstructed -- TJPF_GRAY output of a YCbCr file is its Y plane -- the chroma blocks
// are entropy-decoded (their codes must be consumed) and dropped.
constexpr int JD = SVSR_JPEG_DESC_INTS;
constexpr int HT_BYTES = SVSR_JPEG_HUFF_BYTES;
// Huffman table image: int32 maxcode[18] | int32 valoff[17] | u8 vals[256] | u16 look[512] (9-bit lookahead: len<<8 | symbol)
constexpr int HT_VALOFF = 72, HT_VALS = 140, HT_LOOK = 396;

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  unsigned long long buf;  // valid bits are the top `cnt`
  int cnt;
  bool marker;  // a marker was reached: feed zero bits from here on (libjpeg's behaviour on truncated data)
  __device__ void fill() {
    // fast path: four bytes at once when none of them is 0xFF (no stuffing / marker handling needed); the four loads are
    // independent, so one L1 latency covers 32 bits instead of 8
    if (cnt > 32) return;  // every decode step needs at most 16 + 15 bits
    if (!marker && p + 4 <= end) {
      const unsigned b0 = p[0], b1 = p[1], b2 = p[2], b3 = p[3];
      if (b0 != 0xFF && b1 != 0xFF && b2 != 0xFF && b3 != 0xFF) {
        const unsigned w = (b0 << 24) | (b1 << 16) | (b2 << 8) | b3;
        buf |= (unsigned long long)w << (32 - cnt);
        cnt += 32, p += 4;
        return;
      }
    }
    while (cnt <= 56) {
      unsigned b = 0;
      if (!marker && p < end) {
        b = *p;
        if (b == 0xFF) {
          const unsigned b2 = (p + 1 < end) ? p[1] : 0xD9;
          if (b2 == 0) p += 2;  // stuffed zero
          else marker = true, b = 0;
        } else {
          ++p;
        }
      }
      buf |= (unsigned long long)b << (56 - cnt);
      cnt += 8;
    }
  }
  __device__ unsigned peek(int n) const { return (unsigned)(buf >> (64 - n)); }
  __device__ void skip(int n) { buf <<= n, cnt -= n; }
  __device__ int receive_extend(int s) {  // s in 1..16: Figure F.12 of the JPEG standard
    const int r = (int)peek(s);
    skip(s);
    return r < (1 << (s - 1)) ? r - (1 << s) + 1 : r;
  }
  __device__ void restart() {  // byte-align, step over the RSTn marker
    buf = 0, cnt = 0;
    if (p + 1 < end && p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7) p += 2;
    marker = false;
  }
};

__device__ __forceinline__ int huff_decode(BitReader& br, const uint8_t* ht) {
  const unsigned e = reinterpret_cast<const uint16_t*>(ht + HT_LOOK)[br.peek(9)];
  if (e) {
    br.skip((int)(e >> 8));
    return (int)(e & 255u);
  }
  const int* maxcode = reinterpret_cast<const int*>(ht);
  const int* valoff = reinterpret_cast<const int*>(ht + HT_VALOFF);
  int l = 10;
  int code = (int)br.peek(10);
  while (l <= 16 && code > maxcode[l]) ++l, code = (int)br.peek(l);
  if (l > 16) return 0;  // corrupt code: libjpeg substitutes a zero symbol
  br.skip(l);
  return ht[HT_VALS + ((code + valoff[l]) & 255)];
}

// One thread per frame: entropy-decode every block of the scan, keep the luminance coefficients (natural order).
// coefs: int16 [n][blocks_per_frame][64]
__global__ void __launch_bounds__(32)
jpeg_huffman_kernel(const uint8_t* __restrict__ blob, const int* __restrict__ desc, int n, const uint8_t* __restrict__ htabs,
                    short* __restrict__ coefs, int blocks_w, int blocks_h) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const int* d = desc + (long long)f * JD;
  BitReader br;
  br.p = blob + d[0], br.end = br.p + d[1], br.buf = 0, br.cnt = 0, br.marker = false;
  const int W = d[2], H = d[3], ri = d[4], nc = d[5];
  int hs[3], vs[3], pred[3] = {0, 0, 0};
  const uint8_t *dct[3], *act[3];
  int hmax = 1, vmax = 1;
  for (int c = 0; c < nc; ++c) {
    hs[c] = nc == 1 ? 1 : d[6 + 5 * c], vs[c] = nc == 1 ? 1 : d[7 + 5 * c];
    dct[c] = htabs + (long long)d[9 + 5 * c] * HT_BYTES, act[c] = htabs + (long long)d[10 + 5 * c] * HT_BYTES;
    hmax = max(hmax, hs[c]), vmax = max(vmax, vs[c]);
  }
  const int mcus_x = (W + 8 * hmax - 1) / (8 * hmax), mcus_y = (H + 8 * vmax - 1) / (8 * vmax);
  short* out = coefs + (long long)f * blocks_w * blocks_h * 64;
  int until_restart = ri;
  for (int my = 0; my < mcus_y; ++my)
    for (int mx = 0; mx < mcus_x; ++mx) {
      if (ri) {
        if (until_restart == 0) {
          br.restart();
          pred[0] = pred[1] = pred[2] = 0;
          until_restart = ri;
        }
        --until_restart;
      }
      for (int c = 0; c < nc; ++c)
        for (int by = 0; by < vs[c]; ++by)
          for (int bx = 0; bx < hs[c]; ++bx) {
            short blk[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) blk[i] = 0;
            br.fill();
            int s = huff_decode(br, dct[c]);
            if (s) br.fill(), pred[c] += br.receive_extend(s);
            blk[0] = (short)pred[c];
            for (int k = 1; k < 64;) {
              br.fill();
              const int rs = huff_decode(br, act[c]);
              const int r = rs >> 4;
              s = rs & 15;
              if (s) {
                k += r;
                const int v = br.receive_extend(s);
                blk[c_zigzag[k & 63]] = (short)v;
                ++k;
              } else {
                if (r != 15) break;
                k += 16;
              }
            }
            if (c == 0) {
              const int gx = mx * hs[0] + bx, gy = my * vs[0] + by;
              if (gx < blocks_w && gy < blocks_h) {
                uint4* o = reinterpret_cast<uint4*>(out + ((long long)gy * blocks_w + gx) * 64);
                const uint4* b4 = reinterpret_cast<const uint4*>(blk);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = b4[i];
              }
            }
          }
    }
}

// libjpeg's accurate integer inverse DCT (jidctint.c "islow": Loeffler-Ligtenberg-Moschytz, CONST_BITS 13, PASS1_BITS 2)
__device__ __forceinline__ void idct_islow_1d(const int (&in)[8], int (&o)[8], int shift) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 + z3 * (-15137);
  int tmp3 = z1 + z2 * 6270;
  z2 = in[0], z3 = in[4];
  int tmp0 = (z2 + z3) << 13;
  int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7], tmp1 = in[5], tmp2 = in[3], tmp3 = in[1];
  z1 = tmp0 + tmp3, z2 = tmp1 + tmp2, z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446, tmp1 *= 16819, tmp2 *= 25172, tmp3 *= 12299;
  z1 *= -7373, z2 *= -20995, z3 *= -16069, z4 *= -3196;
  z3 += z5, z4 += z5;
  tmp0 += z1 + z3, tmp1 += z2 + z4, tmp2 += z2 + z3, tmp3 += z1 + z4;
  const int rnd = 1 << (shift - 1);
  o[0] = (tmp10 + tmp3 + rnd) >> shift, o[7] = (tmp10 - tmp3 + rnd) >> shift;
  o[1] = (tmp11 + tmp2 + rnd) >> shift, o[6] = (tmp11 - tmp2 + rnd) >> shift;
  o[2] = (tmp12 + tmp1 + rnd) >> shift, o[5] = (tmp12 - tmp1 + rnd) >> shift;
  o[3] = (tmp13 + tmp0 + rnd) >> shift, o[4] = (tmp13 - tmp0 + rnd) >> shift;
}

__device__ __forceinline__ unsigned range_limit(int x) {  // sample_range_limit[(x & RANGE_MASK)] with the IDCT centring
  const int m = x & 1023;
  return m < 128 ? (unsigned)(m + 128) : m < 512 ? 255u : m < 896 ? 0u : (unsigned)(m - 896);
}

// One thread per luminance block: dequantise, 2-D islow IDCT, range-limit, store the part inside the image.
__global__ void __launch_bounds__(128)
jpeg_idct_kernel(const short* __restrict__ coefs, const int* __restrict__ desc, const unsigned short* __restrict__ qtabs,
                 uint8_t* __restrict__ out, int n, int blocks_w, int blocks_h, int W, int H) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = blocks_w * blocks_h;
  if (gid >= (long long)n * per) return;
  const int f = (int)(gid / per), b = (int)(gid - (long long)f * per);
  const int by = b / blocks_w, bx = b - by * blocks_w;
  const unsigned short* q = qtabs + (long long)desc[(long long)f * JD + 8] * 64;
  const short* cf = coefs + gid * 64;
  int ws[64];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const uint4 raw = reinterpret_cast<const uint4*>(cf)[r];
    const short* s8 = reinterpret_cast<const short*>(&raw);
#pragma unroll
    for (int c = 0; c < 8; ++c) ws[r * 8 + c] = (int)s8[c] * (int)q[r * 8 + c];
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {  // pass 1: columns
    int in[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) in[r] = ws[r * 8 + c];
    idct_islow_1d(in, o, 13 - 2);
#pragma unroll
    for (int r = 0; r < 8; ++r) ws[r * 8 + c] = o[r];
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {  // pass 2: rows
    int in[8], o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) in[c] = ws[r * 8 + c];
    idct_islow_1d(in, o, 13 + 2 + 3);
    const int y = by * 8 + r;
    if (y < H) {
      uint8_t* dst = out + ((long long)f * H + y) * W + bx * 8;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (bx * 8 + c < W) dst[c] = (uint8_t)range_limit(o[c]);
    }
  }
}

}  // namespace
}  // namespace svsr

using namespace svsr;

// ---------------------------------------------------------------------------------------- JPEG host-side parser --
namespace {

struct HuffSpec {
  uint8_t bits[17];
  uint8_t vals[256];
  int nvals;
};

void build_huff_image(const HuffSpec& h, uint8_t* img) {
  memset(img, 0, HT_BYTES);
  int* maxcode = reinterpret_cast<int*>(img);
  int* valoff = reinterpret_cast<int*>(img + HT_VALOFF);
  uint16_t* look = reinterpret_cast<uint16_t*>(img + HT_LOOK);
  memcpy(img + HT_VALS, h.vals, 256);
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    valoff[l] = k - code;
    for (int i = 0; i < h.bits[l]; ++i, ++k, ++code)
      if (l <= 9)
        for (int e = 0; e < (1 << (9 - l)); ++e) look[(code << (9 - l)) + e] = (uint16_t)((l << 8) | h.vals[k]);
    maxcode[l] = h.bits[l] ? code - 1 : -1;
    code <<= 1;
  }
  maxcode[17] = 0x7fffffff;
}

template <class T>
int pool_index(T* pool, int* count, int cap, const T& item) {
  for (int i = 0; i < *count; ++i)
    if (memcmp(&pool[i], &item, sizeof(T)) == 0) return i;
  if (*count >= cap) return -1;
  pool[*count] = item;
  return (*count)++;
}

struct QTab {
  uint16_t q[64];
};
struct HImg {
  uint8_t b[HT_BYTES];
};

}  // namespace

static const uint8_t h_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

extern "C" {

int svsr_video_transform(const uint8_t* frames, const int* xform, float* out, double* clip_sum, int B, int T, int H, int W,
                         int OH, int OW, float mean, float stdv, int time_mask, void* stream) {
  SVSR_REQUIRE(frames && xform && out, "video_transform: null pointer");
  SVSR_REQUIRE(B > 0 && T > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "video_transform: empty geometry");
  SVSR_REQUIRE(!time_mask || clip_sum, "video_transform: TimeMask needs the clip_sum scratch (B doubles)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (time_mask) SVSR_CHECK_CUDA(cudaMemsetAsync(clip_sum, 0, sizeof(double) * B, st));
  const long long per_clip = (long long)T * OH * OW;
  dim3 grid((unsigned)((per_clip + 255) / 256), (unsigned)B);
  video_transform_kernel<<<grid, 256, 0, st>>>(frames, xform, out, time_mask ? clip_sum : nullptr, T, H, W, OH, OW, mean,
                                               stdv);
  note_launch();
  if (time_mask) {
    video_timemask_kernel<<<dim3(32, (unsigned)B), 256, 0, st>>>(xform, out, clip_sum, T, OH, OW, mean, stdv);
    note_launch();
  }
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

int svsr_jpeg_parse(const uint8_t* blob, const int64_t* offsets, int n, int32_t* desc, uint16_t* qtabs, int qcap, int* n_q,
                    uint8_t* htabs, int hcap, int* n_h) {
  SVSR_REQUIRE(blob && offsets && desc && qtabs && n_q && htabs && n_h, "jpeg_parse: null pointer");
  QTab* qpool = reinterpret_cast<QTab*>(qtabs);
  HImg* hpool = reinterpret_cast<HImg*>(htabs);
  *n_q = 0, *n_h = 0;
  // consecutive frames almost always repeat the same tables: remember the raw DHT payload last seen in each slot
  uint8_t last_raw[8][17 + 256];
  int last_len[8] = {0, 0, 0, 0, 0, 0, 0, 0}, last_idx[8];
  for (int f = 0; f < n; ++f) {
    const uint8_t* p = blob + offsets[f];
    const uint8_t* end = blob + offsets[f + 1];
    SVSR_REQUIRE(end - p >= 4 && p[0] == 0xFF && p[1] == 0xD8, "jpeg_parse: frame %d does not start with SOI", f);
    SVSR_REQUIRE(offsets[f + 1] < (int64_t)1 << 31, "jpeg_parse: blob larger than 2 GiB");
    p += 2;
    int qidx[4] = {-1, -1, -1, -1}, dcidx[4] = {-1, -1, -1, -1}, acidx[4] = {-1, -1, -1, -1};
    int W = 0, H = 0, ncomp_frame = 0, ri = 0;
    int comp_id[3], comp_h[3], comp_v[3], comp_q[3];
    int32_t* d = desc + (int64_t)f * JD;
    bool done = false;
    while (!done) {
      SVSR_REQUIRE(p + 4 <= end && p[0] == 0xFF, "jpeg_parse: frame %d: marker expected", f);
      while (p < end && p[1] == 0xFF) ++p;  // fill bytes
      const int m = p[1];
      const int len = (p[2] << 8) | p[3];
      const uint8_t* seg = p + 4;
      const uint8_t* seg_end = p + 2 + len;
      SVSR_REQUIRE(seg_end <= end, "jpeg_parse: frame %d: truncated segment", f);
      if (m == 0xDB) {  // DQT
        while (seg < seg_end) {
          const int pq = seg[0] >> 4, tq = seg[0] & 15;
          SVSR_REQUIRE(tq < 4, "jpeg_parse: frame %d: bad quantisation table id", f);
          QTab t;
          ++seg;
          for (int i = 0; i < 64; ++i) {
            t.q[h_zigzag[i]] = pq ? (uint16_t)((seg[0] << 8) | seg[1]) : seg[0];
            seg += pq ? 2 : 1;
          }
          qidx[tq] = pool_index(qpool, n_q, qcap, t);
          SVSR_REQUIRE(qidx[tq] >= 0, "jpeg_parse: more than %d distinct quantisation tables", qcap);
        }
      } else if (m == 0xC4) {  // DHT
        while (seg < seg_end) {
          const int tc = seg[0] >> 4, th = seg[0] & 15;
          SVSR_REQUIRE(tc < 2 && th < 4, "jpeg_parse: frame %d: bad Huffman table id", f);
          HuffSpec h;
          memset(&h, 0, sizeof(h));
          int total = 0;
          for (int l = 1; l <= 16; ++l) h.bits[l] = seg[l], total += seg[l];
          SVSR_REQUIRE(total <= 256 && seg + 17 + total <= seg_end, "jpeg_parse: frame %d: bad Huffman table", f);
          const int slot = tc * 4 + th, raw_len = 17 + total;
          int idx;
          if (last_len[slot] == raw_len && memcmp(last_raw[slot], seg, raw_len) == 0) {
            idx = last_idx[slot];
          } else {
            memcpy(h.vals, seg + 17, total);
            h.nvals = total;
            HImg img;
            build_huff_image(h, img.b);
            idx = pool_index(hpool, n_h, hcap, img);
            SVSR_REQUIRE(idx >= 0, "jpeg_parse: more than %d distinct Huffman tables", hcap);
            memcpy(last_raw[slot], seg, raw_len);
            last_len[slot] = raw_len, last_idx[slot] = idx;
          }
          seg += raw_len;
          (tc ? acidx : dcidx)[th] = idx;
        }
      } else if (m == 0xC0 || m == 0xC1) {  // SOF0 / SOF1: sequential Huffman
        SVSR_REQUIRE(seg[0] == 8, "jpeg_parse: frame %d: only 8-bit samples are supported", f);
        H = (seg[1] << 8) | seg[2], W = (seg[3] << 8) | seg[4];
        ncomp_frame = seg[5];
        SVSR_REQUIRE(ncomp_frame == 1 || ncomp_frame == 3, "jpeg_parse: frame %d: %d components unsupported", f, ncomp_frame);
        for (int c = 0; c < ncomp_frame; ++c)
          comp_id[c] = seg[6 + 3 * c], comp_h[c] = seg[7 + 3 * c] >> 4, comp_v[c] = seg[7 + 3 * c] & 15,
          comp_q[c] = seg[8 + 3 * c];
      } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
        SVSR_REQUIRE(false, "jpeg_parse: frame %d: progressive / lossless / arithmetic JPEG (SOF%d) is not supported", f,
                     m - 0xC0);
      } else if (m == 0xDD) {  // DRI
        ri = (seg[0] << 8) | seg[1];
      } else if (m == 0xDA) {  // SOS
        const int ns = seg[0];
        SVSR_REQUIRE(W > 0 && ns == ncomp_frame, "jpeg_parse: frame %d: one scan with all components expected", f);
        d[2] = W, d[3] = H, d[4] = ri, d[5] = ns;
        for (int c = 0; c < ns; ++c) {
          const int cid = seg[1 + 2 * c], td = seg[2 + 2 * c] >> 4, ta = seg[2 + 2 * c] & 15;
          SVSR_REQUIRE(cid == comp_id[c], "jpeg_parse: frame %d: scan component order differs from the frame header", f);
          SVSR_REQUIRE(td < 4 && ta < 4 && dcidx[td] >= 0 && acidx[ta] >= 0 && comp_q[c] < 4 && qidx[comp_q[c]] >= 0,
                       "jpeg_parse: frame %d: scan refers to an undefined table", f);
          d[6 + 5 * c] = comp_h[c], d[7 + 5 * c] = comp_v[c], d[8 + 5 * c] = qidx[comp_q[c]];
          d[9 + 5 * c] = dcidx[td], d[10 + 5 * c] = acidx[ta];
        }
        d[0] = (int32_t)(seg_end - blob), d[1] = (int32_t)(end - seg_end);
        done = true;
      }
      p = seg_end;
    }
  }
  return SVSR_OK;
}

int svsr_jpeg_decode_gray(const uint8_t* blob_dev, const int32_t* desc_dev, int n, const uint16_t* qtabs_dev,
                          const uint8_t* htabs_dev, int16_t* coef_scratch, uint8_t* out, int W, int H, int blocks_w,
                          int blocks_h, void* stream) {
  SVSR_REQUIRE(blob_dev && desc_dev && qtabs_dev && htabs_dev && coef_scratch && out, "jpeg_decode_gray: null pointer");
  SVSR_REQUIRE(n > 0 && W > 0 && H > 0 && blocks_w * 8 >= W && blocks_h * 8 >= H, "jpeg_decode_gray: bad geometry");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // One frame per warp (a single active lane): every frame is a serial bit-stream walk whose length differs from its
  // neighbours', so lanes sharing a warp would serialise each other's branches (measured: 8 frames per warp executed
  // 4.1 of 8 lanes per instruction and took 2x longer); 1856 independent warps keep ~3 resident per scheduler.
  jpeg_huffman_kernel<<<n, 1, 0, st>>>(blob_dev, desc_dev, n, htabs_dev, coef_scratch, blocks_w, blocks_h);
  note_launch();
  const long long nblk = (long long)n * blocks_w * blocks_h;
  jpeg_idct_kernel<<<(unsigned)((nblk + 127) / 128), 128, 0, st>>>(coef_scratch, desc_dev, qtabs_dev, out, n, blocks_w,
                                                                  blocks_h, W, H);
  note_launch();
  SVSR_CHECK_CUDA(cudaGetLastError());
  return SVSR_OK;
}

}  // extern "C"
