// deform_conv.cu -- DeformableConvolution forward / backward (operator_cxx/deformable_convolution-inl.h:91-233): deformable
// sampling (nn/deformable_im2col.cuh:216-262, bilinear :77-113) followed by a per-group GEMM W[g].col[g].
// Compiled with -fmad=false so the bilinear arithmetic is bit-comparable with oracle/oracle_c.c and with the reference's own
// kernels compiled the same way (the libref_deform.so checker, tests/test_gpu_refpin.py).
//
// Work decomposition (NOT the reference's thread-per-column-element): the four tap offsets and bilinear weights of a sample
// depend on (deformable group, kernel tap, output position) only, so a CTA evaluates that SAMPLE TABLE once into shared
// memory and streams the group's channels through it.  Three forms of the forward:
//   NCHW fp32 in  -> col fp32 [C*kh*kw, Ho*Wo]          the reference's column buffer (rn_deform_im2col; fp32 GEMM path)
//   NHWC fp32 in  -> colT fp16 [Ho*Wo, kh*kw*C]         tap-major K: the 8 channels a thread samples are 16 contiguous bytes of
//   NHWC bf16 in  -> colT fp16 (the trunk's layout)     the K-major B operand of the tcgen05 GEMM (gemm_tc.cu), bias / relu fused
// Measured at 512 -> 512, 3x3 dilated, 38 x 63 (11.3 GFLOP / layer): sampler 19.5 us (16-position tiles: 600 CTAs) + GEMM =
// 48 us per layer = 235 TFLOP/s (round 1: 163 us).  The 22 MB fp16 column tile still round-trips HBM/L2 once each way; gathering
// the K-slab straight into the SWIZZLE_128B tile that feeds tcgen05.mma (true implicit GEMM) is the next step.
#include "common.cuh"
#include "gemm_tc.cuh"
#include <cuda_bf16.h>

namespace rn {

// Sampling geometry of one (deformable group, kernel tap, output position): the four tap offsets into a channel plane and
// the four bilinear weights, or all-zero when the sample falls outside the image.  It does not depend on the channel: a CTA
// evaluates it ONCE per (group, tap, position) into shared memory and streams the group's C/dg channels through it (the
// reference recomputes it, and re-reads the two offsets, for every channel).  Arithmetic of the entries:
// nn/deformable_im2col.cuh:77-113 (bilinear) + :240-254 (position / validity), operation for operation.
struct DcSample { int o1, o2, o3, o4; float w1, w2, w3, w4; };

struct DcGeom { int H, W, kh, kw, pad_h, pad_w, sh, sw, dil_h, dil_w, Ho, Wo; };

__device__ __forceinline__ DcSample dc_sample(const DcGeom& q, const float* __restrict__ off_g, int tap, int pos) {
  const int w_col = pos % q.Wo, h_col = pos / q.Wo, i = tap / q.kw, j = tap - i * q.kw;
  const int h_in = h_col * q.sh - q.pad_h, w_in = w_col * q.sw - q.pad_w;
  const float oh = __ldg(off_g + ((size_t)(2 * tap) * q.Ho + h_col) * q.Wo + w_col);
  const float ow = __ldg(off_g + ((size_t)(2 * tap + 1) * q.Ho + h_col) * q.Wo + w_col);
  DcSample t = {0, 0, 0, 0, 0.f, 0.f, 0.f, 0.f};
  const float h_im = h_in + i * q.dil_h + oh, w_im = w_in + j * q.dil_w + ow;
  if (h_im >= 0 && w_im >= 0 && h_im < q.H && w_im < q.W) {
    // bilinear in the window that starts at (h_in, w_in): coordinates relative to it, extent (H - h_in, W - w_in)
    float h = i * q.dil_h + oh, w = j * q.dil_w + ow;
    const int height = q.H - h_in, width = q.W - w_in;
    int h_low = (int)floorf(h), w_low = (int)floorf(w), h_high, w_high;
    if (h_low >= height - 1) { h_high = h_low = height - 1; h = (float)h_low; } else h_high = h_low + 1;
    if (w_low >= width - 1) { w_high = w_low = width - 1; w = (float)w_low; } else w_high = w_low + 1;
    const float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
    t.o1 = (h_in + h_low) * q.W + (w_in + w_low); t.o2 = (h_in + h_low) * q.W + (w_in + w_high);
    t.o3 = (h_in + h_high) * q.W + (w_in + w_low); t.o4 = (h_in + h_high) * q.W + (w_in + w_high);
    t.w1 = hh * hw; t.w2 = hh * lw; t.w3 = lh * hw; t.w4 = lh * lw;
  }
  return t;
}

constexpr int kDcTile = 64;                  // output positions per CTA (NCHW form and the backward kernels)
// channels-last forms: 16 positions per CTA -> 4x the CTAs (600 at 38 x 63, 4 per SM).  The walk is a chain of table lookup ->
// four dependent 16-byte taps per item, i.e. latency bound: at 64 positions the grid was 152 CTAs = 8 warps per SM and the
// sampler took 31 us for 22 MB of output (profiles/r02_launches_deform.csv)
constexpr int kDcTileCl = 16;

// MODE 0: data NCHW fp32 -> col fp32 [C*kh*kw, Ho*Wo] (K index c*kh*kw + tap: the reference's column buffer, bit-comparable)
// MODE 1: data NHWC fp32 / MODE 2: data NHWC bf16 -> colT fp16 [Ho*Wo, ldk] with K index tap*C + c (tap-major: the 8
//         channels a thread samples are 16 contiguous bytes of the GEMM's K-major B operand; the weight is packed the same way)
// grid (position tiles, deformable groups); dynamic smem = kh*kw*kDcTile table entries
template <int MODE>
__global__ void __launch_bounds__(256) deform_sample_kernel(DcGeom q, int C, int cpg, const void* __restrict__ im_,
                                                            const float* __restrict__ off, float* __restrict__ col,
                                                            __half* __restrict__ colT, int ldk) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  DcSample* tab = reinterpret_cast<DcSample*>(dc_smem);
  const int taps = q.kh * q.kw, Nsp = q.Ho * q.Wo;
  constexpr int kTile = MODE == 0 ? kDcTile : kDcTileCl;
  const int p0 = blockIdx.x * kTile, g = blockIdx.y;
  const int npos = min(kTile, Nsp - p0);
  const float* off_g = off + (size_t)g * 2 * taps * Nsp;
  for (int e = threadIdx.x; e < taps * kTile; e += blockDim.x) {
    const int tap = e / kTile, px = e - tap * kTile;
    if (px < npos) tab[e] = dc_sample(q, off_g, tap, p0 + px);
  }
  __syncthreads();
  if (MODE == 0) {
    const float* im = reinterpret_cast<const float*>(im_);
    const int px = threadIdx.x % kDcTile;                  // consecutive threads: consecutive positions (coalesced col rows)
    if (px >= npos) return;
    for (int cl = threadIdx.x / kDcTile; cl < cpg; cl += blockDim.x / kDcTile) {
      const int c = g * cpg + cl;
      const float* d = im + (size_t)c * q.H * q.W;
      float* dst = col + ((size_t)c * taps) * Nsp + p0 + px;
      for (int tap = 0; tap < taps; ++tap) {
        const DcSample t = tab[tap * kTile + px];
        dst[(size_t)tap * Nsp] = t.w1 * __ldg(d + t.o1) + t.w2 * __ldg(d + t.o2) + t.w3 * __ldg(d + t.o3) + t.w4 * __ldg(d + t.o4);
      }
    }
  } else {
    const int nv = cpg / 8;                                // 8-channel vectors of the group
    for (int item = threadIdx.x; item < npos * taps * nv; item += blockDim.x) {
      const int cv = item % nv, pt = item / nv, tap = pt % taps, px = pt / taps;
      const DcSample t = tab[tap * kTile + px];
      const int c0 = g * cpg + cv * 8;
      float v[4][8];
      const int o[4] = {t.o1, t.o2, t.o3, t.o4};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == 1) {
          const float4* s4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(im_) + (size_t)o[k] * C + c0);
          const float4 a = __ldg(s4), b = __ldg(s4 + 1);
          v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w; v[k][4] = b.x; v[k][5] = b.y; v[k][6] = b.z; v[k][7] = b.w;
        } else {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(im_) + (size_t)o[k] * C + c0));
          const __nv_bfloat162* pu = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int z = 0; z < 4; ++z) { const float2 f = __bfloat1622float2(pu[z]); v[k][2 * z] = f.x; v[k][2 * z + 1] = f.y; }
        }
      }
      __half2 h[4];
#pragma unroll
      for (int z = 0; z < 4; ++z) {
        const float a = t.w1 * v[0][2 * z] + t.w2 * v[1][2 * z] + t.w3 * v[2][2 * z] + t.w4 * v[3][2 * z];
        const float b = t.w1 * v[0][2 * z + 1] + t.w2 * v[1][2 * z + 1] + t.w3 * v[2][2 * z + 1] + t.w4 * v[3][2 * z + 1];
        h[z] = __floats2half2_rn(a, b);
      }
      *reinterpret_cast<uint4*>(colT + (size_t)(p0 + px) * ldk + (size_t)tap * C + c0) = *reinterpret_cast<const uint4*>(h);
    }
  }
}

// [C, HW] fp32 -> [HW, C] fp32 (32 x 32 tiles through shared memory): lets the NCHW entry point use the channels-last sampler
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, int C, int HW, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (c0 + r < C && p0 + tx < HW) tile[r][tx] = in[(size_t)(c0 + r) * HW + p0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (p0 + r < HW && c0 + tx < C) out[(size_t)(p0 + r) * C + c0 + tx] = tile[tx][r];
}

// weight [Co, C, kh*kw] fp32 -> fp16 [Co, ldk] with K index tap*C + c (the order deform_sample_kernel<1/2> writes colT in)
__global__ void __launch_bounds__(256) deform_pack_weight_kernel(const float* __restrict__ W, int Co, int C, int taps, int ldk,
                                                                 __half* __restrict__ out) {
  const size_t total = (size_t)Co * ldk;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % ldk, co = i / ldk;
    float v = 0.f;
    if (k < taps * C) { const int tap = k / C, c = k - tap * C; v = W[((size_t)co * C + c) * taps + tap]; }
    out[i] = __float2half_rn(v);
  }
}

__global__ void add_channel_bias_kernel(float* __restrict__ y, const float* __restrict__ bias, int Co, size_t spatial) {
  size_t total = (size_t)Co * spatial;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    y[i] += bias[i / spatial];
}

// ---- backward (training): operator_cxx/deformable_convolution-inl.h:145-233 ------------------------------------------
// Same decomposition as the forward: the geometry of a (deformable group, tap, position) is evaluated once per CTA into a
// shared-memory table and the group's channels stream through it.
//   data gradient  (deformable_col2im, nn/deformable_im2col.cuh:315-360 with get_gradient_weight :116-158): the table holds
//                  the <= 4 image cells that receive weight from the sample and their weights; a thread owns a
//                  (channel, position) and issues red.global.add per (tap, cell).  grad_im zero on entry.
//   offset gradient (deformable_col2im_coord, :407-458 with get_coordinate_weight :161-207): the table holds the four tap
//                  offsets and the four signed interpolation coefficients of each direction; a thread owns a
//                  (tap, direction, position) and sums over the group's channels in channel order (bit-comparable).
struct DcScatter { int n; int off[4]; float w[4]; };
struct DcCoord { int valid; int o00, o01, o10, o11; float kh_[4]; float kw_[4]; };

__device__ __forceinline__ float dc_gradient_weight(float ah, float aw, int h, int w, int height, int width) {   // :116-158
  if (ah < 0 || ah > height || aw < 0 || aw > width) return 0.f;
  ah = fmaxf(ah, 0.f); aw = fmaxf(aw, 0.f);
  int hl = (int)ah, wl = (int)aw, hh, wh;
  if (hl >= height - 1) { hh = hl = height - 1; ah = (float)hl; } else hh = hl + 1;
  if (wl >= width - 1) { wh = wl = width - 1; aw = (float)wl; } else wh = wl + 1;
  float weight = 0.f;
  if (h == hl) {
    if (w == wl) weight = (h + 1 - ah) * (w + 1 - aw);
    else if (w == wh) weight = (h + 1 - ah) * (aw + 1 - w);
  } else if (h == hh) {
    if (w == wl) weight = (ah + 1 - h) * (w + 1 - aw);
    else if (w == wh) weight = (ah + 1 - h) * (aw + 1 - w);
  }
  return weight;
}

__global__ void __launch_bounds__(256) deform_col2im_kernel(DcGeom q, int cpg, const float* __restrict__ col,
                                                            const float* __restrict__ off, float* __restrict__ grad_im) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  DcScatter* tab = reinterpret_cast<DcScatter*>(dc_smem);
  const int taps = q.kh * q.kw, Nsp = q.Ho * q.Wo;
  const int p0 = blockIdx.x * kDcTile, g = blockIdx.y;
  const int npos = min(kDcTile, Nsp - p0);
  const float* off_g = off + (size_t)g * 2 * taps * Nsp;
  for (int e = threadIdx.x; e < taps * kDcTile; e += blockDim.x) {
    const int tap = e / kDcTile, px = e - tap * kDcTile;
    if (px >= npos) continue;
    const int pos = p0 + px, w_out = pos % q.Wo, h_out = pos / q.Wo, i = tap / q.kw, j = tap - i * q.kw;
    const int w_in = w_out * q.sw - q.pad_w, h_in = h_out * q.sh - q.pad_h;
    const float oh = __ldg(off_g + ((size_t)(2 * tap) * q.Ho + h_out) * q.Wo + w_out);
    const float ow = __ldg(off_g + ((size_t)(2 * tap + 1) * q.Ho + h_out) * q.Wo + w_out);
    const float ih = h_in + i * q.dil_h + oh, iw = w_in + j * q.dil_w + ow;
    DcScatter t; t.n = 0;
    const int ch = (int)ih, cw = (int)iw;
    for (int dy = -1; dy <= 1; ++dy)                       // only the cells within one pixel of the sample can carry weight
      for (int dx = -1; dx <= 1; ++dx) {
        const int y = ch + dy, x = cw + dx;
        if (y >= 0 && y < q.H && x >= 0 && x < q.W && fabsf(ih - y) < 1 && fabsf(iw - x) < 1) {
          const float wgt = dc_gradient_weight(ih, iw, y, x, q.H, q.W);
          if (wgt != 0.f && t.n < 4) { t.off[t.n] = y * q.W + x; t.w[t.n] = wgt; ++t.n; }
        }
      }
    tab[e] = t;
  }
  __syncthreads();
  const int px = threadIdx.x % kDcTile;
  if (px >= npos) return;
  for (int cl = threadIdx.x / kDcTile; cl < cpg; cl += blockDim.x / kDcTile) {
    const int c = g * cpg + cl;
    float* gi = grad_im + (size_t)c * q.H * q.W;
    const float* src = col + ((size_t)c * taps) * Nsp + p0 + px;
    for (int tap = 0; tap < taps; ++tap) {
      const DcScatter& t = tab[tap * kDcTile + px];
      if (t.n == 0) continue;
      const float top = __ldg(src + (size_t)tap * Nsp);
      for (int k = 0; k < t.n; ++k) atomicAdd(gi + t.off[k], t.w[k] * top);
    }
  }
}

__global__ void __launch_bounds__(256) deform_col2im_coord_kernel(DcGeom q, int cpg, const float* __restrict__ col,
                                                                  const float* __restrict__ im, const float* __restrict__ off,
                                                                  float* __restrict__ grad_off) {
  extern __shared__ __align__(16) unsigned char dc_smem[];
  DcCoord* tab = reinterpret_cast<DcCoord*>(dc_smem);
  const int taps = q.kh * q.kw, Nsp = q.Ho * q.Wo;
  const int p0 = blockIdx.x * kDcTile, g = blockIdx.y;
  const int npos = min(kDcTile, Nsp - p0);
  const float* off_g = off + (size_t)g * 2 * taps * Nsp;
  for (int e = threadIdx.x; e < taps * kDcTile; e += blockDim.x) {
    const int tap = e / kDcTile, px = e - tap * kDcTile;
    if (px >= npos) continue;
    const int pos = p0 + px, w = pos % q.Wo, h = pos / q.Wo, i = tap / q.kw, j = tap - i * q.kw;
    const int w_in = w * q.sw - q.pad_w, h_in = h * q.sh - q.pad_h;
    const float oh = __ldg(off_g + ((size_t)(2 * tap) * q.Ho + h) * q.Wo + w);
    const float ow = __ldg(off_g + ((size_t)(2 * tap + 1) * q.Ho + h) * q.Wo + w);
    float ah = h_in + i * q.dil_h + oh, aw = w_in + j * q.dil_w + ow;
    if (ah < 0 || aw < 0 || ah >= q.H || aw >= q.W) ah = aw = -1;            // :440-442
    DcCoord t; t.valid = !(ah < 0 || ah > q.H || aw < 0 || aw > q.W);        // get_coordinate_weight :165-168
    t.o00 = t.o01 = t.o10 = t.o11 = 0;
    for (int k = 0; k < 4; ++k) { t.kh_[k] = 0.f; t.kw_[k] = 0.f; }
    if (t.valid) {
      int hl = (int)ah, wl = (int)aw, hh, wh;
      if (hl >= q.H - 1) { hh = hl = q.H - 1; ah = (float)hl; } else hh = hl + 1;
      if (wl >= q.W - 1) { wh = wl = q.W - 1; aw = (float)wl; } else wh = wl + 1;
      t.o00 = hl * q.W + wl; t.o01 = hl * q.W + wh; t.o10 = hh * q.W + wl; t.o11 = hh * q.W + wh;
      t.kh_[0] = -1 * (wl + 1 - aw); t.kh_[1] = -1 * (aw - wl); t.kh_[2] = (wl + 1 - aw); t.kh_[3] = (aw - wl);    // bp_dir 0
      t.kw_[0] = -1 * (hl + 1 - ah); t.kw_[1] = (hl + 1 - ah); t.kw_[2] = -1 * (ah - hl); t.kw_[3] = (ah - hl);    // bp_dir 1
    }
    tab[e] = t;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < npos * taps * 2; item += blockDim.x) {
    const int px = item % npos, td = item / npos, dir = td & 1, tap = td >> 1;
    const DcCoord& t = tab[tap * kDcTile + px];
    float val = 0.f;
    if (t.valid) {
      const float* k4 = dir == 0 ? t.kh_ : t.kw_;
      for (int cl = 0; cl < cpg; ++cl) {                   // channel order of the reference's loop (:447-455)
        const int c = g * cpg + cl;
        const float* d = im + (size_t)c * q.H * q.W;
        float weight = 0.f;
        weight += k4[0] * __ldg(d + t.o00);
        weight += k4[1] * __ldg(d + t.o01);
        weight += k4[2] * __ldg(d + t.o10);
        weight += k4[3] * __ldg(d + t.o11);
        val += weight * __ldg(col + ((size_t)c * taps + tap) * Nsp + p0 + px);
      }
    }
    grad_off[((size_t)g * 2 * taps + 2 * tap + dir) * Nsp + p0 + px] = val;
  }
}

// plain im2col (MXNet nn/im2col.h, not in tree) -- the reference's dWeight uses it (deformable_convolution-inl.h:215)
__global__ void __launch_bounds__(256) im2col_kernel(size_t n, const float* __restrict__ im, int H, int W, int kh, int kw,
                                                     int pad_h, int pad_w, int sh, int sw, int dil_h, int dil_w, int Ho,
                                                     int Wo, float* __restrict__ col) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < n; index += (size_t)gridDim.x * blockDim.x) {
    const int wo = index % Wo, ho = (index / Wo) % Ho;
    const int j = (index / Wo / Ho) % kw, i = (index / Wo / Ho / kw) % kh;
    const int c = index / Wo / Ho / kw / kh;
    const int h = ho * sh - pad_h + i * dil_h, w = wo * sw - pad_w + j * dil_w;
    col[index] = (h >= 0 && h < H && w >= 0 && w < W) ? __ldg(im + ((size_t)c * H + h) * W + w) : 0.f;
  }
}

// dbias[co] = sum over images and positions of dout
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ dout, int B, int Co, int Nsp,
                                                          float* __restrict__ dbias) {
  __shared__ float red[256];
  const int co = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b)
    for (int i = threadIdx.x; i < Nsp; i += 256) s += dout[((size_t)b * Co + co) * Nsp + i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dbias[co] = red[0];
}

static void out_hw(const rn_deform_conv_desc* d, int* Ho, int* Wo) {
  *Ho = (d->H + 2 * d->pad_h - (d->dil_h * (d->kh - 1) + 1)) / d->stride_h + 1;
  *Wo = (d->W + 2 * d->pad_w - (d->dil_w * (d->kw - 1) + 1)) / d->stride_w + 1;
}

static int check(const rn_deform_conv_desc* d) {
  RN_CHECK_ARG(d, "rn_deform_conv: null descriptor");
  RN_CHECK_ARG(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0 && d->Co > 0 && d->kh > 0 && d->kw > 0, "rn_deform_conv: bad sizes");
  RN_CHECK_ARG(d->num_group > 0 && d->C % d->num_group == 0 && d->Co % d->num_group == 0, "rn_deform_conv: bad num_group");
  RN_CHECK_ARG(d->num_deformable_group > 0 && d->C % d->num_deformable_group == 0, "rn_deform_conv: bad num_deformable_group");
  RN_CHECK_ARG(d->stride_h > 0 && d->stride_w > 0 && d->dil_h > 0 && d->dil_w > 0, "rn_deform_conv: bad stride/dilate");
  return RN_OK;
}

static DcGeom dc_geom(const rn_deform_conv_desc* d) {
  DcGeom q;
  q.H = d->H; q.W = d->W; q.kh = d->kh; q.kw = d->kw; q.pad_h = d->pad_h; q.pad_w = d->pad_w; q.sh = d->stride_h;
  q.sw = d->stride_w; q.dil_h = d->dil_h; q.dil_w = d->dil_w;
  out_hw(d, &q.Ho, &q.Wo);
  return q;
}

// mode 0: im NCHW fp32 -> col fp32; mode 1 / 2: im NHWC fp32 / bf16 -> colT fp16 (tap-major K, pitch ldk)
static int launch_sample(const rn_deform_conv_desc* d, int mode, const void* im, const float* off, float* col, __half* colT,
                         int ldk, cudaStream_t st) {
  const DcGeom q = dc_geom(d);
  const int taps = d->kh * d->kw, cpg = d->C / d->num_deformable_group;
  const int tile = mode == 0 ? kDcTile : kDcTileCl;
  const size_t smem = sizeof(DcSample) * taps * tile;
  RN_CHECK_ARG(smem <= 96 * 1024, "deformable conv: kernel %dx%d too large for the sample table", d->kh, d->kw);
  RN_CHECK_ARG(mode == 0 || (cpg % 8 == 0 && d->C % 8 == 0), "deformable conv (channels-last): C / num_deformable_group must be a multiple of 8");
  const dim3 grid(cdiv(q.Ho * q.Wo, tile), d->num_deformable_group);
  if (smem > 48 * 1024) {
    RN_CUDA(cudaFuncSetAttribute(deform_sample_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RN_CUDA(cudaFuncSetAttribute(deform_sample_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RN_CUDA(cudaFuncSetAttribute(deform_sample_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  if (mode == 0) deform_sample_kernel<0><<<grid, 256, smem, st>>>(q, d->C, cpg, im, off, col, nullptr, 0);
  else if (mode == 1) deform_sample_kernel<1><<<grid, 256, smem, st>>>(q, d->C, cpg, im, off, nullptr, colT, ldk);
  else deform_sample_kernel<2><<<grid, 256, smem, st>>>(q, d->C, cpg, im, off, nullptr, colT, ldk);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

static int launch_im2col(const rn_deform_conv_desc* d, const float* im, const float* off, float* col, cudaStream_t st) {
  return launch_sample(d, 0, im, off, col, nullptr, 0, st);
}

}  // namespace rn

extern "C" size_t rn_deform_conv_workspace_bytes(const rn_deform_conv_desc* d) {
  if (!d) return 0;
  int Ho, Wo; rn::out_hw(d, &Ho, &Wo);
  const size_t K = (size_t)d->C * d->kh * d->kw, K8 = rn::align_up(K / d->num_group, 8);
  const size_t f32 = rn::ws_slice(K * Ho * Wo, 4);
  const size_t f16 = rn::ws_slice((size_t)Ho * Wo * K8, 2) + rn::ws_slice((size_t)d->Co * K8, 2) +
                     rn::ws_slice((size_t)d->C * d->H * d->W, 4) +
                     std::max(rn::gemm_tc_workspace_bytes(d->Co / d->num_group, Ho * Wo, (int)K8),
                              rn::gemm_tc_workspace_bytes(Ho * Wo, d->Co / d->num_group, (int)K8)) + 512;
  return (f32 > f16 ? f32 : f16) + 256;
}

extern "C" int rn_deform_im2col(const rn_deform_conv_desc* d, const float* data_b, const float* offset_b, float* col,
                                rn_stream_t stream) {
  int r = rn::check(d);
  if (r) return r;
  RN_CHECK_ARG(data_b && offset_b && col, "rn_deform_im2col: null pointer");
  return rn::launch_im2col(d, data_b, offset_b, col, (cudaStream_t)stream);
}

extern "C" int rn_deform_conv_fwd(const rn_deform_conv_desc* d, const float* data, const float* offset,
                                  const float* weight, const float* bias, float* out, void* wsp, size_t ws_bytes,
                                  rn_stream_t stream) {
  int r = rn::check(d);
  if (r) return r;
  RN_CHECK_ARG(data && offset && weight && out && wsp, "rn_deform_conv_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int Ho, Wo; rn::out_hw(d, &Ho, &Wo);
  const int K = d->C * d->kh * d->kw, Nsp = Ho * Wo, G = d->num_group;
  if (d->precision == RN_PREC_F16 && G == 1 && rn::is_sm100() && d->C % 8 == 0 && (d->C / d->num_deformable_group) % 8 == 0) {
    // tensor-core path: channels-last copy of the image -> fp16 K-major column buffer (tap-major K) -> tcgen05 GEMM
    const int K8 = (int)rn::align_up(K, 8);
    rn::Workspace w2(wsp, ws_bytes);
    __half* colT = w2.take<__half>((size_t)Nsp * K8);
    __half* w16 = w2.take<__half>((size_t)d->Co * K8);
    float* nhwc = w2.take<float>((size_t)d->C * d->H * d->W);
    if (!nhwc) { rn::set_error("rn_deform_conv_fwd(F16): workspace too small"); return RN_ERR_WORKSPACE; }
    rn::deform_pack_weight_kernel<<<rn::sm_count() > 0 ? rn::sm_count() * 4 : 592, 256, 0, st>>>(weight, d->Co, d->C, d->kh * d->kw, K8, w16);
    RN_LAUNCH_CHECK();
    if (K8 != K) RN_CUDA(cudaMemsetAsync(colT, 0, (size_t)Nsp * K8 * 2, st));
    const size_t off_per16 = (size_t)d->num_deformable_group * 2 * d->kh * d->kw * Nsp;
    for (int b = 0; b < d->B; ++b) {
      rn::nchw_to_nhwc_kernel<<<dim3(rn::cdiv(d->H * d->W, 32), rn::cdiv(d->C, 32)), 256, 0, st>>>(
          data + (size_t)b * d->C * d->H * d->W, d->C, d->H * d->W, nhwc);
      RN_LAUNCH_CHECK();
      if ((r = rn::launch_sample(d, 1, nhwc, offset + b * off_per16, nullptr, colT, K8, st))) return r;
      float* ob = out + (size_t)b * d->Co * Nsp;
      if ((r = rn::gemm_tc(st, w16, K8, colT, K8, d->Co, Nsp, K8, bias, 1, 0, ob, Nsp, nullptr, 0, w2.base + w2.off,
                           w2.size - w2.off))) return r;
    }
    return RN_OK;
  }
  rn::Workspace ws(wsp, ws_bytes);
  float* col = ws.take<float>((size_t)K * Nsp);
  if (!col) { rn::set_error("rn_deform_conv_fwd: workspace too small"); return RN_ERR_WORKSPACE; }
  const size_t off_per = (size_t)d->num_deformable_group * 2 * d->kh * d->kw * Nsp;
  for (int b = 0; b < d->B; ++b) {
    if ((r = rn::launch_im2col(d, data + (size_t)b * d->C * d->H * d->W, offset + b * off_per, col, st))) return r;
    float* ob = out + (size_t)b * d->Co * Nsp;
    // out[g] (Co/G x Nsp) = W[g] (Co/G x K/G) . col[g] (K/G x Nsp)
    if ((r = rn::sgemm_nn(st, d->Co / G, Nsp, K / G, weight, K / G, col, Nsp, ob, Nsp, G,
                          (long long)(d->Co / G) * (K / G), (long long)(K / G) * Nsp, (long long)(d->Co / G) * Nsp)))
      return r;
    if (bias) {
      rn::add_channel_bias_kernel<<<rn::cdiv(d->Co * Nsp, 256), 256, 0, st>>>(ob, bias, d->Co, (size_t)Nsp);
      RN_LAUNCH_CHECK();
    }
  }
  return RN_OK;
}

// ---- channels-last fast path (the Deformable Faster-RCNN res5 layers inside the bf16 channels_last trunk) ---------------
extern "C" size_t rn_deform_conv_packed_bytes(const rn_deform_conv_desc* d) {
  if (!d) return 0;
  return rn::ws_slice((size_t)d->Co * rn::align_up((size_t)d->C * d->kh * d->kw, 8), 2);
}

extern "C" int rn_deform_conv_pack(const rn_deform_conv_desc* d, const float* weight, void* packed, rn_stream_t stream) {
  int r = rn::check(d);
  if (r) return r;
  RN_CHECK_ARG(weight && packed && d->num_group == 1, "rn_deform_conv_pack: null pointer / num_group != 1");
  const int K8 = (int)rn::align_up((size_t)d->C * d->kh * d->kw, 8);
  rn::deform_pack_weight_kernel<<<rn::sm_count() > 0 ? rn::sm_count() * 4 : 592, 256, 0, (cudaStream_t)stream>>>(
      weight, d->Co, d->C, d->kh * d->kw, K8, (__half*)packed);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

// data NHWC (fp32 / bf16), offset [dg*2*kh*kw, Ho, Wo] fp32, packed weight -> out NHWC: out32 [Ho*Wo, Co] fp32 and / or
// out16 fp16 (either may be NULL), bias per output channel, optional relu in the GEMM epilogue.  One image (B == 1).
extern "C" int rn_deform_conv_nhwc_fwd(const rn_deform_conv_desc* d, const void* data_nhwc, int32_t data_is_bf16,
                                       const float* offset, const void* packed_weight, const float* bias, int32_t relu,
                                       float* out32, void* out16, void* wsp, size_t ws_bytes, rn_stream_t stream) {
  int r = rn::check(d);
  if (r) return r;
  RN_CHECK_ARG(data_nhwc && offset && packed_weight && (out32 || out16) && wsp, "rn_deform_conv_nhwc_fwd: null pointer");
  RN_CHECK_ARG(d->B == 1 && d->num_group == 1, "rn_deform_conv_nhwc_fwd: one image, num_group == 1");
  RN_CHECK_ARG(rn::is_sm100(), "rn_deform_conv_nhwc_fwd: needs an sm_100 device (tcgen05 GEMM)");
  cudaStream_t st = (cudaStream_t)stream;
  int Ho, Wo; rn::out_hw(d, &Ho, &Wo);
  const int K = d->C * d->kh * d->kw, K8 = (int)rn::align_up(K, 8), Nsp = Ho * Wo;
  rn::Workspace ws(wsp, ws_bytes);
  __half* colT = ws.take<__half>((size_t)Nsp * K8);
  if (!colT) { rn::set_error("rn_deform_conv_nhwc_fwd: workspace too small"); return RN_ERR_WORKSPACE; }
  if (K8 != K) RN_CUDA(cudaMemsetAsync(colT, 0, (size_t)Nsp * K8 * 2, st));
  if ((r = rn::launch_sample(d, data_is_bf16 ? 2 : 1, data_nhwc, offset, nullptr, colT, K8, st))) return r;
  // out[pos, co] = colT[pos, :] . W16[co, :]^T  (+ bias[co], relu): rows = positions -> channels-last output
  return rn::gemm_tc(st, colT, K8, (const __half*)packed_weight, K8, Nsp, d->Co, K8, bias, 0, relu, out32, d->Co,
                     (__half*)out16, d->Co, ws.base + ws.off, ws.size - ws.off);
}

// Backward.  ddata [B,C,H,W], doffset (shape of offset), dweight (shape of weight), dbias [Co] or NULL -- all OVERWRITTEN.
// weight_grad_deformed = 0 reproduces the reference (dWeight from the plain im2col of data, deformable_convolution-inl.h
// :215); 1 uses the deformed sampling (the mathematically consistent gradient, as later MXNet releases do).
extern "C" int rn_deform_conv_bwd(const rn_deform_conv_desc* d, const float* dout, const float* data, const float* offset,
                                  const float* weight, int32_t weight_grad_deformed, float* ddata, float* doffset,
                                  float* dweight, float* dbias, void* wsp, size_t ws_bytes, rn_stream_t stream) {
  int r = rn::check(d);
  if (r) return r;
  RN_CHECK_ARG(dout && data && offset && weight && ddata && doffset && dweight && wsp, "rn_deform_conv_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int Ho, Wo; rn::out_hw(d, &Ho, &Wo);
  const int K = d->C * d->kh * d->kw, Nsp = Ho * Wo, G = d->num_group, Cog = d->Co / G, Kg = K / G;
  rn::Workspace ws(wsp, ws_bytes);
  float* col = ws.take<float>((size_t)K * Nsp);
  if (!col) { rn::set_error("rn_deform_conv_bwd: workspace too small"); return RN_ERR_WORKSPACE; }
  const size_t off_per = (size_t)d->num_deformable_group * 2 * d->kh * d->kw * Nsp;
  const size_t im_per = (size_t)d->C * d->H * d->W;
  const size_t cap = (size_t)(rn::sm_count() > 0 ? rn::sm_count() : 148) * 16;
  auto grid = [&](size_t n) { size_t b = (n + 255) / 256; return (int)(b < cap ? b : cap); };
  RN_CUDA(cudaMemsetAsync(ddata, 0, sizeof(float) * d->B * im_per, st));
  for (int b = 0; b < d->B; ++b) {
    const float* dob = dout + (size_t)b * d->Co * Nsp;
    // col[g] (Kg x Nsp) = W[g]^T (Kg x Cog) . dout[b][g] (Cog x Nsp)
    if ((r = rn::sgemm_rm(st, true, false, Kg, Nsp, Cog, 1.f, weight, Kg, dob, Nsp, 0.f, col, Nsp, G, (long long)Cog * Kg,
                          (long long)Cog * Nsp, (long long)Kg * Nsp))) return r;
    {
      const rn::DcGeom q = rn::dc_geom(d);
      const int taps = d->kh * d->kw, cpg = d->C / d->num_deformable_group;
      const dim3 g2(rn::cdiv(Nsp, rn::kDcTile), d->num_deformable_group);
      const size_t sm_c = sizeof(rn::DcCoord) * taps * rn::kDcTile, sm_s = sizeof(rn::DcScatter) * taps * rn::kDcTile;
      RN_CHECK_ARG(sm_c <= 48 * 1024 && sm_s <= 48 * 1024, "rn_deform_conv_bwd: kernel %dx%d too large for the sample table", d->kh, d->kw);
      rn::deform_col2im_coord_kernel<<<g2, 256, sm_c, st>>>(q, cpg, col, data + b * im_per, offset + b * off_per,
                                                           doffset + b * off_per);
      RN_LAUNCH_CHECK();
      rn::deform_col2im_kernel<<<g2, 256, sm_s, st>>>(q, cpg, col, offset + b * off_per, ddata + b * im_per);
      RN_LAUNCH_CHECK();
    }
    const size_t n_col = (size_t)K * Nsp;
    if (weight_grad_deformed) {
      if ((r = rn::launch_im2col(d, data + b * im_per, offset + b * off_per, col, st))) return r;
    } else {
      rn::im2col_kernel<<<grid(n_col), 256, 0, st>>>(n_col, data + b * im_per, d->H, d->W, d->kh, d->kw, d->pad_h, d->pad_w,
                                                    d->stride_h, d->stride_w, d->dil_h, d->dil_w, Ho, Wo, col);
      RN_LAUNCH_CHECK();
    }
    // dW[g] (Cog x Kg) (+)= dout[b][g] (Cog x Nsp) . col[g]^T (Nsp x Kg)
    if ((r = rn::sgemm_rm(st, false, true, Cog, Kg, Nsp, 1.f, dob, Nsp, col, Nsp, b ? 1.f : 0.f, dweight, Kg, G,
                          (long long)Cog * Nsp, (long long)Kg * Nsp, (long long)Cog * Kg))) return r;
  }
  if (dbias) {
    rn::channel_sum_kernel<<<d->Co, 256, 0, st>>>(dout, d->B, d->Co, Nsp, dbias);
    RN_LAUNCH_CHECK();
  }
  return RN_OK;
}
