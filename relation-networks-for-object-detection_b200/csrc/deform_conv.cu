// deform_conv.cu -- DeformableConvolution forward (operator_cxx/deformable_convolution-inl.h:91-144):
// deformable im2col (nn/deformable_im2col.cuh:216-262, bilinear :77-113) followed by a per-group GEMM W[g].col[g].
// Compiled with -fmad=false so the bilinear arithmetic is bit-comparable with oracle/oracle_c.c.
//
// HBM layout: data [B,C,H,W], offset [B, dg*2*kh*kw, Ho, Wo], weight [Co, C/g*kh*kw], col workspace [C*kh*kw, Ho*Wo]
// (44 MB at 512ch/38x63, written once and re-read by the GEMM -- round-1 shape of the op; the implicit-GEMM form that
// never materialises col is listed under "next" in DESIGN.md).
// Roofline: GEMM 2*Co*C*kh*kw*Ho*Wo FLOP (11.3 GFLOP/layer) on the tensor pipe; im2col is HBM-write bound (4*C*kh*kw*Ho*Wo B).
#include "common.cuh"
#include "gemm_tc.cuh"

namespace rn {

__device__ __forceinline__ float dim2col_bilinear(const float* __restrict__ d, int data_width, int height, int width,
                                                  float h, float w) {
  int h_low = (int)floorf(h), w_low = (int)floorf(w), h_high, w_high;
  if (h_low >= height - 1) { h_high = h_low = height - 1; h = (float)h_low; } else h_high = h_low + 1;
  if (w_low >= width - 1) { w_high = w_low = width - 1; w = (float)w_low; } else w_high = w_low + 1;
  const float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
  const float v1 = __ldg(d + h_low * data_width + w_low), v2 = __ldg(d + h_low * data_width + w_high);
  const float v3 = __ldg(d + h_high * data_width + w_low), v4 = __ldg(d + h_high * data_width + w_high);
  const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

// one thread = one (c_im, h_col, w_col); writes kh*kw column entries (coalesced along w_col)
__global__ void __launch_bounds__(256) deform_im2col_kernel(size_t n, const float* __restrict__ im,
                                                            const float* __restrict__ off, int H, int W, int kh, int kw,
                                                            int pad_h, int pad_w, int sh, int sw, int dil_h, int dil_w,
                                                            int cpg, int Ho, int Wo, float* __restrict__ col) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < n; index += (size_t)gridDim.x * blockDim.x) {
    const int w_col = index % Wo, h_col = (index / Wo) % Ho, c_im = (index / Wo) / Ho;
    const int g = c_im / cpg;
    const int h_in = h_col * sh - pad_h, w_in = w_col * sw - pad_w;
    float* col_ptr = col + (((size_t)c_im * kh * kw) * Ho + h_col) * Wo + w_col;
    const float* im_ptr = im + ((ptrdiff_t)c_im * H + h_in) * W + w_in;
    const float* off_ptr = off + (size_t)g * 2 * kh * kw * Ho * Wo;
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j) {
        const float oh = __ldg(off_ptr + ((size_t)(2 * (i * kw + j)) * Ho + h_col) * Wo + w_col);
        const float ow = __ldg(off_ptr + ((size_t)(2 * (i * kw + j) + 1) * Ho + h_col) * Wo + w_col);
        float val = 0.f;
        const float h_im = h_in + i * dil_h + oh, w_im = w_in + j * dil_w + ow;
        if (h_im >= 0 && w_im >= 0 && h_im < H && w_im < W) {
          const float map_h = i * dil_h + oh, map_w = j * dil_w + ow;
          val = dim2col_bilinear(im_ptr, W, H - h_in, W - w_in, map_h, map_w);
        }
        *col_ptr = val;
        col_ptr += (size_t)Ho * Wo;
      }
  }
}

// RN_PREC_F16 form: the column buffer is written TRANSPOSED and in fp16, colT[pos][c*kh*kw + k] (K contiguous), which
// is exactly the K-major B operand of the tcgen05 GEMM  out[Co, pos] = W16[Co, K] . colT[pos, K]^T  -- half the bytes of
// the fp32 col buffer and no SGEMM.  Same bilinear arithmetic as above, rounded to fp16 at the store.
__global__ void __launch_bounds__(256) deform_im2col_t_f16_kernel(size_t n, const float* __restrict__ im,
                                                                  const float* __restrict__ off, int H, int W, int kh,
                                                                  int kw, int pad_h, int pad_w, int sh, int sw, int dil_h,
                                                                  int dil_w, int cpg, int Ho, int Wo, int C, int ldk,
                                                                  __half* __restrict__ colT) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < n; index += (size_t)gridDim.x * blockDim.x) {
    const int c_im = index % C;                         // channel fastest: consecutive threads write 2*kh*kw B apart
    const int pos = index / C, w_col = pos % Wo, h_col = pos / Wo;
    const int g = c_im / cpg;
    const int h_in = h_col * sh - pad_h, w_in = w_col * sw - pad_w;
    const float* im_ptr = im + ((ptrdiff_t)c_im * H + h_in) * W + w_in;
    const float* off_ptr = off + (size_t)g * 2 * kh * kw * Ho * Wo;
    __half* dst = colT + (size_t)pos * ldk + (size_t)c_im * kh * kw;
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j) {
        const float oh = __ldg(off_ptr + ((size_t)(2 * (i * kw + j)) * Ho + h_col) * Wo + w_col);
        const float ow = __ldg(off_ptr + ((size_t)(2 * (i * kw + j) + 1) * Ho + h_col) * Wo + w_col);
        float val = 0.f;
        const float h_im = h_in + i * dil_h + oh, w_im = w_in + j * dil_w + ow;
        if (h_im >= 0 && w_im >= 0 && h_im < H && w_im < W)
          val = dim2col_bilinear(im_ptr, W, H - h_in, W - w_in, i * dil_h + oh, j * dil_w + ow);
        dst[i * kw + j] = __float2half_rn(val);
      }
  }
}

__global__ void add_channel_bias_kernel(float* __restrict__ y, const float* __restrict__ bias, int Co, size_t spatial) {
  size_t total = (size_t)Co * spatial;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    y[i] += bias[i / spatial];
}

// ---- backward (training): operator_cxx/deformable_convolution-inl.h:145-233 ------------------------------------------
// get_gradient_weight, nn/deformable_im2col.cuh:116-158
__device__ __forceinline__ float dcol_gradient_weight(float ah, float aw, int h, int w, int height, int width) {
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

// get_coordinate_weight, nn/deformable_im2col.cuh:161-207
__device__ __forceinline__ float dcol_coordinate_weight(float ah, float aw, int height, int width,
                                                        const float* __restrict__ im, int data_width, int bp_dir) {
  if (ah < 0 || ah > height || aw < 0 || aw > width) return 0.f;
  int hl = (int)ah, wl = (int)aw, hh, wh;
  if (hl >= height - 1) { hh = hl = height - 1; ah = (float)hl; } else hh = hl + 1;
  if (wl >= width - 1) { wh = wl = width - 1; aw = (float)wl; } else wh = wl + 1;
  const float v00 = __ldg(im + hl * data_width + wl), v01 = __ldg(im + hl * data_width + wh);
  const float v10 = __ldg(im + hh * data_width + wl), v11 = __ldg(im + hh * data_width + wh);
  float weight = 0.f;
  if (bp_dir == 0) {
    weight += -1 * (wl + 1 - aw) * v00;
    weight += -1 * (aw - wl) * v01;
    weight += (wl + 1 - aw) * v10;
    weight += (aw - wl) * v11;
  } else {
    weight += -1 * (hl + 1 - ah) * v00;
    weight += (hl + 1 - ah) * v01;
    weight += -1 * (ah - hl) * v10;
    weight += (ah - hl) * v11;
  }
  return weight;
}

// deformable_col2im_gpu_kernel (nn/deformable_im2col.cuh:315-360): one thread per column element, bilinear scatter into
// grad_im [C,H,W] with red.global.add; grad_im zero on entry.  Only the 2x2 neighbourhood of the sample can carry weight,
// so the reference's 5x5 scan is reduced to the four candidate taps (same predicate, same weights).
__global__ void __launch_bounds__(256) deform_col2im_kernel(size_t n, const float* __restrict__ col,
                                                            const float* __restrict__ off, int H, int W, int kh, int kw,
                                                            int pad_h, int pad_w, int sh, int sw, int dil_h, int dil_w,
                                                            int cpg, int Ho, int Wo, float* __restrict__ grad_im) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < n; index += (size_t)gridDim.x * blockDim.x) {
    const int w_out = index % Wo, h_out = (index / Wo) % Ho;
    const int j = (index / Wo / Ho) % kw, i = (index / Wo / Ho / kw) % kh;
    const int c = index / Wo / Ho / kw / kh;
    const int g = c / cpg;
    const int w_in = w_out * sw - pad_w, h_in = h_out * sh - pad_h;
    const float* off_ptr = off + (size_t)g * 2 * kh * kw * Ho * Wo;
    const float oh = __ldg(off_ptr + ((size_t)(2 * (i * kw + j)) * Ho + h_out) * Wo + w_out);
    const float ow = __ldg(off_ptr + ((size_t)(2 * (i * kw + j) + 1) * Ho + h_out) * Wo + w_out);
    const float ih = h_in + i * dil_h + oh, iw = w_in + j * dil_w + ow;
    const float top = col[index];
    const int ch = (int)ih, cw = (int)iw;
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        const int y = ch + dy, x = cw + dx;
        if (y >= 0 && y < H && x >= 0 && x < W && fabsf(ih - y) < 1 && fabsf(iw - x) < 1) {
          const float wgt = dcol_gradient_weight(ih, iw, y, x, H, W);
          if (wgt != 0.f) atomicAdd(grad_im + ((size_t)c * H + y) * W + x, wgt * top);
        }
      }
  }
}

// deformable_col2im_coord_gpu_kernel (nn/deformable_im2col.cuh:407-458): one thread per offset element, sum over the
// channels of its deformable group.
__global__ void __launch_bounds__(256) deform_col2im_coord_kernel(size_t n, const float* __restrict__ col,
                                                                  const float* __restrict__ im,
                                                                  const float* __restrict__ off, int H, int W, int kh,
                                                                  int kw, int pad_h, int pad_w, int sh, int sw, int dil_h,
                                                                  int dil_w, int cpg_col, int Ho, int Wo,
                                                                  float* __restrict__ grad_off) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < n; index += (size_t)gridDim.x * blockDim.x) {
    const int w = index % Wo, h = (index / Wo) % Ho;
    const int c = index / Wo / Ho;
    const int g = c / (2 * kh * kw);
    const int offset_c = c - g * 2 * kh * kw;
    const int kpos = offset_c / 2, bp_dir = offset_c % 2;
    const int i = kpos / kw, j = kpos % kw;
    const float* col_ptr = col + (size_t)g * cpg_col * Wo * Ho;
    const float* im_ptr = im + (size_t)g * (cpg_col / kh / kw) * H * W;
    const float* off_ptr = off + (size_t)g * 2 * kh * kw * Ho * Wo;
    const int w_in = w * sw - pad_w, h_in = h * sh - pad_h;
    const float oh = __ldg(off_ptr + ((size_t)(2 * kpos) * Ho + h) * Wo + w);
    const float ow = __ldg(off_ptr + ((size_t)(2 * kpos + 1) * Ho + h) * Wo + w);
    float inv_h = h_in + i * dil_h + oh, inv_w = w_in + j * dil_w + ow;
    if (inv_h < 0 || inv_w < 0 || inv_h >= H || inv_w >= W) inv_h = inv_w = -1;
    float val = 0.f;
    int cnt = 0;
    for (int col_c = kpos; col_c < cpg_col; col_c += kh * kw, ++cnt) {
      const float wgt = dcol_coordinate_weight(inv_h, inv_w, H, W, im_ptr + (size_t)cnt * H * W, W, bp_dir);
      val += wgt * __ldg(col_ptr + ((size_t)col_c * Ho + h) * Wo + w);
    }
    grad_off[index] = val;
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

static int launch_im2col(const rn_deform_conv_desc* d, const float* im, const float* off, float* col, cudaStream_t st) {
  int Ho, Wo; out_hw(d, &Ho, &Wo);
  size_t n = (size_t)d->C * Ho * Wo;
  size_t blocks = (n + 255) / 256, cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 16;
  deform_im2col_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(
      n, im, off, d->H, d->W, d->kh, d->kw, d->pad_h, d->pad_w, d->stride_h, d->stride_w, d->dil_h, d->dil_w,
      d->C / d->num_deformable_group, Ho, Wo, col);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn

extern "C" size_t rn_deform_conv_workspace_bytes(const rn_deform_conv_desc* d) {
  if (!d) return 0;
  int Ho, Wo; rn::out_hw(d, &Ho, &Wo);
  const size_t K = (size_t)d->C * d->kh * d->kw, K8 = rn::align_up(K / d->num_group, 8);
  const size_t f32 = rn::ws_slice(K * Ho * Wo, 4);
  const size_t f16 = rn::ws_slice((size_t)Ho * Wo * K8, 2) + rn::ws_slice((size_t)d->Co * K8, 2) +
                     rn::gemm_tc_workspace_bytes(d->Co / d->num_group, Ho * Wo, (int)K8) + 512;
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
  if (d->precision == RN_PREC_F16 && G == 1 && rn::is_sm100()) {
    // tensor-core path: fp16 transposed column buffer + tcgen05 GEMM (fp32 accumulate, fp32 output)
    const int K8 = (int)rn::align_up(K, 8);
    rn::Workspace w2(wsp, ws_bytes);
    __half* colT = w2.take<__half>((size_t)Nsp * K8);
    __half* w16 = w2.take<__half>((size_t)d->Co * K8);
    if (!w16) { rn::set_error("rn_deform_conv_fwd(F16): workspace too small"); return RN_ERR_WORKSPACE; }
    if ((r = rn::cast_rows_f16(st, weight, w16, d->Co, K, K8))) return r;
    if (K8 != K) RN_CUDA(cudaMemsetAsync(colT, 0, (size_t)Nsp * K8 * 2, st));
    const size_t off_per16 = (size_t)d->num_deformable_group * 2 * d->kh * d->kw * Nsp;
    for (int b = 0; b < d->B; ++b) {
      const size_t n = (size_t)d->C * Nsp;
      size_t blocks = (n + 255) / 256, cap = (size_t)(rn::sm_count() > 0 ? rn::sm_count() : 148) * 16;
      rn::deform_im2col_t_f16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, st>>>(
          n, data + (size_t)b * d->C * d->H * d->W, offset + b * off_per16, d->H, d->W, d->kh, d->kw, d->pad_h, d->pad_w,
          d->stride_h, d->stride_w, d->dil_h, d->dil_w, d->C / d->num_deformable_group, Ho, Wo, d->C, K8, colT);
      RN_LAUNCH_CHECK();
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
    const size_t n_off = off_per;
    rn::deform_col2im_coord_kernel<<<grid(n_off), 256, 0, st>>>(
        n_off, col, data + b * im_per, offset + b * off_per, d->H, d->W, d->kh, d->kw, d->pad_h, d->pad_w, d->stride_h,
        d->stride_w, d->dil_h, d->dil_w, K / d->num_deformable_group, Ho, Wo, doffset + b * off_per);
    RN_LAUNCH_CHECK();
    const size_t n_col = (size_t)K * Nsp;
    rn::deform_col2im_kernel<<<grid(n_col), 256, 0, st>>>(n_col, col, offset + b * off_per, d->H, d->W, d->kh, d->kw,
                                                         d->pad_h, d->pad_w, d->stride_h, d->stride_w, d->dil_h, d->dil_w,
                                                         d->C / d->num_deformable_group, Ho, Wo, ddata + b * im_per);
    RN_LAUNCH_CHECK();
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
