// rois.cu -- ROIPooling (max, MXNet semantics) and DeformablePSROIPooling forward (= average ROIAlign when no_trans).
// Compiled with -fmad=false: the float/double op order of the reference kernels is kept literally so results are
// bit-comparable with the plain-C oracle (oracle/oracle_c.c).
//
// HBM layout: data [B,C,H,W] fp32 (the 38x63x256 map is 2.4 MB -> L2 resident), out [R,C,P,P] fp32.
// Roofline: HBM-write bound -- algorithmic bytes = R*C*P*P*4 (out) + the feature map once; the gathers hit L2.
// Grid: one thread per output element, blocks of 256, grid-stride capped at 148*8 CTAs.
#include "common.cuh"
#include <cfloat>
#include <cuda_bf16.h>

namespace rn {

// Restates MXNet 1.1.0 src/operator/roi_pooling.cu ROIPoolForwardKernel (not in the reference tree; call site SYM_REL:252)
__global__ void __launch_bounds__(256) roi_pool_fwd_kernel(const float* __restrict__ data, const float* __restrict__ rois,
                                                           size_t count, int C, int H, int W, int PH, int PW,
                                                           float spatial_scale, float* __restrict__ out,
                                                           int* __restrict__ argmax) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < count;
       index += (size_t)gridDim.x * blockDim.x) {
    const int pw = index % PW, ph = (index / PW) % PH;
    const int c = (index / PW / PH) % C;
    const int n = index / PW / PH / C;
    const float* roi = rois + 5 * n;
    const int b = (int)roi[0];
    const int rsw = (int)roundf(roi[1] * spatial_scale), rsh = (int)roundf(roi[2] * spatial_scale);
    const int rew = (int)roundf(roi[3] * spatial_scale), reh = (int)roundf(roi[4] * spatial_scale);
    const int rh = max(reh - rsh + 1, 1), rw = max(rew - rsw + 1, 1);
    const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
    int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
    hs = min(max(hs + rsh, 0), H); he = min(max(he + rsh, 0), H);
    ws = min(max(ws + rsw, 0), W); we = min(max(we + rsw, 0), W);
    const bool empty = (he <= hs) || (we <= ws);
    float m = empty ? 0.f : -FLT_MAX;
    int mi = -1;
    const float* d = data + ((size_t)b * C + c) * H * W;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const float v = __ldg(d + h * W + w);
        if (v > m) { m = v; mi = h * W + w; }
      }
    out[index] = m;
    if (argmax) argmax[index] = mi;
  }
}

// Same op for the fused head: channels-last input (what the cuDNN trunk produces) and fp16 output in [R, PH*PW, C] order
// (the K order of the permuted fc_new_1 weight, see rn_linear_pack_chw_to_hwc) -- one CTA per (roi, bin), threads along
// channels: every load and store is coalesced, the bin window is shared by the CTA, and the 15 MB fp32 pooled tensor
// plus its fp16 cast kernel disappear.  Same bin arithmetic as above -> same maxima (then rounded to fp16).
__global__ void __launch_bounds__(128) roi_pool_nhwc_f16_kernel(const float* __restrict__ data, const float* __restrict__ rois,
                                                                int C, int H, int W, int PH, int PW, float spatial_scale,
                                                                __half* __restrict__ out) {
  const int bin = blockIdx.x, n = blockIdx.y;
  const int ph = bin / PW, pw = bin % PW;
  const float* roi = rois + 5 * n;
  const int b = (int)roi[0];
  const int rsw = (int)roundf(roi[1] * spatial_scale), rsh = (int)roundf(roi[2] * spatial_scale);
  const int rew = (int)roundf(roi[3] * spatial_scale), reh = (int)roundf(roi[4] * spatial_scale);
  const int rh = max(reh - rsh + 1, 1), rw = max(rew - rsw + 1, 1);
  const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
  int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
  int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
  hs = min(max(hs + rsh, 0), H); he = min(max(he + rsh, 0), H);
  ws = min(max(ws + rsw, 0), W); we = min(max(we + rsw, 0), W);
  const bool empty = (he <= hs) || (we <= ws);
  const float* d = data + (size_t)b * H * W * C;
  for (int c = threadIdx.x * 2; c < C; c += blockDim.x * 2) {
    float m0 = empty ? 0.f : -FLT_MAX, m1 = m0;
    for (int h = hs; h < he; ++h)
      for (int w = ws; w < we; ++w) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(d + ((size_t)h * W + w) * C + c));
        m0 = fmaxf(m0, v.x); m1 = fmaxf(m1, v.y);
      }
    *reinterpret_cast<__half2*>(out + ((size_t)n * PH * PW + bin) * C + c) = __floats2half2_rn(m0, m1);
  }
}

// bf16 channels-last input (the cuDNN trunk's native output: no fp32 conversion pass, half the L2 reads).  CTA = one roi
// x one row of bins; warp w takes bins pw = w, w + 4; lane l owns channels 8l..8l+7 of a 256-channel slab (16-byte loads,
// 512 B per cell per warp, fully coalesced), max in bf16x2 (exact), result converted bf16 -> fp16 (exact in fp16's range:
// a bf16 value has 8 significant bits), i.e. bitwise the value the fp32 kernel above produces.
__global__ void __launch_bounds__(128) roi_pool_nhwc_bf16in_f16_kernel(const __nv_bfloat16* __restrict__ data,
                                                                       const float* __restrict__ rois, int C, int H, int W,
                                                                       int PH, int PW, float spatial_scale,
                                                                       __half* __restrict__ out) {
  const int ph = blockIdx.x, n = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* roi = rois + 5 * n;
  const int b = (int)roi[0];
  const int rsw = (int)roundf(roi[1] * spatial_scale), rsh = (int)roundf(roi[2] * spatial_scale);
  const int rew = (int)roundf(roi[3] * spatial_scale), reh = (int)roundf(roi[4] * spatial_scale);
  const int rh = max(reh - rsh + 1, 1), rw = max(rew - rsw + 1, 1);
  const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
  int hs = (int)floorf((float)ph * bh), he = (int)ceilf((float)(ph + 1) * bh);
  hs = min(max(hs + rsh, 0), H); he = min(max(he + rsh, 0), H);
  const __nv_bfloat16* d = data + (size_t)b * H * W * C;
  for (int pw = warp; pw < PW; pw += 4) {
    int ws = (int)floorf((float)pw * bw), we = (int)ceilf((float)(pw + 1) * bw);
    ws = min(max(ws + rsw, 0), W); we = min(max(we + rsw, 0), W);
    const bool empty = (he <= hs) || (we <= ws);
    for (int c = lane * 8; c < C; c += 256) {
      const uint32_t ninf = 0xFF80FF80u;                                    // (-inf, -inf) in bf16
      __nv_bfloat162 m[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) m[q] = *reinterpret_cast<const __nv_bfloat162*>(&ninf);
      for (int h = hs; h < he; ++h)
        for (int w = ws; w < we; ++w) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(d + ((size_t)h * W + w) * C + c));
          const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], pv[q]);
        }
      __half2 o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(m[q]);
        o[q] = empty ? __floats2half2_rn(0.f, 0.f) : __floats2half2_rn(f.x, f.y);
      }
      *reinterpret_cast<uint4*>(out + ((size_t)n * PH * PW + ph * PW + pw) * C + c) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// operator_cxx/deformable_psroi_pooling.cu:29-49
__device__ __forceinline__ float psroi_bilinear(const float* __restrict__ data, float x, float y, int width) {
  const int x1 = (int)floorf(x), x2 = (int)ceilf(x), y1 = (int)floorf(y), y2 = (int)ceilf(y);
  const float dx = x - (float)x1, dy = y - (float)y1;
  const float v11 = __ldg(data + y1 * width + x1), v12 = __ldg(data + y2 * width + x1);
  const float v21 = __ldg(data + y1 * width + x2), v22 = __ldg(data + y2 * width + x2);
  return (1 - dx) * (1 - dy) * v11 + (1 - dx) * dy * v12 + dx * (1 - dy) * v21 + dx * dy * v22;
}

// operator_cxx/deformable_psroi_pooling.cu:52-138, float/double mixing kept literally
__global__ void __launch_bounds__(256) deform_psroi_fwd_kernel(rn_psroi_desc p, size_t count, const float* __restrict__ data,
                                                               const float* __restrict__ rois,
                                                               const float* __restrict__ trans, float* __restrict__ out,
                                                               float* __restrict__ top_count) {
  const int pooled = p.pooled_size, part_size = p.part_size, spp = p.sample_per_part, gs = p.group_size;
  const int H = p.H, W = p.W;
  const int num_classes = p.no_trans ? 1 : p.num_classes;
  const int channels_each_class = p.no_trans ? p.output_dim : p.output_dim / num_classes;
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < count;
       index += (size_t)gridDim.x * blockDim.x) {
    const int pw = index % pooled, ph = (index / pooled) % pooled;
    const int ctop = (index / pooled / pooled) % p.output_dim;
    const int n = index / pooled / pooled / p.output_dim;
    const float* roi = rois + 5 * n;
    const int b = (int)roi[0];
    const float rsw = (float)((double)(roundf(roi[1]) * p.spatial_scale) - 0.5);
    const float rsh = (float)((double)(roundf(roi[2]) * p.spatial_scale) - 0.5);
    const float rew = (float)((double)((float)((double)roundf(roi[3]) + 1.) * p.spatial_scale) - 0.5);
    const float reh = (float)((double)((float)((double)roundf(roi[4]) + 1.) * p.spatial_scale) - 0.5);
    const float roi_w = (float)fmax((double)(rew - rsw), 0.1);
    const float roi_h = (float)fmax((double)(reh - rsh), 0.1);
    const float bin_h = roi_h / (float)pooled, bin_w = roi_w / (float)pooled;
    const float sub_h = bin_h / (float)spp, sub_w = bin_w / (float)spp;
    const int part_h = (int)floorf((float)ph / pooled * part_size);
    const int part_w = (int)floorf((float)pw / pooled * part_size);
    const int class_id = ctop / channels_each_class;
    const float tx = p.no_trans ? 0.f
        : trans[(((size_t)(n * num_classes + class_id) * 2) * part_size + part_h) * part_size + part_w] * p.trans_std;
    const float ty = p.no_trans ? 0.f
        : trans[(((size_t)(n * num_classes + class_id) * 2 + 1) * part_size + part_h) * part_size + part_w] * p.trans_std;
    float wstart = (float)pw * bin_w + rsw; wstart += tx * roi_w;
    float hstart = (float)ph * bin_h + rsh; hstart += ty * roi_h;
    float sum = 0.f; int cnt = 0;
    int gw = (int)floorf((float)pw * gs / pooled), gh = (int)floorf((float)ph * gs / pooled);
    gw = min(max(gw, 0), gs - 1); gh = min(max(gh, 0), gs - 1);
    const float* d0 = data + (size_t)b * p.channels * H * W;
    const int c = (ctop * gs + gh) * gs + gw;
    for (int ih = 0; ih < spp; ++ih)
      for (int iw = 0; iw < spp; ++iw) {
        float w = wstart + iw * sub_w, h = hstart + ih * sub_h;
        if ((double)w < -0.5 || (double)w > W - 0.5 || (double)h < -0.5 || (double)h > H - 0.5) continue;
        w = (float)fmin(fmax((double)w, 0.), W - 1.);
        h = (float)fmin(fmax((double)h, 0.), H - 1.);
        sum += psroi_bilinear(d0 + (size_t)c * H * W, w, h, W);
        cnt++;
      }
    out[index] = cnt == 0 ? 0.f : sum / cnt;
    if (top_count) top_count[index] = (float)cnt;
  }
}

// ---- backward (training) -----------------------------------------------------------------------------------------
// ROIPooling backward (MXNet 1.1.0 roi_pooling.cu ROIPoolBackwardAcc, not in tree): each pooled cell sends its gradient
// to the argmax element the forward recorded.  Scatter form (one thread per pooled cell, red.global.add) instead of the
// reference's gather-over-all-rois form: work is R*C*PH*PW instead of B*C*H*W*R.
__global__ void __launch_bounds__(256) roi_pool_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ argmax,
                                                           const float* __restrict__ rois, size_t count, int C, int H,
                                                           int W, int PHW, float* __restrict__ ddata) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < count;
       index += (size_t)gridDim.x * blockDim.x) {
    const int a = argmax[index];
    if (a < 0) continue;
    const int c = (index / PHW) % C;
    const int n = index / PHW / C;
    const int b = (int)rois[5 * n];
    atomicAdd(ddata + ((size_t)b * C + c) * H * W + a, dout[index]);
  }
}

// DeformablePSROIPoolBackwardAccKernel, operator_cxx/deformable_psroi_pooling.cu:177-289 (same float/double mixing as the
// forward above); ddata / dtrans must be zero on entry.
__global__ void __launch_bounds__(256) deform_psroi_bwd_kernel(rn_psroi_desc p, size_t count, const float* __restrict__ dout,
                                                               const float* __restrict__ top_count,
                                                               const float* __restrict__ data, const float* __restrict__ rois,
                                                               const float* __restrict__ trans, float* __restrict__ ddata,
                                                               float* __restrict__ dtrans) {
  const int pooled = p.pooled_size, part_size = p.part_size, spp = p.sample_per_part, gs = p.group_size;
  const int H = p.H, W = p.W;
  const int num_classes = p.no_trans ? 1 : p.num_classes;
  const int channels_each_class = p.no_trans ? p.output_dim : p.output_dim / num_classes;
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < count;
       index += (size_t)gridDim.x * blockDim.x) {
    const float tc = top_count[index];
    if (tc <= 0) continue;
    const int pw = index % pooled, ph = (index / pooled) % pooled;
    const int ctop = (index / pooled / pooled) % p.output_dim;
    const int n = index / pooled / pooled / p.output_dim;
    const float* roi = rois + 5 * n;
    const int b = (int)roi[0];
    const float rsw = (float)((double)(roundf(roi[1]) * p.spatial_scale) - 0.5);
    const float rsh = (float)((double)(roundf(roi[2]) * p.spatial_scale) - 0.5);
    const float rew = (float)((double)((float)((double)roundf(roi[3]) + 1.) * p.spatial_scale) - 0.5);
    const float reh = (float)((double)((float)((double)roundf(roi[4]) + 1.) * p.spatial_scale) - 0.5);
    const float roi_w = (float)fmax((double)(rew - rsw), 0.1);
    const float roi_h = (float)fmax((double)(reh - rsh), 0.1);
    const float bin_h = roi_h / (float)pooled, bin_w = roi_w / (float)pooled;
    const float sub_h = bin_h / (float)spp, sub_w = bin_w / (float)spp;
    const int part_h = (int)floorf((float)ph / pooled * part_size);
    const int part_w = (int)floorf((float)pw / pooled * part_size);
    const int class_id = ctop / channels_each_class;
    const size_t tix = (((size_t)(n * num_classes + class_id) * 2) * part_size + part_h) * part_size + part_w;
    const size_t tiy = (((size_t)(n * num_classes + class_id) * 2 + 1) * part_size + part_h) * part_size + part_w;
    const float tx = p.no_trans ? 0.f : trans[tix] * p.trans_std;
    const float ty = p.no_trans ? 0.f : trans[tiy] * p.trans_std;
    float wstart = (float)pw * bin_w + rsw; wstart += tx * roi_w;
    float hstart = (float)ph * bin_h + rsh; hstart += ty * roi_h;
    const float diff_val = dout[index] / tc;
    int gw = (int)floorf((float)pw * gs / pooled), gh = (int)floorf((float)ph * gs / pooled);
    gw = min(max(gw, 0), gs - 1); gh = min(max(gh, 0), gs - 1);
    const int c = (ctop * gs + gh) * gs + gw;
    const float* d0 = data + ((size_t)b * p.channels + c) * H * W;
    float* g0 = ddata + ((size_t)b * p.channels + c) * H * W;
    float acc_x = 0.f, acc_y = 0.f;
    for (int ih = 0; ih < spp; ++ih)
      for (int iw = 0; iw < spp; ++iw) {
        float w = wstart + iw * sub_w, h = hstart + ih * sub_h;
        if ((double)w < -0.5 || (double)w > W - 0.5 || (double)h < -0.5 || (double)h > H - 0.5) continue;
        w = (float)fmin(fmax((double)w, 0.), W - 1.);
        h = (float)fmin(fmax((double)h, 0.), H - 1.);
        const int x0 = (int)floorf(w), x1 = (int)ceilf(w), y0 = (int)floorf(h), y1 = (int)ceilf(h);
        const float dx = w - x0, dy = h - y0;
        atomicAdd(g0 + y0 * W + x0, (1 - dx) * (1 - dy) * diff_val);
        atomicAdd(g0 + y1 * W + x0, (1 - dx) * dy * diff_val);
        atomicAdd(g0 + y0 * W + x1, dx * (1 - dy) * diff_val);
        atomicAdd(g0 + y1 * W + x1, dx * dy * diff_val);
        if (p.no_trans) continue;
        const float U00 = __ldg(d0 + y0 * W + x0), U01 = __ldg(d0 + y1 * W + x0);
        const float U10 = __ldg(d0 + y0 * W + x1), U11 = __ldg(d0 + y1 * W + x1);
        float diff_x = (U11 * dy + U10 * (1 - dy) - U01 * dy - U00 * (1 - dy)) * p.trans_std * diff_val;
        diff_x *= roi_w;
        float diff_y = (U11 * dx + U01 * (1 - dx) - U10 * dx - U00 * (1 - dx)) * p.trans_std * diff_val;
        diff_y *= roi_h;
        acc_x += diff_x; acc_y += diff_y;
      }
    if (!p.no_trans) {      // one atomic pair per pooled cell (the reference issues one per sample)
      atomicAdd(dtrans + tix, acc_x);
      atomicAdd(dtrans + tiy, acc_y);
    }
  }
}

static inline int grid_for(size_t count) {
  size_t blocks = (count + 255) / 256;
  size_t cap = (size_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace rn

extern "C" int rn_roi_pool_fwd(const float* data, const float* rois, int32_t R, int32_t C, int32_t H, int32_t W,
                               int32_t PH, int32_t PW, float spatial_scale, float* out, int32_t* argmax,
                               rn_stream_t stream) {
  RN_CHECK_ARG(R >= 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0, "rn_roi_pool_fwd: bad arguments");
  if (R == 0) return RN_OK;
  RN_CHECK_ARG(data && rois && out, "rn_roi_pool_fwd: null pointer");
  size_t count = (size_t)R * C * PH * PW;
  rn::roi_pool_fwd_kernel<<<rn::grid_for(count), 256, 0, (cudaStream_t)stream>>>(data, rois, count, C, H, W, PH, PW,
                                                                                spatial_scale, out, argmax);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_deform_psroi_pool_fwd(const rn_psroi_desc* desc, const float* data, const float* rois,
                                        const float* trans, float* out, float* top_count, rn_stream_t stream) {
  RN_CHECK_ARG(desc, "rn_deform_psroi_pool_fwd: null descriptor");
  if (desc->R == 0) return RN_OK;
  RN_CHECK_ARG(data && rois && out, "rn_deform_psroi_pool_fwd: null argument");
  rn_psroi_desc p = *desc;
  if (p.part_size == 0) p.part_size = p.pooled_size;
  RN_CHECK_ARG(p.no_trans || trans, "rn_deform_psroi_pool_fwd: trans required when no_trans == 0");
  RN_CHECK_ARG(p.group_size > 0 && p.pooled_size > 0 && p.sample_per_part > 0 && p.output_dim > 0,
               "rn_deform_psroi_pool_fwd: bad geometry");
  RN_CHECK_ARG(p.channels == p.output_dim * p.group_size * p.group_size,
               "rn_deform_psroi_pool_fwd: channels %d != output_dim*group_size^2 = %d", p.channels,
               p.output_dim * p.group_size * p.group_size);
  if (!p.no_trans) RN_CHECK_ARG(p.num_classes > 0 && p.output_dim % p.num_classes == 0, "rn_deform_psroi_pool_fwd: bad num_classes");
  if (p.R == 0) return RN_OK;
  size_t count = (size_t)p.R * p.output_dim * p.pooled_size * p.pooled_size;
  rn::deform_psroi_fwd_kernel<<<rn::grid_for(count), 256, 0, (cudaStream_t)stream>>>(p, count, data, rois, trans, out,
                                                                                    top_count);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_roi_pool_nhwc_f16_fwd(const float* data_nhwc, const float* rois, int32_t R, int32_t C, int32_t H,
                                        int32_t W, int32_t PH, int32_t PW, float spatial_scale, void* out_f16,
                                        rn_stream_t stream) {
  RN_CHECK_ARG(data_nhwc && rois && out_f16 && R >= 0 && C > 0 && (C % 2) == 0 && H > 0 && W > 0 && PH > 0 && PW > 0,
               "rn_roi_pool_nhwc_f16_fwd: bad arguments (C must be even)");
  if (R == 0) return RN_OK;
  rn::roi_pool_nhwc_f16_kernel<<<dim3(PH * PW, R), 128, 0, (cudaStream_t)stream>>>(data_nhwc, rois, C, H, W, PH, PW,
                                                                                 spatial_scale, (__half*)out_f16);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_roi_pool_bwd(const float* dout, const int32_t* argmax, const float* rois, int32_t R, int32_t B, int32_t C,
                               int32_t H, int32_t W, int32_t PH, int32_t PW, float* ddata, rn_stream_t stream) {
  RN_CHECK_ARG(R >= 0 && B > 0 && C > 0 && H > 0 && W > 0 && PH > 0 && PW > 0 && ddata, "rn_roi_pool_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  RN_CUDA(cudaMemsetAsync(ddata, 0, sizeof(float) * (size_t)B * C * H * W, st));
  if (R == 0) return RN_OK;
  RN_CHECK_ARG(dout && argmax && rois, "rn_roi_pool_bwd: null pointer");
  size_t count = (size_t)R * C * PH * PW;
  rn::roi_pool_bwd_kernel<<<rn::grid_for(count), 256, 0, st>>>(dout, argmax, rois, count, C, H, W, PH * PW, ddata);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_deform_psroi_pool_bwd(const rn_psroi_desc* desc, int32_t B, const float* dout, const float* top_count,
                                        const float* data, const float* rois, const float* trans, float* ddata,
                                        float* dtrans, rn_stream_t stream) {
  RN_CHECK_ARG(desc && B > 0 && ddata, "rn_deform_psroi_pool_bwd: bad arguments");
  rn_psroi_desc p = *desc;
  if (p.part_size == 0) p.part_size = p.pooled_size;
  RN_CHECK_ARG(p.no_trans || (trans && dtrans), "rn_deform_psroi_pool_bwd: trans / dtrans required when no_trans == 0");
  RN_CHECK_ARG(p.group_size > 0 && p.pooled_size > 0 && p.sample_per_part > 0 && p.output_dim > 0,
               "rn_deform_psroi_pool_bwd: bad geometry");
  RN_CHECK_ARG(p.channels == p.output_dim * p.group_size * p.group_size,
               "rn_deform_psroi_pool_bwd: channels %d != output_dim*group_size^2", p.channels);
  if (!p.no_trans) RN_CHECK_ARG(p.num_classes > 0 && p.output_dim % p.num_classes == 0, "rn_deform_psroi_pool_bwd: bad num_classes");
  cudaStream_t st = (cudaStream_t)stream;
  RN_CUDA(cudaMemsetAsync(ddata, 0, sizeof(float) * (size_t)B * p.channels * p.H * p.W, st));
  if (!p.no_trans && p.R > 0)
    RN_CUDA(cudaMemsetAsync(dtrans, 0, sizeof(float) * (size_t)p.R * 2 * p.num_classes * p.part_size * p.part_size, st));
  if (p.R == 0) return RN_OK;
  RN_CHECK_ARG(dout && top_count && data && rois, "rn_deform_psroi_pool_bwd: null pointer");
  size_t count = (size_t)p.R * p.output_dim * p.pooled_size * p.pooled_size;
  rn::deform_psroi_bwd_kernel<<<rn::grid_for(count), 256, 0, st>>>(p, count, dout, top_count, data, rois, trans, ddata, dtrans);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_roi_pool_nhwc_bf16in_f16_fwd(const void* data_nhwc_bf16, const float* rois, int32_t R, int32_t C, int32_t H,
                                               int32_t W, int32_t PH, int32_t PW, float spatial_scale, void* out_f16,
                                               rn_stream_t stream) {
  RN_CHECK_ARG(data_nhwc_bf16 && rois && out_f16 && R >= 0 && C > 0 && (C % 8) == 0 && H > 0 && W > 0 && PH > 0 && PW > 0,
               "rn_roi_pool_nhwc_bf16in_f16_fwd: bad arguments (C must be a multiple of 8)");
  if (R == 0) return RN_OK;
  rn::roi_pool_nhwc_bf16in_f16_kernel<<<dim3(PH, R), 128, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)data_nhwc_bf16, rois, C, H, W, PH, PW, spatial_scale, (__half*)out_f16);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
