// rois.cu -- ROIPooling (max, MXNet semantics) forward / backward.  (DeformablePSROIPooling lives in psroi.cu.)
// Compiled with -fmad=false: the float op order of the MXNet kernel is kept literally so results are bit-comparable with the
// plain-C oracle (oracle/oracle_c.c) and with torchvision.ops.roi_pool (tests/test_gpu_refpin.py).
//
// HBM layout: data [B,C,H,W] fp32 (reference layout) or channels-last fp32 / bf16 (the trunk's layout: the hot path), out
// [R,C,P,P] fp32 or fp16 [R, P*P, C] (the K order of the permuted fc_new_1 weight).
// Roofline: HBM-write bound -- algorithmic bytes = out (7.5 MB fp16 at R = 300) + the feature map once; the gathers hit L2.
// Measured: bf16-in / fp16-out form 10.5 us inside the step (profiles/r02_timeline_step.json).
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
    const int b = max((int)roi[0], 0);
    const bool padded = (int)roi[0] < 0;           // MXNet: roi_batch_ind < 0 marks a padding roi -> zeros, argmax -1
    const int rsw = (int)roundf(roi[1] * spatial_scale), rsh = (int)roundf(roi[2] * spatial_scale);
    const int rew = (int)roundf(roi[3] * spatial_scale), reh = (int)roundf(roi[4] * spatial_scale);
    const int rh = max(reh - rsh + 1, 1), rw = max(rew - rsw + 1, 1);
    const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
    int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
    int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
    hs = min(max(hs + rsh, 0), H); he = min(max(he + rsh, 0), H);
    ws = min(max(ws + rsw, 0), W); we = min(max(we + rsw, 0), W);
    const bool empty = padded || (he <= hs) || (we <= ws);
    if (padded) { he = hs; we = ws; }
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
  const int b = max((int)roi[0], 0);
  const bool padded = (int)roi[0] < 0;             // padding roi (MXNet semantics): zeros
  const int rsw = (int)roundf(roi[1] * spatial_scale), rsh = (int)roundf(roi[2] * spatial_scale);
  const int rew = (int)roundf(roi[3] * spatial_scale), reh = (int)roundf(roi[4] * spatial_scale);
  const int rh = max(reh - rsh + 1, 1), rw = max(rew - rsw + 1, 1);
  const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
  int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
  int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
  hs = min(max(hs + rsh, 0), H); he = min(max(he + rsh, 0), H);
  ws = min(max(ws + rsw, 0), W); we = min(max(we + rsw, 0), W);
  const bool empty = padded || (he <= hs) || (we <= ws);
  if (padded) { he = hs; we = ws; }
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
  const int b = max((int)roi[0], 0);
  const bool padded = (int)roi[0] < 0;             // padding roi (MXNet semantics): zeros
  const int rsw = (int)roundf(roi[1] * spatial_scale), rsh = (int)roundf(roi[2] * spatial_scale);
  const int rew = (int)roundf(roi[3] * spatial_scale), reh = (int)roundf(roi[4] * spatial_scale);
  const int rh = max(reh - rsh + 1, 1), rw = max(rew - rsw + 1, 1);
  const float bh = (float)rh / (float)PH, bw = (float)rw / (float)PW;
  int hs = (int)floorf((float)ph * bh), he = (int)ceilf((float)(ph + 1) * bh);
  hs = min(max(hs + rsh, 0), H); he = min(max(he + rsh, 0), H);
  if (padded) he = hs;
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

// ---- backward (training) -----------------------------------------------------------------------------------------
// ROIPooling backward (MXNet 1.1.0 roi_pooling.cu ROIPoolBackwardAcc, not in tree): each pooled cell sends its gradient
// to the argmax element the forward recorded.  Scatter form (one thread per pooled cell, red.global.add) instead of the
// reference's gather-over-all-rois form: work is R*C*PH*PW instead of B*C*H*W*R.
__global__ void __launch_bounds__(256) roi_pool_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ argmax,
                                                           const float* __restrict__ rois, size_t count, int B, int C, int H,
                                                           int W, int PHW, float* __restrict__ ddata) {
  for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < count;
       index += (size_t)gridDim.x * blockDim.x) {
    const int a = argmax[index];
    if (a < 0) continue;
    const int c = (index / PHW) % C;
    const int n = index / PHW / C;
    const int b = (int)rois[5 * n];
    if (b < 0 || b >= B) continue;                 // padding roi / index outside the gradient buffer
    atomicAdd(ddata + ((size_t)b * C + c) * H * W + a, dout[index]);
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
  rn::roi_pool_bwd_kernel<<<rn::grid_for(count), 256, 0, st>>>(dout, argmax, rois, count, B, C, H, W, PH * PW, ddata);
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
