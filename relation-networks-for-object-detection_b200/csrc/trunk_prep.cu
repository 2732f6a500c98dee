// trunk_prep.cu -- the two memory-bound steps either side of the (library, cuDNN) stem convolution, hand-written because
// the stock kernels were 10x off the HBM roofline in the step timeline (tools/timeline.py: 7x7/2 conv on a 3-channel input
// 149 us incl. cuDNN's channel padding, ATen max-pool 47 us):
//   rn_image_s2d_bf16      fp32 [3,H,W] image -> bf16 channels-last [Hs, Ws, 16] space-to-depth(2) of the 3-pixel zero-
//                          padded image: the 7x7 stride-2 stem conv (resnet_v1_101_rcnn_base.py: conv1) becomes a 4x4
//                          stride-1 conv over 12 (+4 zero) channels -- same arithmetic, tensor-core friendly K = 256
//   rn_maxpool3x3s2_nhwc_bf16   3x3 / stride 2 / ceil-mode max pool (pool1, same file) on a channels-last bf16 map
// HBM roofline: bytes in + bytes out (7.2 MB + 4.9 MB; 19.2 MB + 4.8 MB); one thread handles 16 B of output channels.
#include "common.cuh"
#include "gemm_tc.cuh"
#include <cuda_bf16.h>

namespace rn {

// y[I][J][c*4 + r*2 + s] = xpad[c][2I + r][2J + s], xpad = image zero-padded by `pad` on every side; channels 12..15 = 0
__global__ void __launch_bounds__(256) image_s2d_bf16_kernel(const float* __restrict__ img, int H, int W, int pad, int Hs,
                                                             int Ws, __nv_bfloat16* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= Hs * Ws) return;
  const int I = t / Ws, J = t - I * Ws;
  __align__(16) __nv_bfloat16 v[16];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int y = 2 * I + r - pad, x = 2 * J + s - pad;
        const float f = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + ((size_t)c * H + y) * W + x) : 0.f;
        v[c * 4 + r * 2 + s] = __float2bfloat16_rn(f);
      }
#pragma unroll
  for (int k = 12; k < 16; ++k) v[k] = __float2bfloat16_rn(0.f);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)t * 16);
  dst[0] = reinterpret_cast<const uint4*>(v)[0];
  dst[1] = reinterpret_cast<const uint4*>(v)[1];
}

// one thread = one output pixel x 8 channels (16 B); window clipped at the bottom/right edge (ceil mode, no padding)
__global__ void __launch_bounds__(256) maxpool3x3s2_nhwc_bf16_kernel(const __nv_bfloat16* __restrict__ in, int H, int W, int C8,
                                                                     int Ho, int Wo, __nv_bfloat16* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)Ho * Wo * C8;
  if (t >= total) return;
  const int c8 = (int)(t % C8);
  const int ow = (int)((t / C8) % Wo), oh = (int)(t / ((size_t)C8 * Wo));
  const int h0 = oh * 2, w0 = ow * 2;
  const int h1 = min(h0 + 3, H), w1 = min(w0 + 3, W);
  __nv_bfloat162 m[4];
  bool first = true;
  for (int h = h0; h < h1; ++h)
    for (int w = w0; w < w1; ++w) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((size_t)h * W + w) * C8 * 8) + c8);
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
      if (first) { m[0] = p[0]; m[1] = p[1]; m[2] = p[2]; m[3] = p[3]; first = false; }
      else { m[0] = __hmax2(m[0], p[0]); m[1] = __hmax2(m[1], p[1]); m[2] = __hmax2(m[2], p[2]); m[3] = __hmax2(m[3], p[3]); }
    }
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&m[0]); o.y = *reinterpret_cast<uint32_t*>(&m[1]);
  o.z = *reinterpret_cast<uint32_t*>(&m[2]); o.w = *reinterpret_cast<uint32_t*>(&m[3]);
  reinterpret_cast<uint4*>(out + ((size_t)oh * Wo + ow) * C8 * 8)[c8] = o;
}

// RPN head (resnet_v1_101_rcnn_base.py:685-693 + the 2-way softmax of get_symbol): the two 1x1 convolutions rpn_cls_score
// (2A) / rpn_bbox_pred (4A) over the 512-channel rpn_conv map, the {bg, fg} softmax per anchor and the fp32 NCHW outputs
// `proposal` reads -- one kernel instead of ten library launches (2 GEMMs, 2 bias adds, casts, reshapes, softmax: 33 us on the
// critical branch of the step, tools/timeline.py).  CTA = 32 positions x all 6A outputs; K streamed through shared memory in
// chunks of 64; fp32 accumulation.  r: channels-last bf16 [HW, Cin]; Wc [2A, Cin], Wb [4A, Cin], biases bf16.
constexpr int kRpnPos = 32, kRpnKc = 64, kRpnMaxO = 12;     // up to 8 * 12 = 96 outputs per position (A <= 16)
__global__ void __launch_bounds__(256) rpn_head_kernel(const __nv_bfloat16* __restrict__ r, int HW, int Cin, int A,
                                                       const __nv_bfloat16* __restrict__ Wc, const __nv_bfloat16* __restrict__ bc,
                                                       const __nv_bfloat16* __restrict__ Wb, const __nv_bfloat16* __restrict__ bb,
                                                       float* __restrict__ prob, float* __restrict__ bbox) {
  extern __shared__ float sm_rpn[];
  const int NO = 6 * A;                                   // outputs per position: 2A scores then 4A deltas
  float* Ws = sm_rpn;                                     // [NO][kRpnKc + 1]
  float* Rs = Ws + NO * (kRpnKc + 1);                     // [kRpnPos][kRpnKc + 1]
  float* Ss = Rs + kRpnPos * (kRpnKc + 1);                // [NO][kRpnPos + 1]  accumulated outputs
  const int p0 = blockIdx.x * kRpnPos, tid = threadIdx.x;
  // thread -> (position p = lane, output group og = warp): outputs og, og + 8, ... ; a warp reads one weight (broadcast) and
  // 32 different positions (row stride 65 words: conflict-free)
  const int p = tid & 31, og = tid >> 5;
  float acc[kRpnMaxO];
#pragma unroll
  for (int j = 0; j < kRpnMaxO; ++j) acc[j] = 0.f;
  for (int k0 = 0; k0 < Cin; k0 += kRpnKc) {
    for (int i = tid; i < NO * kRpnKc; i += 256) {
      const int oo = i / kRpnKc, kk = i % kRpnKc;
      const __nv_bfloat16 w = oo < 2 * A ? Wc[(size_t)oo * Cin + k0 + kk] : Wb[(size_t)(oo - 2 * A) * Cin + k0 + kk];
      Ws[oo * (kRpnKc + 1) + kk] = __bfloat162float(w);
    }
    for (int i = tid; i < kRpnPos * kRpnKc; i += 256) {
      const int pp = i / kRpnKc, kk = i % kRpnKc;
      Rs[pp * (kRpnKc + 1) + kk] = p0 + pp < HW ? __bfloat162float(r[(size_t)(p0 + pp) * Cin + k0 + kk]) : 0.f;
    }
    __syncthreads();
    const float* rrow = Rs + p * (kRpnKc + 1);
#pragma unroll 4
    for (int kk = 0; kk < kRpnKc; ++kk) {
      const float rv = rrow[kk];
#pragma unroll
      for (int j = 0; j < kRpnMaxO; ++j) {
        const int o = og + 8 * j;
        if (o < NO) acc[j] = fmaf(Ws[o * (kRpnKc + 1) + kk], rv, acc[j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < kRpnMaxO; ++j) {
    const int o = og + 8 * j;
    if (o < NO) Ss[o * (kRpnPos + 1) + p] = acc[j] + __bfloat162float(o < 2 * A ? bc[o] : bb[o - 2 * A]);
  }
  __syncthreads();
  // outputs, position-fastest (NCHW): prob[c][p] = softmax over {c, c +- A}; bbox[c][p]
  for (int i = tid; i < NO * kRpnPos; i += 256) {
    const int oo = i / kRpnPos, pp = i % kRpnPos;
    if (p0 + pp >= HW) continue;
    const float v = Ss[oo * (kRpnPos + 1) + pp];
    if (oo < 2 * A) {
      const float other = Ss[(oo < A ? oo + A : oo - A) * (kRpnPos + 1) + pp];
      const float m = fmaxf(v, other);
      const float e = expf(v - m), eo = expf(other - m);
      prob[(size_t)oo * HW + p0 + pp] = e / (e + eo);
    } else {
      bbox[(size_t)(oo - 2 * A) * HW + p0 + pp] = v;
    }
  }
}

// ---- tensor-core form of the same head (the hot path).  The SIMT kernel above is 79 us at 38 x 63 x 512 -- alone on the
// critical branch rpn_conv -> proposal (profiles/r02_timeline_step.json) -- because every FMA costs two shared loads; the
// head is a [HW, Cin] x [Cin, 6A] GEMM, so it goes through the tcgen05 GEMM with bf16 operands exactly as cuDNN left them
// (weights of both 1x1 convs concatenated once into one [6A, Cin] block), then one small kernel does the {bg, fg} softmax
// and the position-fastest (NCHW) fp32 layout `proposal` reads.
__global__ void rpn_pack_kernel(const __nv_bfloat16* __restrict__ Wc, const __nv_bfloat16* __restrict__ bc,
                                const __nv_bfloat16* __restrict__ Wb, const __nv_bfloat16* __restrict__ bb, int Cin, int A,
                                __nv_bfloat16* __restrict__ W, float* __restrict__ bias) {
  const int NO = 6 * A;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NO * Cin; i += gridDim.x * blockDim.x) {
    const int o = i / Cin, k = i - o * Cin;
    W[i] = o < 2 * A ? Wc[(size_t)o * Cin + k] : Wb[(size_t)(o - 2 * A) * Cin + k];
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < NO; o += gridDim.x * blockDim.x)
    bias[o] = __bfloat162float(o < 2 * A ? bc[o] : bb[o - 2 * A]);
}

// y [HW, NO] fp32 (bias added) -> prob [2A, HW] (softmax over {c, c +- A}), bbox [4A, HW]; 32 positions per CTA through smem
__global__ void __launch_bounds__(256) rpn_finish_kernel(const float* __restrict__ y, int HW, int A, float* __restrict__ prob,
                                                         float* __restrict__ bbox) {
  extern __shared__ float ys[];                             // [32][NO + 1]
  const int NO = 6 * A, p0 = blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * NO; i += blockDim.x) {
    const int pp = i / NO, o = i - pp * NO;
    ys[pp * (NO + 1) + o] = p0 + pp < HW ? y[(size_t)(p0 + pp) * NO + o] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NO * 32; i += blockDim.x) {
    const int o = i >> 5, pp = i & 31;
    if (p0 + pp >= HW) continue;
    const float v = ys[pp * (NO + 1) + o];
    if (o < 2 * A) {
      const float other = ys[pp * (NO + 1) + (o < A ? o + A : o - A)];
      const float m = fmaxf(v, other);
      const float e = expf(v - m), eo = expf(other - m);
      prob[(size_t)o * HW + p0 + pp] = e / (e + eo);
    } else {
      bbox[(size_t)(o - 2 * A) * HW + p0 + pp] = v;
    }
  }
}

}  // namespace rn

extern "C" size_t rn_rpn_head_packed_bytes(int32_t Cin, int32_t A) {
  if (Cin <= 0 || A <= 0) return 0;
  return rn::ws_slice((size_t)6 * A * Cin, 2) + rn::ws_slice((size_t)6 * A, 4);
}

extern "C" int rn_rpn_head_pack(const void* Wcls_bf16, const void* bcls_bf16, const void* Wbbox_bf16, const void* bbbox_bf16,
                                int32_t Cin, int32_t A, void* packed, rn_stream_t stream) {
  RN_CHECK_ARG(Wcls_bf16 && bcls_bf16 && Wbbox_bf16 && bbbox_bf16 && packed && Cin > 0 && A >= 1, "rn_rpn_head_pack: bad arguments");
  __nv_bfloat16* W = (__nv_bfloat16*)packed;
  float* bias = (float*)((char*)packed + rn::ws_slice((size_t)6 * A * Cin, 2));
  rn::rpn_pack_kernel<<<rn::cdiv(6 * A * Cin, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)Wcls_bf16, (const __nv_bfloat16*)bcls_bf16, (const __nv_bfloat16*)Wbbox_bf16,
      (const __nv_bfloat16*)bbbox_bf16, Cin, A, W, bias);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" size_t rn_rpn_head_workspace_bytes(int32_t HW, int32_t Cin, int32_t A) {
  if (HW <= 0 || Cin <= 0 || A <= 0) return 0;
  return rn::ws_slice((size_t)HW * 6 * A, 4) + rn::gemm_tc_workspace_bytes(HW, 6 * A, Cin) + 512;
}

extern "C" int rn_rpn_head_packed_fwd(const void* r_nhwc_bf16, int32_t HW, int32_t Cin, int32_t A, const void* packed, float* prob,
                                      float* bbox, void* wsp, size_t ws_bytes, rn_stream_t stream) {
  using namespace rn;
  RN_CHECK_ARG(r_nhwc_bf16 && packed && prob && bbox && wsp, "rn_rpn_head_packed_fwd: null pointer");
  RN_CHECK_ARG(HW > 0 && Cin > 0 && Cin % 8 == 0 && A >= 1 && A <= 64, "rn_rpn_head_packed_fwd: need Cin %% 8 == 0 and A <= 64");
  RN_CHECK_ARG(is_sm100(), "rn_rpn_head_packed_fwd: tcgen05 GEMM needs an sm_100 device (rn_rpn_head_fwd is the portable form)");
  const int NO = 6 * A;
  Workspace ws(wsp, ws_bytes);
  float* y = ws.take<float>((size_t)HW * NO);
  if (!y) { set_error("rn_rpn_head_packed_fwd: workspace too small (%zu < %zu)", ws_bytes, rn_rpn_head_workspace_bytes(HW, Cin, A)); return RN_ERR_WORKSPACE; }
  const float* bias = (const float*)((const char*)packed + ws_slice((size_t)NO * Cin, 2));
  cudaStream_t st = (cudaStream_t)stream;
  int r = gemm_tc(st, (const __half*)r_nhwc_bf16, Cin, (const __half*)packed, Cin, HW, NO, Cin, bias, 0, 0, y, NO, nullptr, 0,
                  ws.base + ws.off, ws.size - ws.off, nullptr, /*bf16=*/true);
  if (r) return r;
  rpn_finish_kernel<<<cdiv(HW, 32), 256, (size_t)32 * (NO + 1) * sizeof(float), st>>>(y, HW, A, prob, bbox);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_rpn_head_fwd(const void* r_nhwc_bf16, int32_t HW, int32_t Cin, int32_t A, const void* Wcls_bf16,
                               const void* bcls_bf16, const void* Wbbox_bf16, const void* bbbox_bf16, float* prob,
                               float* bbox, rn_stream_t stream) {
  RN_CHECK_ARG(r_nhwc_bf16 && Wcls_bf16 && bcls_bf16 && Wbbox_bf16 && bbbox_bf16 && prob && bbox, "rn_rpn_head_fwd: null pointer");
  RN_CHECK_ARG(HW > 0 && Cin > 0 && Cin % rn::kRpnKc == 0 && A >= 1 && 6 * A <= 8 * rn::kRpnMaxO, "rn_rpn_head_fwd: need Cin %% 64 == 0 and A <= 16");
  const int NO = 6 * A;
  const size_t smem = ((size_t)NO * (rn::kRpnKc + 1) + (size_t)rn::kRpnPos * (rn::kRpnKc + 1) + (size_t)NO * (rn::kRpnPos + 1)) * sizeof(float);
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RN_CUDA(cudaFuncSetAttribute(rn::rpn_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  rn::rpn_head_kernel<<<rn::cdiv(HW, rn::kRpnPos), 256, smem, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)r_nhwc_bf16, HW, Cin, A, (const __nv_bfloat16*)Wcls_bf16, (const __nv_bfloat16*)bcls_bf16,
      (const __nv_bfloat16*)Wbbox_bf16, (const __nv_bfloat16*)bbbox_bf16, prob, bbox);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_image_s2d_bf16(const float* image_chw, int32_t H, int32_t W, int32_t pad, void* out_nhwc16_bf16,
                                 rn_stream_t stream) {
  RN_CHECK_ARG(image_chw && out_nhwc16_bf16 && H > 0 && W > 0 && pad >= 0, "rn_image_s2d_bf16: bad arguments");
  RN_CHECK_ARG(((H + 2 * pad) % 2) == 0 && ((W + 2 * pad) % 2) == 0, "rn_image_s2d_bf16: padded size must be even (%d x %d, pad %d)", H, W, pad);
  const int Hs = (H + 2 * pad) / 2, Ws = (W + 2 * pad) / 2;
  rn::image_s2d_bf16_kernel<<<rn::cdiv(Hs * Ws, 256), 256, 0, (cudaStream_t)stream>>>(image_chw, H, W, pad, Hs, Ws,
                                                                                     (__nv_bfloat16*)out_nhwc16_bf16);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_maxpool3x3s2_nhwc_bf16(const void* in_nhwc, int32_t H, int32_t W, int32_t C, void* out_nhwc,
                                         rn_stream_t stream) {
  RN_CHECK_ARG(in_nhwc && out_nhwc && H >= 3 && W >= 3 && C > 0 && (C % 8) == 0, "rn_maxpool3x3s2_nhwc_bf16: bad arguments (C %% 8)");
  const int Ho = (H - 3 + 1) / 2 + 1, Wo = (W - 3 + 1) / 2 + 1;      // ceil((H - 3) / 2) + 1
  const size_t total = (size_t)Ho * Wo * (C / 8);
  rn::maxpool3x3s2_nhwc_bf16_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in_nhwc, H, W, C / 8, Ho, Wo, (__nv_bfloat16*)out_nhwc);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
