// psroi.cu -- DeformablePSROIPooling forward / backward (the average ROIAlign of the Deformable Faster-RCNN head when
// no_trans and group_size = 1).  Arithmetic: operator_cxx/deformable_psroi_pooling.cu:29-138 (forward), :177-289
// (backward) -- the per-sample float expressions are kept operation for operation (compiled with -fmad=false) so the
// result is bit-identical to the reference kernels compiled the same way (tests/test_gpu_refpin.py).
//
// Work decomposition (NOT the reference's thread-per-output): the positions, validity and bilinear weights of the
// pooled_size x sample_per_part^2 samples of one bin row depend only on (roi, bin row, class), not on the channel.  One
// CTA owns a (roi, bin row): it evaluates that SAMPLE TABLE once into shared memory and then streams the channels
// through it --
//   * channels-last maps (the hot path: the trunk's conv_new_1 output, fp32 or bf16): a thread owns a (bin, 4- or
//     8-channel vector); every tap is one 16-byte load and consecutive threads read consecutive channels, so a warp
//     fetches whole 128-byte lines and the 16 overlapping samples of a bin hit L1;
//   * NCHW fp32 maps (the reference layout of the C ABI): a thread owns a (channel, bin), consecutive threads consecutive
//     bins of one channel (their taps share cache lines of that channel's plane); samples are added in the reference's order.
// Outputs are staged per bin row and written as runs of pooled_size floats.
#include "common.cuh"
#include <cuda_bf16.h>

namespace rn {

struct PsSample { int o11, o12, o21, o22; float dx, dy; int valid; int pad; };     // tap offsets y*W + x, weights

struct PsRoi { float rsw, rsh, roi_w, roi_h, bin_w, bin_h, sub_w, sub_h; int b; };

__device__ __forceinline__ PsRoi ps_roi(const rn_psroi_desc& p, const float* __restrict__ roi) {
  PsRoi r;                                                      // deformable_psroi_pooling.cu:86-103, float/double mixing kept
  r.b = (int)roi[0];
  r.rsw = (float)((double)(roundf(roi[1]) * p.spatial_scale) - 0.5);
  r.rsh = (float)((double)(roundf(roi[2]) * p.spatial_scale) - 0.5);
  const float rew = (float)((double)((float)((double)roundf(roi[3]) + 1.) * p.spatial_scale) - 0.5);
  const float reh = (float)((double)((float)((double)roundf(roi[4]) + 1.) * p.spatial_scale) - 0.5);
  r.roi_w = (float)fmax((double)(rew - r.rsw), 0.1);
  r.roi_h = (float)fmax((double)(reh - r.rsh), 0.1);
  r.bin_h = r.roi_h / (float)p.pooled_size; r.bin_w = r.roi_w / (float)p.pooled_size;
  r.sub_h = r.bin_h / (float)p.sample_per_part; r.sub_w = r.bin_w / (float)p.sample_per_part;
  return r;
}

// sample table of bin row ph for class `cls`: entry e = pw * spp^2 + ih * spp + iw (:104-131)
__device__ __forceinline__ void ps_build_table(const rn_psroi_desc& p, const PsRoi& r, const float* __restrict__ trans, int n,
                                               int ph, int cls, int num_classes, PsSample* __restrict__ tab) {
  const int spp = p.sample_per_part, spp2 = spp * spp, pooled = p.pooled_size, part = p.part_size;
  for (int e = threadIdx.x; e < pooled * spp2; e += blockDim.x) {
    const int pw = e / spp2, s = e - pw * spp2, ih = s / spp, iw = s - ih * spp;
    const int part_h = (int)floorf((float)ph / pooled * part), part_w = (int)floorf((float)pw / pooled * part);
    const float tx = p.no_trans ? 0.f : trans[(((size_t)(n * num_classes + cls) * 2) * part + part_h) * part + part_w] * p.trans_std;
    const float ty = p.no_trans ? 0.f : trans[(((size_t)(n * num_classes + cls) * 2 + 1) * part + part_h) * part + part_w] * p.trans_std;
    float wstart = (float)pw * r.bin_w + r.rsw; wstart += tx * r.roi_w;
    float hstart = (float)ph * r.bin_h + r.rsh; hstart += ty * r.roi_h;
    float w = wstart + iw * r.sub_w, h = hstart + ih * r.sub_h;
    PsSample t;
    t.valid = !((double)w < -0.5 || (double)w > p.W - 0.5 || (double)h < -0.5 || (double)h > p.H - 0.5);
    w = (float)fmin(fmax((double)w, 0.), p.W - 1.);
    h = (float)fmin(fmax((double)h, 0.), p.H - 1.);
    const int x1 = (int)floorf(w), x2 = (int)ceilf(w), y1 = (int)floorf(h), y2 = (int)ceilf(h);
    t.dx = w - (float)x1; t.dy = h - (float)y1;
    t.o11 = y1 * p.W + x1; t.o12 = y2 * p.W + x1; t.o21 = y1 * p.W + x2; t.o22 = y2 * p.W + x2; t.pad = 0;
    tab[e] = t;
  }
}

// Merged-tap form of the channels-last paths.  The 16 samples of a bin are 0.1 .. 2 cells apart, so their 64 taps fall on a
// handful of distinct feature-map cells (9 .. 16 for a 200-pixel roi): the sample loop is regrouped per CELL,
//   sum_s sum_tap w(s,tap) v(tap)  =  sum_cell (sum_{(s,tap) on cell} w(s,tap)) v(cell),
// the per-cell weights being accumulated once per (roi, bin) into a kPsG x kPsG window and then shared by all channels:
// 4 .. 7 times fewer 16-byte loads and multiply-adds per output than walking the samples (same value up to fp32
// re-association, ~1e-7).  Bins wider than the window (rois over ~800 pixels) keep the sample walk.
constexpr int kPsG = 8;

__device__ __forceinline__ float ps_interp(const PsSample& t, float v11, float v12, float v21, float v22) {   // :44-46
  return (1 - t.dx) * (1 - t.dy) * v11 + (1 - t.dx) * t.dy * v12 + t.dx * (1 - t.dy) * v21 + t.dx * t.dy * v22;
}

// LAYOUT 0: data NCHW fp32; 1: NHWC fp32; 2: NHWC bf16.  out / top_count [R, output_dim, pooled, pooled] fp32.
template <int LAYOUT>
__global__ void __launch_bounds__(256) psroi_fwd_kernel(rn_psroi_desc p, const void* __restrict__ data_,
                                                        const float* __restrict__ rois, const float* __restrict__ trans,
                                                        float* __restrict__ out, float* __restrict__ top_count) {
  extern __shared__ __align__(16) unsigned char ps_smem[];
  const int pooled = p.pooled_size, spp = p.sample_per_part, spp2 = spp * spp, gs = p.group_size;
  PsSample* tab = reinterpret_cast<PsSample*>(ps_smem);                                  // [pooled * spp2]
  int* cnt = reinterpret_cast<int*>(tab + pooled * spp2);                                // [pooled]
  float* stage = reinterpret_cast<float*>(cnt + ((pooled + 3) & ~3));                    // [cec][pooled]
  const int n = blockIdx.y, ph = blockIdx.x;
  // channels-last forms: per bin, the bilinear weights of all its samples merged per feature-map cell (kPsG x kPsG window)
  const int cec_all = p.no_trans ? p.output_dim : p.output_dim / (p.no_trans ? 1 : p.num_classes);
  float* wgrid = stage + (size_t)cec_all * pooled;                                       // [pooled][kPsG * kPsG]
  int* wfoot = reinterpret_cast<int*>(wgrid + pooled * kPsG * kPsG);                     // [pooled][4]: x_lo, y_lo, gw, gh (gw < 0: no grid)
  const int num_classes = p.no_trans ? 1 : p.num_classes;
  const int cec = p.no_trans ? p.output_dim : p.output_dim / num_classes;                // channels of one class
  const PsRoi r = ps_roi(p, rois + 5 * n);
  const int HW = p.H * p.W;
  int gh = (int)floorf((float)ph * gs / pooled);
  gh = min(max(gh, 0), gs - 1);
  for (int cls = 0; cls < num_classes; ++cls) {
    __syncthreads();
    ps_build_table(p, r, trans, n, ph, cls, num_classes, tab);
    __syncthreads();
    if (threadIdx.x < pooled) {
      int c = 0;
      for (int s = 0; s < spp2; ++s) c += tab[threadIdx.x * spp2 + s].valid;
      cnt[threadIdx.x] = c;
    }
    if (LAYOUT != 0) {
      for (int i = threadIdx.x; i < pooled * kPsG * kPsG; i += blockDim.x) wgrid[i] = 0.f;
      if (threadIdx.x < pooled) {
        wfoot[threadIdx.x * 4 + 0] = 1 << 30; wfoot[threadIdx.x * 4 + 1] = 1 << 30;      // x_lo, y_lo
        wfoot[threadIdx.x * 4 + 2] = -1; wfoot[threadIdx.x * 4 + 3] = -1;                // x_hi, y_hi (turned into gw, gh below)
      }
      __syncthreads();
      for (int e = threadIdx.x; e < pooled * spp2; e += blockDim.x) {                    // footprint of every bin, all samples at once
        const PsSample t = tab[e];
        if (!t.valid) continue;
        const int pw = e / spp2;
        const int y1 = t.o11 / p.W, x1 = t.o11 - y1 * p.W, y2 = t.o22 / p.W, x2 = t.o22 - y2 * p.W;
        atomicMin(&wfoot[pw * 4 + 0], x1); atomicMin(&wfoot[pw * 4 + 1], y1);
        atomicMax(&wfoot[pw * 4 + 2], x2); atomicMax(&wfoot[pw * 4 + 3], y2);
      }
      __syncthreads();
      if (threadIdx.x < pooled) {
        const int xl = wfoot[threadIdx.x * 4], yl = wfoot[threadIdx.x * 4 + 1], xh = wfoot[threadIdx.x * 4 + 2], yh = wfoot[threadIdx.x * 4 + 3];
        const int gw = xh - xl + 1, gh2 = yh - yl + 1;
        const bool ok = xh >= 0 && gw <= kPsG && gh2 <= kPsG;
        wfoot[threadIdx.x * 4 + 2] = ok ? gw : (xh < 0 ? 0 : -1); wfoot[threadIdx.x * 4 + 3] = ok ? gh2 : 0;
      }
      __syncthreads();
      for (int e = threadIdx.x; e < pooled * spp2; e += blockDim.x) {
        const int pw = e / spp2;
        const PsSample t = tab[e];
        const int gw = wfoot[pw * 4 + 2];
        if (!t.valid || gw <= 0) continue;
        const int xl = wfoot[pw * 4], yl = wfoot[pw * 4 + 1];
        float* g = wgrid + pw * kPsG * kPsG;
        const int o[4] = {t.o11, t.o12, t.o21, t.o22};
        const float w[4] = {(1 - t.dx) * (1 - t.dy), (1 - t.dx) * t.dy, t.dx * (1 - t.dy), t.dx * t.dy};     // :44-46
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int y = o[k] / p.W, x = o[k] - y * p.W;
          atomicAdd(g + (y - yl) * kPsG + (x - xl), w[k]);
        }
      }
    }
    __syncthreads();
    if (LAYOUT == 0) {
      // NCHW: a thread owns a (channel, bin); consecutive threads take consecutive bins of one channel, whose taps share
      // cache lines of that channel's plane.  The geometry comes from the table (no per-channel recomputation).
      const float* data = reinterpret_cast<const float*>(data_) + (size_t)r.b * p.channels * HW;
      for (int item = threadIdx.x; item < cec * pooled; item += blockDim.x) {
        const int cl = item / pooled, pw = item - cl * pooled;
        int gw = (int)floorf((float)pw * gs / pooled);
        gw = min(max(gw, 0), gs - 1);
        const float* d0 = data + (size_t)(((cls * cec + cl) * gs + gh) * gs + gw) * HW;
        float sum = 0.f;
        for (int s = 0; s < spp2; ++s) {
          const PsSample t = tab[pw * spp2 + s];
          if (t.valid) sum += ps_interp(t, __ldg(d0 + t.o11), __ldg(d0 + t.o12), __ldg(d0 + t.o21), __ldg(d0 + t.o22));
        }
        stage[cl * pooled + pw] = cnt[pw] == 0 ? 0.f : sum / cnt[pw];
      }
    } else {
      constexpr int VEC = LAYOUT == 1 ? 4 : 8;
      const bool vec_ok = gs == 1 && (cec % VEC) == 0 && (p.channels % VEC) == 0;
      if (vec_ok) {
        const int nv = cec / VEC;
        for (int item = threadIdx.x; item < nv * pooled; item += blockDim.x) {
          const int pw = item / nv, cv = item - pw * nv;
          const int c0 = cls * cec + cv * VEC;
          float sum[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) sum[k] = 0.f;
          const int fgw = wfoot[pw * 4 + 2], fgh = wfoot[pw * 4 + 3];
          if (fgw >= 0) {                       // merged-tap form: one load per distinct cell of the bin's footprint
            const int xl = wfoot[pw * 4], yl = wfoot[pw * 4 + 1];
            const float* g = wgrid + pw * kPsG * kPsG;
            const size_t base = (size_t)r.b * HW + (size_t)yl * p.W + xl;
            const int ncell = fgw * fgh;
            // four cells per round: their loads are issued back to back (the walk is latency bound, not bandwidth bound:
            // one dependent L2 round trip per cell and 8 warps per CTA otherwise); zero-weight cells of the window are loaded
            // too -- they are inside the bin's footprint, hence inside the map
            for (int i0 = 0; i0 < ncell; i0 += 4) {
              float wv[4];
              size_t cell[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = min(i0 + u, ncell - 1), cy = i / fgw, cx = i - cy * fgw;
                wv[u] = i0 + u < ncell ? g[cy * kPsG + cx] : 0.f;
                cell[u] = base + (size_t)cy * p.W + cx;
              }
              if (LAYOUT == 1) {
                float4 a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  a[u] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(data_) + cell[u] * p.channels + c0));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  sum[0] += wv[u] * a[u].x; sum[1] += wv[u] * a[u].y; sum[2] += wv[u] * a[u].z; sum[3] += wv[u] * a[u].w;
                }
              } else {
                uint4 a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  a[u] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(data_) + cell[u] * p.channels + c0));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a[u]);
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float2 fa = __bfloat1622float2(pa[q]);
                    sum[2 * q] += wv[u] * fa.x; sum[2 * q + 1] += wv[u] * fa.y;
                  }
                }
              }
            }
          } else
          for (int s = 0; s < spp2; ++s) {
            const PsSample t = tab[pw * spp2 + s];
            if (!t.valid) continue;
            float v11[VEC], v12[VEC], v21[VEC], v22[VEC];
            if (LAYOUT == 1) {
              const float* d = reinterpret_cast<const float*>(data_) + (size_t)r.b * HW * p.channels + c0;
              const float4 a = __ldg(reinterpret_cast<const float4*>(d + (size_t)t.o11 * p.channels));
              const float4 b = __ldg(reinterpret_cast<const float4*>(d + (size_t)t.o12 * p.channels));
              const float4 c = __ldg(reinterpret_cast<const float4*>(d + (size_t)t.o21 * p.channels));
              const float4 e = __ldg(reinterpret_cast<const float4*>(d + (size_t)t.o22 * p.channels));
              v11[0] = a.x; v11[1] = a.y; v11[2] = a.z; v11[3] = a.w; v12[0] = b.x; v12[1] = b.y; v12[2] = b.z; v12[3] = b.w;
              v21[0] = c.x; v21[1] = c.y; v21[2] = c.z; v21[3] = c.w; v22[0] = e.x; v22[1] = e.y; v22[2] = e.z; v22[3] = e.w;
            } else {
              const __nv_bfloat16* d = reinterpret_cast<const __nv_bfloat16*>(data_) + (size_t)r.b * HW * p.channels + c0;
              const uint4 a = __ldg(reinterpret_cast<const uint4*>(d + (size_t)t.o11 * p.channels));
              const uint4 b = __ldg(reinterpret_cast<const uint4*>(d + (size_t)t.o12 * p.channels));
              const uint4 c = __ldg(reinterpret_cast<const uint4*>(d + (size_t)t.o21 * p.channels));
              const uint4 e = __ldg(reinterpret_cast<const uint4*>(d + (size_t)t.o22 * p.channels));
              const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
              const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
              const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&c);
              const __nv_bfloat162* pe = reinterpret_cast<const __nv_bfloat162*>(&e);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 fa = __bfloat1622float2(pa[q]), fb = __bfloat1622float2(pb[q]);
                const float2 fc = __bfloat1622float2(pc[q]), fe = __bfloat1622float2(pe[q]);
                v11[2 * q] = fa.x; v11[2 * q + 1] = fa.y; v12[2 * q] = fb.x; v12[2 * q + 1] = fb.y;
                v21[2 * q] = fc.x; v21[2 * q + 1] = fc.y; v22[2 * q] = fe.x; v22[2 * q + 1] = fe.y;
              }
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) sum[k] += ps_interp(t, v11[k], v12[k], v21[k], v22[k]);
          }
#pragma unroll
          for (int k = 0; k < VEC; ++k) stage[(cv * VEC + k) * pooled + pw] = cnt[pw] == 0 ? 0.f : sum[k] / cnt[pw];
        }
      } else {
        // grouped / odd channel counts on a channels-last map: one (channel, bin) per thread through the same table
        for (int item = threadIdx.x; item < cec * pooled; item += blockDim.x) {
          const int cl = item / pooled, pw = item - cl * pooled;
          int gw = (int)floorf((float)pw * gs / pooled);
          gw = min(max(gw, 0), gs - 1);
          const int c = ((cls * cec + cl) * gs + gh) * gs + gw;
          float sum = 0.f;
          for (int s = 0; s < spp2; ++s) {
            const PsSample t = tab[pw * spp2 + s];
            if (!t.valid) continue;
            float v[4];
            const int o[4] = {t.o11, t.o12, t.o21, t.o22};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const size_t a = ((size_t)r.b * HW + o[k]) * p.channels + c;
              v[k] = LAYOUT == 1 ? __ldg(reinterpret_cast<const float*>(data_) + a)
                                 : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(data_)[a]);
            }
            sum += ps_interp(t, v[0], v[1], v[2], v[3]);
          }
          stage[cl * pooled + pw] = cnt[pw] == 0 ? 0.f : sum / cnt[pw];
        }
      }
    }
    __syncthreads();
    // the bin row of every channel of the class: runs of `pooled` contiguous floats
    for (int it = threadIdx.x; it < cec * pooled; it += blockDim.x) {
      const int cl = it / pooled, pw = it - cl * pooled;
      const size_t o = (((size_t)n * p.output_dim + cls * cec + cl) * pooled + ph) * pooled + pw;
      out[o] = stage[it];
      if (top_count) top_count[o] = (float)cnt[pw];
    }
  }
}

// Backward (:177-289): one CTA per (roi, bin row) with the same sample table; a thread owns a (channel, bin), spreads
// dout / count over the 4 taps of every valid sample with atomics (NCHW fp32 gradient, as the reference) and keeps the
// offset gradient of its cell in registers (one atomic pair per cell instead of one per sample).
__global__ void __launch_bounds__(256) psroi_bwd_kernel(rn_psroi_desc p, const float* __restrict__ dout,
                                                        const float* __restrict__ top_count, const float* __restrict__ data,
                                                        const float* __restrict__ rois, const float* __restrict__ trans,
                                                        float* __restrict__ ddata, float* __restrict__ dtrans) {
  extern __shared__ __align__(16) unsigned char ps_smem[];
  const int pooled = p.pooled_size, spp = p.sample_per_part, spp2 = spp * spp, gs = p.group_size, part = p.part_size;
  PsSample* tab = reinterpret_cast<PsSample*>(ps_smem);
  const int n = blockIdx.y, ph = blockIdx.x;
  const int num_classes = p.no_trans ? 1 : p.num_classes;
  const int cec = p.no_trans ? p.output_dim : p.output_dim / num_classes;
  const PsRoi r = ps_roi(p, rois + 5 * n);
  const int HW = p.H * p.W;
  int gh = (int)floorf((float)ph * gs / pooled);
  gh = min(max(gh, 0), gs - 1);
  const int part_h = (int)floorf((float)ph / pooled * part);
  for (int cls = 0; cls < num_classes; ++cls) {
    __syncthreads();
    ps_build_table(p, r, trans, n, ph, cls, num_classes, tab);
    __syncthreads();
    for (int item = threadIdx.x; item < cec * pooled; item += blockDim.x) {
      const int cl = item / pooled, pw = item - cl * pooled;
      const int ctop = cls * cec + cl;
      const size_t oi = (((size_t)n * p.output_dim + ctop) * pooled + ph) * pooled + pw;
      const float tc = top_count[oi];
      if (tc <= 0) continue;
      const float diff_val = dout[oi] / tc;
      int gw = (int)floorf((float)pw * gs / pooled);
      gw = min(max(gw, 0), gs - 1);
      const int c = (ctop * gs + gh) * gs + gw;
      const float* d0 = data + ((size_t)r.b * p.channels + c) * HW;
      float* g0 = ddata + ((size_t)r.b * p.channels + c) * HW;
      float acc_x = 0.f, acc_y = 0.f;
      for (int s = 0; s < spp2; ++s) {
        const PsSample t = tab[pw * spp2 + s];
        if (!t.valid) continue;
        atomicAdd(g0 + t.o11, (1 - t.dx) * (1 - t.dy) * diff_val);
        atomicAdd(g0 + t.o12, (1 - t.dx) * t.dy * diff_val);
        atomicAdd(g0 + t.o21, t.dx * (1 - t.dy) * diff_val);
        atomicAdd(g0 + t.o22, t.dx * t.dy * diff_val);
        if (p.no_trans) continue;
        const float U00 = __ldg(d0 + t.o11), U01 = __ldg(d0 + t.o12), U10 = __ldg(d0 + t.o21), U11 = __ldg(d0 + t.o22);
        float diff_x = (U11 * t.dy + U10 * (1 - t.dy) - U01 * t.dy - U00 * (1 - t.dy)) * p.trans_std * diff_val;
        diff_x *= r.roi_w;
        float diff_y = (U11 * t.dx + U01 * (1 - t.dx) - U10 * t.dx - U00 * (1 - t.dx)) * p.trans_std * diff_val;
        diff_y *= r.roi_h;
        acc_x += diff_x; acc_y += diff_y;
      }
      if (!p.no_trans) {
        const int part_w = (int)floorf((float)pw / pooled * part);
        atomicAdd(dtrans + (((size_t)(n * num_classes + cls) * 2) * part + part_h) * part + part_w, acc_x);
        atomicAdd(dtrans + (((size_t)(n * num_classes + cls) * 2 + 1) * part + part_h) * part + part_w, acc_y);
      }
    }
  }
}

static int psroi_check(const rn_psroi_desc* desc, rn_psroi_desc* p, const void* data, const float* rois, const float* trans,
                       const char* who) {
  RN_CHECK_ARG(desc, "%s: null descriptor", who);
  *p = *desc;
  if (p->part_size == 0) p->part_size = p->pooled_size;
  if (p->R == 0) return RN_OK;
  RN_CHECK_ARG(data && rois, "%s: null argument", who);
  RN_CHECK_ARG(p->no_trans || trans, "%s: trans required when no_trans == 0", who);
  RN_CHECK_ARG(p->group_size > 0 && p->pooled_size > 0 && p->pooled_size <= 32 && p->sample_per_part > 0 &&
               p->sample_per_part <= 8 && p->output_dim > 0, "%s: bad geometry", who);
  RN_CHECK_ARG(p->channels == p->output_dim * p->group_size * p->group_size, "%s: channels %d != output_dim*group_size^2 = %d",
               who, p->channels, p->output_dim * p->group_size * p->group_size);
  if (!p->no_trans) RN_CHECK_ARG(p->num_classes > 0 && p->output_dim % p->num_classes == 0, "%s: bad num_classes", who);
  return RN_OK;
}

static size_t psroi_smem(const rn_psroi_desc& p) {
  const int ncls = p.no_trans ? 1 : p.num_classes, cec = p.no_trans ? p.output_dim : p.output_dim / ncls;
  return sizeof(PsSample) * p.pooled_size * p.sample_per_part * p.sample_per_part + sizeof(int) * ((p.pooled_size + 3) & ~3) +
         sizeof(float) * (size_t)cec * p.pooled_size + sizeof(float) * p.pooled_size * kPsG * kPsG + sizeof(int) * p.pooled_size * 4;
}

template <int LAYOUT>
static int psroi_launch(const rn_psroi_desc& p, const void* data, const float* rois, const float* trans, float* out,
                        float* top_count, cudaStream_t st) {
  const size_t smem = psroi_smem(p);
  RN_CHECK_ARG(smem <= 200 * 1024, "deform_psroi_pool: %zu bytes of shared memory for one class's bin row (output_dim too large)", smem);
  if (smem > 48 * 1024) RN_CUDA(cudaFuncSetAttribute(psroi_fwd_kernel<LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  psroi_fwd_kernel<LAYOUT><<<dim3(p.pooled_size, p.R), 256, smem, st>>>(p, data, rois, trans, out, top_count);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn

extern "C" int rn_deform_psroi_pool_fwd(const rn_psroi_desc* desc, const float* data, const float* rois,
                                        const float* trans, float* out, float* top_count, rn_stream_t stream) {
  rn_psroi_desc p;
  int r = rn::psroi_check(desc, &p, data, rois, trans, "rn_deform_psroi_pool_fwd");
  if (r || p.R == 0) return r;
  RN_CHECK_ARG(out, "rn_deform_psroi_pool_fwd: null output");
  return rn::psroi_launch<0>(p, data, rois, trans, out, top_count, (cudaStream_t)stream);
}

extern "C" int rn_deform_psroi_pool_nhwc_fwd(const rn_psroi_desc* desc, const void* data_nhwc, int32_t data_is_bf16,
                                             const float* rois, const float* trans, float* out, float* top_count,
                                             rn_stream_t stream) {
  rn_psroi_desc p;
  int r = rn::psroi_check(desc, &p, data_nhwc, rois, trans, "rn_deform_psroi_pool_nhwc_fwd");
  if (r || p.R == 0) return r;
  RN_CHECK_ARG(out, "rn_deform_psroi_pool_nhwc_fwd: null output");
  if (data_is_bf16) return rn::psroi_launch<2>(p, data_nhwc, rois, trans, out, top_count, (cudaStream_t)stream);
  return rn::psroi_launch<1>(p, data_nhwc, rois, trans, out, top_count, (cudaStream_t)stream);
}

extern "C" int rn_deform_psroi_pool_bwd(const rn_psroi_desc* desc, int32_t B, const float* dout, const float* top_count,
                                        const float* data, const float* rois, const float* trans, float* ddata,
                                        float* dtrans, rn_stream_t stream) {
  RN_CHECK_ARG(desc && B > 0 && ddata, "rn_deform_psroi_pool_bwd: bad arguments");
  rn_psroi_desc p;
  int r = rn::psroi_check(desc, &p, desc->R ? (const void*)data : (const void*)ddata, desc->R ? rois : (const float*)ddata, trans,
                          "rn_deform_psroi_pool_bwd");
  if (r) return r;
  RN_CHECK_ARG(p.no_trans || dtrans, "rn_deform_psroi_pool_bwd: dtrans required when no_trans == 0");
  cudaStream_t st = (cudaStream_t)stream;
  RN_CUDA(cudaMemsetAsync(ddata, 0, sizeof(float) * (size_t)B * p.channels * p.H * p.W, st));
  if (!p.no_trans && p.R > 0)
    RN_CUDA(cudaMemsetAsync(dtrans, 0, sizeof(float) * (size_t)p.R * 2 * p.num_classes * p.part_size * p.part_size, st));
  if (p.R == 0) return RN_OK;
  RN_CHECK_ARG(dout && top_count, "rn_deform_psroi_pool_bwd: null pointer");
  const size_t smem = sizeof(rn::PsSample) * p.pooled_size * p.sample_per_part * p.sample_per_part;
  rn::psroi_bwd_kernel<<<dim3(p.pooled_size, p.R), 256, smem, st>>>(p, dout, top_count, data, rois, trans, ddata, dtrans);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
