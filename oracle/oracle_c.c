/* Plain-C CPU restatement of the ROI / deformable / NMS kernels of the hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): built by oracle/Makefile into oracle/_build/liboracle_c.so and
 * loaded by oracle/rois_np.py; never linked into or called from the product library.
 *
 * Each function cites the reference lines it restates.  The reference has NO CPU implementation of the two
 * operator_cxx kernels (deformable_im2col.h:93-97 is LOG(FATAL); deformable_psroi_pooling.cc:21-37 is an empty
 * stub) and ROIPooling is an MXNet built-in that is not in the tree, so these are restatements of the CUDA
 * kernels / the Fast-RCNN definition: "parity unpinned" (DESIGN.md).
 *
 * Compile with -ffp-contract=off so that no FMA is formed: the CUDA kernels they check are compiled with
 * -fmad=false for the same reason (bit-comparable float arithmetic).
 */
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdint.h>

/* ---------------------------------------------------------------------------------------------------------
 * ROIPooling (max), MXNet 1.1.0 semantics (src/operator/roi_pooling.cc, Caffe-derived); call sites SYM_REL:252-253.
 * data [B,C,H,W], rois [R,5] (batch, x1,y1,x2,y2 image px) -> out [R,C,PH,PW], argmax [R,C,PH,PW] (int, -1 = empty) */
void oracle_roi_pool(const float* data, const float* rois, int R, int C, int H, int W, int PH, int PW,
                     float spatial_scale, float* out, int* argmax) {
  for (int n = 0; n < R; ++n) {
    const float* roi = rois + 5 * n;
    int b = (int)roi[0];
    int rsw = (int)roundf(roi[1] * spatial_scale);
    int rsh = (int)roundf(roi[2] * spatial_scale);
    int rew = (int)roundf(roi[3] * spatial_scale);
    int reh = (int)roundf(roi[4] * spatial_scale);
    int rh = reh - rsh + 1; if (rh < 1) rh = 1;
    int rw = rew - rsw + 1; if (rw < 1) rw = 1;
    float bh = (float)rh / (float)PH;
    float bw = (float)rw / (float)PW;
    for (int c = 0; c < C; ++c) {
      const float* d = data + ((size_t)b * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          int hs = (int)floorf((float)ph * bh), ws = (int)floorf((float)pw * bw);
          int he = (int)ceilf((float)(ph + 1) * bh), we = (int)ceilf((float)(pw + 1) * bw);
          hs = hs + rsh; he = he + rsh; ws = ws + rsw; we = we + rsw;
          if (hs < 0) hs = 0; if (hs > H) hs = H; if (he < 0) he = 0; if (he > H) he = H;
          if (ws < 0) ws = 0; if (ws > W) ws = W; if (we < 0) we = 0; if (we > W) we = W;
          int empty = (he <= hs) || (we <= ws);
          float m = empty ? 0.f : -FLT_MAX;
          int mi = -1;
          for (int h = hs; h < he; ++h)
            for (int w = ws; w < we; ++w)
              if (d[h * W + w] > m) { m = d[h * W + w]; mi = h * W + w; }
          size_t o = (((size_t)n * C + c) * PH + ph) * PW + pw;
          out[o] = m; argmax[o] = mi;
        }
    }
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * DeformablePSROIPoolForwardKernel, relation_rcnn/operator_cxx/deformable_psroi_pooling.cu:52-138 and its bilinear
 * helper :29-49.  float/double mixing follows the CUDA source literally (0.5, 0.1, 0., 1. are double literals). */
static float psroi_bilinear(const float* data, float x, float y, int width, int height) {
  int x1 = (int)floorf(x), x2 = (int)ceilf(x), y1 = (int)floorf(y), y2 = (int)ceilf(y);
  float dx = x - (float)x1, dy = y - (float)y1;
  float v11 = data[y1 * width + x1], v12 = data[y2 * width + x1];
  float v21 = data[y1 * width + x2], v22 = data[y2 * width + x2];
  (void)height;
  return (1 - dx) * (1 - dy) * v11 + (1 - dx) * dy * v12 + dx * (1 - dy) * v21 + dx * dy * v22;
}

void oracle_deform_psroi_pool(const float* data, const float* rois, const float* trans, int R, int channels, int H,
                              int W, int no_trans, float spatial_scale, int output_dim, int group_size,
                              int pooled, int part_size, int sample_per_part, float trans_std, int num_classes,
                              float* out, float* top_count) {
  int channels_each_class = no_trans ? output_dim : output_dim / num_classes;
  if (no_trans) num_classes = 1;
  size_t count = (size_t)R * output_dim * pooled * pooled;
  for (size_t index = 0; index < count; ++index) {
    int pw = index % pooled, ph = (index / pooled) % pooled;
    int ctop = (index / pooled / pooled) % output_dim;
    int n = index / pooled / pooled / output_dim;
    const float* roi = rois + 5 * n;
    int b = (int)roi[0];
    float rsw = (float)((double)(roundf(roi[1]) * spatial_scale) - 0.5);
    float rsh = (float)((double)(roundf(roi[2]) * spatial_scale) - 0.5);
    float rew = (float)((double)((float)((double)roundf(roi[3]) + 1.) * spatial_scale) - 0.5);
    float reh = (float)((double)((float)((double)roundf(roi[4]) + 1.) * spatial_scale) - 0.5);
    float roi_w = (float)fmax((double)(rew - rsw), 0.1);
    float roi_h = (float)fmax((double)(reh - rsh), 0.1);
    float bin_h = roi_h / (float)pooled, bin_w = roi_w / (float)pooled;
    float sub_h = bin_h / (float)sample_per_part, sub_w = bin_w / (float)sample_per_part;
    int part_h = (int)floorf((float)ph / pooled * part_size);
    int part_w = (int)floorf((float)pw / pooled * part_size);
    int class_id = ctop / channels_each_class;
    float tx = no_trans ? 0.f
        : trans[(((size_t)(n * num_classes + class_id) * 2) * part_size + part_h) * part_size + part_w] * trans_std;
    float ty = no_trans ? 0.f
        : trans[(((size_t)(n * num_classes + class_id) * 2 + 1) * part_size + part_h) * part_size + part_w] * trans_std;
    float wstart = (float)pw * bin_w + rsw; wstart += tx * roi_w;
    float hstart = (float)ph * bin_h + rsh; hstart += ty * roi_h;
    float sum = 0.f; int cnt = 0;
    int gw = (int)floorf((float)pw * group_size / pooled), gh = (int)floorf((float)ph * group_size / pooled);
    if (gw < 0) gw = 0; if (gw > group_size - 1) gw = group_size - 1;
    if (gh < 0) gh = 0; if (gh > group_size - 1) gh = group_size - 1;
    const float* d0 = data + (size_t)b * channels * H * W;
    for (int ih = 0; ih < sample_per_part; ++ih)
      for (int iw = 0; iw < sample_per_part; ++iw) {
        float w = wstart + iw * sub_w, h = hstart + ih * sub_h;
        if ((double)w < -0.5 || (double)w > W - 0.5 || (double)h < -0.5 || (double)h > H - 0.5) continue;
        w = (float)fmin(fmax((double)w, 0.), W - 1.);
        h = (float)fmin(fmax((double)h, 0.), H - 1.);
        int c = (ctop * group_size + gh) * group_size + gw;
        sum += psroi_bilinear(d0 + (size_t)c * H * W, w, h, W, H);
        cnt++;
      }
    out[index] = cnt == 0 ? 0.f : sum / cnt;
    top_count[index] = (float)cnt;
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * deformable_im2col_gpu_kernel, relation_rcnn/operator_cxx/nn/deformable_im2col.cuh:216-262 + bilinear :77-113.
 * data_im [C,H,W], offset [dg*2*kh*kw, Ho, Wo] -> col [C*kh*kw, Ho, Wo] */
static float dim2col_bilinear(const float* d, int data_width, int height, int width, float h, float w) {
  int h_low = (int)floorf(h), w_low = (int)floorf(w), h_high, w_high;
  if (h_low >= height - 1) { h_high = h_low = height - 1; h = (float)h_low; } else h_high = h_low + 1;
  if (w_low >= width - 1) { w_high = w_low = width - 1; w = (float)w_low; } else w_high = w_low + 1;
  float lh = h - h_low, lw = w - w_low, hh = 1 - lh, hw = 1 - lw;
  float v1 = d[h_low * data_width + w_low], v2 = d[h_low * data_width + w_high];
  float v3 = d[h_high * data_width + w_low], v4 = d[h_high * data_width + w_high];
  float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
  return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
}

void oracle_deform_im2col(const float* im, const float* off, int C, int H, int W, int kh, int kw, int pad_h, int pad_w,
                          int sh, int sw, int dil_h, int dil_w, int dg, int Ho, int Wo, float* col) {
  int cpg = C / dg;
  for (int c_im = 0; c_im < C; ++c_im)
    for (int h_col = 0; h_col < Ho; ++h_col)
      for (int w_col = 0; w_col < Wo; ++w_col) {
        int g = c_im / cpg;
        int h_in = h_col * sh - pad_h, w_in = w_col * sw - pad_w;
        const float* im_ptr = im + ((size_t)c_im * H + h_in) * W + w_in;   /* may point before the row: only dereferenced in-bounds */
        const float* off_ptr = off + (size_t)g * 2 * kh * kw * Ho * Wo;
        for (int i = 0; i < kh; ++i)
          for (int j = 0; j < kw; ++j) {
            float oh = off_ptr[((size_t)(2 * (i * kw + j)) * Ho + h_col) * Wo + w_col];
            float ow = off_ptr[((size_t)(2 * (i * kw + j) + 1) * Ho + h_col) * Wo + w_col];
            float val = 0.f;
            float h_im = h_in + i * dil_h + oh, w_im = w_in + j * dil_w + ow;
            if (h_im >= 0 && w_im >= 0 && h_im < H && w_im < W) {
              float map_h = i * dil_h + oh, map_w = j * dil_w + ow;
              val = dim2col_bilinear(im_ptr, W, H - h_in, W - w_in, map_h, map_w);
            }
            col[(((size_t)c_im * kh * kw + i * kw + j) * Ho + h_col) * Wo + w_col] = val;
          }
      }
}

/* ---------------------------------------------------------------------------------------------------------
 * GPU NMS semantics: lib/nms/nms_kernel.cu:24-32 (devIoU, float32), :61-77 (IoU > thresh), :124-139 (sweep).
 * boxes [n,5] already sorted by score; returns number kept, indices in keep. */
static float dev_iou(const float* a, const float* b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1, 0.f), height = fmaxf(bottom - top + 1, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
  float Sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  return interS / (Sa + Sb - interS);
}

int oracle_nms_sorted(const float* boxes, int n, int box_dim, float thresh, int* keep) {
  unsigned char* removed = (unsigned char*)__builtin_alloca((size_t)n);
  memset(removed, 0, (size_t)n);
  int nk = 0;
  for (int i = 0; i < n; ++i) {
    if (removed[i]) continue;
    keep[nk++] = i;
    for (int j = i + 1; j < n; ++j)
      if (!removed[j] && dev_iou(boxes + (size_t)i * box_dim, boxes + (size_t)j * box_dim) > thresh) removed[j] = 1;
  }
  return nk;
}

/* =========================================================================================================
 * Backward kernels (training).  Sequential double-free float restatements; the CUDA kernels accumulate with atomics,
 * so GPU parity is to a float tolerance (summation order), not bit-exact.
 * ========================================================================================================= */

/* ROIPooling backward (MXNet 1.1.0 roi_pooling.cu ROIPoolBackwardAcc, not in tree): every pooled cell routes its
 * gradient to the argmax element recorded by the forward; ddata is OVERWRITTEN (zeroed first). */
void oracle_roi_pool_bwd(const float* dout, const int* argmax, const float* rois, int R, int B, int C, int H, int W,
                         int PH, int PW, float* ddata) {
  memset(ddata, 0, sizeof(float) * (size_t)B * C * H * W);
  for (int n = 0; n < R; ++n) {
    int b = (int)rois[5 * n];
    for (int c = 0; c < C; ++c)
      for (int p = 0; p < PH * PW; ++p) {
        size_t o = ((size_t)n * C + c) * PH * PW + p;
        if (argmax[o] >= 0) ddata[((size_t)b * C + c) * H * W + argmax[o]] += dout[o];
      }
  }
}

/* DeformablePSROIPoolBackwardAccKernel, relation_rcnn/operator_cxx/deformable_psroi_pooling.cu:177-289.
 * ddata [B,channels,H,W] and dtrans [R,2*num_classes,part,part] are zeroed then accumulated (req = write). */
void oracle_deform_psroi_pool_bwd(const float* dout, const float* top_count, const float* data, const float* rois,
                                  const float* trans, int R, int B, int channels, int H, int W, int no_trans,
                                  float spatial_scale, int output_dim, int group_size, int pooled, int part_size,
                                  int sample_per_part, float trans_std, int num_classes, float* ddata, float* dtrans) {
  int channels_each_class = no_trans ? output_dim : output_dim / num_classes;
  if (no_trans) num_classes = 1;
  memset(ddata, 0, sizeof(float) * (size_t)B * channels * H * W);
  if (!no_trans) memset(dtrans, 0, sizeof(float) * (size_t)R * 2 * num_classes * part_size * part_size);
  size_t count = (size_t)R * output_dim * pooled * pooled;
  for (size_t index = 0; index < count; ++index) {
    int pw = index % pooled, ph = (index / pooled) % pooled;
    int ctop = (index / pooled / pooled) % output_dim;
    int n = index / pooled / pooled / output_dim;
    const float* roi = rois + 5 * n;
    int b = (int)roi[0];
    float rsw = (float)((double)(roundf(roi[1]) * spatial_scale) - 0.5);
    float rsh = (float)((double)(roundf(roi[2]) * spatial_scale) - 0.5);
    float rew = (float)((double)((float)((double)roundf(roi[3]) + 1.) * spatial_scale) - 0.5);
    float reh = (float)((double)((float)((double)roundf(roi[4]) + 1.) * spatial_scale) - 0.5);
    float roi_w = (float)fmax((double)(rew - rsw), 0.1);
    float roi_h = (float)fmax((double)(reh - rsh), 0.1);
    float bin_h = roi_h / (float)pooled, bin_w = roi_w / (float)pooled;
    float sub_h = bin_h / (float)sample_per_part, sub_w = bin_w / (float)sample_per_part;
    int part_h = (int)floorf((float)ph / pooled * part_size);
    int part_w = (int)floorf((float)pw / pooled * part_size);
    int class_id = ctop / channels_each_class;
    size_t tix = (((size_t)(n * num_classes + class_id) * 2) * part_size + part_h) * part_size + part_w;
    size_t tiy = (((size_t)(n * num_classes + class_id) * 2 + 1) * part_size + part_h) * part_size + part_w;
    float tx = no_trans ? 0.f : trans[tix] * trans_std;
    float ty = no_trans ? 0.f : trans[tiy] * trans_std;
    float wstart = (float)pw * bin_w + rsw; wstart += tx * roi_w;
    float hstart = (float)ph * bin_h + rsh; hstart += ty * roi_h;
    if (top_count[index] <= 0) continue;
    float diff_val = dout[index] / top_count[index];
    int gw = (int)floorf((float)pw * group_size / pooled), gh = (int)floorf((float)ph * group_size / pooled);
    if (gw < 0) gw = 0; if (gw > group_size - 1) gw = group_size - 1;
    if (gh < 0) gh = 0; if (gh > group_size - 1) gh = group_size - 1;
    const float* d0 = data + (size_t)b * channels * H * W;
    float* g0 = ddata + (size_t)b * channels * H * W;
    for (int ih = 0; ih < sample_per_part; ++ih)
      for (int iw = 0; iw < sample_per_part; ++iw) {
        float w = wstart + iw * sub_w, h = hstart + ih * sub_h;
        if ((double)w < -0.5 || (double)w > W - 0.5 || (double)h < -0.5 || (double)h > H - 0.5) continue;
        w = (float)fmin(fmax((double)w, 0.), W - 1.);
        h = (float)fmin(fmax((double)h, 0.), H - 1.);
        int c = (ctop * group_size + gh) * group_size + gw;
        int x0 = (int)floorf(w), x1 = (int)ceilf(w), y0 = (int)floorf(h), y1 = (int)ceilf(h);
        float dx = w - x0, dy = h - y0;
        float q00 = (1 - dx) * (1 - dy), q01 = (1 - dx) * dy, q10 = dx * (1 - dy), q11 = dx * dy;
        size_t base = (size_t)c * H * W;
        g0[base + y0 * W + x0] += q00 * diff_val;
        g0[base + y1 * W + x0] += q01 * diff_val;
        g0[base + y0 * W + x1] += q10 * diff_val;
        g0[base + y1 * W + x1] += q11 * diff_val;
        if (no_trans) continue;
        float U00 = d0[base + y0 * W + x0], U01 = d0[base + y1 * W + x0];
        float U10 = d0[base + y0 * W + x1], U11 = d0[base + y1 * W + x1];
        float diff_x = (U11 * dy + U10 * (1 - dy) - U01 * dy - U00 * (1 - dy)) * trans_std * diff_val;
        diff_x *= roi_w;
        float diff_y = (U11 * dx + U01 * (1 - dx) - U10 * dx - U00 * (1 - dx)) * trans_std * diff_val;
        diff_y *= roi_h;
        dtrans[tix] += diff_x;
        dtrans[tiy] += diff_y;
      }
  }
}

/* get_gradient_weight, nn/deformable_im2col.cuh:116-158 */
static float dcol_gradient_weight(float ah, float aw, int h, int w, int height, int width) {
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

/* get_coordinate_weight, nn/deformable_im2col.cuh:161-207 */
static float dcol_coordinate_weight(float ah, float aw, int height, int width, const float* im, int data_width,
                                    int bp_dir) {
  if (ah < 0 || ah > height || aw < 0 || aw > width) return 0.f;
  if (ah < 0) ah = 0;
  if (aw < 0) aw = 0;
  int hl = (int)ah, wl = (int)aw, hh, wh;
  if (hl >= height - 1) { hh = hl = height - 1; ah = (float)hl; } else hh = hl + 1;
  if (wl >= width - 1) { wh = wl = width - 1; aw = (float)wl; } else wh = wl + 1;
  float weight = 0.f;
  if (bp_dir == 0) {
    weight += -1 * (wl + 1 - aw) * im[hl * data_width + wl];
    weight += -1 * (aw - wl) * im[hl * data_width + wh];
    weight += (wl + 1 - aw) * im[hh * data_width + wl];
    weight += (aw - wl) * im[hh * data_width + wh];
  } else {
    weight += -1 * (hl + 1 - ah) * im[hl * data_width + wl];
    weight += (hl + 1 - ah) * im[hl * data_width + wh];
    weight += -1 * (ah - hl) * im[hh * data_width + wl];
    weight += (ah - hl) * im[hh * data_width + wh];
  }
  return weight;
}

/* deformable_col2im_gpu_kernel, nn/deformable_im2col.cuh:315-360: col [C*kh*kw, Ho, Wo] -> grad_im [C,H,W] (+=) */
void oracle_deform_col2im(const float* col, const float* off, int C, int H, int W, int kh, int kw, int pad_h, int pad_w,
                          int sh, int sw, int dil_h, int dil_w, int dg, int Ho, int Wo, float* grad_im) {
  int cpg = C / dg;
  size_t n = (size_t)C * kh * kw * Ho * Wo;
  for (size_t index = 0; index < n; ++index) {
    int j = (index / Wo / Ho) % kw, i = (index / Wo / Ho / kw) % kh;
    int c = index / Wo / Ho / kw / kh;
    int g = c / cpg;
    int w_out = index % Wo, h_out = (index / Wo) % Ho;
    int w_in = w_out * sw - pad_w, h_in = h_out * sh - pad_h;
    const float* off_ptr = off + (size_t)g * 2 * kh * kw * Ho * Wo;
    float oh = off_ptr[((size_t)(2 * (i * kw + j)) * Ho + h_out) * Wo + w_out];
    float ow = off_ptr[((size_t)(2 * (i * kw + j) + 1) * Ho + h_out) * Wo + w_out];
    float ih = h_in + i * dil_h + oh, iw = w_in + j * dil_w + ow;
    float top = col[index];
    int ch = (int)ih, cw = (int)iw;
    for (int dy = -2; dy <= 2; ++dy)
      for (int dx = -2; dx <= 2; ++dx)
        if (ch + dy >= 0 && ch + dy < H && cw + dx >= 0 && cw + dx < W && fabsf(ih - (ch + dy)) < 1 &&
            fabsf(iw - (cw + dx)) < 1) {
          float wgt = dcol_gradient_weight(ih, iw, ch + dy, cw + dx, H, W);
          grad_im[((size_t)c * H + ch + dy) * W + cw + dx] += wgt * top;
        }
  }
}

/* deformable_col2im_coord_gpu_kernel, nn/deformable_im2col.cuh:407-458: -> grad_offset [dg*2*kh*kw, Ho, Wo] (=) */
void oracle_deform_col2im_coord(const float* col, const float* im, const float* off, int C, int H, int W, int kh,
                                int kw, int pad_h, int pad_w, int sh, int sw, int dil_h, int dil_w, int dg, int Ho,
                                int Wo, float* grad_off) {
  int cpg_col = C * kh * kw / dg;          /* channel_per_deformable_group of the COLUMN buffer */
  size_t n = (size_t)Ho * Wo * 2 * kh * kw * dg;
  for (size_t index = 0; index < n; ++index) {
    float val = 0.f;
    int w = index % Wo, h = (index / Wo) % Ho;
    int c = index / Wo / Ho;
    int g = c / (2 * kh * kw);
    int col_step = kh * kw, cnt = 0;
    const float* col_ptr = col + (size_t)g * cpg_col * Wo * Ho;
    const float* im_ptr = im + (size_t)g * cpg_col / kh / kw * H * W;
    const float* off_ptr = off + (size_t)g * 2 * kh * kw * Ho * Wo;
    int offset_c = c - g * 2 * kh * kw;
    for (int col_c = offset_c / 2; col_c < cpg_col; col_c += col_step) {
      size_t col_pos = ((size_t)col_c * Ho + h) * Wo + w;
      int bp_dir = offset_c % 2;
      int j = (col_pos / Wo / Ho) % kw, i = (col_pos / Wo / Ho / kw) % kh;
      int w_out = col_pos % Wo, h_out = (col_pos / Wo) % Ho;
      int w_in = w_out * sw - pad_w, h_in = h_out * sh - pad_h;
      float oh = off_ptr[((size_t)(2 * (i * kw + j)) * Ho + h_out) * Wo + w_out];
      float ow = off_ptr[((size_t)(2 * (i * kw + j) + 1) * Ho + h_out) * Wo + w_out];
      float inv_h = h_in + i * dil_h + oh, inv_w = w_in + j * dil_w + ow;
      if (inv_h < 0 || inv_w < 0 || inv_h >= H || inv_w >= W) inv_h = inv_w = -1;
      float wgt = dcol_coordinate_weight(inv_h, inv_w, H, W, im_ptr + (size_t)cnt * H * W, W, bp_dir);
      val += wgt * col_ptr[col_pos];
      cnt += 1;
    }
    grad_off[index] = val;
  }
}

/* plain im2col (MXNet src/operator/nn/im2col.h, not in tree; used for dWeight at deformable_convolution-inl.h:215) */
void oracle_im2col(const float* im, int C, int H, int W, int kh, int kw, int pad_h, int pad_w, int sh, int sw,
                   int dil_h, int dil_w, int Ho, int Wo, float* col) {
  for (int c = 0; c < C; ++c)
    for (int i = 0; i < kh; ++i)
      for (int j = 0; j < kw; ++j)
        for (int ho = 0; ho < Ho; ++ho)
          for (int wo = 0; wo < Wo; ++wo) {
            int h = ho * sh - pad_h + i * dil_h, w = wo * sw - pad_w + j * dil_w;
            col[(((size_t)c * kh * kw + i * kw + j) * Ho + ho) * Wo + wo] =
                (h >= 0 && h < H && w >= 0 && w < W) ? im[((size_t)c * H + h) * W + w] : 0.f;
          }
}
