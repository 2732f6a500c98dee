// learn_nms.cu -- learned-NMS duplicate-removal head, device resident (operator_py/learn_nms.py:238-401 = LNMS).
// The reference op syncs with the host twice (LNMS:296, :382) and builds its index arrays in numpy; here every step
// stays on the stream:
//   prep     softmax over classes (LNMS:289) + refine_bbox_nd (LNMS:175-217)               1 thread / roi
//   sort     per-class descending sort of the fg probabilities (LNMS:291, :306-308)        1 CTA / class (bitonic)
//   valid    class pruning threshold min(class_thresh, global max) (LNMS:298-303)          1 warp
//   gather   sorted_bbox (LNMS:311-323) and per-class features emb[idx] + rank_feat (LNMS:339-345)
//   relation per-class object relation (LNMS:347-352 = nms_attention_nd) via rn_relation_fwd, batch = classes
//   logits   128 -> T logits, sigmoid, times sorted score, pruned classes -> 0, merge over thresholds
//            (LNMS:357-381, SYM_REL_NMS:553-560)
// Layouts: prob [Rn,C], refined [Rn,4,K], rank_idx [n,C] int32, feat_cls [C,n,128], boxes_cls [C,n,4].
#include "common.cuh"
#include "relation.cuh"
#include "learn_nms.cuh"

namespace rn {


// one warp per roi: the class softmax is evaluated in float32 with MXNet's order (max, then a sequential sum of
// exp(x - max)); lanes compute the exps in parallel, lane 0 adds them in class order so the sum matches bit for bit
__global__ void __launch_bounds__(128) lnms_prep_kernel(rn_learn_nms_desc d, int Rn, const int* __restrict__ sel,
                                                        const float* __restrict__ cls_score,
                                                        const float* __restrict__ bbox_pred, const float* __restrict__ rois,
                                                        const float* __restrict__ im_info, float* __restrict__ prob,
                                                        float* __restrict__ refined) {
  extern __shared__ float ex[];                          // [4 warps][num_classes]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + warp;
  if (r >= Rn) return;
  const int src = sel ? sel[r] : r;
  const int NC = d.num_classes, C = NC - 1;
  const float* s = cls_score + (size_t)src * NC;
  float* e = ex + warp * NC;
  float mx = -INFINITY;
  for (int c = lane; c < NC; c += 32) mx = fmaxf(mx, s[c]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  for (int c = lane; c < NC; c += 32) e[c] = expf(s[c] - mx);
  __syncwarp();
  float sum = 0.f;
  if (lane == 0) for (int c = 0; c < NC; ++c) sum += e[c];
  sum = __shfl_sync(0xffffffffu, sum, 0);
  for (int c = 1 + lane; c < NC; c += 32) prob[(size_t)r * C + (c - 1)] = e[c] / sum;
  // refine_bbox_nd, LNMS:175-217
  const float* b = rois + (size_t)src * 5 + 1;
  const float w = b[2] - b[0] + 1.f, h = b[3] - b[1] + 1.f;
  const float cx = 0.5f * (b[0] + b[2]), cy = 0.5f * (b[1] + b[3]);
  const int K = d.num_reg_classes - 1;
  const float lim_w = im_info[1] - 1.f, lim_h = im_info[0] - 1.f;
  for (int k = lane; k < K; k += 32) {
    const float* dl = bbox_pred + (size_t)src * 4 * d.num_reg_classes + 4 + 4 * k;
    float dx = dl[0], dy = dl[1], dw = dl[2], dh = dl[3];
    if (d.has_means_stds) {
      dx = dx * d.stds[0] + d.means[0]; dy = dy * d.stds[1] + d.means[1];
      dw = dw * d.stds[2] + d.means[2]; dh = dh * d.stds[3] + d.means[3];
    }
    const float rcx = cx + w * dx, rcy = cy + h * dy;
    const float rw = w * expf(dw), rh = h * expf(dh);
    const float wo = 0.5f * (rw - 1.f), ho = 0.5f * (rh - 1.f);
    float o[4] = {rcx - wo, rcy - ho, rcx + wo, rcy + ho};
    o[0] = fmaxf(fminf(o[0], lim_w), 0.f); o[1] = fmaxf(fminf(o[1], lim_h), 0.f);
    o[2] = fmaxf(fminf(o[2], lim_w), 0.f); o[3] = fmaxf(fminf(o[3], lim_h), 0.f);
    for (int j = 0; j < 4; ++j) refined[((size_t)r * 4 + j) * K + k] = o[j];
  }
}

// one CTA per fg class: descending bitonic sort of prob[:, c] with index, ties -> lower roi index first
__global__ void __launch_bounds__(512) lnms_sort_kernel(const float* __restrict__ prob, int Rn, int C, int n,
                                                        float* __restrict__ sorted_score, int* __restrict__ rank_idx,
                                                        float* __restrict__ cmax) {
  extern __shared__ uint8_t sm[];
  int P = 1;
  while (P < Rn) P <<= 1;
  float* k = reinterpret_cast<float*>(sm);
  int* ix = reinterpret_cast<int*>(sm + (size_t)P * 4);
  const int c = blockIdx.x;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    k[i] = i < Rn ? prob[(size_t)i * C + c] : -INFINITY;
    ix[i] = i < Rn ? i : 0x7fffffff;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const float ka = k[lo], kb = k[hi];
        const int ia = ix[lo], ib = ix[hi];
        const bool a_before_b = (ka > kb) || (ka == kb && ia < ib);
        if (a_before_b != desc) { k[lo] = kb; k[hi] = ka; ix[lo] = ib; ix[hi] = ia; }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    sorted_score[(size_t)i * C + c] = k[i];
    rank_idx[(size_t)i * C + c] = ix[i];
  }
  if (threadIdx.x == 0) cmax[c] = k[0];
}

__global__ void lnms_valid_kernel(const float* __restrict__ cmax, int C, double class_thresh, int* __restrict__ valid) {
  float g = -INFINITY;
  for (int c = threadIdx.x; c < C; c += 32) g = fmaxf(g, cmax[c]);
  for (int o = 16; o; o >>= 1) g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, o));
  const double th = fmin(class_thresh, (double)g);       // np.minimum(python float, float32 max) -> float64
  for (int c = threadIdx.x; c < C; c += 32) valid[c] = ((double)cmax[c] >= th) ? 1 : 0;
}

// extract_rank_embedding_nd, LNMS:129-140: [n, 1024] = [sin(r / 1000^(2k/1024)), cos(...)], k < 512
__global__ void lnms_rank_embed_kernel(int n, float* __restrict__ emb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = kRankDim / 2;
  if (i >= n * half) return;
  const int r = i / half, k = i % half;
  const float dim = powf(1000.0f, (2.0f / (float)kRankDim) * (float)k);
  float s, c;
  sincosf((float)r / dim, &s, &c);
  emb[(size_t)r * kRankDim + k] = s;
  emb[(size_t)r * kRankDim + half + k] = c;
}

// sorted_bbox [n,C,4]; feat_cls [C,n,128] = emb[rank_idx] + rank_feat; boxes_cls [C,n,4]
__global__ void lnms_gather_kernel(int n, int C, int K, int class_agnostic, const int* __restrict__ rank_idx,
                                   const float* __restrict__ refined, const float* __restrict__ emb,
                                   const float* __restrict__ rank_feat, float* __restrict__ sorted_bbox,
                                   float* __restrict__ feat_cls, float* __restrict__ boxes_cls) {
  const int i = blockIdx.x, c = blockIdx.y;
  const int src = rank_idx[(size_t)i * C + c];
  if (threadIdx.x < 4) {
    const int k = class_agnostic ? 0 : c;
    const float v = refined[((size_t)src * 4 + threadIdx.x) * K + k];
    sorted_bbox[((size_t)i * C + c) * 4 + threadIdx.x] = v;
    boxes_cls[((size_t)c * n + i) * 4 + threadIdx.x] = v;
  }
  for (int j = threadIdx.x; j < kNmsFeat; j += blockDim.x)
    feat_cls[((size_t)c * n + i) * kNmsFeat + j] = emb[(size_t)src * kNmsFeat + j] + rank_feat[(size_t)i * kNmsFeat + j];
}

// one warp per (i, c): logits, sigmoid, score product, class pruning, merge
__global__ void __launch_bounds__(128) lnms_logit_kernel(int n, int C, int T, const float* __restrict__ feat_out,
                                                         const float* __restrict__ Wl, const float* __restrict__ bl,
                                                         const float* __restrict__ sorted_score,
                                                         const int* __restrict__ valid, int merge_method,
                                                         float* __restrict__ multi, float* __restrict__ final_score) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n * C) return;
  const int i = warp / C, c = warp % C;
  const float* f = feat_out + ((size_t)c * n + i) * kNmsFeat;
  float x[kNmsFeat / 32];
#pragma unroll
  for (int j = 0; j < kNmsFeat / 32; ++j) x[j] = f[lane + 32 * j];
  const float sc = sorted_score[(size_t)i * C + c];
  const bool ok = valid[c] != 0;
  float acc_mean = 0.f, acc_max = -INFINITY, pick = 0.f;
  for (int t = 0; t < T; ++t) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kNmsFeat / 32; ++j) s = fmaf(x[j], Wl[(size_t)t * kNmsFeat + lane + 32 * j], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s += bl[t];
    const float cond = ok ? 1.f / (1.f + expf(-s)) : 0.f;
    const float m = sc * cond;
    if (lane == 0) multi[((size_t)i * C + c) * T + t] = m;
    acc_mean += m; acc_max = fmaxf(acc_max, m);
    if (t == merge_method) pick = m;
  }
  if (lane == 0 && final_score)
    final_score[(size_t)i * C + c] = merge_method == -1 ? acc_mean / (float)T : merge_method == -2 ? acc_max : pick;
}

// nms_multi_target (operator_py/nms_multi_target.py:24-74): learn-NMS training labels.  One CTA per fg class.
// qual(i,g,t) = (first-argmax_g IoU(i,.) == g) && IoU(i,g) > th[t]; winner(g,t) = first argmax_i (qual ? score : 0)
// (box 0 when nobody qualifies); out[winner] = 1 iff max_g IoU(winner,.) > th[t].  IoU in float64 (bbox.pyx:15-55).
__device__ __forceinline__ double lnms_iou_f64(const double* b, const double* q) {
  const double iw = fmin(b[2], q[2]) - fmax(b[0], q[0]) + 1;
  if (iw <= 0) return 0.0;
  const double ih = fmin(b[3], q[3]) - fmax(b[1], q[1]) + 1;
  if (ih <= 0) return 0.0;
  const double qa = (q[2] - q[0] + 1) * (q[3] - q[1] + 1);
  const double ua = (b[2] - b[0] + 1) * (b[3] - b[1] + 1) + qa - iw * ih;
  return iw * ih / ua;
}

struct ThreshSet { double v[16]; int T; };

__global__ void __launch_bounds__(128) nms_multi_target_kernel(const float* __restrict__ bbox, const float* __restrict__ gt,
                                                               const float* __restrict__ score, int n, int C, int G,
                                                               ThreshSet th, float* __restrict__ out) {
  extern __shared__ double sm_d[];                 // ov_best[n] | then int best_gt[n] | float sc[n] | uchar flag[n*T]
  double* ov_best = sm_d;
  int* best_gt = reinterpret_cast<int*>(ov_best + n);
  float* sc = reinterpret_cast<float*>(best_gt + n);
  unsigned char* flag = reinterpret_cast<unsigned char*>(sc + n);
  __shared__ int cls_gt[256];
  __shared__ int n_cls_gt;
  const int c = blockIdx.x, T = th.T;
  if (threadIdx.x == 0) {
    int k = 0;
    for (int g = 0; g < G && k < 256; ++g)
      if ((int)gt[(size_t)g * 5 + 4] == c + 1) cls_gt[k++] = g;
    n_cls_gt = k;
  }
  for (int i = threadIdx.x; i < n * T; i += blockDim.x) flag[i] = 0;
  __syncthreads();
  const int Gc = n_cls_gt;
  if (Gc > 0) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float* b = bbox + ((size_t)i * C + c) * 4;
      const double bd[4] = {(double)b[0], (double)b[1], (double)b[2], (double)b[3]};
      double best = -1.0; int arg = 0;
      for (int k = 0; k < Gc; ++k) {
        const float* q = gt + (size_t)cls_gt[k] * 5;
        const double qd[4] = {(double)q[0], (double)q[1], (double)q[2], (double)q[3]};
        const double ov = lnms_iou_f64(bd, qd);
        if (ov > best) { best = ov; arg = k; }
      }
      ov_best[i] = best; best_gt[i] = arg; sc[i] = score[(size_t)i * C + c];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int pair = warp; pair < Gc * T; pair += 4) {
      const int k = pair / T, t = pair % T;
      float bv = -1.f; int bi = 0x7fffffff;
      for (int i = lane; i < n; i += 32) {
        const float v = (best_gt[i] == k && ov_best[i] > th.v[t]) ? sc[i] : 0.f;
        if (v > bv) { bv = v; bi = i; }                 // strict '>' keeps the first maximum of this lane's stride
      }
      for (int o = 16; o; o >>= 1) {
        const float ov_ = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov_ > bv || (ov_ == bv && oi < bi)) { bv = ov_; bi = oi; }
      }
      if (lane == 0 && bi < n && ov_best[bi] > th.v[t]) flag[bi * T + t] = 1;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n * T; i += blockDim.x) out[((size_t)(i / T) * C + c) * T + (i % T)] = flag[i] ? 1.f : 0.f;
}

rn_relation_desc lnms_inner_desc(const rn_learn_nms_desc* d) {
  rn_relation_desc r;
  r.batch = d->num_classes - 1; r.N = d->first_n; r.M = d->first_n; r.d = kNmsFeat; r.dq = 1024; r.dout = kNmsFeat;
  r.H = 16; r.E = 64; r.wave_length = 1000.f; r.fuse_residual_relu = 1; r.precision = d->precision;
  return r;
}

size_t lnms_carve(const rn_learn_nms_desc* d, int Rn, void* base, size_t bytes, LnmsWs* w) {
  const size_t C = d->num_classes - 1, n = d->first_n, K = d->num_reg_classes - 1;
  rn_relation_desc rd = lnms_inner_desc(d);
  const size_t rel = rn_relation_workspace_bytes(&rd) + relation_tc_lnms_extra_bytes(&rd, d->R);
  size_t need = ws_slice((size_t)Rn * C, 4) + ws_slice((size_t)Rn * 4 * K, 4) + ws_slice(C, 4) +
                ws_slice(n * kRankDim, 4) + ws_slice(n * kNmsFeat, 4) + ws_slice((size_t)d->R * kNmsFeat, 4) +
                ws_slice(C * n * kNmsFeat, 4) + ws_slice(C * n * 4, 4) + ws_slice(C * n * kNmsFeat, 4) +
                ws_slice(n * C, 4) + ws_slice(C, 4) + ws_slice((size_t)16 * Rn * align_up(Rn, 4), 4) + align_up(rel, 256);
  if (!w) return need;
  Workspace ws(base, bytes);
  w->prob = ws.take<float>((size_t)Rn * C);
  w->refined = ws.take<float>((size_t)Rn * 4 * K);
  w->cmax = ws.take<float>(C);
  w->rank_emb = ws.take<float>(n * kRankDim);
  w->rank_feat = ws.take<float>(n * kNmsFeat);
  w->emb = ws.take<float>((size_t)d->R * kNmsFeat);
  w->feat_cls = ws.take<float>(C * n * kNmsFeat);
  w->boxes_cls = ws.take<float>(C * n * 4);
  w->feat_out = ws.take<float>(C * n * kNmsFeat);
  w->rank_idx = ws.take<int>(n * C);
  w->valid = ws.take<int>(C);
  w->lg_roi = ws.take<float>((size_t)16 * Rn * align_up(Rn, 4));
  w->rel_ws = ws.take<char>(rel);
  w->rel_ws_bytes = rel;
  return w->rel_ws ? need : 0;
}

int lnms_selected_rows(const rn_learn_nms_desc* d, const int* non_gt_index) {
  if (d->nongt_dim > 0) return d->nongt_dim;
  if (non_gt_index) return d->num_non_gt;
  return d->R;
}

}  // namespace rn

extern "C" size_t rn_learn_nms_workspace_bytes(const rn_learn_nms_desc* d) {
  if (!d) return 0;
  return rn::lnms_carve(d, d->R, nullptr, 0, nullptr) + 256;
}

namespace rn {

// prepared (weight-only) state of the F16 class-agnostic path -- rn_learn_nms_pack
struct LnmsPrepared {
  const float* rank_feat;     // [n, 128] fp32
  const void* w_emb16;        // roi_feat_embedding weight, fp16 [128, ceil8(feat_dim)]
  const void* rel;            // relation_tc_lnms prepared blob (packed fp16 Q/K/V' weights + RQKV)
};

static bool lnms_fast_path(const rn_learn_nms_desc* d) {
  rn_relation_desc rd = lnms_inner_desc(d);
  return d->precision == RN_PREC_F16 && d->class_agnostic && d->first_n <= 512 && relation_tc_shape_ok(&rd) && is_sm100();
}

static size_t lnms_prepared_layout(const rn_learn_nms_desc* d, size_t* o_rank, size_t* o_emb, size_t* o_rel) {
  rn_relation_desc rd = lnms_inner_desc(d);
  size_t off = 0;
  *o_rank = off; off += ws_slice((size_t)d->first_n * kNmsFeat, 4);
  *o_emb = off; off += align_up(linear_tc_packed_bytes(d->feat_dim, kNmsFeat), 256);
  *o_rel = off; off += align_up(relation_tc_lnms_prepared_bytes(&rd), 256);
  return off;
}

static int lnms_forward(const rn_learn_nms_desc* d, const float* cls_score, const float* bbox_pred, const float* rois,
                        const float* im_info, const float* feat, const rn_learn_nms_weights* w, const LnmsPrepared* prep,
                        const void* feat_f16, const float* emb_ext, const int32_t* non_gt_index, float* nms_multi_score, float* sorted_bbox, float* sorted_score,
                        float* final_score, void* wsp, size_t ws_bytes, rn_stream_t stream) {
  RN_CHECK_ARG(d && cls_score && bbox_pred && rois && im_info && feat && w && nms_multi_score && sorted_bbox &&
                   sorted_score && wsp, "rn_learn_nms_fwd: null argument");
  const int C = d->num_classes - 1, n = d->first_n, T = d->num_thresh, K = d->num_reg_classes - 1;
  const int Rn = lnms_selected_rows(d, non_gt_index);
  RN_CHECK_ARG(C >= 1 && n >= 1 && T >= 1 && K >= 1, "rn_learn_nms_fwd: bad sizes C=%d n=%d T=%d K=%d", C, n, T, K);
  RN_CHECK_ARG(Rn >= n && Rn <= d->R, "rn_learn_nms_fwd: need first_n=%d <= non-gt rois=%d <= R=%d", n, Rn, d->R);
  RN_CHECK_ARG(Rn <= 8192, "rn_learn_nms_fwd: %d rois exceed the per-class sort capacity 8192", Rn);
  RN_CHECK_ARG(d->class_agnostic || K == C, "rn_learn_nms_fwd: class-specific boxes need num_reg_classes == num_classes");
  // class-agnostic boxes: one foreground regressor (the reference's Reshape(0,0,0) at LNMS:331-333 only works for K == 1)
  RN_CHECK_ARG(!d->class_agnostic || K == 1, "rn_learn_nms_fwd: class_agnostic needs num_reg_classes == 2 (got %d)", d->num_reg_classes);
  RN_CHECK_ARG(d->merge_method >= -2 && d->merge_method < T, "rn_learn_nms_fwd: unknown merge method %d", d->merge_method);
  cudaStream_t st = (cudaStream_t)stream;
  LnmsWs W;
  if (!lnms_carve(d, Rn, wsp, ws_bytes, &W)) { set_error("rn_learn_nms_fwd: workspace too small (%zu < %zu)", ws_bytes, rn_learn_nms_workspace_bytes(d)); return RN_ERR_WORKSPACE; }
  const int* sel = d->nongt_dim > 0 ? nullptr : non_gt_index;
  int r;
  lnms_prep_kernel<<<cdiv(Rn, 4), 128, (size_t)4 * d->num_classes * sizeof(float), st>>>(*d, Rn, sel, cls_score, bbox_pred, rois, im_info, W.prob, W.refined);
  RN_LAUNCH_CHECK();
  int P = 1; while (P < Rn) P <<= 1;
  static thread_local bool sort_cfg = false;
  if (!sort_cfg) {
    RN_CUDA(cudaFuncSetAttribute(lnms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
    sort_cfg = true;
  }
  lnms_sort_kernel<<<C, 512, (size_t)P * 8, st>>>(W.prob, Rn, C, n, sorted_score, W.rank_idx, W.cmax);
  RN_LAUNCH_CHECK();
  lnms_valid_kernel<<<1, 32, 0, st>>>(W.cmax, C, d->class_thresh, W.valid);
  RN_LAUNCH_CHECK();
  // rank feature (depends on weights only: taken from the prepared blob when there is one) and roi feature embedding
  const float* rank_feat = W.rank_feat;
  const float* emb = emb_ext ? emb_ext : W.emb;      // emb_ext: roi_feat_embedding already evaluated by the caller's merged GEMM
  if (prep) {
    rank_feat = prep->rank_feat;
    if (emb_ext) {
    } else if (feat_f16) {     // the producer's fp16 copy of feat: no cast launch
      if ((r = linear_tc_packed_f16in(feat_f16, prep->w_emb16, w->roi_feat_embedding_bias, W.emb, nullptr, d->R, d->feat_dim,
                                      kNmsFeat, 0, W.rel_ws, W.rel_ws_bytes, st))) return r;
    } else if ((r = linear_tc_packed(feat, prep->w_emb16, w->roi_feat_embedding_bias, W.emb, d->R, d->feat_dim, kNmsFeat, 0,
                                     W.rel_ws, W.rel_ws_bytes, st))) return r;
  } else {
    lnms_rank_embed_kernel<<<cdiv(n * kRankDim / 2, 256), 256, 0, st>>>(n, W.rank_emb);
    RN_LAUNCH_CHECK();
    if ((r = rn_linear_fwd(W.rank_emb, w->nms_rank_weight, w->nms_rank_bias, W.rank_feat, n, kRankDim, kNmsFeat, 0,
                           d->precision, W.rel_ws, W.rel_ws_bytes, stream))) return r;
    if ((r = rn_linear_fwd(feat, w->roi_feat_embedding_weight, w->roi_feat_embedding_bias, W.emb, d->R, d->feat_dim,
                           kNmsFeat, 0, d->precision, W.rel_ws, W.rel_ws_bytes, stream))) return r;
  }
  lnms_gather_kernel<<<dim3(n, C), 128, 0, st>>>(n, C, K, d->class_agnostic, W.rank_idx, W.refined, emb, rank_feat,
                                                 sorted_bbox, W.feat_cls, W.boxes_cls);
  RN_LAUNCH_CHECK();
  rn_relation_desc rd = lnms_inner_desc(d);
  if (lnms_fast_path(d)) {
    // class-agnostic boxes: every class sorts the SAME refined rois, so the per-class pair geometry is a gather from
    // one [16, Rn, Rn] table (LNMS:332 builds [C, n, n, 4] position matrices -- 9x the pairs at C = 80, n = 100)
    const int ldr = (int)align_up(Rn, 4);
    if ((r = launch_geom_weight_log2_T(st, W.refined, Rn, 16, 64, 1000.f, w->nms_pair_pos_fc1_1_weight,
                                       w->nms_pair_pos_fc1_1_bias, W.lg_roi, ldr))) return r;
    GeomGather gg = {W.lg_roi, ldr, Rn, W.rank_idx, C, 1, nullptr, W.valid};     // pruned classes cost nothing
    if ((r = relation_tc_lnms(&rd, W.feat_cls, emb, d->R, rank_feat, &gg, w->nms_query_1_weight, w->nms_query_1_bias,
                              w->nms_key_1_weight, w->nms_key_1_bias, w->nms_linear_out_1_weight,
                              w->nms_linear_out_1_bias, W.feat_out, W.rel_ws, W.rel_ws_bytes, st,
                              prep ? prep->rel : nullptr))) return r;
  } else if ((r = rn_relation_fwd(&rd, W.feat_cls, W.boxes_cls, nullptr, w->nms_query_1_weight, w->nms_query_1_bias,
                           w->nms_key_1_weight, w->nms_key_1_bias, w->nms_pair_pos_fc1_1_weight,
                           w->nms_pair_pos_fc1_1_bias, w->nms_linear_out_1_weight, w->nms_linear_out_1_bias, W.feat_out,
                           nullptr, W.rel_ws, W.rel_ws_bytes, stream))) return r;
  lnms_logit_kernel<<<cdiv(n * C * 32, 128), 128, 0, st>>>(n, C, T, W.feat_out, w->nms_logit_weight, w->nms_logit_bias,
                                                           sorted_score, W.valid, d->merge_method, nms_multi_score,
                                                           final_score);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

}  // namespace rn

extern "C" int rn_learn_nms_fwd(const rn_learn_nms_desc* d, const float* cls_score, const float* bbox_pred,
                                const float* rois, const float* im_info, const float* feat,
                                const rn_learn_nms_weights* w, const int32_t* non_gt_index, float* nms_multi_score,
                                float* sorted_bbox, float* sorted_score, float* final_score, void* wsp, size_t ws_bytes,
                                rn_stream_t stream) {
  // RN_PREC_TF32: the general kernels of RN_PREC_FP32 with every GEMM on the library's tcgen05 tf32 engine (the forward that
  // rn_learn_nms_bwd recomputes under RN_PREC_F16); inside a backward the caller's engine stays in force
  RN_CHECK_ARG(d && (d->precision != RN_PREC_TF32 || rn::is_sm100()), "rn_learn_nms_fwd: RN_PREC_TF32 needs an sm_100 device");
  rn::GemmBackendScope backend(d->precision == RN_PREC_TF32 ? 1 : rn::gemm_backend());
  return rn::lnms_forward(d, cls_score, bbox_pred, rois, im_info, feat, w, nullptr, nullptr, nullptr, non_gt_index, nms_multi_score,
                          sorted_bbox, sorted_score, final_score, wsp, ws_bytes, stream);
}

// ---- weight-only work done once per weight update (RN_PREC_F16, class-agnostic): rank embedding -> nms_rank FC, fp16
// copies of the embedding / Q / K / V' weights, and the rank half of the factored Q/K/V' projection
extern "C" size_t rn_learn_nms_packed_bytes(const rn_learn_nms_desc* d) {
  if (!d || !rn::lnms_fast_path(d)) return 0;
  size_t a, b, c;
  return rn::lnms_prepared_layout(d, &a, &b, &c);
}

extern "C" int rn_learn_nms_pack(const rn_learn_nms_desc* d, const rn_learn_nms_weights* w, void* packed, void* wsp,
                                 size_t ws_bytes, rn_stream_t stream) {
  using namespace rn;
  RN_CHECK_ARG(d && w && packed && wsp, "rn_learn_nms_pack: null argument");
  RN_CHECK_ARG(lnms_fast_path(d), "rn_learn_nms_pack: only the RN_PREC_F16 class-agnostic path has a packed form");
  cudaStream_t st = (cudaStream_t)stream;
  size_t o_rank, o_emb, o_rel;
  lnms_prepared_layout(d, &o_rank, &o_emb, &o_rel);
  char* base = (char*)packed;
  float* rank_feat = (float*)(base + o_rank);
  const int n = d->first_n;
  Workspace ws(wsp, ws_bytes);
  float* rank_emb = ws.take<float>((size_t)n * kRankDim);
  if (!rank_emb) { set_error("rn_learn_nms_pack: workspace too small"); return RN_ERR_WORKSPACE; }
  int r;
  lnms_rank_embed_kernel<<<cdiv(n * kRankDim / 2, 256), 256, 0, st>>>(n, rank_emb);
  RN_LAUNCH_CHECK();
  void* gws = ws.base + ws.off; const size_t gws_bytes = ws.size - ws.off;
  if ((r = rn_linear_fwd(rank_emb, w->nms_rank_weight, w->nms_rank_bias, rank_feat, n, kRankDim, kNmsFeat, 0, RN_PREC_F16,
                         gws, gws_bytes, stream))) return r;
  if ((r = linear_tc_pack(w->roi_feat_embedding_weight, d->feat_dim, kNmsFeat, base + o_emb, st))) return r;
  rn_relation_desc rd = lnms_inner_desc(d);
  return relation_tc_lnms_prepare(&rd, rank_feat, w->nms_query_1_weight, w->nms_query_1_bias, w->nms_key_1_weight,
                                  w->nms_key_1_bias, w->nms_linear_out_1_weight, w->nms_linear_out_1_bias, base + o_rel,
                                  gws, gws_bytes, st);
}

extern "C" int rn_learn_nms_packed_fwd(const rn_learn_nms_desc* d, const float* cls_score, const float* bbox_pred,
                                       const float* rois, const float* im_info, const float* feat, const void* feat_f16,
                                       const float* emb, const rn_learn_nms_weights* w, const void* packed, const int32_t* non_gt_index,
                                       float* nms_multi_score, float* sorted_bbox, float* sorted_score, float* final_score,
                                       void* wsp, size_t ws_bytes, rn_stream_t stream) {
  using namespace rn;
  RN_CHECK_ARG(d && packed, "rn_learn_nms_packed_fwd: null argument");
  RN_CHECK_ARG(lnms_fast_path(d), "rn_learn_nms_packed_fwd: only the RN_PREC_F16 class-agnostic path has a packed form");
  size_t o_rank, o_emb, o_rel;
  lnms_prepared_layout(d, &o_rank, &o_emb, &o_rel);
  const char* base = (const char*)packed;
  LnmsPrepared prep = {(const float*)(base + o_rank), base + o_emb, base + o_rel};
  return lnms_forward(d, cls_score, bbox_pred, rois, im_info, feat, w, &prep, feat_f16, emb, non_gt_index, nms_multi_score, sorted_bbox,
                      sorted_score, final_score, wsp, ws_bytes, stream);
}

extern "C" int rn_nms_multi_target_fwd(const float* bbox, const float* gt_boxes, const float* score, int32_t n, int32_t C,
                                       int32_t G, const double* target_thresh_host, int32_t T, float* out,
                                       rn_stream_t stream) {
  using namespace rn;
  RN_CHECK_ARG(bbox && score && out && n >= 1 && n <= 4096 && C >= 1 && G >= 0 && T >= 1 && T <= 16 && target_thresh_host,
               "rn_nms_multi_target_fwd: bad arguments (n <= 4096, T <= 16)");
  RN_CHECK_ARG(G == 0 || gt_boxes, "rn_nms_multi_target_fwd: null gt_boxes");
  RN_CHECK_ARG(G <= 256, "rn_nms_multi_target_fwd: %d gt boxes exceed the per-class list capacity 256", G);
  ThreshSet th; th.T = T;
  for (int i = 0; i < T; ++i) th.v[i] = target_thresh_host[i];
  const size_t smem = (size_t)n * (8 + 4 + 4) + (size_t)n * T + 16;
  static thread_local size_t configured = 40 * 1024;
  if (smem > configured) {
    RN_CUDA(cudaFuncSetAttribute(nms_multi_target_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  nms_multi_target_kernel<<<C, 128, smem, (cudaStream_t)stream>>>(bbox, gt_boxes, score, n, C, G, th, out);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
