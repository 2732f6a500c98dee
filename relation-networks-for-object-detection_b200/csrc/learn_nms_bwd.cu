// learn_nms_bwd.cu -- training side of the learned-NMS head:
//   rn_learn_nms_bwd   gradients of nms_multi_score w.r.t. the 14 head weights, fc_all_2_relu (feat) and cls_score
//   rn_nms_loss        the positive / negative cross-entropy terms and d loss / d nms_multi_score
//   rn_box_annotator_ohem   'BoxAnnotatorOHEM' CustomOp forward (operator_py/box_annotator_ohem.py:26-53)
//
// The reference differentiates the train graph (resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16_
// learn_nms.py:424-501, "SYM_REL_NMS") with MXNet autograd; there is no backward source to cite.  What carries gradient:
// the sorted scores through the per-class take (the sort order does not), roi_feat_embedding / nms_rank FCs, the relation
// module, nms_logit.  What does not: bbox_pred (BlockGrad, :428), rois / im_info.
//
// Nothing is saved by the forward: rn_learn_nms_fwd is re-run in fp32 into the caller's workspace (same carve), then
//   logits:   ds[t] = dm[t] * score * cond[t] (1 - cond[t]),  d score = sum_t dm[t] cond[t]
//             dW_logit = dS^T F, db_logit = colsum dS, dF = dS W_logit            (F = relu(f + attention), [C*n,128])
//   relation: rn::relation_bwd over the C per-class problems (batch = C)          -> d f [C*n,128] + 8 weight gradients
//   gather:   d emb[r] = sum_c d f[c, pos_c(r)]   (pos_c = inverse of the per-class rank index; deterministic, no atomics)
//             d rank_feat[i] = sum_c d f[c, i]
//   FCs:      roi_feat_embedding / nms_rank weight + bias gradients, d feat = d emb . W_emb
//   scores:   d prob[r,c] = d score[pos_c(r), c], then the class-softmax backward per roi -> d cls_score
#include "common.cuh"
#include "learn_nms.cuh"
#include "relation.cuh"
#include <algorithm>

namespace rn {

// one warp per (i, c): recompute logits from F, write dS [(c*n+i), T] and d sorted_score [i, c]
__global__ void __launch_bounds__(128) lnms_logit_bwd_kernel(int n, int C, int T, const float* __restrict__ feat_out,
                                                             const float* __restrict__ Wl, const float* __restrict__ bl,
                                                             const float* __restrict__ sorted_score,
                                                             const int* __restrict__ valid,
                                                             const float* __restrict__ d_multi, float* __restrict__ dS,
                                                             float* __restrict__ d_score) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n * C) return;
  const int i = warp / C, c = warp % C;
  const float* f = feat_out + ((size_t)c * n + i) * kNmsFeat;
  float x[kNmsFeat / 32];
#pragma unroll
  for (int j = 0; j < kNmsFeat / 32; ++j) x[j] = f[lane + 32 * j];
  const float sc = sorted_score[(size_t)i * C + c];
  const bool ok = valid[c] != 0;
  float dsc = 0.f;
  for (int t = 0; t < T; ++t) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kNmsFeat / 32; ++j) s = fmaf(x[j], Wl[(size_t)t * kNmsFeat + lane + 32 * j], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    s += bl[t];
    const float cond = ok ? 1.f / (1.f + expf(-s)) : 0.f;
    const float dm = d_multi[((size_t)i * C + c) * T + t];
    dsc += dm * cond;
    if (lane == 0) dS[((size_t)c * n + i) * T + t] = ok ? dm * sc * cond * (1.f - cond) : 0.f;
  }
  if (lane == 0) d_score[(size_t)i * C + c] = dsc;
}

// pos[c][r] = rank of roi r in class c's top-n list, -1 when it is not in it
__global__ void lnms_inverse_rank_kernel(int n, int C, int Rn, const int* __restrict__ rank_idx, int* __restrict__ pos) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * C) return;
  const int i = t / C, c = t % C;
  pos[(size_t)c * Rn + rank_idx[t]] = i;
}

// d_emb [R,128]: row r = sum over classes of d_f[c, pos[c][r]]  (rows >= Rn are zero)
__global__ void __launch_bounds__(kNmsFeat) lnms_gather_bwd_emb_kernel(int n, int C, int Rn, const int* __restrict__ pos,
                                                                       const float* __restrict__ d_f, float* __restrict__ d_emb) {
  const int r = blockIdx.x, j = threadIdx.x;
  float s = 0.f;
  if (r < Rn)
    for (int c = 0; c < C; ++c) {
      const int i = pos[(size_t)c * Rn + r];
      if (i >= 0) s += d_f[((size_t)c * n + i) * kNmsFeat + j];
    }
  d_emb[(size_t)r * kNmsFeat + j] = s;
}

// d_rank_feat [n,128]: row i = sum over classes of d_f[c, i]
__global__ void __launch_bounds__(kNmsFeat) lnms_gather_bwd_rank_kernel(int n, int C, const float* __restrict__ d_f,
                                                                        float* __restrict__ d_rank) {
  const int i = blockIdx.x, j = threadIdx.x;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += d_f[((size_t)c * n + i) * kNmsFeat + j];
  d_rank[(size_t)i * kNmsFeat + j] = s;
}

// one warp per selected roi: class softmax (as lnms_prep_kernel) and its backward with d prob gathered from d score
__global__ void __launch_bounds__(128) lnms_score_bwd_kernel(int NC, int Rn, int n, const int* __restrict__ sel,
                                                             const float* __restrict__ cls_score,
                                                             const int* __restrict__ pos, const float* __restrict__ d_score,
                                                             float* __restrict__ d_cls_score) {
  extern __shared__ float sh[];                          // [4 warps][2*NC]: p, dp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + warp;
  if (r >= Rn) return;
  const int src = sel ? sel[r] : r;
  const int C = NC - 1;
  const float* s = cls_score + (size_t)src * NC;
  float* p = sh + warp * 2 * NC;
  float* dp = p + NC;
  float mx = -INFINITY;
  for (int c = lane; c < NC; c += 32) mx = fmaxf(mx, s[c]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int c = lane; c < NC; c += 32) { const float e = expf(s[c] - mx); p[c] = e; sum += e; }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float dot = 0.f;
  for (int c = lane; c < NC; c += 32) {
    const float pc = p[c] / sum;
    float d = 0.f;
    if (c >= 1) {
      const int i = pos[(size_t)(c - 1) * Rn + r];
      if (i >= 0) d = d_score[(size_t)i * C + (c - 1)];
    }
    p[c] = pc; dp[c] = d;
    dot = fmaf(pc, d, dot);
  }
  for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int c = lane; c < NC; c += 32) d_cls_score[(size_t)src * NC + c] = p[c] * (dp[c] - dot);
}

// SYM_REL_NMS:539-547 elementwise; d_multi = pos_grad_scale * d pos/d m + d neg/d m  (MakeLoss: gradient = grad_scale)
__global__ void nms_loss_kernel(const float* __restrict__ multi, const float* __restrict__ target, size_t count, float k,
                                float eps, float pos_grad_scale, float* __restrict__ pos_loss, float* __restrict__ neg_loss,
                                float* __restrict__ d_multi) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const float m = multi[i], t = target[i];
    const float a = m + eps, b = 1.0f - m + eps;
    if (pos_loss) pos_loss[i] = -(t * logf(a)) * k;
    if (neg_loss) neg_loss[i] = -((1.0f - t) * logf(b)) * k;
    if (d_multi) d_multi[i] = pos_grad_scale * (-(t / a) * k) + ((1.0f - t) / b) * k;
  }
}

// BoxAnnotatorOHEM: per-roi loss = -log(softmax[label] + 1e-14) + sum(weights * smooth_l1(pred - target)), one warp per roi
__global__ void __launch_bounds__(128) ohem_loss_kernel(int R, int NC, int D, const float* __restrict__ cls_score,
                                                        const float* __restrict__ bbox_pred, const float* __restrict__ labels,
                                                        const float* __restrict__ bbox_targets,
                                                        const float* __restrict__ bbox_weights, float* __restrict__ loss) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  const float* s = cls_score + (size_t)warp * NC;
  float mx = -INFINITY;
  for (int c = lane; c < NC; c += 32) mx = fmaxf(mx, s[c]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  // MXNet SoftmaxActivation: exp(x - max) summed in class order (sequential, lane 0) so the float sum matches the oracle
  float sum = 0.f;
  if (lane == 0) for (int c = 0; c < NC; ++c) sum += expf(s[c] - mx);
  sum = __shfl_sync(0xffffffffu, sum, 0);
  const int lab = (int)labels[warp];
  const float pl = expf(s[lab] - mx) / sum + 1e-14f;
  float lb = 0.f;
  if (lane == 0)
    for (int j = 0; j < D; ++j) {       // mx.nd.sum(axis=1): sequential float accumulation
      const float x = bbox_pred[(size_t)warp * D + j] - bbox_targets[(size_t)warp * D + j];
      const float ax = fabsf(x);
      const float sl = ax < 1.0f ? 0.5f * x * x : ax - 0.5f;             // smooth_l1, scalar = 1
      lb += bbox_weights[(size_t)warp * D + j] * sl;
    }
  if (lane == 0) loss[warp] = -1.f * logf(pl) + lb;
}

// keep the roi_per_img rois of largest loss; np.argsort(loss)[::-1] order: ties -> LARGER index first (stable ascending
// sort reversed).  rank(r) = #{q : loss[q] > loss[r] or (loss[q] == loss[r] and q > r)}; rois with rank >= keep are dropped.
__global__ void ohem_select_kernel(int R, int D, int keep, const float* __restrict__ loss, const float* __restrict__ labels,
                                   const float* __restrict__ bbox_weights, float* __restrict__ labels_out,
                                   float* __restrict__ weights_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float l = loss[r];
  int rank = 0;
  for (int q = 0; q < R; ++q) {
    const float lq = loss[q];
    rank += (lq > l || (lq == l && q > r)) ? 1 : 0;
  }
  const bool drop = rank >= keep;
  labels_out[r] = drop ? -1.f : labels[r];
  for (int j = 0; j < D; ++j) weights_out[(size_t)r * D + j] = drop ? 0.f : bbox_weights[(size_t)r * D + j];
}

static size_t lnms_bwd_extra(const rn_learn_nms_desc* d) {
  const size_t C = d->num_classes - 1, n = d->first_n, T = d->num_thresh, R = d->R;
  rn_relation_desc rd = lnms_inner_desc(d);
  rd.precision = RN_PREC_FP32;
  size_t t = 0;
  t += ws_slice(n * C * T, 4) + ws_slice(n * C * 4, 4) + 2 * ws_slice(n * C, 4);    // multi, sorted_bbox, sorted_score, final
  t += ws_slice(C * n * T, 4);                 // dS
  t += ws_slice(n * C, 4);                     // d sorted_score
  t += 2 * ws_slice(C * n * kNmsFeat, 4);      // dF, d f
  t += ws_slice(C * R, 4);                     // pos
  t += ws_slice(R * kNmsFeat, 4);              // d emb
  t += ws_slice(n * kNmsFeat, 4);              // d rank_feat
  t += align_up(relation_bwd_ws_bytes(&rd) + 256, 256);
  return t;
}

}  // namespace rn

extern "C" size_t rn_learn_nms_bwd_workspace_bytes(const rn_learn_nms_desc* d) {
  if (!d) return 0;
  rn_learn_nms_desc f = *d;
  f.precision = RN_PREC_FP32;
  return rn_learn_nms_workspace_bytes(&f) + rn::lnms_bwd_extra(d) + 256;
}

extern "C" int rn_learn_nms_bwd(const rn_learn_nms_desc* desc, const float* cls_score, const float* bbox_pred,
                                const float* rois, const float* im_info, const float* feat, const rn_learn_nms_weights* w,
                                const int32_t* non_gt_index, const float* d_multi, const rn_learn_nms_grads* g,
                                float* d_cls_score, float* d_feat, void* wsp, size_t ws_bytes, rn_stream_t stream) {
  using namespace rn;
  RN_CHECK_ARG(desc && cls_score && bbox_pred && rois && im_info && feat && w && d_multi && g && d_cls_score && d_feat && wsp,
               "rn_learn_nms_bwd: null argument");
  RN_CHECK_ARG(g->nms_rank_weight && g->nms_rank_bias && g->roi_feat_embedding_weight && g->roi_feat_embedding_bias &&
                   g->nms_pair_pos_fc1_1_weight && g->nms_pair_pos_fc1_1_bias && g->nms_query_1_weight &&
                   g->nms_query_1_bias && g->nms_key_1_weight && g->nms_key_1_bias && g->nms_linear_out_1_weight &&
                   g->nms_linear_out_1_bias && g->nms_logit_weight && g->nms_logit_bias,
               "rn_learn_nms_bwd: null gradient pointer");
  rn_learn_nms_desc d = *desc;
  // desc->precision selects the contraction engine of the whole backward (the recomputed forward included):
  // RN_PREC_F16 = the repo's tcgen05 tf32 GEMM (gemm_tf32.cu), RN_PREC_FP32 = cuBLAS fp32 (pedantic) parity mode
  GemmBackendScope backend(desc->precision == RN_PREC_F16 && is_sm100() ? 1 : 0);
  d.precision = RN_PREC_FP32;          // kernel layout of the recomputed forward (materialised P, g): the general fp32 form
  const int C = d.num_classes - 1, n = d.first_n, T = d.num_thresh, R = d.R, NC = d.num_classes;
  const int Rn = lnms_selected_rows(&d, non_gt_index);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t fwd_bytes = rn_learn_nms_workspace_bytes(&d);
  if (ws_bytes < fwd_bytes + lnms_bwd_extra(&d)) {
    set_error("rn_learn_nms_bwd: workspace too small (%zu < %zu)", ws_bytes, rn_learn_nms_bwd_workspace_bytes(desc));
    return RN_ERR_WORKSPACE;
  }
  Workspace ws((char*)wsp + fwd_bytes, ws_bytes - fwd_bytes);
  float* multi = ws.take<float>((size_t)n * C * T);
  float* sorted_bbox = ws.take<float>((size_t)n * C * 4);
  float* sorted_score = ws.take<float>((size_t)n * C);
  float* final_score = ws.take<float>((size_t)n * C);
  float* dS = ws.take<float>((size_t)C * n * T);
  float* d_score = ws.take<float>((size_t)n * C);
  float* dF = ws.take<float>((size_t)C * n * kNmsFeat);
  float* d_f = ws.take<float>((size_t)C * n * kNmsFeat);
  int* pos = ws.take<int>((size_t)C * R);
  float* d_emb = ws.take<float>((size_t)R * kNmsFeat);
  float* d_rank = ws.take<float>((size_t)n * kNmsFeat);
  rn_relation_desc rd = lnms_inner_desc(&d);
  const size_t rel_bytes = relation_bwd_ws_bytes(&rd) + 256;
  char* rel_ws = ws.take<char>(rel_bytes);
  if (!rel_ws) { set_error("rn_learn_nms_bwd: workspace too small"); return RN_ERR_WORKSPACE; }
  int r;
  // 1. recompute the forward (fp32) -- intermediates stay in the first fwd_bytes of the workspace
  if ((r = rn_learn_nms_fwd(&d, cls_score, bbox_pred, rois, im_info, feat, w, non_gt_index, multi, sorted_bbox, sorted_score,
                            final_score, wsp, fwd_bytes, stream))) return r;
  LnmsWs W;
  if (!lnms_carve(&d, Rn, wsp, fwd_bytes, &W)) { set_error("rn_learn_nms_bwd: carve failed"); return RN_ERR_WORKSPACE; }
  // 2. logits
  lnms_logit_bwd_kernel<<<cdiv(n * C * 32, 128), 128, 0, st>>>(n, C, T, W.feat_out, w->nms_logit_weight, w->nms_logit_bias,
                                                               sorted_score, W.valid, d_multi, dS, d_score);
  RN_LAUNCH_CHECK();
  if ((r = sgemm_rm(st, true, false, T, kNmsFeat, C * n, 1.f, dS, T, W.feat_out, kNmsFeat, 0.f, g->nms_logit_weight, kNmsFeat))) return r;
  if ((r = launch_colsum(st, dS, C * n, T, g->nms_logit_bias))) return r;
  if ((r = sgemm_rm(st, false, false, C * n, kNmsFeat, T, 1.f, dS, T, w->nms_logit_weight, kNmsFeat, 0.f, dF, kNmsFeat))) return r;
  // 3. relation module over the C per-class problems; its forward intermediates are the ones step 1 left in W.rel_ws
  Fp32State fwd_state;
  if (!relation_fp32_carve(&rd, W.rel_ws, W.rel_ws_bytes, &fwd_state)) { set_error("rn_learn_nms_bwd: relation state carve failed"); return RN_ERR_WORKSPACE; }
  if ((r = relation_bwd(&rd, W.feat_cls, W.boxes_cls, nullptr, w->nms_query_1_weight, w->nms_query_1_bias,
                        w->nms_key_1_weight, w->nms_key_1_bias, w->nms_pair_pos_fc1_1_weight, w->nms_pair_pos_fc1_1_bias,
                        w->nms_linear_out_1_weight, w->nms_linear_out_1_bias, dF, d_f, g->nms_query_1_weight,
                        g->nms_query_1_bias, g->nms_key_1_weight, g->nms_key_1_bias, g->nms_pair_pos_fc1_1_weight,
                        g->nms_pair_pos_fc1_1_bias, g->nms_linear_out_1_weight, g->nms_linear_out_1_bias, rel_ws, rel_bytes, st,
                        &fwd_state, W.feat_out)))
    return r;
  // 4. gather
  RN_CUDA(cudaMemsetAsync(pos, 0xff, sizeof(int) * (size_t)C * Rn, st));
  lnms_inverse_rank_kernel<<<cdiv(n * C, 256), 256, 0, st>>>(n, C, Rn, W.rank_idx, pos);
  RN_LAUNCH_CHECK();
  lnms_gather_bwd_emb_kernel<<<R, kNmsFeat, 0, st>>>(n, C, Rn, pos, d_f, d_emb);
  RN_LAUNCH_CHECK();
  lnms_gather_bwd_rank_kernel<<<n, kNmsFeat, 0, st>>>(n, C, d_f, d_rank);
  RN_LAUNCH_CHECK();
  // 5. the two FCs
  if ((r = sgemm_rm(st, true, false, kNmsFeat, d.feat_dim, R, 1.f, d_emb, kNmsFeat, feat, d.feat_dim, 0.f,
                    g->roi_feat_embedding_weight, d.feat_dim))) return r;
  if ((r = launch_colsum(st, d_emb, R, kNmsFeat, g->roi_feat_embedding_bias))) return r;
  if ((r = sgemm_rm(st, false, false, R, d.feat_dim, kNmsFeat, 1.f, d_emb, kNmsFeat, w->roi_feat_embedding_weight,
                    d.feat_dim, 0.f, d_feat, d.feat_dim))) return r;
  if ((r = sgemm_rm(st, true, false, kNmsFeat, kRankDim, n, 1.f, d_rank, kNmsFeat, W.rank_emb, kRankDim, 0.f,
                    g->nms_rank_weight, kRankDim))) return r;
  if ((r = launch_colsum(st, d_rank, n, kNmsFeat, g->nms_rank_bias))) return r;
  // 6. class scores
  RN_CUDA(cudaMemsetAsync(d_cls_score, 0, sizeof(float) * (size_t)R * NC, st));
  const int* sel = d.nongt_dim > 0 ? nullptr : non_gt_index;
  lnms_score_bwd_kernel<<<cdiv(Rn, 4), 128, (size_t)4 * 2 * NC * sizeof(float), st>>>(NC, Rn, n, sel, cls_score, pos, d_score,
                                                                                      d_cls_score);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_nms_loss(const float* nms_multi_score, const float* nms_multi_target, int32_t first_n, int32_t C,
                           int32_t num_thresh, float loss_scale, float pos_grad_scale, float eps, float* pos_loss,
                           float* neg_loss, float* d_multi, rn_stream_t stream) {
  RN_CHECK_ARG(first_n > 0 && C > 0 && num_thresh > 0, "rn_nms_loss: bad sizes");
  RN_CHECK_ARG(nms_multi_score && nms_multi_target, "rn_nms_loss: null input");
  const size_t count = (size_t)first_n * C * num_thresh;
  const float k = loss_scale / (float)(first_n * num_thresh);
  rn::nms_loss_kernel<<<(int)std::min<size_t>((count + 255) / 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(
      nms_multi_score, nms_multi_target, count, k, eps, pos_grad_scale, pos_loss, neg_loss, d_multi);
  RN_LAUNCH_CHECK();
  return RN_OK;
}

extern "C" int rn_box_annotator_ohem(const float* cls_score, const float* bbox_pred, const float* labels,
                                     const float* bbox_targets, const float* bbox_weights, int32_t R, int32_t num_classes,
                                     int32_t num_reg_classes, int32_t roi_per_img, float* labels_ohem,
                                     float* bbox_weights_ohem, float* per_roi_loss, rn_stream_t stream) {
  RN_CHECK_ARG(R >= 0 && num_classes > 0 && num_reg_classes > 0 && roi_per_img >= 0, "rn_box_annotator_ohem: bad sizes");
  if (R == 0) return RN_OK;
  RN_CHECK_ARG(cls_score && bbox_pred && labels && bbox_targets && bbox_weights && labels_ohem && bbox_weights_ohem &&
                   per_roi_loss, "rn_box_annotator_ohem: null pointer (per_roi_loss [R] is required scratch/output)");
  cudaStream_t st = (cudaStream_t)stream;
  const int D = 4 * num_reg_classes;
  rn::ohem_loss_kernel<<<rn::cdiv(R * 32, 128), 128, 0, st>>>(R, num_classes, D, cls_score, bbox_pred, labels, bbox_targets,
                                                              bbox_weights, per_roi_loss);
  RN_LAUNCH_CHECK();
  rn::ohem_select_kernel<<<rn::cdiv(R, 128), 128, 0, st>>>(R, D, roi_per_img, per_roi_loss, labels, bbox_weights, labels_ohem,
                                                          bbox_weights_ohem);
  RN_LAUNCH_CHECK();
  return RN_OK;
}
