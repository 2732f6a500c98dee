// learn_nms.cuh -- workspace layout of rn_learn_nms_fwd, shared with the backward (learn_nms_bwd.cu)
#pragma once
#include "common.cuh"

namespace rn {

constexpr int kNmsFeat = 128;    // nms_attention_feat_dim (LNMS:223)
constexpr int kRankDim = 1024;   // rank embedding dim (LNMS:328)

struct LnmsWs {
  float *prob, *refined, *cmax, *rank_emb, *rank_feat, *emb, *feat_cls, *boxes_cls, *feat_out, *lg_roi;
  int *rank_idx, *valid;
  void* rel_ws; size_t rel_ws_bytes;
};

rn_relation_desc lnms_inner_desc(const rn_learn_nms_desc* d);
// returns the bytes needed; fills *w when w != nullptr (0 when `bytes` is too small)
size_t lnms_carve(const rn_learn_nms_desc* d, int Rn, void* base, size_t bytes, LnmsWs* w);
int lnms_selected_rows(const rn_learn_nms_desc* d, const int* non_gt_index);

}  // namespace rn
