"""torch restatement of the learn-NMS TRAIN graph with autograd -- the oracle of rn_learn_nms_bwd / rn_nms_loss
(TEST INFRASTRUCTURE).

Follows resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16_learn_nms.py:424-501 (SYM_REL_NMS): softmax
scores -> per-class sort / take (values carry gradient through the take, the sort order does not), box refinement under
BlockGrad (:428), roi_feat_embedding + nms_rank FCs, the relation module per class, residual + relu, nms_logit, sigmoid,
times the sorted score.  The forward must equal oracle/learn_nms_np.learn_nms_forward (pinned by reference execution);
tests/test_oracle_golden.py checks that.  The loss is :539-551."""
import numpy as np
import torch
from . import learn_nms_np as LN
from . import relation_torch as RT


def learn_nms_forward(cls_score, bbox_pred, rois, im_info, feat, P, first_n=100, num_fg_classes=80, num_thresh=5,
                      class_thresh=0.0, class_agnostic=True, means=None, stds=None, nongt_dim=None):
    """cls_score, feat and the entries of P are torch tensors (any may require grad); bbox_pred / rois / im_info numpy.
    Returns nms_multi_score [n,C,T], sorted_score [n,C], order [n,C] (numpy)."""
    f = cls_score.dtype
    npf = np.float64 if f == torch.float64 else np.float32
    Rn = int(nongt_dim) if nongt_dim is not None else cls_score.shape[0]
    n, C, T = first_n, num_fg_classes, num_thresh
    refined = LN.refine_boxes(np.asarray(rois)[:Rn, 1:], np.asarray(bbox_pred)[:Rn, 4:], im_info, means, stds, npf)
    prob = torch.softmax(cls_score[:Rn], dim=1)[:, 1:]
    # ordering is decided on the float32 scores (what the op under test sees); ties -> lower index first
    p32 = torch.softmax(cls_score[:Rn].detach().float(), dim=1)[:, 1:].numpy()
    order = np.argsort(-p32, axis=0, kind='stable')[:n]
    ot = torch.from_numpy(order)
    sorted_score = torch.gather(prob, 0, ot)
    cmax = sorted_score.detach().max(dim=0).values
    valid = (cmax >= min(class_thresh, float(cmax.max()))).numpy()
    if class_agnostic:
        sorted_bbox = refined[:, :, 0][order]
    else:
        sorted_bbox = np.stack([refined[order[:, c], :, c] for c in range(C)], axis=1)
    rank_feat = torch.tensor(LN.rank_embedding(n, 1024, dtype=npf), dtype=f) @ P['nms_rank_weight'].T + P['nms_rank_bias']
    emb = feat @ P['roi_feat_embedding_weight'].T + P['roi_feat_embedding_bias']
    cond = []
    for c in range(C):
        if not valid[c]:
            cond.append(torch.zeros(n, T, dtype=f))
            continue
        fc = emb[ot[:, c]] + rank_feat
        o = RT.relation_forward(fc, torch.tensor(sorted_bbox[:, c, :], dtype=f), P['nms_query_1_weight'],
                                P['nms_query_1_bias'], P['nms_key_1_weight'], P['nms_key_1_bias'],
                                P['nms_pair_pos_fc1_1_weight'], P['nms_pair_pos_fc1_1_bias'],
                                P['nms_linear_out_1_weight'].reshape(128, -1), P['nms_linear_out_1_bias'], group=16,
                                residual_relu=True)
        cond.append(torch.sigmoid(o @ P['nms_logit_weight'].T + P['nms_logit_bias']))
    cond = torch.stack(cond, dim=1)                                   # [n,C,T]
    return sorted_score[:, :, None] * cond, sorted_score, order


def nms_loss(multi, target, first_n, num_thresh, loss_scale=1.0, eps=1e-8):
    """SYM_REL_NMS:539-547: elementwise positive / negative cross-entropy terms (MakeLoss sums them implicitly)."""
    normalizer = first_n * num_thresh
    pos = -(target * torch.log(multi + eps)) * loss_scale / normalizer
    neg = -((1.0 - target) * torch.log(1.0 - multi + eps)) * loss_scale / normalizer
    return pos, neg
