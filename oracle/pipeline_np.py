"""numpy/C oracle of the whole hot path for one image (TEST INFRASTRUCTURE; also the timed CPU baseline of bench.py).
Mirrors relation-networks-for-object-detection_b200/pipeline.py step for step (SYM_REL_NMS:324-565, test graph)."""
import numpy as np
from . import proposal_np as P, relation_np as R, learn_nms_np as L, rois_np as RO

NMS_NAMES = list(L.NMS_PARAM_SHAPES)


def head_forward(prm, rpn_cls_prob, rpn_bbox_pred, conv_feat, im_info, post_nms_top_n=300, pre_nms_top_n=6000,
                 first_n=100, class_thresh=0.01, dtype=np.float32):
    f = dtype
    rois, _ = P.proposal_forward(rpn_cls_prob, rpn_bbox_pred, im_info, pre_nms_top_n=pre_nms_top_n,
                                 post_nms_top_n=post_nms_top_n)
    pooled, _ = RO.roi_pool(conv_feat, rois, (7, 7), 0.0625)
    boxes = rois[:, 1:]
    x = pooled.reshape(pooled.shape[0], -1).astype(f)
    fc1 = x @ prm['fc_new_1_weight'].T.astype(f) + prm['fc_new_1_bias'].astype(f)

    def rel(x_, i):
        return R.relation_forward(x_, boxes, prm['query_%d_weight' % i], prm['query_%d_bias' % i],
                                  prm['key_%d_weight' % i], prm['key_%d_bias' % i], prm['pair_pos_fc1_%d_weight' % i],
                                  prm['pair_pos_fc1_%d_bias' % i], prm['linear_out_%d_weight' % i].reshape(1024, -1),
                                  prm['linear_out_%d_bias' % i], key_index=post_nms_top_n, group=16, residual_relu=True,
                                  dtype=f)
    a1 = rel(fc1, 1)
    fc2 = a1 @ prm['fc_new_2_weight'].T.astype(f) + prm['fc_new_2_bias'].astype(f)
    a2 = rel(fc2, 2)
    cls_score = a2 @ prm['cls_score_weight'].T.astype(f) + prm['cls_score_bias'].astype(f)
    bbox_pred = a2 @ prm['bbox_pred_weight'].T.astype(f) + prm['bbox_pred_bias'].astype(f)
    multi, sbbox, sscore, final = L.learn_nms_forward(cls_score, bbox_pred, rois, im_info, a2,
                                                      {k: prm[k] for k in NMS_NAMES}, first_n=first_n,
                                                      class_thresh=class_thresh, nongt_dim=post_nms_top_n, dtype=f)
    return dict(rois=rois, cls_score=cls_score, bbox_pred=bbox_pred, fc_all_2_relu=a2, nms_multi_score=multi,
                learn_nms_sorted_bbox=sbbox, sorted_score=sscore, nms_final_score_output=final)


def _tail(prm, rois, fc1, im_info, nongt, first_n, class_thresh, f, key_index=None):
    """relation#1 -> fc_new_2 -> relation#2 -> cls/bbox -> learn_nms (SYM_REL_NMS:346-565), shared by the three heads"""
    boxes = rois[:, 1:]
    ki = nongt if key_index is None else key_index

    def rel(x_, i):
        return R.relation_forward(x_, boxes, prm['query_%d_weight' % i], prm['query_%d_bias' % i],
                                  prm['key_%d_weight' % i], prm['key_%d_bias' % i], prm['pair_pos_fc1_%d_weight' % i],
                                  prm['pair_pos_fc1_%d_bias' % i], prm['linear_out_%d_weight' % i].reshape(1024, -1),
                                  prm['linear_out_%d_bias' % i], key_index=ki, group=16, residual_relu=True, dtype=f)
    a1 = rel(fc1, 1)
    fc2 = a1 @ prm['fc_new_2_weight'].T.astype(f) + prm['fc_new_2_bias'].astype(f)
    a2 = rel(fc2, 2)
    cls_score = a2 @ prm['cls_score_weight'].T.astype(f) + prm['cls_score_bias'].astype(f)
    bbox_pred = a2 @ prm['bbox_pred_weight'].T.astype(f) + prm['bbox_pred_bias'].astype(f)
    kw = dict(nongt_dim=nongt) if key_index is None else dict(non_gt_index=key_index)
    multi, sbbox, sscore, final = L.learn_nms_forward(cls_score, bbox_pred, rois, im_info, a2,
                                                      {k: prm[k] for k in NMS_NAMES}, first_n=first_n,
                                                      class_thresh=class_thresh, dtype=f, **kw)
    return dict(rois=rois, cls_score=cls_score, bbox_pred=bbox_pred, fc_all_2_relu=a2, nms_multi_score=multi,
                learn_nms_sorted_bbox=sbbox, sorted_score=sscore, nms_final_score_output=final)


def head_forward_dcn(prm, rois, conv_feat, im_info, first_n=100, class_thresh=0.01, dtype=np.float32):
    """Deformable Faster-RCNN head (SYM_DCN_REL_NMS:1073-1080 + the common tail); prm carries offset_weight / offset_bias."""
    f = dtype
    kw = dict(spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7, part_size=7, sample_per_part=4)
    offset_t, _ = RO.deform_psroi_pool(conv_feat, rois, None, **kw)
    offset = offset_t.reshape(offset_t.shape[0], -1).astype(f) @ prm['offset_weight'].T.astype(f) + prm['offset_bias'].astype(f)
    pooled, _ = RO.deform_psroi_pool(conv_feat, rois, offset.reshape(-1, 2, 7, 7).astype(np.float32), trans_std=0.1, **kw)
    fc1 = pooled.reshape(pooled.shape[0], -1).astype(f) @ prm['fc_new_1_weight'].T.astype(f) + prm['fc_new_1_bias'].astype(f)
    return _tail(prm, rois, fc1, im_info, rois.shape[0], first_n, class_thresh, f)


def fpn_level(rois, k_min=2, k_max=5):
    w = rois[:, 3] - rois[:, 1] + 1.0
    h = rois[:, 4] - rois[:, 2] + 1.0
    return (np.clip(np.floor(2.0 + np.log2(np.sqrt(w * h) / 224.0)), k_min, k_max) - k_min).astype(np.int64)


def head_forward_fpn(prm, rois_sorted, counts, feats, im_info, first_n=150, class_thresh=0.01, non_gt_index=None,
                     dtype=np.float32):
    """FPN head (SYM_FPN_REL_NMS:1061-1141): per-level ROIPooling (strides 4..32) concatenated in level order + the tail."""
    f = dtype
    parts, start = [], 0
    for l, n in enumerate(counts):
        if n:
            parts.append(RO.roi_pool(feats[l], rois_sorted[start:start + n], (7, 7), 1.0 / (4, 8, 16, 32)[l])[0])
        start += n
    pooled = np.concatenate(parts, 0)
    fc1 = pooled.reshape(pooled.shape[0], -1).astype(f) @ prm['fc_new_1_weight'].T.astype(f) + prm['fc_new_1_bias'].astype(f)
    return _tail(prm, rois_sorted, fc1, im_info, rois_sorted.shape[0], first_n, class_thresh, f, key_index=non_gt_index)
