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
