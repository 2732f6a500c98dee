"""numpy oracle of the learned-NMS duplicate-removal head (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates relation_rcnn/operator_py/learn_nms.py (LNMS):
  refine_boxes          LNMS:175-217 (refine_bbox_nd; symbol twin SYM_NMS_BASE:29-90)
  rank_embedding        LNMS:129-140 (extract_rank_embedding_nd)
  learn_nms_forward     LNMS:238-401 (LearnNmsOperator.forward) incl. nms_attention_nd LNMS:45-127 and the
                        test-time merge SYM_REL_NMS:553-560

Closed form per class c (valid classes only), n = first_n, feature f[i,c,:] = emb[idx[i,c]] + rank_feat[i]:
  relation module (oracle.relation_np) over the n sorted boxes of class c with d=128, dq=1024, H=16, no M<N
  slicing, then relu(f + attn), 128->T logits, sigmoid, times the sorted score.

Sorting: per class column, descending, ties -> lower roi index first (MXNet leaves it unspecified; the CUDA path
uses the same rule).  Invalid classes (max score < min(class_thresh, global max)) get zero conditional score.
"""
import numpy as np
from . import relation_np as R


def refine_boxes(rois4, deltas, im_info=None, means=None, stds=None, dtype=np.float32):
    """rois4 [R,4], deltas [R,4*K] -> [R,4,K] (x1,y1,x2,y2 on axis 1), clipped to [0, im_wh-1] when im_info given."""
    f = dtype
    b = np.asarray(rois4, f)
    d = np.asarray(deltas, f).reshape(b.shape[0], -1, 4)             # [R,K,4]
    w = (b[:, 2] - b[:, 0] + f(1))[:, None]; h = (b[:, 3] - b[:, 1] + f(1))[:, None]
    cx = (f(0.5) * (b[:, 0] + b[:, 2]))[:, None]; cy = (f(0.5) * (b[:, 1] + b[:, 3]))[:, None]
    dx, dy, dw, dh = d[:, :, 0], d[:, :, 1], d[:, :, 2], d[:, :, 3]
    if means is not None and stds is not None:
        dx = dx * f(stds[0]) + f(means[0]); dy = dy * f(stds[1]) + f(means[1])
        dw = dw * f(stds[2]) + f(means[2]); dh = dh * f(stds[3]) + f(means[3])
    rcx = cx + w * dx; rcy = cy + h * dy
    rw = w * np.exp(dw); rh = h * np.exp(dh)
    wo = f(0.5) * (rw - f(1)); ho = f(0.5) * (rh - f(1))
    out = np.stack([rcx - wo, rcy - ho, rcx + wo, rcy + ho], axis=1)  # [R,4,K]
    if im_info is not None:
        info = np.asarray(im_info, f).reshape(-1)
        lim = np.array([info[1] - 1, info[0] - 1, info[1] - 1, info[0] - 1], f).reshape(1, 4, 1)
        out = np.maximum(np.minimum(out, lim), f(0))
    return out.astype(f)


def rank_embedding(n, feat_dim=1024, wave_length=1000.0, dtype=np.float32):
    f = dtype
    rng_ = np.arange(n, dtype=f)
    dim = np.power(f(wave_length), f(2.0 / feat_dim) * np.arange(feat_dim // 2, dtype=f)).astype(f)
    div = rng_[:, None] / dim[None, :]
    return np.concatenate([np.sin(div), np.cos(div)], axis=1).astype(f)


def learn_nms_forward(cls_score, bbox_pred, rois, im_info, feat, P, first_n=100, num_fg_classes=80, num_thresh=5,
                      class_thresh=0.01, class_agnostic=True, means=None, stds=None, nongt_dim=None,
                      non_gt_index=None, merge_method=-1, dtype=np.float32, return_all=False):
    """Inputs as LNMS:429-441; P is a dict with the 14 weights by their checkpoint names.

    Returns nms_multi_score [n,C,T], sorted_bbox [n,C,4], sorted_score [n,C], final_score [n,C].
    """
    f = dtype
    cls_score = np.asarray(cls_score, f); bbox_pred = np.asarray(bbox_pred, f)
    rois = np.asarray(rois, f); feat = np.asarray(feat, f)
    if nongt_dim is not None:
        sel = np.arange(int(nongt_dim))
    elif non_gt_index is not None:
        sel = np.asarray(non_gt_index).astype(np.int64)
    else:
        sel = np.arange(cls_score.shape[0])
    cs = cls_score[sel]; bp = bbox_pred[sel]; r4 = rois[sel, 1:]
    refined = refine_boxes(r4, bp[:, 4:], im_info, means, stds, f)           # [R,4,K]
    e = np.exp(cs - cs.max(axis=1, keepdims=True))
    prob = (e / e.sum(axis=1, keepdims=True))[:, 1:]                          # [R,C]
    n, C, T = first_n, num_fg_classes, num_thresh
    order = np.argsort(-prob, axis=0, kind='stable')[:n]                      # [n,C] roi index per rank
    sorted_score = np.take_along_axis(prob, order, axis=0)                    # [n,C]
    cmax = sorted_score.max(axis=0)
    th = min(class_thresh, float(cmax.max()))
    valid = cmax >= th
    if class_agnostic:
        sorted_bbox = refined[:, :, 0][order]                                 # [n,C,4]
    else:
        sorted_bbox = np.stack([refined[order[:, c], :, c] for c in range(C)], axis=1)
    rank_feat = rank_embedding(n, 1024, dtype=f) @ np.asarray(P['nms_rank_weight'], f).T \
        + np.asarray(P['nms_rank_bias'], f)                                   # [n,128]
    emb = feat @ np.asarray(P['roi_feat_embedding_weight'], f).T + np.asarray(P['roi_feat_embedding_bias'], f)
    # NB: the reference takes roi_feat_embedding rows by the rank indices computed on the non-gt slice, i.e. it
    # indexes fc_all_2_relu directly with those indices (LNMS:339) -- identical when the non-gt rois are a prefix.
    cond = np.zeros((n, C, T), f)
    attn_all = {}
    for c in np.where(valid)[0]:
        fc = emb[order[:, c]] + rank_feat                                     # [n,128]
        o = R.relation_forward(fc, sorted_bbox[:, c, :],
                               P['nms_query_1_weight'], P['nms_query_1_bias'],
                               P['nms_key_1_weight'], P['nms_key_1_bias'],
                               P['nms_pair_pos_fc1_1_weight'], P['nms_pair_pos_fc1_1_bias'],
                               np.asarray(P['nms_linear_out_1_weight']).reshape(128, -1), P['nms_linear_out_1_bias'],
                               group=16, residual_relu=True, dtype=f)
        logit = o @ np.asarray(P['nms_logit_weight'], f).T + np.asarray(P['nms_logit_bias'], f)
        cond[:, c, :] = f(1) / (f(1) + np.exp(-logit))
        if return_all:
            attn_all[int(c)] = o
    multi = (sorted_score[:, :, None] * cond).astype(f)
    if merge_method == -1:
        final = multi.mean(axis=2, dtype=f)
    elif merge_method == -2:
        final = multi.max(axis=2)
    else:
        final = multi[:, :, merge_method]
    if return_all:
        return dict(nms_multi_score=multi, sorted_bbox=sorted_bbox.astype(f), sorted_score=sorted_score.astype(f),
                    final_score=final.astype(f), order=order, valid=valid, feat_all=attn_all)
    return multi, sorted_bbox.astype(f), sorted_score.astype(f), final.astype(f)


NMS_PARAM_SHAPES = dict(
    nms_rank_weight=(128, 1024), nms_rank_bias=(128,),
    roi_feat_embedding_weight=(128, 1024), roi_feat_embedding_bias=(128,),
    nms_pair_pos_fc1_1_weight=(16, 64), nms_pair_pos_fc1_1_bias=(16,),
    nms_query_1_weight=(1024, 128), nms_query_1_bias=(1024,),
    nms_key_1_weight=(1024, 128), nms_key_1_bias=(1024,),
    nms_linear_out_1_weight=(128, 128, 1, 1), nms_linear_out_1_bias=(128,),
    nms_logit_weight=(5, 128), nms_logit_bias=(5,))


def make_learn_nms_case(seed, R=300, C=80, d=1024, init='fan_in', n_peaky=12):
    """Synthetic head outputs: a few 'present' classes with peaked scores, the rest near-uniform low scores so
    that class pruning (LNMS:298-303) is exercised; scores have no exact ties."""
    rng = np.random.default_rng(seed)
    from .relation_np import make_boxes
    boxes = make_boxes(rng, R)
    rois = np.hstack([np.zeros((R, 1), np.float32), boxes]).astype(np.float32)
    cls_score = (rng.standard_normal((R, C + 1)) * 0.3).astype(np.float32)
    cls_score[:, 0] += 6.0                                                    # background dominates
    present = rng.choice(np.arange(1, C + 1), size=min(n_peaky, max(1, C // 2)), replace=False)
    for c in present:
        hot = rng.choice(R, size=rng.integers(5, min(60, R)), replace=False)
        cls_score[hot, c] += rng.uniform(4.0, 10.0, size=hot.size).astype(np.float32)
    bbox_pred = (rng.standard_normal((R, 8)) * 0.1).astype(np.float32)
    feat = np.maximum(rng.standard_normal((R, d)) * 0.5, 0).astype(np.float32)
    P = {}
    for k, shp in NMS_PARAM_SHAPES.items():
        if k.endswith('_bias'):
            P[k] = (rng.standard_normal(shp) * 0.05).astype(np.float32)
        else:
            fan = int(np.prod(shp[1:]))
            sd = 0.01 if init == 'ref' else 1.0 / np.sqrt(fan)
            P[k] = (rng.standard_normal(shp) * sd).astype(np.float32)
    P['nms_pair_pos_fc1_1_weight'] = (rng.standard_normal((16, 64)) * (0.01 if init == 'ref' else 0.125)).astype(np.float32)
    P['nms_pair_pos_fc1_1_bias'] = rng.uniform(0, 0.5, 16).astype(np.float32)
    P['nms_logit_bias'] = np.full(5, -3.0, np.float32) if init == 'ref' else P['nms_logit_bias']
    im_info = np.array([[600.0, 1000.0, 1.0]], np.float32)
    return dict(cls_score=cls_score, bbox_pred=bbox_pred, rois=rois, im_info=im_info, feat=feat, P=P)


def nms_multi_target(bbox, gt_box, score, target_thresh):
    """Learn-NMS training labels (relation_rcnn/operator_py/nms_multi_target.py:24-74), restated.

    bbox [n,C,4], gt_box [1,G,5] (x1,y1,x2,y2,cls), score [n,C] -> [n,C,T] float32.  Per class and threshold: every gt of
    that class claims the highest-scoring box among those whose IoU with it exceeds the threshold AND whose best-matching
    gt (first argmax) it is; a gt nobody qualifies for "claims" box 0 (argmax of zeros), which only counts if box 0
    overlaps some gt of the class above the threshold -- kept literally.
    """
    from .proposal_np import bbox_overlaps
    bbox = np.asarray(bbox, np.float32); score = np.asarray(score, np.float32); gt_box = np.asarray(gt_box, np.float32)
    n, C = bbox.shape[0], bbox.shape[1]
    T = len(target_thresh)
    out = np.zeros((n, C, T), np.float32)
    for c in range(C):
        gts = gt_box[0, gt_box[0, :, -1].astype(np.int32) == c + 1, :4]
        if len(gts) == 0:
            continue
        ov = bbox_overlaps(bbox[:, c, :].astype(np.float64), gts.astype(np.float64))      # [n,G]
        best_gt = ov.argmax(axis=1)
        for t, th in enumerate(target_thresh):
            mask = ov > th
            qual = mask & (best_gt[:, None] == np.arange(len(gts))[None, :])
            s = np.where(qual, score[:, c:c + 1], np.float32(0)).astype(np.float32)
            winners = s.argmax(axis=0)
            any_ov = mask.any(axis=1)
            for i in winners:
                if any_ov[i]:
                    out[i, c, t] = 1.0
    return out
