"""numpy oracle of the training-only pieces next to the hot path (TEST INFRASTRUCTURE, see oracle/__init__.py):

  box_annotator_ohem   relation_rcnn/operator_py/box_annotator_ohem.py:26-53 ('BoxAnnotatorOHEM' CustomOp forward)
  nms_loss             resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16_learn_nms.py:539-551 (pos / neg
                       cross-entropy terms of the learn-NMS head and the gradient MakeLoss sends back)

Tie rule of the OHEM ranking: the reference takes np.argsort(loss)[::-1] with numpy's default (unstable) sort, i.e. ties
are unspecified; the oracle (and the CUDA path) use a STABLE ascending sort reversed -> among equal losses the larger roi
index ranks first.  Float32 throughout: class softmax with a sequential sum (mshadow), log in float32, box term summed
sequentially over the 4*num_reg_classes columns."""
import numpy as np

F = np.float32


def box_annotator_ohem(cls_score, bbox_pred, labels, bbox_targets, bbox_weights, roi_per_img):
    cls_score = np.asarray(cls_score, F); labels = np.asarray(labels, F)
    e = np.exp(cls_score - cls_score.max(axis=1, keepdims=True))
    s = np.zeros(e.shape[0], F)
    for c in range(e.shape[1]):
        s = (s + e[:, c]).astype(F)
    p = (e / s[:, None]).astype(F) + F(1e-14)
    loss_cls = F(-1) * np.log(p[np.arange(p.shape[0]), labels.astype(int)])
    x = np.asarray(bbox_pred, F) - np.asarray(bbox_targets, F)
    ax = np.abs(x)
    sl = np.where(ax < F(1), F(0.5) * (x * x), ax - F(0.5)).astype(F)
    t = (np.asarray(bbox_weights, F) * sl).astype(F)
    loss_box = np.zeros(t.shape[0], F)
    for j in range(t.shape[1]):
        loss_box = (loss_box + t[:, j]).astype(F)
    loss = (loss_cls + loss_box).astype(F)
    order = np.argsort(loss, kind='stable')[::-1]
    drop = order[int(roi_per_img):]
    lab = labels.copy(); lab[drop] = -1
    w = np.asarray(bbox_weights, F).copy(); w[drop] = 0
    return lab, w, loss


def nms_loss(multi, target, first_n, num_thresh, loss_scale=1.0, pos_grad_scale=4.0, eps=1e-8):
    """-> pos_loss, neg_loss [n,C,T] and d(sum of MakeLoss outputs)/d multi (grad_scale = pos_grad_scale on the pos term)"""
    m = np.asarray(multi, F); t = np.asarray(target, F)
    k = F(loss_scale) / F(first_n * num_thresh)
    a = m + F(eps); b = F(1) - m + F(eps)
    pos = -(t * np.log(a)) * k
    neg = -((F(1) - t) * np.log(b)) * k
    d = F(pos_grad_scale) * (-(t / a) * k) + ((F(1) - t) / b) * k
    return pos.astype(F), neg.astype(F), d.astype(F)
