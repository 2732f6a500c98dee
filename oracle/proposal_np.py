"""numpy oracle of proposal / GPU-NMS / bbox_overlaps / proposal_target (TEST INFRASTRUCTURE).

Restates
  generate_anchors        lib/rpn/generate_anchor.py:22-86
  decode / clip           lib/bbox/bbox_transform.py:103-140 (nonlinear_pred, float64), :45-60 (clip_boxes)
  proposal_forward        relation_rcnn/operator_py/proposal.py:51-168
  gpu_nms                 lib/nms/gpu_nms.pyx:16-31 + lib/nms/nms_kernel.cu:24-32 (fp32 IoU), :61-77 ('>' thresh),
                          :124-139 (greedy sweep)
  bbox_overlaps           lib/bbox/bbox.pyx:15-55 (float64, +1 areas)
  proposal_target_forward relation_rcnn/operator_py/proposal_target.py:44-93 -> core/rcnn.py:288-325
                          (sample_rois_v2, the BATCH_ROIS=-1 path) -> bbox_transform.py:74-100 (encode) ->
                          bbox_regression.py:120-140 (expand)

Tie rule: numpy's ``argsort()[::-1]`` leaves the order of equal scores unspecified (introsort).  The oracle and the
CUDA path both define it as "score descending, then index DESCENDING" (what a stable ascending sort reversed
gives).  Padding when fewer than post_nms_top_n boxes survive: the reference draws ``npr.choice`` (random,
proposal.py:154-156); oracle and CUDA path both pad deterministically with keep[i % len(keep)].
"""
import numpy as np


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32)):
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    w = h = float(base_size)
    xc = yc = 0.5 * (base_size - 1)
    ws = np.round(np.sqrt(w * h / ratios))
    hs = np.round(ws * ratios)
    out = []
    for wr, hr in zip(ws, hs):
        for s in scales:
            W, Hh = wr * s, hr * s
            out.append([xc - 0.5 * (W - 1), yc - 0.5 * (Hh - 1), xc + 0.5 * (W - 1), yc + 0.5 * (Hh - 1)])
    return np.array(out, dtype=np.float64)


def decode_boxes(boxes, deltas):
    """nonlinear_pred (bbox_transform.py:114-138): float64 arithmetic on float32 deltas.  boxes [N,4], deltas [N,4].

    In the reference ``np.exp(dw)`` runs on the float32 deltas, i.e. a float32 exp whose last bit depends on the numpy
    build (SIMD expf is not correctly rounded).  The oracle (and the CUDA path) define it as the correctly rounded
    float32 exp: float32(exp(float64(dw))).
    """
    boxes = boxes.astype(np.float64)
    deltas = np.asarray(deltas, np.float32)
    w = boxes[:, 2] - boxes[:, 0] + 1.0
    h = boxes[:, 3] - boxes[:, 1] + 1.0
    cx = boxes[:, 0] + 0.5 * (w - 1.0)
    cy = boxes[:, 1] + 0.5 * (h - 1.0)
    dx, dy, dw, dh = (deltas[:, i] for i in range(4))
    pcx = dx * w + cx
    pcy = dy * h + cy
    pw = np.exp(dw.astype(np.float64)).astype(np.float32) * w
    ph = np.exp(dh.astype(np.float64)).astype(np.float32) * h
    out = np.zeros(deltas.shape, dtype=np.float64)
    out[:, 0] = pcx - 0.5 * (pw - 1.0)
    out[:, 1] = pcy - 0.5 * (ph - 1.0)
    out[:, 2] = pcx + 0.5 * (pw - 1.0)
    out[:, 3] = pcy + 0.5 * (ph - 1.0)
    return out


def iou_f32(a, b):
    """devIoU of nms_kernel.cu:24-32, float32 op by op.  a [4], b [K,4] float32."""
    f = np.float32
    left = np.maximum(a[0], b[:, 0]); right = np.minimum(a[2], b[:, 2])
    top = np.maximum(a[1], b[:, 1]); bottom = np.minimum(a[3], b[:, 3])
    width = np.maximum(right - left + f(1), f(0))
    height = np.maximum(bottom - top + f(1), f(0))
    inter = width * height
    sa = (a[2] - a[0] + f(1)) * (a[3] - a[1] + f(1))
    sb = (b[:, 2] - b[:, 0] + f(1)) * (b[:, 3] - b[:, 1] + f(1))
    return inter / (sa + sb - inter)


def nms_sorted(boxes, thresh, max_keep=None):
    """Greedy sweep over boxes already sorted by score (nms_kernel.cu semantics: suppress j>i when IoU > thresh)."""
    boxes = np.ascontiguousarray(boxes[:, :4], dtype=np.float32)
    n = boxes.shape[0]
    removed = np.zeros(n, dtype=bool)
    keep = []
    th = np.float32(thresh)
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        if max_keep is not None and len(keep) >= max_keep:
            break
        if i + 1 < n:
            ov = iou_f32(boxes[i], boxes[i + 1:])
            removed[i + 1:] |= ov > th
    return np.asarray(keep, dtype=np.int64)


def order_desc(scores):
    """score descending, ties -> larger index first (== argsort(kind='stable')[::-1])."""
    return np.argsort(scores, kind='stable')[::-1]


def gpu_nms(dets, thresh):
    """gpu_nms.pyx:16-31: re-sort by score (col 4) descending, NMS, map back.  dets [n,5] float32."""
    dets = np.asarray(dets, dtype=np.float32)
    order = order_desc(dets[:, 4])
    keep = nms_sorted(dets[order], thresh)
    return list(order[keep])


def proposal_forward(cls_prob, bbox_pred, im_info, feat_stride=16, scales=(4, 8, 16, 32), ratios=(0.5, 1, 2),
                     pre_nms_top_n=6000, post_nms_top_n=300, thresh=0.7, min_size=0, return_aux=False):
    """cls_prob [1,2A,H,W], bbox_pred [1,4A,H,W] float32, im_info [1,3] -> rois [post,5], scores [post,1]."""
    anchors0 = generate_anchors(feat_stride, ratios, scales)
    A = anchors0.shape[0]
    cls_prob = np.asarray(cls_prob, np.float32); bbox_pred = np.asarray(bbox_pred, np.float32)
    info = np.asarray(im_info, np.float32).reshape(-1)
    height, width = int(info[0] / feat_stride), int(info[1] / feat_stride)
    scores = cls_prob[:, A:, :height, :width]
    deltas = bbox_pred[:, :, :height, :width]
    sx = np.arange(0, width) * feat_stride
    sy = np.arange(0, height) * feat_stride
    sx, sy = np.meshgrid(sx, sy)
    shifts = np.stack([sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel()], axis=1)            # [K,4]
    anchors = (anchors0[None, :, :] + shifts[:, None, :]).reshape(-1, 4)                  # (h,w,a) order
    deltas = deltas.transpose(0, 2, 3, 1).reshape(-1, 4)
    scores = scores.transpose(0, 2, 3, 1).reshape(-1)
    props = decode_boxes(anchors, deltas)                                                 # float64
    im_h, im_w = info[0], info[1]                                                         # float32 scalars
    props[:, 0] = np.maximum(np.minimum(props[:, 0], im_w - 1), 0)
    props[:, 1] = np.maximum(np.minimum(props[:, 1], im_h - 1), 0)
    props[:, 2] = np.maximum(np.minimum(props[:, 2], im_w - 1), 0)
    props[:, 3] = np.maximum(np.minimum(props[:, 3], im_h - 1), 0)
    ms = min_size * info[2]
    ws = props[:, 2] - props[:, 0] + 1
    hs = props[:, 3] - props[:, 1] + 1
    keep0 = np.where((ws >= ms) & (hs >= ms))[0]
    props = props[keep0]; scores = scores[keep0]
    order = order_desc(scores)
    if pre_nms_top_n > 0:
        order = order[:pre_nms_top_n]
    props = props[order]; scores = scores[order]
    det = np.hstack([props, scores[:, None]]).astype(np.float32)
    keep = np.asarray(gpu_nms(det, thresh), dtype=np.int64)
    if post_nms_top_n > 0:
        keep = keep[:post_nms_top_n]
    n_kept = len(keep)
    if n_kept < post_nms_top_n:
        pad = keep[np.arange(post_nms_top_n - n_kept) % n_kept]
        keep = np.hstack([keep, pad])
    rois = np.hstack([np.zeros((len(keep), 1), np.float32), props[keep].astype(np.float32)])
    sc = scores[keep].astype(np.float32)[:, None]
    if return_aux:
        return rois, sc, dict(pre_nms_index=keep0[order], keep=keep, n_kept=n_kept, det=det)
    return rois, sc


def bbox_overlaps(boxes, query_boxes):
    """bbox.pyx:15-55, float64; zero when iw<=0 or ih<=0.  boxes [N,4], query [K,4] -> [N,K]."""
    b = np.asarray(boxes, np.float64); q = np.asarray(query_boxes, np.float64)
    N, K = b.shape[0], q.shape[0]
    if N == 0 or K == 0:
        return np.zeros((N, K), np.float64)
    qa = (q[:, 2] - q[:, 0] + 1) * (q[:, 3] - q[:, 1] + 1)
    iw = np.minimum(b[:, None, 2], q[None, :, 2]) - np.maximum(b[:, None, 0], q[None, :, 0]) + 1
    ih = np.minimum(b[:, None, 3], q[None, :, 3]) - np.maximum(b[:, None, 1], q[None, :, 1]) + 1
    ba = ((b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1))[:, None]
    ua = ba + qa[None, :] - iw * ih
    ov = np.where((iw > 0) & (ih > 0), iw * ih / ua, 0.0)
    return ov


def encode_boxes(ex, gt):
    """nonlinear_transform, bbox_transform.py:74-100 (float arithmetic in the dtype of the inputs)."""
    ew = ex[:, 2] - ex[:, 0] + 1.0; eh = ex[:, 3] - ex[:, 1] + 1.0
    ecx = ex[:, 0] + 0.5 * (ew - 1.0); ecy = ex[:, 1] + 0.5 * (eh - 1.0)
    gw = gt[:, 2] - gt[:, 0] + 1.0; gh = gt[:, 3] - gt[:, 1] + 1.0
    gcx = gt[:, 0] + 0.5 * (gw - 1.0); gcy = gt[:, 1] + 0.5 * (gh - 1.0)
    return np.vstack(((gcx - ecx) / (ew + 1e-14), (gcy - ecy) / (eh + 1e-14),
                      np.log(gw / ew), np.log(gh / eh))).transpose()


def proposal_target_forward(rois, gt_boxes, num_reg_classes=2, class_agnostic=True, bg_thresh_hi=0.5,
                            means=(0.0, 0.0, 0.0, 0.0), stds=(0.1, 0.1, 0.2, 0.2), normalize=True,
                            bbox_weights=(1.0, 1.0, 1.0, 1.0)):
    """BATCH_ROIS=-1 path: every roi (+ every gt box appended) is kept, labelled and given targets.

    rois [N,5], gt_boxes [G,5] (x1,y1,x2,y2,cls) float32 -> rois' [N+G,5], label [N+G], target [N+G,4R], weight.
    float32 rois/gt arithmetic in encode (numpy float32 in the reference: rois[:,1:] and gt are float32),
    float64 IoU, first-max argmax.
    """
    rois = np.asarray(rois, np.float32); gt = np.asarray(gt_boxes, np.float32)
    all_rois = np.vstack((rois, np.hstack((np.zeros((gt.shape[0], 1), gt.dtype), gt[:, :-1]))))
    ov = bbox_overlaps(all_rois[:, 1:].astype(np.float64), gt[:, :4].astype(np.float64))
    assign = ov.argmax(axis=1)
    mx = ov.max(axis=1)
    labels = gt[assign, 4].copy()
    labels[mx < bg_thresh_hi] = 0
    targets = encode_boxes(all_rois[:, 1:], gt[assign, :4])                     # float32 in, float32 arithmetic
    if normalize:
        targets = (targets - np.array(means)) / np.array(stds)                  # -> float64
    R = 2 if class_agnostic else num_reg_classes
    bt = np.zeros((labels.size, 4 * R), np.float32)
    bw = np.zeros_like(bt)
    for i in np.where(labels > 0)[0]:
        s = 4 if class_agnostic else int(4 * labels[i])
        bt[i, s:s + 4] = targets[i]
        bw[i, s:s + 4] = bbox_weights
    return all_rois, labels, bt, bw


# ------------------------------------------------------------------------------------------------
def make_proposal_case(seed, H=38, W=63, A=12, im_info=(600.0, 1000.0, 1.0), delta_scale=0.2):
    """Synthetic RPN outputs: fg scores ~ U(0,1) made unique so the sort has no
    ties; deltas ~ N(0, delta_scale)."""
    rng = np.random.default_rng(seed)
    n = A * H * W
    fg = rng.permutation(n).astype(np.float64)
    fg = ((fg + 0.5) / n).astype(np.float32)              # unique float32 values in (0,1)
    assert np.unique(fg).size == n
    fg = fg.reshape(1, A, H, W)
    cls_prob = np.concatenate([1 - fg, fg], axis=1).astype(np.float32)
    bbox_pred = (rng.standard_normal((1, 4 * A, H, W)) * delta_scale).astype(np.float32)
    return cls_prob, bbox_pred, np.asarray([im_info], np.float32)
