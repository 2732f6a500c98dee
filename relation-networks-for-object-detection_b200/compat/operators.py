"""The reference's Python CustomOp protocol (mx.operator.CustomOp / CustomOpProp) over torch tensors.

Each ``*Prop`` takes the reference's STRING kwargs (they arrive as strings from ``mx.sym.Custom``; parsing mirrors
proposal.py:205-212, proposal_target.py:103-109, learn_nms.py:410-427), exposes ``list_arguments`` / ``list_outputs`` /
``infer_shape`` / ``create_operator``; each operator has ``forward(is_train, req, in_data, out_data, aux)`` writing with
``assign(dst, req, src)`` and a ``backward`` that assigns zero gradients, exactly like the reference ops.
``Custom(op_type=..., **kwargs)`` is the ``mx.sym.Custom`` call form.
"""
import numpy as np
import torch
from .. import ops

REGISTRY = {}


def register(name):
    def deco(cls):
        REGISTRY[name] = cls
        return cls
    return deco


class CustomOp(object):
    def assign(self, dst, req, src):
        if req == 'null':
            return
        if req == 'add':
            dst.add_(src.reshape(dst.shape))
        else:                                   # 'write' / 'inplace'
            dst.copy_(src.reshape(dst.shape))

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        for g, r in zip(in_grad, req):          # all reference ops of this path return zero input gradients
            if r != 'null' and g is not None:
                g.zero_()


class CustomOpProp(object):
    def __init__(self, need_top_grad=False):
        self.need_top_grad = need_top_grad

    def declare_backward_dependency(self, out_grad, in_data, out_data):
        return []


def _tuple_of_floats(s):
    if isinstance(s, (tuple, list)):
        return tuple(float(v) for v in s)
    return tuple(float(v) for v in str(s).strip('()[] ').replace(',', ' ').split())


# ------------------------------------------------------------------------------------------------------- proposal
class ProposalOperator(CustomOp):
    """operator_py/proposal.py:31-173"""

    def __init__(self, feat_stride, scales, ratios, output_score, rpn_pre_nms_top_n, rpn_post_nms_top_n, threshold,
                 rpn_min_size):
        self._feat_stride = int(feat_stride)
        self._scales = _tuple_of_floats(scales)
        self._ratios = _tuple_of_floats(ratios)
        self._output_score = output_score
        self._rpn_pre_nms_top_n, self._rpn_post_nms_top_n = int(rpn_pre_nms_top_n), int(rpn_post_nms_top_n)
        self._threshold, self._rpn_min_size = float(threshold), float(rpn_min_size)

    def forward(self, is_train, req, in_data, out_data, aux):
        if in_data[0].shape[0] > 1:
            raise ValueError("Sorry, multiple images each device is not implemented")
        rois, scores = ops.proposal(in_data[0], in_data[1], in_data[2], self._feat_stride, self._scales, self._ratios,
                                    self._rpn_pre_nms_top_n, self._rpn_post_nms_top_n, self._threshold,
                                    self._rpn_min_size)
        self.assign(out_data[0], req[0], rois)
        if self._output_score:
            self.assign(out_data[1], req[1], scores)


@register('proposal')
class ProposalProp(CustomOpProp):
    """operator_py/proposal.py:200-243"""

    def __init__(self, feat_stride='16', scales='(8, 16, 32)', ratios='(0.5, 1, 2)', output_score='False',
                 rpn_pre_nms_top_n='6000', rpn_post_nms_top_n='300', threshold='0.3', rpn_min_size='16'):
        super(ProposalProp, self).__init__(need_top_grad=False)
        self._feat_stride = int(feat_stride)
        self._scales, self._ratios = scales, ratios
        self._output_score = str(output_score) in ('True', 'true', '1')
        self._rpn_pre_nms_top_n, self._rpn_post_nms_top_n = int(rpn_pre_nms_top_n), int(rpn_post_nms_top_n)
        self._threshold, self._rpn_min_size = float(threshold), int(float(rpn_min_size))

    def list_arguments(self):
        return ['cls_prob', 'bbox_pred', 'im_info']

    def list_outputs(self):
        return ['output', 'score'] if self._output_score else ['output']

    def infer_shape(self, in_shape):
        cls_prob_shape, bbox_pred_shape = in_shape[0], in_shape[1]
        assert cls_prob_shape[0] == bbox_pred_shape[0], 'ROI number does not equal in cls and reg'
        batch_size = cls_prob_shape[0]
        im_info_shape = (batch_size, 3)
        output_shape, score_shape = (self._rpn_post_nms_top_n, 5), (self._rpn_post_nms_top_n, 1)
        outs = [output_shape, score_shape] if self._output_score else [output_shape]
        return [cls_prob_shape, bbox_pred_shape, im_info_shape], outs

    def create_operator(self, ctx=None, shapes=None, dtypes=None):
        return ProposalOperator(self._feat_stride, self._scales, self._ratios, self._output_score,
                                self._rpn_pre_nms_top_n, self._rpn_post_nms_top_n, self._threshold, self._rpn_min_size)


# ------------------------------------------------------------------------------------------------ proposal_target
class ProposalTargetOperator(CustomOp):
    """operator_py/proposal_target.py:30-97; only the BATCH_ROIS == -1 path works end to end in the reference
     and is the one implemented."""

    def __init__(self, num_classes, batch_images, batch_rois, cfg, fg_fraction):
        self._num_classes, self._batch_images, self._batch_rois = num_classes, batch_images, batch_rois
        self._cfg, self._fg_fraction = cfg, fg_fraction

    def forward(self, is_train, req, in_data, out_data, aux):
        assert self._batch_rois == -1 or self._batch_rois % self._batch_images == 0, \
            'batchimages {} must devide batch_rois {}'.format(self._batch_images, self._batch_rois)
        if self._batch_rois != -1:
            raise NotImplementedError('proposal_target: only BATCH_ROIS == -1 (sample_rois_v2) is supported')
        cfg = self._cfg
        tr = cfg['TRAIN'] if isinstance(cfg, dict) else cfg.TRAIN
        get = (lambda o, k: o[k]) if isinstance(tr, dict) else getattr
        agnostic = bool(cfg['CLASS_AGNOSTIC'] if isinstance(cfg, dict) else cfg.CLASS_AGNOSTIC)
        rois, label, bt, bw = ops.proposal_target(
            in_data[0], in_data[1], num_reg_classes=self._num_classes, class_agnostic=agnostic,
            bg_thresh_hi=float(get(tr, 'BG_THRESH_HI')), normalize=bool(get(tr, 'BBOX_NORMALIZATION_PRECOMPUTED')),
            means=tuple(get(tr, 'BBOX_MEANS')), stds=tuple(get(tr, 'BBOX_STDS')),
            bbox_weights=tuple(float(v) for v in get(tr, 'BBOX_WEIGHTS')))
        assert bool((rois[:, 0] == 0).all()), 'Only single item batches are supported'
        for ind, val in enumerate([rois, label, bt, bw]):
            self.assign(out_data[ind], req[ind], val)


@register('proposal_target')
class ProposalTargetProp(CustomOpProp):
    """operator_py/proposal_target.py:100-143 (``cfg`` may be the pickled string the reference passes, or a dict)."""

    def __init__(self, num_classes, batch_images, batch_rois, cfg, fg_fraction='0.25'):
        super(ProposalTargetProp, self).__init__(need_top_grad=False)
        self._num_classes, self._batch_images, self._batch_rois = int(num_classes), int(batch_images), int(batch_rois)
        if isinstance(cfg, (bytes, str)):
            import pickle
            cfg = pickle.loads(cfg if isinstance(cfg, bytes) else cfg.encode('latin1'))
        self._cfg, self._fg_fraction = cfg, float(fg_fraction)

    def list_arguments(self):
        return ['rois', 'gt_boxes']

    def list_outputs(self):
        return ['rois_output', 'label', 'bbox_target', 'bbox_weight']

    def infer_shape(self, in_shape):
        rpn_rois_shape, gt_boxes_shape = in_shape[0], in_shape[1]
        rois = rpn_rois_shape[0] + gt_boxes_shape[0] if self._batch_rois == -1 else self._batch_rois
        return [rpn_rois_shape, gt_boxes_shape], [(rois, 5), (rois,), (rois, self._num_classes * 4),
                                                   (rois, self._num_classes * 4)]

    def create_operator(self, ctx=None, shapes=None, dtypes=None):
        return ProposalTargetOperator(self._num_classes, self._batch_images, self._batch_rois, self._cfg,
                                      self._fg_fraction)


# ------------------------------------------------------------------------------------------------------ learn_nms
class LearnNmsOperator(CustomOp):
    """operator_py/learn_nms.py:219-406"""

    def __init__(self, num_fg_classes, bbox_means, bbox_stds, first_n, class_agnostic, num_thresh, class_thresh,
                 nongt_dim=None, has_non_gt_index=False, precision=None, merge_method=-1):
        self.num_fg_classes, self.nongt_dim, self.has_non_gt_index = num_fg_classes, nongt_dim, has_non_gt_index
        self.bbox_means, self.bbox_stds, self.first_n = bbox_means, bbox_stds, first_n
        self.class_agnostic, self.num_thresh, self.class_thresh = class_agnostic, num_thresh, class_thresh
        self.precision, self.merge_method = precision, merge_method
        self.nms_final_score = None

    def forward(self, is_train, req, in_data, out_data, aux):
        names = LearnNmsProp.ARGS + (['non_gt_index'] if self.has_non_gt_index else [])
        t = dict(zip(names, in_data))
        weights = {k: t[k] for k in LearnNmsProp.ARGS[5:]}
        multi, sbbox, sscore, final = ops.learn_nms(
            t['cls_score'], t['bbox_pred'], t['rois'], t['im_info'], t['fc_all_2_relu'], weights, first_n=self.first_n,
            num_thresh=self.num_thresh, class_thresh=self.class_thresh, class_agnostic=self.class_agnostic,
            means=self.bbox_means, stds=self.bbox_stds, nongt_dim=self.nongt_dim,
            non_gt_index=t.get('non_gt_index') if self.nongt_dim is None else None, merge_method=self.merge_method,
            precision=self.precision)
        self.nms_final_score = final           # SYM_REL_NMS:553-560 merge, fused into the same call
        self.assign(out_data[0], req[0], multi)
        self.assign(out_data[1], req[1], sbbox)
        self.assign(out_data[2], req[2], sscore)


@register('learn_nms')
class LearnNmsProp(CustomOpProp):
    """operator_py/learn_nms.py:408-457"""
    ARGS = ['cls_score', 'bbox_pred', 'rois', 'im_info', 'fc_all_2_relu', 'nms_rank_weight', 'nms_rank_bias',
            'roi_feat_embedding_weight', 'roi_feat_embedding_bias', 'nms_pair_pos_fc1_1_weight',
            'nms_pair_pos_fc1_1_bias', 'nms_query_1_weight', 'nms_query_1_bias', 'nms_key_1_weight', 'nms_key_1_bias',
            'nms_linear_out_1_weight', 'nms_linear_out_1_bias', 'nms_logit_weight', 'nms_logit_bias']

    def __init__(self, num_fg_classes, bbox_means, bbox_stds, first_n, class_agnostic, num_thresh, class_thresh,
                 nongt_dim, has_non_gt_index):
        super(LearnNmsProp, self).__init__(need_top_grad=False)
        self.num_fg_classes = int(num_fg_classes)
        self.nongt_dim = int(nongt_dim) if str(nongt_dim) != 'None' else None
        self.class_thresh = float(class_thresh)
        bbox_means, bbox_stds = str(bbox_means), str(bbox_stds)
        # gluon customops use , to separate elements, make sure this doesn't happen (learn_nms.py:416-417)
        assert ',' not in bbox_means and ',' not in bbox_stds
        if bbox_means == 'None' or bbox_stds == 'None':
            self.bbox_means = self.bbox_stds = None
        else:
            self.bbox_means = np.array(bbox_means[1:-1].split(), dtype=float)
            self.bbox_stds = np.array(bbox_stds[1:-1].split(), dtype=float)
        self.first_n = int(first_n)
        self.class_agnostic = str(class_agnostic) == 'True'
        self.num_thresh = int(num_thresh)
        self.has_non_gt_index = str(has_non_gt_index) == 'True'

    def list_arguments(self):
        return self.ARGS + (['non_gt_index'] if self.has_non_gt_index else [])

    def list_outputs(self):
        return ['nms_multi_score', 'sorted_bbox', 'sorted_score']

    def infer_shape(self, in_shape):
        return in_shape, [(self.first_n, self.num_fg_classes, self.num_thresh), (self.first_n, self.num_fg_classes, 4),
                          (self.first_n, self.num_fg_classes)]

    def create_operator(self, ctx=None, shapes=None, dtypes=None):
        return LearnNmsOperator(self.num_fg_classes, self.bbox_means, self.bbox_stds, self.first_n, self.class_agnostic,
                                self.num_thresh, self.class_thresh, self.nongt_dim, self.has_non_gt_index)


# ----------------------------------------------------------------------------------------------- nms_multi_target
class NmsMultiTargetOp(CustomOp):
    """operator_py/nms_multi_target.py:18-79"""

    def __init__(self, target_thresh):
        self._target_thresh = target_thresh
        self._num_thresh = len(target_thresh)

    def forward(self, is_train, req, in_data, out_data, aux):
        self.assign(out_data[0], req[0], ops.nms_multi_target(in_data[0], in_data[1], in_data[2], self._target_thresh))


@register('nms_multi_target')
class NmsMultiTargetProp(CustomOpProp):
    """operator_py/nms_multi_target.py:82-112 (target_thresh arrives as the string '[0.5 0.6 0.7 0.8 0.9]')"""

    def __init__(self, target_thresh):
        super(NmsMultiTargetProp, self).__init__(need_top_grad=False)
        if isinstance(target_thresh, str):
            target_thresh = [float(v) for v in target_thresh.strip('[]() ').replace(',', ' ').split()]
        self._target_thresh = np.asarray(target_thresh, dtype=float)
        self._num_thresh = len(self._target_thresh)

    def list_arguments(self):
        return ['bbox', 'gt_bbox', 'score']

    def list_outputs(self):
        return ['nms_multi_target']

    def infer_shape(self, in_shape):
        bbox_shape, score_shape = in_shape[0], in_shape[2]
        assert bbox_shape[0] == score_shape[0], 'ROI number should be same for bbox and score'
        return in_shape, [(bbox_shape[0], bbox_shape[1], self._num_thresh)]

    def create_operator(self, ctx=None, shapes=None, dtypes=None):
        return NmsMultiTargetOp(self._target_thresh)


# ----------------------------------------------------------------------------------------------- BoxAnnotatorOHEM
class BoxAnnotatorOHEMOperator(CustomOp):
    """operator_py/box_annotator_ohem.py:19-58"""

    def __init__(self, num_classes, num_reg_classes, roi_per_img):
        self._num_classes, self._num_reg_classes, self._roi_per_img = num_classes, num_reg_classes, roi_per_img

    def forward(self, is_train, req, in_data, out_data, aux):
        lab, w = ops.box_annotator_ohem(in_data[0], in_data[1], in_data[2], in_data[3], in_data[4], self._num_classes,
                                        self._num_reg_classes, self._roi_per_img)
        for ind, val in enumerate([lab, w]):
            self.assign(out_data[ind], req[ind], val)


@register('BoxAnnotatorOHEM')
class BoxAnnotatorOHEMProp(CustomOpProp):
    """operator_py/box_annotator_ohem.py:61-88 (kwargs arrive as strings)"""

    def __init__(self, num_classes, num_reg_classes, roi_per_img):
        super(BoxAnnotatorOHEMProp, self).__init__(need_top_grad=False)
        self._num_classes = int(num_classes)
        self._num_reg_classes = int(num_reg_classes)
        self._roi_per_img = int(roi_per_img)

    def list_arguments(self):
        return ['cls_score', 'bbox_pred', 'labels', 'bbox_targets', 'bbox_weights']

    def list_outputs(self):
        return ['labels_ohem', 'bbox_weights_ohem']

    def infer_shape(self, in_shape):
        return in_shape, [in_shape[2], in_shape[4]]

    def create_operator(self, ctx=None, shapes=None, dtypes=None):
        return BoxAnnotatorOHEMOperator(self._num_classes, self._num_reg_classes, self._roi_per_img)


def Custom(op_type, name=None, **kwargs):
    """``mx.sym.Custom(op_type=..., tensor kwargs..., string kwargs...)`` evaluated eagerly on torch tensors."""
    tensors = {k: v for k, v in kwargs.items() if isinstance(v, torch.Tensor)}
    attrs = {k: (v if isinstance(v, (bytes, dict)) or hasattr(v, 'TRAIN') else str(v))
             for k, v in kwargs.items() if not isinstance(v, torch.Tensor)}
    prop = REGISTRY[op_type](**attrs)
    in_data = [tensors[k] for k in prop.list_arguments()]
    _, out_shapes = prop.infer_shape([tuple(t.shape) for t in in_data])
    out_data = [torch.empty(s, dtype=torch.float32, device=in_data[0].device) for s in out_shapes]
    op = prop.create_operator(None, None, None)
    op.forward(False, ['write'] * len(out_data), in_data, out_data, [])
    return out_data[0] if len(out_data) == 1 else out_data


# --------------------------------------------------------------------------- the C symbol and the operator_cxx ops
def gpu_nms(dets, thresh, device_id=0):
    """lib/nms/gpu_nms.pyx:16-31: dets [n,5] (x1,y1,x2,y2,score) -> list of kept indices, highest score first.
    dets may be a numpy array (the reference's calling convention) or a CUDA tensor."""
    t = torch.as_tensor(np.asarray(dets, dtype=np.float32)) if not isinstance(dets, torch.Tensor) else dets.float()
    t = t.cuda(device_id) if not t.is_cuda else t
    order = torch.flip(torch.sort(t[:, 4], stable=True).indices, dims=[0])       # == argsort(kind='stable')[::-1]
    keep, num = ops.nms(t[order].contiguous(), float(thresh))
    return order[keep[:int(num.item())].long()].tolist()


def bbox_overlaps_cython(boxes, query_boxes):
    """lib/bbox/bbox.pyx:15-55: float64 [N,4] x [K,4] -> [N,K] (numpy in, numpy out like the reference)."""
    b = torch.as_tensor(np.asarray(boxes, dtype=np.float64)).cuda()
    q = torch.as_tensor(np.asarray(query_boxes, dtype=np.float64)).cuda()
    return ops.bbox_overlaps(b, q).cpu().numpy()


def DeformableConvolution(data, offset, weight, bias=None, kernel=(3, 3), stride=(1, 1), dilate=(1, 1), pad=(0, 0),
                          num_filter=None, num_group=1, num_deformable_group=1, no_bias=False, name=None, **kw):
    """mx.contrib.sym.DeformableConvolution (operator_cxx/deformable_convolution-inl.h:39-76 parameters)."""
    assert num_filter is None or num_filter == weight.shape[0]
    return ops.deform_conv(data, offset, weight, None if no_bias else bias, kernel, pad, stride, dilate, num_group,
                           num_deformable_group)


def DeformablePSROIPooling(data, rois, trans=None, spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7,
                           part_size=0, sample_per_part=1, trans_std=0.0, no_trans=False, name=None, **kw):
    """mx.contrib.sym.DeformablePSROIPooling (operator_cxx/deformable_psroi_pooling-inl.h:32-54 parameters)."""
    return ops.deform_psroi_pool(data, rois, None if no_trans else trans, spatial_scale, output_dim, group_size,
                                 pooled_size, part_size, sample_per_part, trans_std, no_trans)


def ROIPooling(data, rois, pooled_size=(7, 7), spatial_scale=0.0625, name=None, **kw):
    """mx.symbol.ROIPooling (call sites SYM_REL:252-253)."""
    return ops.roi_pool(data, rois, tuple(pooled_size), spatial_scale)
