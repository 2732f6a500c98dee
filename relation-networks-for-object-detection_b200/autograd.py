"""torch.autograd bindings of the hand-written forward/backward pairs -- the role MXNet's autograd plays for the reference
(every op of the path is differentiated by the framework there; here each C-ABI forward has an explicit C-ABI backward and
these Functions tie the two together so the head can be trained with ordinary torch optimizers / GradientBucket).

No arithmetic happens here: forward and backward are single calls into librelnet_b200.so.  Boxes / rois never receive
gradients (zero-gradient custom ops upstream: proposal.py:170-173; BlockGrad at SYM_REL_NMS:428)."""
import torch
from . import ops


class RelationFunction(torch.autograd.Function):
    """relu(X + relation(X)) or relation(X): rn_relation_fwd / rn_relation_bwd"""

    @staticmethod
    def forward(ctx, X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, key_index, M, group, residual_relu, precision,
                grad_precision=None):
        ctx.key_index, ctx.M, ctx.group, ctx.residual_relu, ctx.grad_precision = key_index, M, group, residual_relu, grad_precision
        out = ops.relation(X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, key_index=key_index, M=M, group=group,
                           residual_relu=residual_relu, precision=precision)
        # the executed output rides along: its sign pattern is the relu mask of the backward (as in any autograd graph)
        ctx.save_for_backward(X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, out = ctx.saved_tensors
        g = ops.relation_backward(grad_out.contiguous(), X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout,
                                  key_index=ctx.key_index, M=ctx.M, group=ctx.group, residual_relu=ctx.residual_relu,
                                  precision=ctx.grad_precision, forward_out=out if ctx.residual_relu else None)
        return (g['X'], None, g['Wq'], g['bq'], g['Wk'], g['bk'], g['Wg'], g['bg'], g['Wout'], g['bout'],
                None, None, None, None, None, None)


def relation(X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, key_index=None, M=None, group=16, residual_relu=False,
             precision='fp32', grad_precision=None):
    """differentiable object-relation module; `precision` selects the forward kernels, `grad_precision` the contraction
    engine of the backward ('f16' = tcgen05 tf32 GEMM, the default on sm_100; 'fp32' = cuBLAS fp32); gradients are fp32"""
    return RelationFunction.apply(X, boxes, Wq, bq, Wk, bk, Wg, bg, Wout, bout, key_index, M, group, residual_relu, precision,
                                  grad_precision)


class LearnNmsFunction(torch.autograd.Function):
    """nms_multi_score of the learn-NMS train graph: rn_learn_nms_fwd (class_thresh 0: no class pruning) / rn_learn_nms_bwd.
    Also returns sorted_bbox / sorted_score (non-differentiable outputs used by nms_multi_target)."""

    @staticmethod
    def forward(ctx, cls_score, bbox_pred, rois, im_info, feat, names, kw, *weights):
        ctx.save_for_backward(cls_score, bbox_pred, rois, im_info, feat, *weights)
        ctx.names, ctx.kw = names, kw
        multi, sbbox, sscore, _ = ops.learn_nms(cls_score, bbox_pred, rois, im_info, feat, dict(zip(names, weights)),
                                                precision=kw.get('precision', 'fp32'),
                                                **{k: v for k, v in kw.items() if k not in ('precision', 'grad_precision')})
        ctx.mark_non_differentiable(sbbox, sscore)
        return multi, sbbox, sscore

    @staticmethod
    def backward(ctx, g_multi, _g_bbox, _g_score):
        cls_score, bbox_pred, rois, im_info, feat = ctx.saved_tensors[:5]
        weights = dict(zip(ctx.names, ctx.saved_tensors[5:]))
        kw = {k: v for k, v in ctx.kw.items() if k not in ('precision', 'grad_precision')}
        grads, d_cls, d_feat = ops.learn_nms_backward(g_multi.contiguous(), cls_score, bbox_pred, rois, im_info, feat,
                                                      weights, precision=ctx.kw.get('grad_precision'), **kw)
        return (d_cls, None, None, None, d_feat, None, None) + tuple(grads[n] for n in ctx.names)


def learn_nms(cls_score, bbox_pred, rois, im_info, feat, weights, first_n=100, num_thresh=5, class_agnostic=True,
              means=None, stds=None, nongt_dim=None, precision='fp32', grad_precision=None):
    """differentiable learn-NMS head (train graph) -> (nms_multi_score, sorted_bbox, sorted_score)"""
    names = tuple(weights.keys())
    kw = dict(first_n=first_n, num_thresh=num_thresh, class_thresh=0.0, class_agnostic=class_agnostic, means=means,
              stds=stds, nongt_dim=nongt_dim, precision=precision, grad_precision=grad_precision)
    return LearnNmsFunction.apply(cls_score, bbox_pred, rois, im_info, feat, names, kw, *[weights[n] for n in names])


class RoiPoolFunction(torch.autograd.Function):
    """ROIPooling (max): rn_roi_pool_fwd / rn_roi_pool_bwd"""

    @staticmethod
    def forward(ctx, data, rois, pooled_size, spatial_scale):
        out, arg = ops.roi_pool(data, rois, pooled_size, spatial_scale, return_argmax=True)
        ctx.save_for_backward(arg, rois)
        ctx.shape = tuple(data.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        arg, rois = ctx.saved_tensors
        return ops.roi_pool_backward(grad_out.contiguous(), arg, rois, ctx.shape), None, None, None


def roi_pool(data, rois, pooled_size=(7, 7), spatial_scale=0.0625):
    return RoiPoolFunction.apply(data, rois, pooled_size, spatial_scale)


class DeformPsroiPoolFunction(torch.autograd.Function):
    """DeformablePSROIPooling: rn_deform_psroi_pool_fwd / _bwd (gradients to data and trans)"""

    @staticmethod
    def forward(ctx, data, rois, trans, kw):
        out, cnt = ops.deform_psroi_pool(data, rois, trans, return_count=True, **kw)
        ctx.save_for_backward(data, rois, cnt, *([trans] if trans is not None else []))
        ctx.kw, ctx.has_trans = kw, trans is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        data, rois, cnt = ctx.saved_tensors[:3]
        trans = ctx.saved_tensors[3] if ctx.has_trans else None
        dd, dt = ops.deform_psroi_pool_backward(grad_out.contiguous(), cnt, data, rois, trans, **ctx.kw)
        return dd, None, dt, None


def deform_psroi_pool(data, rois, trans=None, **kw):
    return DeformPsroiPoolFunction.apply(data, rois, trans, kw)


class DeformConvFunction(torch.autograd.Function):
    """DeformableConvolution: rn_deform_conv_fwd / _bwd.  weight_grad_deformed=False reproduces the reference's dWeight."""

    @staticmethod
    def forward(ctx, data, offset, weight, bias, kw, weight_grad_deformed, precision):
        ctx.save_for_backward(data, offset, weight)
        ctx.kw, ctx.has_bias, ctx.wgd = kw, bias is not None, weight_grad_deformed
        return ops.deform_conv(data, offset, weight, bias, precision=precision, **kw)

    @staticmethod
    def backward(ctx, grad_out):
        data, offset, weight = ctx.saved_tensors
        dd, do, dw, db = ops.deform_conv_backward(grad_out.contiguous(), data, offset, weight, has_bias=ctx.has_bias,
                                                  weight_grad_deformed=ctx.wgd, **ctx.kw)
        return dd, do, dw, db, None, None, None


def deform_conv(data, offset, weight, bias=None, weight_grad_deformed=False, precision='fp32', **kw):
    return DeformConvFunction.apply(data, offset, weight, bias, kw, weight_grad_deformed, precision)
