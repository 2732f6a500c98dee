"""Symbol-class methods of the reference's relation symbols, same names / argument order / asserts.

Reference: relation_rcnn/symbols/resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16.py (SYM_REL) :30-151,
FPN variant resnet_v1_101_rcnn_fpn_attention_..._learn_nms.py (SYM_FPN_REL_NMS) :843-977, learn-NMS variant
(SYM_REL_NMS) :158-238.

The reference materialises position_matrix [N,M,4] and position_embedding [N,M,64] and hands them to the attention
function.  Here both are LAZY handles (they only carry the boxes): the fused kernel recomputes the geometry on the fly,
so a line-for-line copy of a get_symbol body binds unchanged while the 23 MB embedding never exists.  ``materialize()``
gives the real tensor (tests).
"""
import torch
from .. import ops


def _keys_arg(nongt_dim, non_gt_index):
    """The reference has two spellings of the same slot: ``extract_position_matrix(bbox, nongt_dim)`` with an int (SYM_REL:52) and
    ``extract_position_matrix(bbox, non_gt_index)`` with an index array (SYM_FPN_REL_NMS:860; likewise the third argument of
    ``attention_module_multi_head``).  A positional call of the FPN form lands in ``nongt_dim``: an array there is the index list."""
    if nongt_dim is not None and non_gt_index is None and not isinstance(nongt_dim, (int, float)) and getattr(nongt_dim, 'ndim', 0) >= 1:
        return None, nongt_dim
    return nongt_dim, non_gt_index


class PositionMatrix(object):
    """Lazy [num_rois, nongt_dim, 4] (SYM_REL:47-83); ``non_gt_index`` for the FPN form (SYM_FPN_REL_NMS:860-905)."""

    def __init__(self, bbox, nongt_dim=None, non_gt_index=None):
        nongt_dim, non_gt_index = _keys_arg(nongt_dim, non_gt_index)
        self.bbox, self.nongt_dim, self.non_gt_index = bbox, nongt_dim, non_gt_index

    def materialize(self):
        return ops.pos_embed(self.bbox, M=self.nongt_dim, key_index=self.non_gt_index, want_emb=False)[0]


class PositionEmbedding(object):
    """Lazy [num_rois, nongt_dim, feat_dim] (SYM_REL:30-44)."""

    def __init__(self, position_mat, feat_dim, wave_length=1000):
        self.pm, self.feat_dim, self.wave_length = position_mat, feat_dim, wave_length

    def materialize(self):
        return ops.pos_embed(self.pm.bbox, M=self.pm.nongt_dim, key_index=self.pm.non_gt_index, E=self.feat_dim,
                             wave_length=float(self.wave_length), want_eps=False)[1]


class LazySoftmax(object):
    """The second output of attention_module_nms_multi_head (aff_softmax).  The reference returns it from the graph
    function but the learn-NMS head never reads it (SYM_REL_NMS:486 takes only the first output); it is therefore
    computed on demand: ``materialize()`` runs the library's fp32 path, which is the one that writes the softmax."""

    def __init__(self, fn):
        self._fn, self._value = fn, None

    def materialize(self):
        if self._value is None:
            self._value = self._fn()
        return self._value


class RelationSymbols(object):
    """``params``: dict name -> CUDA tensor with the checkpoint names the reference uses
    ('pair_pos_fc1_1_weight', 'query_1_weight', 'key_1_bias', 'linear_out_1_weight', ...)."""

    def __init__(self, params, precision=None):
        self.params = params
        self.precision = precision

    @staticmethod
    def extract_position_embedding(position_mat, feat_dim, wave_length=1000):
        return PositionEmbedding(position_mat, feat_dim, wave_length)

    @staticmethod
    def extract_position_matrix(bbox, nongt_dim=None, non_gt_index=None):
        return PositionMatrix(bbox, nongt_dim, non_gt_index)

    def attention_module_multi_head(self, roi_feat, position_embedding, nongt_dim=None, fc_dim=16, feat_dim=1024,
                                    dim=(1024, 1024, 1024), group=16, index=1, non_gt_index=None, residual_relu=False):
        """SYM_REL:85-151.  Returns the attention output [num_rois, dim[2]]; ``residual_relu=True`` additionally fuses
        the caller's ``relu(roi_feat + attention)`` (SYM_REL:267-268)."""
        assert dim[0] == dim[1], 'Matrix multiply requires same dimensions!'
        assert fc_dim == group, 'fc_dim != group'
        P, i = self.params, str(index)
        pm = position_embedding.pm
        nongt_dim, non_gt_index = _keys_arg(nongt_dim, non_gt_index)
        key_index = non_gt_index if non_gt_index is not None else pm.non_gt_index
        M = None if key_index is not None else (nongt_dim if nongt_dim is not None else pm.nongt_dim)
        return ops.relation(roi_feat, pm.bbox, P['query_' + i + '_weight'], P['query_' + i + '_bias'],
                            P['key_' + i + '_weight'], P['key_' + i + '_bias'], P['pair_pos_fc1_' + i + '_weight'],
                            P['pair_pos_fc1_' + i + '_bias'], P['linear_out_' + i + '_weight'],
                            P['linear_out_' + i + '_bias'], key_index=key_index, M=M, group=group,
                            residual_relu=residual_relu, wave_length=float(position_embedding.wave_length),
                            precision=self.precision)

    def attention_module_nms_multi_head(self, roi_feat, position_mat, num_rois, dim=(1024, 1024, 1024), fc_dim=(64, 16),
                                        feat_dim=1024, group=16, index=1, return_softmax='lazy'):
        """SYM_REL_NMS:158-238 / LNMS:45-127 (same defaults; the call site passes dim=(1024, 1024, 128), feat_dim=128 -- the
        sizes that count are those of the weights).  roi_feat [num_rois, num_fg_classes, feat_dim]; position_mat is built from
        sorted boxes [num_rois, num_fg_classes, 4] (a lazy PositionMatrix over them or the boxes themselves).
        Returns (output [num_rois, num_fg_classes, dim[2]], aff_softmax [num_fg_classes*fc_dim[1], num_rois, num_rois]).
        return_softmax: 'lazy' (default) -> the output comes from the configured precision (tcgen05 under f16) and
        aff_softmax is a LazySoftmax handle; True -> both eagerly from the fp32 path; False -> (output, None)."""
        assert dim[0] == dim[1], 'Matrix multi requires the same dims!'
        assert fc_dim[1] == group, 'Check the dimensions in attention!'
        P, i = self.params, str(index)
        boxes = position_mat.bbox if isinstance(position_mat, PositionMatrix) else position_mat
        X = roi_feat.permute(1, 0, 2).contiguous()          # [C, n, feat]
        B = boxes.permute(1, 0, 2).contiguous()             # [C, n, 4]
        w = (P['nms_query_' + i + '_weight'], P['nms_query_' + i + '_bias'],
             P['nms_key_' + i + '_weight'], P['nms_key_' + i + '_bias'],
             P['nms_pair_pos_fc1_' + i + '_weight'], P['nms_pair_pos_fc1_' + i + '_bias'],
             P['nms_linear_out_' + i + '_weight'], P['nms_linear_out_' + i + '_bias'])

        def eager():
            out, sm = ops.relation(X, B, *w, group=group, precision='fp32', return_softmax=True)
            return out.permute(1, 0, 2).contiguous(), sm.permute(0, 2, 1, 3).reshape(-1, num_rois, num_rois)   # [C*H, n, n]

        if return_softmax is True:
            return eager()
        out = ops.relation(X, B, *w, group=group, precision=self.precision).permute(1, 0, 2).contiguous()
        return out, (LazySoftmax(lambda: eager()[1]) if return_softmax == 'lazy' else None)
