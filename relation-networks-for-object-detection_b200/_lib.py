"""ctypes binding of librelnet_b200.so (the C ABI declared in include/relnet_b200.h)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RELNET_LIB: measurement variants of the same sources (tools/, build.py RELNET_VARIANT); unset = the product library
LIB_PATH = os.environ.get('RELNET_LIB') or os.path.join(HERE, 'librelnet_b200.so')

c_f = C.c_float
c_i = C.c_int32
c_p = C.c_void_p
c_sz = C.c_size_t


class RelationDesc(C.Structure):
    _fields_ = [('batch', c_i), ('N', c_i), ('M', c_i), ('d', c_i), ('dq', c_i), ('dout', c_i), ('H', c_i), ('E', c_i),
                ('wave_length', c_f), ('fuse_residual_relu', c_i), ('precision', c_i)]


class LearnNmsDesc(C.Structure):
    _fields_ = [('R', c_i), ('num_classes', c_i), ('num_reg_classes', c_i), ('feat_dim', c_i), ('first_n', c_i),
                ('num_thresh', c_i), ('class_thresh', C.c_double), ('class_agnostic', c_i), ('has_means_stds', c_i),
                ('means', c_f * 4), ('stds', c_f * 4), ('nongt_dim', c_i), ('num_non_gt', c_i), ('merge_method', c_i),
                ('precision', c_i)]


class LearnNmsWeights(C.Structure):
    NAMES = ['nms_rank_weight', 'nms_rank_bias', 'roi_feat_embedding_weight', 'roi_feat_embedding_bias',
             'nms_pair_pos_fc1_1_weight', 'nms_pair_pos_fc1_1_bias', 'nms_query_1_weight', 'nms_query_1_bias',
             'nms_key_1_weight', 'nms_key_1_bias', 'nms_linear_out_1_weight', 'nms_linear_out_1_bias',
             'nms_logit_weight', 'nms_logit_bias']
    _fields_ = [(n, c_p) for n in NAMES]


class ProposalDesc(C.Structure):
    _fields_ = [('Hf', c_i), ('Wf', c_i), ('feat_stride', c_i), ('num_scales', c_i), ('num_ratios', c_i),
                ('pre_nms_top_n', c_i), ('post_nms_top_n', c_i), ('nms_thresh', c_f), ('min_size', c_f)]


class ProposalTargetDesc(C.Structure):
    _fields_ = [('N', c_i), ('G', c_i), ('num_reg_classes', c_i), ('class_agnostic', c_i), ('bg_thresh_hi', c_f),
                ('normalize', c_i), ('means', C.c_double * 4), ('stds', C.c_double * 4), ('bbox_weights', c_f * 4)]


class PsroiDesc(C.Structure):
    _fields_ = [('R', c_i), ('channels', c_i), ('H', c_i), ('W', c_i), ('spatial_scale', c_f), ('output_dim', c_i),
                ('group_size', c_i), ('pooled_size', c_i), ('part_size', c_i), ('sample_per_part', c_i),
                ('trans_std', c_f), ('no_trans', c_i), ('num_classes', c_i)]


class DeformConvDesc(C.Structure):
    _fields_ = [('B', c_i), ('C', c_i), ('H', c_i), ('W', c_i), ('Co', c_i), ('kh', c_i), ('kw', c_i), ('pad_h', c_i),
                ('pad_w', c_i), ('stride_h', c_i), ('stride_w', c_i), ('dil_h', c_i), ('dil_w', c_i), ('num_group', c_i),
                ('num_deformable_group', c_i), ('precision', c_i)]


# name -> (restype, argtypes); every symbol include/relnet_b200.h declares
SIGNATURES = {
    'rn_last_error': (C.c_char_p, []),
    'rn_version': (C.c_int, []),
    'rn_device_info': (C.c_int, [C.POINTER(C.c_int)] * 3),
    'rn_relation_workspace_bytes': (c_sz, [C.POINTER(RelationDesc)]),
    'rn_relation_fwd': (C.c_int, [C.POINTER(RelationDesc)] + [c_p] * 13 + [c_p, c_sz, c_p]),
    'rn_relation_bwd_workspace_bytes': (c_sz, [C.POINTER(RelationDesc)]),
    'rn_relation_bwd': (C.c_int, [C.POINTER(RelationDesc)] + [c_p] * 21 + [c_p, c_sz, c_p]),
    'rn_relation_bwd_masked': (C.c_int, [C.POINTER(RelationDesc)] + [c_p] * 22 + [c_p, c_sz, c_p]),
    'rn_relation_packed_fwd_f16io': (C.c_int, [C.POINTER(RelationDesc)] + [c_p] * 9 + [c_p, c_sz, c_i, c_p]),
    'rn_linear_multi_packed_bytes': (c_sz, [C.POINTER(c_i), c_i, c_i]),
    'rn_linear_multi_pack': (C.c_int, [C.POINTER(c_p), C.POINTER(c_p), C.POINTER(c_i), c_i, c_i, c_p, c_p]),
    'rn_linear_multi_packed_f16in_fwd': (C.c_int, [c_p, c_p, C.POINTER(c_p), C.POINTER(c_i), c_i, c_i, c_i, c_p, c_sz, c_p]),
    'rn_rpn_head_fwd': (C.c_int, [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    'rn_rpn_head_packed_bytes': (c_sz, [c_i, c_i]),
    'rn_rpn_head_pack': (C.c_int, [c_p] * 4 + [c_i, c_i, c_p, c_p]),
    'rn_rpn_head_workspace_bytes': (c_sz, [c_i, c_i, c_i]),
    'rn_rpn_head_packed_fwd': (C.c_int, [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_sz, c_p]),
    'rn_image_s2d_bf16': (C.c_int, [c_p, c_i, c_i, c_i, c_p, c_p]),
    'rn_maxpool3x3s2_nhwc_bf16': (C.c_int, [c_p, c_i, c_i, c_i, c_p, c_p]),
    'rn_relation_fused_enable': (C.c_int, [c_i]),
    'rn_pos_embed_fwd': (C.c_int, [c_p, c_p, c_i, c_i, c_i, c_f, c_p, c_p, c_p]),
    'rn_geometry_weight_fwd': (C.c_int, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p, c_p]),
    'rn_linear_workspace_bytes': (c_sz, [c_i, c_i, c_i, c_i]),
    'rn_linear_fwd': (C.c_int, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p, c_sz, c_p]),
    'rn_relation_packed_bytes': (c_sz, [C.POINTER(RelationDesc)]),
    'rn_relation_pack': (C.c_int, [C.POINTER(RelationDesc)] + [c_p] * 7 + [c_p]),
    'rn_relation_packed_fwd': (C.c_int, [C.POINTER(RelationDesc)] + [c_p] * 7 + [c_p, c_sz, c_p]),
    'rn_relation_packed_stages': (C.c_int, [C.POINTER(RelationDesc)] + [c_p] * 7 + [c_p, c_sz, c_i, c_p]),
    'rn_linear_packed_bytes': (c_sz, [c_i, c_i]),
    'rn_linear_pack': (C.c_int, [c_p, c_i, c_i, c_p, c_p]),
    'rn_linear_packed_fwd': (C.c_int, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_sz, c_p]),
    'rn_roi_pool_nhwc_f16_fwd': (C.c_int, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p]),
    'rn_roi_pool_nhwc_bf16in_f16_fwd': (C.c_int, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p]),
    'rn_linear_pack_chw_to_hwc': (C.c_int, [c_p, c_i, c_i, c_i, c_p, c_p]),
    'rn_linear_packed_f16in_fwd': (C.c_int, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_sz, c_p]),
    'rn_learn_nms_workspace_bytes': (c_sz, [C.POINTER(LearnNmsDesc)]),
    'rn_learn_nms_fwd': (C.c_int, [C.POINTER(LearnNmsDesc)] + [c_p] * 5 + [C.POINTER(LearnNmsWeights), c_p] + [c_p] * 4 +
                         [c_p, c_sz, c_p]),
    'rn_learn_nms_packed_bytes': (c_sz, [C.POINTER(LearnNmsDesc)]),
    'rn_learn_nms_pack': (C.c_int, [C.POINTER(LearnNmsDesc), C.POINTER(LearnNmsWeights), c_p, c_p, c_sz, c_p]),
    'rn_learn_nms_packed_fwd': (C.c_int, [C.POINTER(LearnNmsDesc)] + [c_p] * 7 + [C.POINTER(LearnNmsWeights), c_p, c_p] + [c_p] * 4 +
                                [c_p, c_sz, c_p]),
    'rn_learn_nms_bwd_workspace_bytes': (c_sz, [C.POINTER(LearnNmsDesc)]),
    'rn_learn_nms_bwd': (C.c_int, [C.POINTER(LearnNmsDesc)] + [c_p] * 5 + [C.POINTER(LearnNmsWeights), c_p, c_p, C.POINTER(LearnNmsWeights), c_p, c_p] +
                         [c_p, c_sz, c_p]),
    'rn_nms_loss': (C.c_int, [c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_f, c_p, c_p, c_p, c_p]),
    'rn_box_annotator_ohem': (C.c_int, [c_p] * 5 + [c_i] * 4 + [c_p] * 3 + [c_p]),
    'rn_nms_multi_target_fwd': (C.c_int, [c_p, c_p, c_p, c_i, c_i, c_i, C.POINTER(C.c_double), c_i, c_p, c_p]),
    'rn_proposal_workspace_bytes': (c_sz, [C.POINTER(ProposalDesc)]),
    'rn_proposal_fwd': (C.c_int, [C.POINTER(ProposalDesc), C.POINTER(c_f), C.POINTER(c_f)] + [c_p] * 6 + [c_p, c_sz, c_p]),
    'rn_nms_workspace_bytes': (c_sz, [c_i]),
    'rn_nms': (C.c_int, [c_p, c_i, c_i, c_f, c_i, c_p, c_p, c_p, c_sz, c_p]),
    'rn_bbox_overlaps': (C.c_int, [c_p, c_p, c_i, c_i, c_p, c_p]),
    'rn_proposal_target_fwd': (C.c_int, [C.POINTER(ProposalTargetDesc)] + [c_p] * 6 + [c_p]),
    'rn_roi_pool_fwd': (C.c_int, [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p]),
    'rn_roi_pool_bwd': (C.c_int, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
    'rn_deform_psroi_pool_bwd': (C.c_int, [C.POINTER(PsroiDesc), c_i] + [c_p] * 7 + [c_p]),
    'rn_deform_conv_bwd': (C.c_int, [C.POINTER(DeformConvDesc)] + [c_p] * 4 + [c_i] + [c_p] * 4 + [c_p, c_sz, c_p]),
    'rn_deform_psroi_pool_fwd': (C.c_int, [C.POINTER(PsroiDesc)] + [c_p] * 5 + [c_p]),
    'rn_deform_psroi_pool_nhwc_fwd': (C.c_int, [C.POINTER(PsroiDesc), c_p, c_i] + [c_p] * 4 + [c_p]),
    'rn_deform_conv_workspace_bytes': (c_sz, [C.POINTER(DeformConvDesc)]),
    'rn_deform_conv_fwd': (C.c_int, [C.POINTER(DeformConvDesc)] + [c_p] * 5 + [c_p, c_sz, c_p]),
    'rn_deform_im2col': (C.c_int, [C.POINTER(DeformConvDesc), c_p, c_p, c_p, c_p]),
    'rn_deform_conv_packed_bytes': (c_sz, [C.POINTER(DeformConvDesc)]),
    'rn_deform_conv_pack': (C.c_int, [C.POINTER(DeformConvDesc), c_p, c_p, c_p]),
    'rn_deform_conv_nhwc_fwd': (C.c_int, [C.POINTER(DeformConvDesc), c_p, c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_sz, c_p]),
    'rn_umma_selftest': (C.c_int, [c_p] * 6 + [c_p]),
    'rn_gemm_tf32': (C.c_int, [c_i] * 5 + [c_f, c_p, c_i, C.c_int64, C.c_int64, c_p, c_i, C.c_int64, C.c_int64, c_f, c_p, c_i,
                                C.c_int64, C.c_int64, c_i, c_i, c_p]),
}

_lib = None


class RelnetError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RelnetError('%s is missing: run `python relation-networks-for-object-detection_b200/build.py` '
                              '(there is no CPU fallback)' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status, what):
    if status != 0:
        raise RelnetError('%s failed (%d): %s' % (what, status, lib().rn_last_error().decode()))
