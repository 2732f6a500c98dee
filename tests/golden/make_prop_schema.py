"""Golden for the operator boundary: what the reference's own CustomOpProp classes answer for list_arguments / list_outputs /
infer_shape under the string kwargs MXNet hands them.  Runs the reference's Python from where it lies (oracle/refexec.py), in
this container only; writes tests/golden/prop_schema.json.

    python tests/golden/make_prop_schema.py
"""
import json
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refexec   # noqa: E402

CFG = {'TRAIN': {'BBOX_NORMALIZATION_PRECOMPUTED': True, 'BBOX_MEANS': [0.0, 0.0, 0.0, 0.0], 'BBOX_STDS': [0.1, 0.1, 0.2, 0.2],
                 'FG_THRESH': 0.5, 'BG_THRESH_HI': 0.5, 'BG_THRESH_LO': 0.0, 'BBOX_WEIGHTS': [1.0, 1.0, 1.0, 1.0]},
       'CLASS_AGNOSTIC': False}

# (registered name, reference module attr, class, kwargs as MXNet passes them: all strings, in_shape)
CASES = [
    ('proposal', 'proposal', 'ProposalProp', {}, [[1, 24, 38, 63], [1, 48, 38, 63], [1, 3]]),
    ('proposal', 'proposal', 'ProposalProp',
     {'feat_stride': '16', 'scales': '(4, 8, 16, 32)', 'ratios': '(0.5, 1, 2)', 'output_score': 'True',
      'rpn_pre_nms_top_n': '12000', 'rpn_post_nms_top_n': '2000', 'threshold': '0.7', 'rpn_min_size': '0'},
     [[1, 24, 38, 63], [1, 48, 38, 63], [1, 3]]),
    ('proposal_target', 'proposal_target', 'ProposalTargetProp',
     {'num_classes': '81', 'batch_images': '1', 'batch_rois': '-1', 'cfg': '<pickle>', 'fg_fraction': '0.25'}, [[300, 5], [7, 5]]),
    ('proposal_target', 'proposal_target', 'ProposalTargetProp',
     {'num_classes': '2', 'batch_images': '1', 'batch_rois': '128', 'cfg': '<pickle>'}, [[2000, 5], [3, 5]]),
    ('learn_nms', 'learn_nms', 'LearnNmsProp',                 # SYM_REL_NMS:530-534: str() of the numpy arrays, has_non_gt_index=False
     {'num_fg_classes': '80', 'bbox_means': '[0. 0. 0. 0.]', 'bbox_stds': '[0.1 0.1 0.2 0.2]', 'first_n': '100',
      'class_agnostic': 'False', 'num_thresh': '5', 'class_thresh': '0.01', 'nongt_dim': '300', 'has_non_gt_index': 'False'},
     [[300, 81], [300, 324], [300, 5], [1, 3], [300, 1024]]),
    ('learn_nms', 'learn_nms', 'LearnNmsProp',                 # FPN form: non_gt_index appended, class-agnostic regression
     {'num_fg_classes': '80', 'bbox_means': 'None', 'bbox_stds': 'None', 'first_n': '150',
      'class_agnostic': 'True', 'num_thresh': '5', 'class_thresh': '0.01', 'nongt_dim': 'None', 'has_non_gt_index': 'True'},
     [[1000, 81], [1000, 8], [1000, 5], [1, 3], [1000, 1024]]),
    ('nms_multi_target', 'nms_multi_target', 'NmsMultiTargetProp', {'target_thresh': '[0.5 0.6 0.7 0.8 0.9]'},
     [[100, 80, 4], [1, 7, 5], [100, 80]]),
    ('BoxAnnotatorOHEM', 'box_annotator_ohem', 'BoxAnnotatorOHEMProp',
     {'num_classes': '81', 'num_reg_classes': '81', 'roi_per_img': '128'}, [[300, 81], [300, 324], [300], [300, 324], [300, 324]]),
]


def tolist(x):
    if isinstance(x, (list, tuple)):
        return [tolist(v) for v in x]
    if isinstance(x, (bool, str, bytes, type(None))):
        return x
    if isinstance(x, float):
        return x if x != int(x) else int(x)
    return int(x) if hasattr(x, '__int__') else x


def run_case(ns_or_mod, cls_name, kwargs, in_shape):
    kw = dict(kwargs)
    if kw.get('cfg') == '<pickle>':
        kw['cfg'] = pickle.dumps(CFG)
    prop = getattr(ns_or_mod, cls_name)(**kw)
    rec = {'list_arguments': list(prop.list_arguments()), 'list_outputs': list(prop.list_outputs()),
           'need_top_grad': bool(getattr(prop, 'need_top_grad', getattr(prop, 'need_top_grad_', False)))}
    rec['attrs'] = {k: tolist(v.tolist() if hasattr(v, 'tolist') else v) for k, v in sorted(vars(prop).items())
                    if isinstance(v, (int, float, bool, str, type(None))) or hasattr(v, 'tolist')}
    if in_shape is not None:
        try:
            res = prop.infer_shape([list(s) for s in in_shape])
            rec['infer_shape'] = tolist(res[:2])
        except Exception as e:                     # the reference's own answer may be "not implemented": record that too
            rec['infer_shape_error'] = type(e).__name__
    return rec


def main():
    assert refexec.available(), 'needs /root/reference (this container only)'
    ns = refexec.load_reference()
    out = []
    for reg, mod, cls, kwargs, in_shape in CASES:
        rec = run_case(getattr(ns, mod), cls, kwargs, in_shape)
        out.append({'op_type': reg, 'class': cls, 'kwargs': kwargs, 'in_shape': in_shape, 'reference': rec,
                    'registered_in_reference': reg in ns.mxshim.REGISTRY})
    # call surfaces of the symbol-class methods of the path (names, order, defaults) in the three symbol files
    import inspect
    sigs = {}
    for modname, cls in [('sym_rel', 'resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16'),
                         ('sym_rel_nms', 'resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16_learn_nms'),
                         ('sym_fpn_rel_nms', 'resnet_v1_101_rcnn_fpn_attention_1024_pairwise_position_multi_head_16_learn_nms')]:
        k = getattr(getattr(ns, modname), cls)
        for fn in ('extract_position_embedding', 'extract_position_matrix', 'attention_module_multi_head',
                   'attention_module_nms_multi_head'):
            if hasattr(k, fn):
                ps = inspect.signature(getattr(k, fn)).parameters
                sigs.setdefault(modname, {})[fn] = [[n, None if q.default is inspect.Parameter.empty else tolist(q.default),
                                                     q.default is not inspect.Parameter.empty] for n, q in ps.items()]
    # dmlc::Parameter fields (name, declared default) of the two C++ operators the path uses
    import re
    cxx = {}
    for op, rel in [('DeformableConvolution', 'relation_rcnn/operator_cxx/deformable_convolution-inl.h'),
                    ('DeformablePSROIPooling', 'relation_rcnn/operator_cxx/deformable_psroi_pooling-inl.h')]:
        src = open(os.path.join(refexec.REF, rel)).read()
        fields = []
        for m in re.finditer(r'DMLC_DECLARE_FIELD\((\w+)\)(.*?);', src, re.S):
            d = re.search(r'\.set_default\(([^()]*(?:\([^()]*\))?[^()]*)\)', m.group(2))
            fields.append([m.group(1), d.group(1).strip() if d else None])
        cxx[op] = fields
    with open(os.path.join(HERE, 'prop_schema.json'), 'w') as f:
        json.dump({'cfg': CFG, 'cases': out, 'symbol_signatures': sigs, 'cxx_params': cxx}, f, indent=1, sort_keys=True)
    print('wrote %d cases' % len(out))


if __name__ == '__main__':
    main()
