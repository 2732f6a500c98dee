"""Reference-facing call surfaces, bound to the C ABI.

  symbols.py    the symbol-class methods of relation_rcnn/symbols/*attention*.py: extract_position_matrix,
                extract_position_embedding, attention_module_multi_head (Faster/DCN and FPN signatures),
                attention_module_nms_multi_head -- same names, argument order and asserts; weights addressed by name
  operators.py  the Python CustomOp protocol (list_arguments / list_outputs / infer_shape / create_operator /
                forward(is_train, req, in_data, out_data, aux)) for 'proposal', 'proposal_target', 'learn_nms',
                'nms_multi_target', 'BoxAnnotatorOHEM', with the
                reference's string kwargs, plus gpu_nms / bbox_overlaps_cython and the two operator_cxx ops as functions

Tensors are torch CUDA tensors where the reference has NDArrays / symbols.
"""
from .symbols import RelationSymbols                                   # noqa: F401
from .operators import (REGISTRY, ProposalProp, ProposalOperator, ProposalTargetProp, ProposalTargetOperator,   # noqa: F401
                        LearnNmsProp, LearnNmsOperator, NmsMultiTargetProp, NmsMultiTargetOp, BoxAnnotatorOHEMProp,
                        BoxAnnotatorOHEMOperator, Custom, gpu_nms, bbox_overlaps_cython,
                        DeformableConvolution, DeformablePSROIPooling, ROIPooling)
