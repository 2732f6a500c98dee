"""Execute the reference's own Python, in this container, to pin the oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference (/root/reference) is Python 2.7 on MXNet
1.1.0; neither exists here.  This module loads reference *source files from where they lie* (never copied into
the repo), applies a tiny in-memory py2->py3 source transform (``print`` statements, ``xrange``, ``cPickle``,
``np.float``), and executes them with stand-in modules:

  mxnet                -> oracle/mxshim.py (numpy float32 restatement of the ~40 MXNet ops used)
  easydict             -> attribute dict
  nms.nms              -> gpu_nms_wrapper bound to oracle.proposal_np.nms_gpu_semantics (the reference's GPU NMS
                          needs a GPU + cython; its semantics are restated from lib/nms/nms_kernel.cu and are
                          pinned separately on the GPU box against oracle/_ref/libref_gpu_nms.so)
  bbox.bbox            -> bbox_overlaps_cython bound to oracle.proposal_np.bbox_overlaps (cython not built here)

It exists so that ``tests/golden/make_golden.py`` can generate golden vectors from the reference's own code
paths; it only runs where /root/reference exists (this container), never on the GPU box.
"""
import os
import re
import sys
import types
import pickle
import numpy as np

REF = os.environ.get('RELNET_REFERENCE', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF, 'relation_rcnn'))


_PRINT = re.compile(r'^(\s*)print\s+(?!\()(.*)$')
_PRINT_PAREN = re.compile(r'^(\s*)print\s+\((.*)$')


def py2to3(src):
    out = []
    for line in src.split('\n'):
        m = _PRINT.match(line)
        if m:
            body, comment = m.group(2), ''
            mc = re.search(r'\s+#[^\'"]*$', body)
            if mc:
                body, comment = body[:mc.start()], body[mc.start():]
            line = '%sprint(%s)%s' % (m.group(1), body.rstrip().rstrip(','), comment)
        out.append(line)
    src = '\n'.join(out)
    src = src.replace('xrange(', 'range(')
    src = re.sub(r'\bimport cPickle\b', 'import pickle as cPickle', src)
    src = re.sub(r'\bnp\.float\b(?!\d|_)', 'np.float64', src)
    src = src.replace('from distutils.util import strtobool', 'strtobool = lambda s: s in ("True", "true", "1")')
    return src


def _load(modname, relpath, extra_globals=None):
    """Load reference file ``relpath`` as module ``modname`` (registered in sys.modules)."""
    path = os.path.join(REF, relpath)
    with open(path) as f:
        src = py2to3(f.read())
    mod = types.ModuleType(modname)
    mod.__file__ = path
    if extra_globals:
        mod.__dict__.update(extra_globals)
    sys.modules[modname] = mod
    exec(compile(src, path, 'exec'), mod.__dict__)
    return mod


class EasyDict(dict):
    def __init__(self, d=None, **kw):
        super(EasyDict, self).__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super(EasyDict, self).__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


_loaded = {}


def load_reference():
    """Install stubs and load the reference modules of the hot path.  Returns a namespace of modules."""
    if _loaded:
        return _loaded['ns']
    from . import mxshim, proposal_np
    mx, nd, operator = mxshim.build_module()
    sys.modules['mxnet'] = mx
    sys.modules['mxnet.nd'] = nd
    sys.modules['mxnet.operator'] = operator
    ed = types.ModuleType('easydict')
    ed.EasyDict = EasyDict
    sys.modules['easydict'] = ed

    # --- lib/bbox (cython bbox.pyx is replaced by the restated fp64 IoU; see module docstring)
    bbox_pkg = types.ModuleType('bbox'); bbox_pkg.__path__ = []
    sys.modules['bbox'] = bbox_pkg
    bbox_cy = types.ModuleType('bbox.bbox')
    bbox_cy.bbox_overlaps_cython = proposal_np.bbox_overlaps
    sys.modules['bbox.bbox'] = bbox_cy
    bbox_pkg.bbox_overlaps_cython = proposal_np.bbox_overlaps
    bt = _load('bbox.bbox_transform', 'lib/bbox/bbox_transform.py',
               {'bbox_overlaps_cython': proposal_np.bbox_overlaps})
    sys.modules['bbox_transform'] = bt           # bbox_regression.py does a py2 implicit-relative import
    br = _load('bbox.bbox_regression', 'lib/bbox/bbox_regression.py')

    # --- lib/rpn, lib/nms
    rpn_pkg = types.ModuleType('rpn'); rpn_pkg.__path__ = []
    sys.modules['rpn'] = rpn_pkg
    ga = _load('rpn.generate_anchor', 'lib/rpn/generate_anchor.py')
    nms_pkg = types.ModuleType('nms'); nms_pkg.__path__ = []
    sys.modules['nms'] = nms_pkg
    nms_mod = types.ModuleType('nms.nms')

    def gpu_nms_wrapper(thresh, device_id):
        def _nms(dets):
            return proposal_np.gpu_nms(dets, thresh)
        return _nms
    nms_mod.gpu_nms_wrapper = gpu_nms_wrapper
    nms_mod.cpu_nms_wrapper = gpu_nms_wrapper
    nms_mod.py_nms_wrapper = gpu_nms_wrapper
    sys.modules['nms.nms'] = nms_mod

    # --- utils (only what core/rcnn.py imports at module level)
    utils_pkg = types.ModuleType('utils'); utils_pkg.__path__ = []
    sys.modules['utils'] = utils_pkg
    uimg = types.ModuleType('utils.image')
    uimg.get_image = uimg.tensor_vstack = None
    sys.modules['utils.image'] = uimg
    usym = _load('utils.symbol', 'lib/utils/symbol.py')

    core_pkg = types.ModuleType('core'); core_pkg.__path__ = []
    sys.modules['core'] = core_pkg
    rcnn = _load('core.rcnn', 'relation_rcnn/core/rcnn.py')

    op_pkg = types.ModuleType('operator_py'); op_pkg.__path__ = []
    sys.modules['operator_py'] = op_pkg
    proposal = _load('operator_py.proposal', 'relation_rcnn/operator_py/proposal.py')
    proposal_target = _load('operator_py.proposal_target', 'relation_rcnn/operator_py/proposal_target.py')
    learn_nms = _load('operator_py.learn_nms', 'relation_rcnn/operator_py/learn_nms.py')
    nms_multi_target = _load('operator_py.nms_multi_target', 'relation_rcnn/operator_py/nms_multi_target.py')
    ohem = _load('operator_py.box_annotator_ohem', 'relation_rcnn/operator_py/box_annotator_ohem.py')

    S = 'relation_rcnn/symbols/'
    base = _load('resnet_v1_101_rcnn_base', S + 'resnet_v1_101_rcnn_base.py')
    nms_base = _load('resnet_v1_101_rcnn_learn_nms_base', S + 'resnet_v1_101_rcnn_learn_nms_base.py')
    sym_rel = _load('sym_rel', S + 'resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16.py')
    sym_rel_nms = _load('sym_rel_nms',
                        S + 'resnet_v1_101_rcnn_attention_1024_pairwise_position_multi_head_16_learn_nms.py')
    sym_fpn_rel_nms = _load('sym_fpn_rel_nms',
                            S + 'resnet_v1_101_rcnn_fpn_attention_1024_pairwise_position_multi_head_16_learn_nms.py')

    ns = types.SimpleNamespace(
        mx=mx, nd=nd, mxshim=mxshim, bbox_transform=bt, bbox_regression=br, generate_anchor=ga, rcnn=rcnn,
        proposal=proposal, proposal_target=proposal_target, learn_nms=learn_nms,
        nms_multi_target=nms_multi_target, box_annotator_ohem=ohem, sym_rel=sym_rel, sym_rel_nms=sym_rel_nms,
        sym_fpn_rel_nms=sym_fpn_rel_nms, nms_base=nms_base, EasyDict=EasyDict)
    _loaded['ns'] = ns
    return ns
