"""One backward call of the relation module / the learn-NMS head at the headline sizes (for an ncu launch list):
    python tools/bwd_one.py relation|learn_nms [fp32|f16]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops, synth
from relnet_b200.pipeline import init_head_params, NMS_NAMES

what = sys.argv[1] if len(sys.argv) > 1 else 'relation'
prec = sys.argv[2] if len(sys.argv) > 2 else 'f16'
dev = torch.device('cuda:0')
torch.manual_seed(0)
if what == 'relation':
    c = synth.make_relation_case(1, 300, 1024, 16)
    t = [torch.from_numpy(c[k]).to(dev) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
    dO = torch.randn(300, 1024, device=dev)
    fn = lambda: ops.relation_backward(dO, *t, group=16, residual_relu=True, precision=prec)
else:
    P = init_head_params(0, dev)
    R, C = 300, 80
    rng = np.random.default_rng(0)
    boxes = synth.make_boxes(rng, R)
    rois = torch.from_numpy(np.hstack([np.zeros((R, 1), np.float32), boxes]).astype(np.float32)).to(dev)
    cls = torch.randn(R, C + 1, device=dev); bbox = torch.randn(R, 8, device=dev) * 0.1
    feat = torch.randn(R, 1024, device=dev).clamp_min(0)
    info = torch.tensor([600.0, 1000.0, 1.0], device=dev)
    dM = torch.randn(100, C, 5, device=dev)
    W = {k: P[k] for k in NMS_NAMES}
    fn = lambda: ops.learn_nms_backward(dM, cls, bbox, rois, info, feat, W, first_n=100, class_thresh=0.0, means=(0, 0, 0, 0),
                                        stds=(0.1, 0.1, 0.2, 0.2), nongt_dim=R, precision=prec)
for _ in range(2):
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
