"""One profiled step for ncu (use with --profile-from-start off): warm up, then cudaProfilerStart / one full step /
cudaProfilerStop.  --what step|hot|relation selects trunk+hot path, hot path only, or one relation module at
N=M=--n, d=1024, H=16."""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument('--what', default='step')
ap.add_argument('--n', type=int, default=300)
ap.add_argument('--reps', type=int, default=1)
a = ap.parse_args()

import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops
from relnet_b200.pipeline import RelationHead, init_head_params
from relnet_b200.trunk import make_trunk
from bench import make_inputs

dev = torch.device('cuda:0')
if a.what in ('step', 'hot'):
    trunk = make_trunk(dev, torch.bfloat16)
    head = RelationHead(init_head_params(0, dev))
    image, im_info = make_inputs()
    image = image.to(dev); im_info = im_info.to(dev)          # fp32 image: the stem converts in its own kernel
    tout = trunk(image)
    fn = (lambda: head.forward(*trunk(image), im_info)) if a.what == 'step' else (lambda: head.forward(*tout, im_info))
else:
    from oracle import relation_np as R
    c = R.make_relation_case(2, a.n, 1024, 16)
    t = [torch.from_numpy(c[k]).to(dev) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
    fn = lambda: ops.relation(*t, group=16, residual_relu=True, precision='f16')
for _ in range(3):
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.reps):
    fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
