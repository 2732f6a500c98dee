"""Driver for an ncu capture of the geometry kernel alone (N = M = 300, H = 16): one full f16 module call, then the
geometry stage five times with an L2 flush in between."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops
from oracle import relation_np as R

dev = torch.device('cuda:0')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
c = R.make_relation_case(1, N, 1024, 16)
t = [torch.from_numpy(c[k]).to(dev) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ops.relation(*t, group=16, residual_relu=True, precision='f16')
for i in range(5):
    flush.fill_(i)
    ops.relation(*t, group=16, residual_relu=True, precision='f16', stage_mask=2)
torch.cuda.synchronize()
