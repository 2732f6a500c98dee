"""Run the relation module a few times at one shape (for ncu):  python tools/fused_one.py N d H [fused=1] [iters=4]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops, synth

N, d, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fused = int(sys.argv[4]) if len(sys.argv) > 4 else 1
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 4
c = synth.make_relation_case(N * 31 + d + H, N, d, H)
t = [torch.from_numpy(c[k]).cuda() for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
ops.relation_fused_enable(fused)
for _ in range(iters):
    out = ops.relation(*t, group=H, residual_relu=True, precision='f16')
torch.cuda.synchronize()
print(float(out.abs().sum()))
