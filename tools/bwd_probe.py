"""Localise differences between the two contraction engines of rn_relation_bwd (cuBLAS fp32 vs tcgen05 tf32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops, synth

N, d, H, M = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (70, 256, 4, 50))]
res = int(sys.argv[5]) if len(sys.argv) > 5 else 1
c = synth.make_relation_case(11, N, d, H, M=M if M != N else None)
t = [torch.from_numpy(c[k]).cuda() for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
torch.manual_seed(0)
dO = torch.randn(N, d, device='cuda')
g32 = ops.relation_backward(dO, *t, M=M, group=H, residual_relu=bool(res), precision='fp32')
g32 = {k: v.clone() for k, v in g32.items()}
g16 = ops.relation_backward(dO, *t, M=M, group=H, residual_relu=bool(res), precision='f16')
torch.cuda.synchronize()
for k in g32:
    a, b = g16[k].double(), g32[k].double()
    e = (a - b).abs()
    print('%-5s rel %.2e  shape %s' % (k, float(e.max() / b.abs().max().clamp_min(1e-30)), tuple(a.shape)))
e = (g16['X'].double() - g32['X'].double()).abs()
thr = 0.01 * float(g32['X'].abs().max())
bad = e > thr
print('dX: %d of %d elements off by > 1%% of max; rows with errors: %s' % (int(bad.sum()), bad.numel(), torch.nonzero(bad.any(1)).flatten().tolist()[:40]))
print('cols with errors (count per 64-col group):', [int(bad[:, i:i + 64].sum()) for i in range(0, d, 64)])
eb = (g16['bout'].double() - g32['bout'].double()).abs()
print('dbout worst cols:', torch.topk(eb, 8).indices.tolist())
# forward in the three modes
o32 = ops.relation(*t, M=M, group=H, residual_relu=bool(res), precision='fp32')
otf = ops.relation(*t, M=M, group=H, residual_relu=bool(res), precision='tf32')
print('forward tf32 vs fp32 rel %.2e; sign flips %d' % (float((otf - o32).abs().max() / o32.abs().max()), int(((otf > 0) != (o32 > 0)).sum())))
