"""Per-launch cost of the head's small kernels when replayed back to back from a CUDA graph (no host launch overhead, warm L2):
20 identical calls captured, time / 20.  Separates the fixed per-kernel cost from the work."""
import json
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


def graph_time(fn, n=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / n)
    return round(sorted(ts)[5], 2)


out = {}
for M, K, N in ((300, 1024, 256), (300, 1024, 1024), (300, 1024, 3072), (300, 128, 3072)):
    x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.03; b = torch.zeros(N, device=dev)
    x16 = x.half()
    out['gemm %dx%dx%d' % (M, N, K)] = graph_time(lambda: ops.linear(x, W, b, precision='f16', x_f16=x16))
c = R.make_relation_case(1, 300, 1024, 16)
t = [torch.from_numpy(c[k]).to(dev) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
ops.relation(*t, group=16, residual_relu=True, precision='f16')
for name, mask in (('relation proj (cast+gemm)', 1), ('relation geometry', 2), ('relation attention (tile+combine)', 4)):
    out[name] = graph_time(lambda: ops.relation(*t, group=16, residual_relu=True, precision='f16', stage_mask=mask))
e = torch.empty(1 << 20, device=dev)
out['torch elementwise 4 MB (reference point)'] = graph_time(lambda: e.add_(1.0))
print(json.dumps(out))
