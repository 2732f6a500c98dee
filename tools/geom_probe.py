"""Geometry stage of the relation module (stage_mask = 2) at N = M in {1000, 3000}: microseconds, L2 flushed."""
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import __graft_entry__ as e; e.build()
import relnet_b200
from relnet_b200 import ops
from oracle import relation_np as R
dev = torch.device('cuda:0')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    evs = []
    for i in range(reps):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) * 1e3 for x, y in evs)
    return round(ts[len(ts) // 2], 1)
for N in (1000, 3000):
    c = R.make_relation_case(N, N, 1024, 16)
    t = [torch.from_numpy(c[k]).to(dev) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
    ops.relation(*t, group=16, residual_relu=True, precision='f16')
    print('geometry stage us at N = M =', N, timeit(lambda: ops.relation(*t, group=16, residual_relu=True, precision='f16', stage_mask=2)))
