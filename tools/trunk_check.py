"""Trunk (library) configuration check: fused cuDNN conv+bias+relu path vs the plain module path -- same outputs within
bf16 rounding, and the time of each (CUDA events, median of 10, CUDA-graph replay like bench.py)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import trunk as TR

dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
t = TR.make_trunk(dev)
img = torch.randn(1, 3, 600, 1000, device=dev) * 50          # fp32 image, as the pipeline feeds it
res = {}
outs = {}
for fused in (False, True):
    TR.FUSED = fused
    for _ in range(3):
        o = t(img)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        t(img)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            o = t(img)
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res['fused' if fused else 'plain'] = round(sorted(ts)[5], 4)
    outs[fused] = [x.float().clone() for x in o]
err = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs[True], outs[False])]
print(json.dumps(dict(trunk_ms=res, rel_diff_fused_vs_plain=dict(rpn_prob=err[0], rpn_bbox=err[1], conv_new_1=err[2]))))
