"""Training-side kernels at the headline sizes: microseconds per call (CUDA events, 256 MB L2 flush between launches,
median of 10).  relation module fwd (fp32 and tcgen05) vs bwd, learn-NMS head fwd vs bwd, ROI / deformable backward ops.
One JSON line per op on stdout (redirect into gpurun_out/)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops
from oracle import relation_np as R, learn_nms_np as LN      # case generators only (a dev tool, not the product)

dev = torch.device('cuda:0')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    evs = []
    for i in range(reps):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) * 1e3 for x, y in evs)
    return round(ts[len(ts) // 2], 1)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


out = []
c = R.make_relation_case(1, 300, 1024, 16)
t = [T(c[k]) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
dO = torch.randn(300, 1024, device=dev)
out.append(dict(op='relation N=300 d=1024 H=16', fwd_f16_us=timeit(lambda: ops.relation(*t, group=16, residual_relu=True, precision='f16')),
                fwd_fp32_us=timeit(lambda: ops.relation(*t, group=16, residual_relu=True, precision='fp32')),
                bwd_fp32_us=timeit(lambda: ops.relation_backward(dO, *t, group=16, residual_relu=True, precision='fp32')),
                bwd_tf32_tc_us=timeit(lambda: ops.relation_backward(dO, *t, group=16, residual_relu=True, precision='f16'))))
l = LN.make_learn_nms_case(2, R=300, C=80, d=1024)
W = {k: T(v) for k, v in l['P'].items()}
la = (T(l['cls_score']), T(l['bbox_pred']), T(l['rois']), T(l['im_info']), T(l['feat']), W)
dM = torch.randn(100, 80, 5, device=dev)
out.append(dict(op='learn_nms R=300 C=80 n=100 (all classes)',
                fwd_f16_us=timeit(lambda: ops.learn_nms(*la, class_thresh=0.0, precision='f16')),
                fwd_fp32_us=timeit(lambda: ops.learn_nms(*la, class_thresh=0.0, precision='fp32')),
                bwd_fp32_us=timeit(lambda: ops.learn_nms_backward(dM, *la, class_thresh=0.0, precision='fp32')),
                bwd_tf32_tc_us=timeit(lambda: ops.learn_nms_backward(dM, *la, class_thresh=0.0, precision='f16'))))
out.append(dict(op='learn_nms R=300 C=80 n=100, class_thresh 0.01 (synthetic scores: ~12 of 80 classes survive the pruning of LNMS:298-303)',
                fwd_f16_us=timeit(lambda: ops.learn_nms(*la, class_thresh=0.01, precision='f16')),
                fwd_fp32_us=timeit(lambda: ops.learn_nms(*la, class_thresh=0.01, precision='fp32'))))
rng = np.random.RandomState(0)
data = T(rng.randn(1, 256, 38, 63).astype(np.float32))
x1 = rng.uniform(0, 800, 300); y1 = rng.uniform(0, 450, 300)
rois = T(np.stack([np.zeros(300), x1, y1, np.minimum(x1 + rng.uniform(8, 500, 300), 999),
                   np.minimum(y1 + rng.uniform(8, 400, 300), 599)], 1).astype(np.float32))
o, arg = ops.roi_pool(data, rois, (7, 7), 0.0625, return_argmax=True)
g = torch.randn_like(o)
out.append(dict(op='roi_pool R=300 C=256', fwd_us=timeit(lambda: ops.roi_pool(data, rois, (7, 7), 0.0625, return_argmax=True)),
                bwd_us=timeit(lambda: ops.roi_pool_backward(g, arg, rois, data.shape))))
trans = torch.randn(300, 2, 7, 7, device=dev)
kw = dict(spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7, sample_per_part=4, trans_std=0.1)
o, cnt = ops.deform_psroi_pool(data, rois, trans, return_count=True, **kw)
out.append(dict(op='deform_psroi_pool R=300 C=256 (with trans)', fwd_us=timeit(lambda: ops.deform_psroi_pool(data, rois, trans, **kw)),
                bwd_us=timeit(lambda: ops.deform_psroi_pool_backward(g, cnt, data, rois, trans, **kw))))
d5 = torch.randn(1, 512, 38, 63, device=dev); off = torch.randn(1, 72, 38, 63, device=dev) * 0.5
w5 = torch.randn(512, 512, 3, 3, device=dev) * 0.02; go = torch.randn(1, 512, 38, 63, device=dev)
out.append(dict(op='deform_conv 512->512 3x3 dil2 38x63', fwd_fp32_us=timeit(lambda: ops.deform_conv(d5, off, w5), 5),
                fwd_f16_us=timeit(lambda: ops.deform_conv(d5, off, w5, precision='f16'), 5),
                bwd_fp32_us=timeit(lambda: ops.deform_conv_backward(go, d5, off, w5), 5)))
for r in out:
    print(json.dumps(r), flush=True)
