"""DeformablePSROIPooling forward at the Deformable Faster-RCNN size (R=300, 256 ch, 38x63): NCHW fp32 / NHWC fp32 / NHWC bf16,
CUDA events, L2 flush, median; algorithmic bytes = 15.05 MB output + feature map."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops, synth
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
R = 300
data = torch.from_numpy(rng.standard_normal((1, 256, 38, 63)).astype(np.float32)).to(dev)
boxes = synth.make_boxes(rng, R)
rois = torch.from_numpy(np.hstack([np.zeros((R, 1), np.float32), boxes]).astype(np.float32)).to(dev)
trans = torch.from_numpy((rng.standard_normal((R, 2, 7, 7)) * 0.1).astype(np.float32)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def t_us(fn, reps=15):
    for _ in range(3):
        fn()
    ts = []
    for i in range(reps):
        flush.fill_(i & 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


cl = data.contiguous(memory_format=torch.channels_last)
bf = cl.to(torch.bfloat16)
for name, d in (('nchw_fp32', data), ('nhwc_fp32', cl), ('nhwc_bf16', bf)):
    for tr in (None, trans):
        us = t_us(lambda: ops.deform_psroi_pool(d, rois, tr, output_dim=256, trans_std=0.1 if tr is not None else 0.0))
        byts = R * 256 * 49 * 4 + d.numel() * d.element_size()
        print(json.dumps(dict(kernel='psroi_fwd', layout=name, trans=tr is not None, us=round(us, 2), algorithmic_mb=round(byts / 1e6, 2),
                              gbs=round(byts / us / 1e3, 1), frac_of_measured_hbm=round(byts / us / 1e3 / peak, 4))), flush=True)
