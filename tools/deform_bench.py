"""Deformable conv (res5 size 512->512, 38x63, 4 deformable groups): NCHW fp32 entry (f16 precision) vs the channels_last bf16
fast path; CUDA events, L2 flush, median.  11.30 GFLOP per layer."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as entry
entry.build()
import relnet_b200
from relnet_b200 import ops
dev = torch.device('cuda:0')
rng = np.random.default_rng(0)
C, H, W, Co = 512, 38, 63, 512
data = torch.from_numpy(rng.standard_normal((1, C, H, W)).astype(np.float32)).to(dev)
off = torch.from_numpy((rng.standard_normal((1, 72, H, W)) * 2.0).astype(np.float32)).to(dev)
wgt = torch.from_numpy((rng.standard_normal((Co, C, 3, 3)) * 0.02).astype(np.float32)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
peak = json.load(open(pk))['bf16_tflops'] if os.path.exists(pk) else 1590.0


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


gf = 2.0 * Co * C * 9 * H * W / 1e9
bf = data.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
for name, fn in (('nchw_fp32_entry_f16', lambda: ops.deform_conv(data, off, wgt, precision='f16')),
                 ('nhwc_bf16_fast_path', lambda: ops.deform_conv_nhwc(bf, off, wgt, relu=True))):
    us = t_us(fn)
    print(json.dumps(dict(kernel='deform_conv_fwd', path=name, us=round(us, 2), gflop=round(gf, 2), tflops=round(gf / us * 1e3, 1),
                          frac_of_measured_bf16_peak=round(gf / us * 1e3 / peak, 4))), flush=True)
