"""BASELINE.json configs[4]: relation-module sweep N=M in {100,300,1000,3000} x d in {256,1024} x H in {4,16}.
For every point: module microseconds (CUDA events, 256 MB L2 flush between launches, median of 20) and, where the fused
tcgen05 kernel applies (d/H == 64), the attention stage alone with achieved TFLOP/s = 4*N*M*d / t against the measured
bf16 peak.  Writes one JSON line per point to stdout (redirect into gpurun_out/)."""
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
from oracle import relation_np as R

dev = torch.device('cuda:0')
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['bf16_tflops'] \
    if os.path.exists('MEASURED_PEAKS.json') else 1590.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
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
    return ts[len(ts) // 2]


for N in (100, 300, 1000, 3000):
    for d in (256, 1024):
        for H in (4, 16):
            c = R.make_relation_case(N * 31 + d + H, N, d, H)
            t = [torch.from_numpy(c[k]).to(dev) for k in ('X', 'boxes', 'Wq', 'bq', 'Wk', 'bk', 'Wg', 'bg', 'Wout', 'bout')]
            tc = ops.relation_tc_supported(d, d, H)
            prec = 'f16' if tc else 'fp32'
            us_mod = timeit(lambda: ops.relation(*t, group=H, residual_relu=True, precision=prec))
            rec = dict(N=N, d=d, H=H, dk=d // H, path='tcgen05 fused' if tc else 'fp32 library GEMMs (dk != 64)',
                       module_us=round(us_mod, 1), F_tc_gflop=round(4.0 * N * N * d / 1e9, 4))
            if tc:
                for name, mask in (('attn_us', 4), ('geom_us', 2), ('proj_us', 1)):
                    rec[name] = round(timeit(lambda: ops.relation(*t, group=H, residual_relu=True, precision='f16',
                                                                   stage_mask=mask)), 1)
                ach = 4.0 * N * N * d / (rec['attn_us'] * 1e-6) / 1e12
                rec.update(attn_tflops=round(ach, 2), attn_frac_of_measured_peak=round(ach / peak, 4),
                           module_tflops_incl_proj=round((4.0 * N * N * d + 2.0 * 3 * N * d * d) / (us_mod * 1e-6) / 1e12, 2))
            print(json.dumps(rec), flush=True)
